"""Opcode histogram per kernel of the shipped library (cuobjdump -sass): the
Blackwell-specific mnemonics that show which engine a kernel uses.
usage: sass_histogram.py [lib.so] > profiles/sass_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "revrand_b200", "lib", "librevrand_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = re.compile(r"^(UTC[A-Z0-9]*MMA|UTCBAR|UTCCP|UTCATOMSWS|LDTM|STTM|UBLKCP|UTMALDG|UTMASTG|UTMAPF|LDGSTS|"
                   r"SYNCS|MUFU|HMMA|IMMA|DMMA|F2I|I2F|RED|ATOM|ATOMG|LDS|STS|STG|LDG|FFMA2?|FADD2?|FMUL2?|PRMT|"
                   r"ELECT|MEMBAR|CCTL|BAR|UCGABAR_ARV|UCGABAR_WAIT)\b")
fn = None
hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(.*", "", fn)
        hist[fn] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
    if m and fn:
        op = m.group(1)
        base = op.split(".")[0]
        hist[fn]["_total"] += 1
        if WATCH.match(base):
            key = op if base.startswith(("UTC", "LDTM", "UBLKCP", "UTMA", "STTM", "MUFU", "RED", "ATOM")) else base
            hist[fn][key] += 1
print("SASS opcode histogram of", os.path.relpath(lib, ROOT), "(sm_100a)")
print("tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP, tensor-map TMA -> UTMALDG/UTMASTG,")
print("cp.async -> LDGSTS, mbarrier -> SYNCS; absent opcodes are not listed\n")
for fn, h in hist.items():
    if not fn.startswith("rr::") and "rr::" not in fn:
        continue
    print("%s   [%d instructions]" % (fn, h["_total"]))
    items = [(k, v) for k, v in h.items() if k != "_total"]
    items.sort(key=lambda kv: (not kv[0].startswith(("UTC", "LDTM", "UBLKCP", "UTMA", "STTM")), kv[0]))
    print("    " + "  ".join("%s=%d" % kv for kv in items))
