"""Debug experiments: time the fused value pass built with each RR_T2_EXP_*
macro (results are garbage; only the timing matters)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "revrand_b200", "csrc")
VARIANTS = {"base": [], "backoff20": ["-DRR_MBAR_BACKOFF=20"], "backoff100": ["-DRR_MBAR_BACKOFF=100"],
            "hint": ["-DRR_MBAR_SUSPEND_HINT"], "base2": []}
def lib(v):
    return os.path.join(ROOT, "revrand_b200", "lib", "librevrand_b200_exp_%s.so" % v)
if "--build" in sys.argv:
    for v, fl in VARIANTS.items():
        subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-lineinfo", "-Xcompiler", "-fPIC", "-shared", "-o", lib(v)] + fl
                              + sorted(glob.glob(os.path.join(CSRC, "*.cu"))), cwd=CSRC)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] in VARIANTS:
    sys.path.insert(0, ROOT)
    import torch
    from revrand_b200 import _cabi
    _cabi.LIB_PATH = lib(sys.argv[1])
    from revrand_b200 import _engine as eng
    from revrand_b200.basis_functions import RandomMatern32
    from revrand_b200.slm import _SLMProblem
    from bench import synthetic
    X, y = synthetic(1000000, 21)
    prob = _SLMProblem(RandomMatern32(nbases=2048, Xdim=21, random_state=1), X, y)
    prob.plan.set_lenscales([4.0])
    ts = []
    for _ in range(8):
        prob.stats.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.slm_suffstats(prob.plan, prob.Xd, prob.yd, prob.stats, engine=prob.engine)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("%-14s N=1e6 suffstats ms: %s" % (sys.argv[1], " ".join("%.2f" % t for t in ts)), flush=True)
else:
    for v in VARIANTS:
        subprocess.call([sys.executable, os.path.abspath(__file__), v])
