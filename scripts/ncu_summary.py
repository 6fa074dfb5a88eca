"""Summarise an ncu report (raw page) into the handful of metrics DESIGN.md /
profiles/ quote.  usage: ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr, units = rows[hi], rows[hi + 1]
KEYS = """gpu__time_duration.sum launch__registers_per_thread launch__grid_size
launch__block_size launch__cluster_size launch__shared_mem_per_block
dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
lts__throughput.avg.pct_of_peak_sustained_elapsed
lts__t_sectors_op_red.sum lts__t_sectors_op_atom.sum
smsp__inst_executed.sum sm__cycles_active.avg sm__cycles_elapsed.max
smsp__mem_tensor_reads_op_utcmma_matrix_c.sum smsp__mem_tensor_writes_op_utcmma.sum""".split()
for r in rows[hi + 2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:100])
    for k in hdr:
        if k in KEYS or ("issue_stalled" in k and k.endswith("per_issue_active.ratio")):
            print("   %-88s %s %s" % (k, d[k], units[hdr.index(k)]))
