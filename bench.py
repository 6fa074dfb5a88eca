#!/usr/bin/env python
"""Benchmark: log-marginal-likelihood evaluations per second.

Metric (BASELINE.json): log-ML evals/sec at N=1e6, d=21, K=2048 random
Fourier features (StandardLinearModel + RandomMatern32(2048), config 2) on
1/2/4/8 B200, next to the reference algorithm on the host CPU.

One "step" = one ``StandardLinearModel._elbo``-equivalent evaluation: value
AND gradients wrt (var, regulariser, lengthscale) -- fixed-point value pass on
the int8 tensor cores, one allreduce (N>1), float64 solve, residual + gradient
pass, second allreduce, host assembly.  Rows of X are sharded contiguously over
the ranks (total N fixed => strong scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload config2|config4|config5|fit]
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "log-ML evals/sec N=1e6 D=21 K=2048 RFF"
EVAL_POINTS = [(1.0, 0.02), (4.0, 0.02), (10.0, 0.02),
               (1.0, 1.0), (4.0, 1.0), (10.0, 1.0)]   # (lengthscale, var)
REG = 1.0
CHECK_POINT = 1      # the evaluation point whose results are printed as "check"


def workload_name(args):
    return ("config2: SLM + RandomMatern32(nbases=%d), N=%d, d=%d, value+grad eval, "
            "isotropic lengthscale" % (args.K, args.N, args.d))


def synthetic(N, d, seed=0):
    """SARCOS-shaped synthetic regression data."""
    rs = np.random.RandomState(seed)
    X = rs.randn(N, d).astype(np.float32)
    w = rs.randn(d)
    y = (np.sin(X.astype(np.float64).dot(w) / 3.0)
         + 0.1 * rs.randn(N)).astype(np.float32)
    return X, y


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(tflops=float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                    hbm=float(j["hbm_gbs"]), src="measured")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                     "--format=csv,noheader,nounits"], capture_output=True,
                    text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(len(r) > 2 + i and r[2 + i].startswith("Active")
                          for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "reasons": reasons}


# ---------------------------------------------------------------------------
# CPU: the reference algorithm on the host cores
# ---------------------------------------------------------------------------

def use_all_host_threads():
    """Give BLAS every host core, whatever OMP_NUM_THREADS says (torchrun sets
    it to 1 for N > 1).  Returns (limiter to keep alive, thread count)."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        lim = threadpool_limits(limits=n)
        got = [p.get("num_threads", 1) for p in threadpool_info()
               if p.get("user_api") == "blas"]
        return lim, (max(got) if got else n)
    except Exception:
        return None, n


def matern32_weights(d, K, seed=1):
    """Student-t(3) frequencies (the sampler of RandomMatern32,
    basis_functions.py:1051-1065); the values do not matter for timing."""
    rs = np.random.RandomState(seed)
    return rs.randn(d, K) * np.sqrt(3.0 / rs.chisquare(3, (K,)))


def cpu_eval(N, d, K, rows, budget):
    """One value+gradient evaluation of the float64 oracle port on ``rows`` rows
    of the workload, inside a time budget (seconds for pass 1, pass 2).  Returns
    (projected seconds for N rows, wall seconds spent, timing dict, fully
    measured?).  Everything but the O(D^3) solve is linear in N."""
    from oracle import oracle as orc
    X, y = synthetic(rows, d)
    X, y = X.astype(np.float64), y.astype(np.float64)
    ls, var = EVAL_POINTS[CHECK_POINT]
    blocks = [dict(kind="trig", W=matern32_weights(d, K), lenscale=ls, cols=None)]
    r = orc.slm_elbo_chunked(X, y, var, [REG], blocks, chunk=20000, budget=budget)
    tm = r["timing"]
    full = rows == N and tm["rows_pass1"] == N and tm["rows_pass2"] == N
    t_eval = (tm["t_pass1"] * N / tm["rows_pass1"] + tm["t_solve"]
              + tm["t_pass2"] * N / max(tm["rows_pass2"], 1))
    wall = tm["t_pass1"] + tm["t_solve"] + tm["t_pass2"]
    return t_eval, wall, tm, full


def unmodified_reference_eval(d, K, rows=100000):
    """Where the reference checkout is importable (the build container, never the
    GPU box): seconds of ONE unmodified ``StandardLinearModel._elbo`` call at the
    largest N whose Phi fits in memory."""
    ref = os.environ.get("REVRAND_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "revrand")):
        return None
    try:
        if not hasattr(np, "asscalar"):
            np.asscalar = lambda a: a.item()   # numpy >= 1.23 shim, utils/base.py:285
        sys.path.insert(0, ref)
        from revrand import StandardLinearModel as RefSLM
        from revrand.basis_functions import RandomMatern32 as RefM32
        X, y = synthetic(rows, d)
        X, y = X.astype(np.float64), y.astype(np.float64)
        slm = RefSLM(basis=RefM32(nbases=K, Xdim=d, random_state=1))
        slm.obj_ = -np.inf
        ls, var = EVAL_POINTS[CHECK_POINT]
        t0 = time.perf_counter()
        slm._elbo(X, y, var, REG, ls)
        return {"rows": rows, "seconds": time.perf_counter() - t0}
    except Exception as e:   # the reference arm must not die on a shim problem
        return {"error": repr(e)}


def run_reference(args):
    """--impl reference: the reference's algorithm for the path (float64 NumPy,
    row-chunked exactly as SURVEY Appendix A because Phi for N=1e6 is 33 GB) on
    the host cores of this box, for the same metric and config.  ONE evaluation
    over ALL N rows is timed; a time budget bounds the run on a slow host, and
    the line says whether any part had to be extrapolated."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    lim, threads = use_all_host_threads()
    b = float(args.ref_budget)
    t_eval, wall, tm, full = cpu_eval(args.N, args.d, args.K, args.N, (0.42 * b, 0.58 * b))
    val = 1.0 / t_eval
    desc = ("oracle port of slm._elbo (value+grad, isotropic lengthscale), float64, "
            "20000-row chunks: pass 1 %d rows in %.1f s, solve (D=%d) %.1f s, pass 2 %d "
            "rows in %.1f s%s" % (tm["rows_pass1"], tm["t_pass1"], 2 * args.K,
                                  tm["t_solve"], tm["rows_pass2"], tm["t_pass2"],
                                  "" if full else "; remaining rows extrapolated linearly"))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": 1, "warmup": 0,
        "requested": {"steps": args.steps, "warmup": args.warmup},
        "ms_per_step": 1e3 * wall, "projected_ms_per_eval": 1e3 * t_eval,
        "fully_measured": bool(full),
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args)},
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": threads,
                         "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    um = unmodified_reference_eval(args.d, args.K) if args.unmodified else None
    if um is not None:
        line["unmodified_reference"] = um
    print(json.dumps(line), flush=True)
    del lim


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------

def assemble(r, N, D, var):
    """-ELBO, d(-ELBO)/dvar from the pieces ``_SLMProblem.evaluate`` returns
    (slm.py:165-171, :183; one regulariser)."""
    lam = r["lam"]
    nelbo = 0.5 * (N * np.log(2 * np.pi * var) + r["sqerr"] / var + r["trgc"] / var
                   + (r["q"] / lam[0]).sum() + r["logdet"] + np.log(lam).sum() - D)
    dvar = -0.5 * (-N + (r["sqerr"] + r["trgc"]) / var) / var
    return float(nelbo), float(dvar)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from revrand_b200 import StandardLinearModel, _cabi, _engine, config
    from revrand_b200.basis_functions import RandomMatern32
    from revrand_b200.slm import _SLMProblem

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local))
    lib = _cabi.load()
    N, d, K = args.N, args.d, args.K
    D = 2 * K
    X, y = synthetic(N, d)
    basis = RandomMatern32(nbases=K, Xdim=d, random_state=1)
    prob = _SLMProblem(basis, X, y)          # shards rows over ranks
    lo, hi = _engine.shard_rows(N, rank, world)
    n_local = hi - lo

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def one_eval(i):
        ls, var = EVAL_POINTS[i % len(EVAL_POINTS)]
        return prob.evaluate(var, [REG], [ls], want_grad=True)

    # ---- device-resident throughput ("value") -------------------------------
    for i in range(args.warmup):
        one_eval(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.rr_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.zero_()                       # evict L2 between timed steps
        ev[i][0].record()
        one_eval(i)
        ev[i][1].record()
    barrier()
    launches = lib.rr_launch_count() - l0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    ms_per_step = ms / args.steps
    value = 1e3 / ms_per_step
    # One nvidia-smi query takes longer than a short timed region (80 ms at 8 GPUs):
    # keep the identical load running, untimed, until the sampler has seen it for
    # ~1.5 s (the same count on every rank: the evaluations hold collectives).
    clock_extra = 0
    if ms < 1500.0:
        clock_extra = int(np.ceil((1500.0 - ms) / ms_per_step))
        for i in range(clock_extra):
            one_eval(i)
        barrier()
    sampler.stop_flag = True

    # ---- results of one evaluation point: identical digits at every N --------
    ls_c, var_c = EVAL_POINTS[CHECK_POINT]
    rc = prob.evaluate(var_c, [REG], [ls_c], want_grad=True)
    nelbo_c, dvar_c = assemble(rc, N, D, var_c)
    check = {"point": {"lenscale": ls_c, "var": var_c}, "neg_elbo": nelbo_c,
             "dvar": dvar_c, "m_norm": float(rc["m"].norm().item()),
             "logdet": float(rc["logdet"]),
             "dl": float(rc["g"][0].sum() / (var_c * ls_c ** 2))}

    # ---- roofline of the dominant pass (fixed-point value pass) --------------
    st = prob.stats
    prob.plan.set_lenscales([ls_c])
    for _ in range(2):
        st.zero_()
        _engine.slm_suffstats(prob.plan, prob.Xd, prob.yd, st, engine=prob.engine,
                              want_yy=False)
    kev = []
    kept = prob._kept_buffer()
    for _ in range(max(3, min(args.steps, 5))):
        st.zero_()
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if kept is not None:      # what a timed evaluation runs: value pass + kept fp16 image
            _engine.slm_suffstats_keep(prob.plan, prob.Xd, prob.yd, st, kept, want_yy=False)
        else:
            _engine.slm_suffstats(prob.plan, prob.Xd, prob.yd, st, engine=prob.engine,
                                  want_yy=False)
        b.record()
        kev.append((a, b))
    torch.cuda.synchronize()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    flops = 2.0 * n_local * D * D + 2.0 * n_local * d * K   # algorithmic, per launch
    pk = peaks()
    achieved = flops / (k_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor",
                "kernel": "value pass: t3_syrk_kernel (tcgen05 kind::i8 Phi^T Phi) with "
                          "t3_digits_kernel overlapped on the helper stream"
                          + (" (also writing the fp16 feature image the gradient pass reads)"
                             if kept is not None else ""),
                "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops"], "traffic": None,
                "peak_source": pk["src"] + " bf16 sustained",
                "ms_per_launch": k_ms,
                "algorithmic_flops_per_launch": flops}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            tb = tj.get("t3_value_pass_bytes_per_launch" if kept is not None
                        else "t3_value_pass_without_kept_image_bytes_per_launch")
            if tb is not None:
                # captured at 1e6 rows on one GPU (profiles/traffic.json); the digit image
                # dominates, so a rank's share scales with its rows
                roofline["traffic"] = int(tb * n_local / 1e6)
                roofline["traffic_note"] = ("ncu dram bytes read+written, all kernels of one "
                                            "value pass at 1e6 rows, scaled to this rank's rows")
        except Exception:
            pass

    # ---- value-only evaluations (random-start phase of fit: no gradient pass,
    #      no explicit inverse) ---------------------------------------------------
    for i in range(2):
        prob.evaluate(EVAL_POINTS[i][1], [REG], [EVAL_POINTS[i][0]], want_grad=False)
    barrier()
    nvo = max(3, min(args.steps, 6))
    evv = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(nvo)]
    for i in range(nvo):
        flush.zero_()
        ls, var = EVAL_POINTS[i % len(EVAL_POINTS)]
        evv[i][0].record()
        prob.evaluate(var, [REG], [ls], want_grad=False)
        evv[i][1].record()
    barrier()
    tv = torch.tensor([sum(a.elapsed_time(b) for a, b in evv)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
    value_only = 1e3 / (float(tv.item()) / nvo)

    # ---- end to end: the call a user makes, StandardLinearModel._elbo(X, y, var,
    #      reg, lenscale) with HOST arrays (pinned), rows uploaded on every call and
    #      the results read back ----------------------------------------------------
    del prob
    Xh = torch.from_numpy(X).pin_memory()
    yh = torch.from_numpy(y).pin_memory()
    Xn, yn = Xh, yh                          # pinned host tensors (array-likes of the API)
    slm = StandardLinearModel(basis=basis)
    slm.obj_ = -np.inf
    old_cache = config.CACHE_DEVICE_DATA
    config.CACHE_DEVICE_DATA = False         # H2D of X and y inside every call

    def e2e_step(i):
        ls, var = EVAL_POINTS[i % len(EVAL_POINTS)]
        nelbo, grads = slm._elbo(Xn, yn, var, REG, ls)
        return nelbo
    for i in range(3):
        e2e_step(i)
    barrier()
    n_e2e = max(3, min(args.steps, 6))
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(n_e2e)]
    for i in range(n_e2e):
        ev2[i][0].record()
        e2e_step(i)
        ev2[i][1].record()
    barrier()
    config.CACHE_DEVICE_DATA = old_cache
    ms2 = sum(a.elapsed_time(b) for a, b in ev2)
    t2 = torch.tensor([ms2], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_val = 1e3 / (float(t2.item()) / n_e2e)
    h2d = int(n_local * d * 4 + n_local * 4 + d * K * 4)
    d2h = int(8 * (4 + 1 + d))

    line = {
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "s8",
        "data": "synthetic",
        "config": {"workload": workload_name(args)},
        "config_detail": {
            "arithmetic": "value pass: 24-bit fixed-point features as three int8 digits, "
                          "tcgen05 kind::i8, exact int32 accumulation in TMEM; gradient pass: "
                          "kind::f16, fp32 accumulate; f64 statistics and solve",
            "rows_per_gpu": n_local, "l2": "flushed between timed steps (256 MiB write)",
            "engine": os.environ.get("REVRAND_B200_ENGINE", "auto"),
            "parallelism": "rows sharded x%d, 2 allreduces/eval" % world},
        "clocks": dict(sampler.summary(), sampled_over_steps=args.steps + clock_extra),
        "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "call": "StandardLinearModel._elbo(X, y, var, reg, lenscale), host arrays"},
        "gpu_launches": int(launches),
        "value_only": {"value": value_only, "unit": "evals/s",
                       "note": "log-ML value without gradients (random-start phase of fit)"},
        "roofline": roofline,
        "check": check,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        lim, threads = use_all_host_threads()
        t_eval, wall, tm, _ = cpu_eval(N, d, K, args.cpu_sample, None)
        line["cpu_baseline"] = {
            "value": 1.0 / t_eval, "unit": "evals/s", "cores": threads,
            "kind": "port",
            "sample": "oracle port of slm._elbo (value+grad) on %d of %d rows: "
                      "%.1f s measured (solve %.1f s), linear extrapolation in N; "
                      "`--impl reference` times all N rows"
                      % (args.cpu_sample, N, wall, tm["t_solve"])}
        del lim
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config4", "config5", "fit"])
    ap.add_argument("--N", type=int, default=None)
    ap.add_argument("--Ks", default="512,1024,2048,4096,8192",
                    help="config5: comma-separated list of nbases to sweep")
    ap.add_argument("--d", type=int, default=21)
    ap.add_argument("--K", type=int, default=2048)
    ap.add_argument("--cpu-sample", type=int, default=40000)
    ap.add_argument("--ref-budget", type=float, default=420.0,
                    help="--impl reference: seconds available for the timed evaluation")
    ap.add_argument("--unmodified", action="store_true",
                    help="--impl reference: also time the unmodified reference where importable")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.N is None:
        args.N = 10000000 if args.workload == "config5" else 1000000
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.workload == "config4":
        if args.steps == 6 and args.warmup == 3:      # this file's defaults suit config 2:
            args.steps, args.warmup = 200, 10         # an SVI step is 0.75 ms
        import bench_glm
        return bench_glm.main(args)
    if args.workload == "config5":
        import bench_config5
        return bench_config5.main(args)
    if args.workload == "fit":
        import bench_fit
        return bench_fit.main(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        import __graft_entry__ as g
        g.build()
        run_ours(args)


if __name__ == "__main__":
    main()
