"""GPU diagnostic: how does tcgen05 round fp32 accumulation?  (see
revrand_b200/csrc/rr_tc_probe.cu).  Prints the loss per accumulating MMA in
units of ulp(accumulator) for diagonal (all-positive) and off-diagonal entries
and compares with a round-toward-zero-per-instruction model."""
import ctypes as C
import json
import sys
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
from revrand_b200 import _cabi

lib = _cabi.load()
rs = np.random.RandomState(0)
th = rs.uniform(0, 2 * np.pi, size=(256, 64))
c = np.cos(th)
hi = c.astype(np.float16)
lo = (c - hi.astype(np.float64)).astype(np.float16)
Ahi, Alo = np.ascontiguousarray(hi[:128]), np.ascontiguousarray(lo[:128])
Bhi, Blo = np.ascontiguousarray(hi), np.ascontiguousarray(lo)   # B rows 0..127 == A rows


def rz32(x):
    y = x.astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(x)
    y[over] = np.nextafter(y[over], np.float32(0))
    return y


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


out = {}
h64, l64 = hi.astype(np.float64), lo.astype(np.float64)
for mode in (0, 1, 2):
    for reps in (8, 32, 128, 512):
        D = np.zeros((128, 256), np.float32)
        X = np.zeros((128, 256), np.float32)
        rc = lib.rr_tcgen05_accum_probe(ptr(Ahi), ptr(Alo), ptr(Bhi), ptr(Blo), reps, mode, ptr(D), ptr(X))
        assert rc == 0, lib.rr_last_error()
        # exact and RZ-per-instruction model
        acc = np.zeros((128, 256), np.float32)
        aux = np.zeros((128, 256), np.float32)
        exact = np.zeros((128, 256))
        for r in range(reps):
            for k in range(4):
                s = slice(16 * k, 16 * k + 16)
                P = h64[:128, s] @ h64[:, s].T
                Q1 = l64[:128, s] @ h64[:, s].T
                Q2 = h64[:128, s] @ l64[:, s].T
                acc = rz32(acc.astype(np.float64) + P)
                exact += P
                if mode == 1:
                    acc = rz32(acc.astype(np.float64) + Q1)
                    acc = rz32(acc.astype(np.float64) + Q2)
                    exact += Q1 + Q2
                elif mode == 2:
                    aux = rz32(aux.astype(np.float64) + Q1)
                    aux = rz32(aux.astype(np.float64) + Q2)
        dg = np.arange(128)
        got = D.astype(np.float64)
        nadd = reps * 4 * (3 if mode == 1 else 1)
        ulp = np.spacing(np.abs(D[dg, dg]).astype(np.float32)).astype(np.float64)
        loss_hw = (exact[dg, dg] - got[dg, dg])
        loss_model = (exact[dg, dg] - acc.astype(np.float64)[dg, dg])
        off = np.ones_like(exact, bool)
        off[dg, dg] = False
        rec = dict(mode=mode, reps=reps, rows=64 * reps, nadd=nadd,
                   diag_rel_hw=float(np.mean(loss_hw / exact[dg, dg])),
                   diag_rel_model=float(np.mean(loss_model / exact[dg, dg])),
                   diag_loss_per_add_in_final_ulp_hw=float(np.mean(loss_hw / ulp) / nadd),
                   diag_loss_per_add_in_final_ulp_model=float(np.mean(loss_model / ulp) / nadd),
                   off_rms_err_hw=float(np.sqrt(np.mean((got - exact)[off] ** 2))),
                   off_rms_err_model=float(np.sqrt(np.mean((acc.astype(np.float64) - exact)[off] ** 2))),
                   off_mean_signed_hw=float(np.mean(((got - exact) * np.sign(exact))[off])),
                   hw_equals_model=bool(np.array_equal(D, acc)))
        if mode == 2:
            rec["aux_rms_err_hw"] = float(np.sqrt(np.mean((X.astype(np.float64) - (reps * sum(
                l64[:128, 16*k:16*k+16] @ h64[:, 16*k:16*k+16].T + h64[:128, 16*k:16*k+16] @ l64[:, 16*k:16*k+16].T
                for k in range(4)))) ** 2)))
            rec["aux_equals_model"] = bool(np.array_equal(X, aux))
        print(json.dumps(rec), flush=True)
