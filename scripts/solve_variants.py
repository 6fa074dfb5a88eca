"""Time candidate formulations of the float64 posterior solve at D = 4096 (and the
per-rank share of a distributed inverse for world = 8) on one GPU."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from revrand_b200 import _engine as eng

D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rs = np.random.RandomState(0)
A = torch.from_numpy(rs.randn(D, 2 * D)).cuda()
G = (A @ A.T) / D
iC = G / 0.02 + torch.eye(D, dtype=torch.float64, device="cuda")
I = torch.eye(D, dtype=torch.float64, device="cuda")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def rec_chol(Am, leaf):
    """In-place recursive Cholesky (lower) of the SPD matrix Am."""
    n = Am.shape[0]
    if n <= leaf:
        Am.copy_(torch.linalg.cholesky_ex(Am)[0])
        return
    h = n // 2
    rec_chol(Am[:h, :h], leaf)
    L11 = Am[:h, :h]
    # L21 = A21 L11^-T
    Am[h:, :h] = torch.linalg.solve_triangular(L11, Am[h:, :h].T, upper=False).T
    Am[h:, h:] -= Am[h:, :h] @ Am[h:, :h].T
    rec_chol(Am[h:, h:], leaf)


print("fp64 GEMM %d^3: %.2f ms" % (D, timed(lambda: A[:, :D] @ A[:, D:])))
print("potrf (cholesky_ex): %.2f ms" % timed(lambda: torch.linalg.cholesky_ex(iC)))
for leaf in (512, 1024, 2048):
    def f():
        W = iC.clone()
        rec_chol(W, leaf)
        return W
    W = f()
    L = torch.linalg.cholesky(iC)
    err = float((torch.tril(W) - L).abs().max() / L.abs().max())
    print("recursive Cholesky leaf %d: %.2f ms (incl. clone; max rel dev %.1e)" % (leaf, timed(f), err))
L = torch.linalg.cholesky(iC)
print("blocked_spd_inverse: %.2f ms" % timed(lambda: eng.blocked_spd_inverse(L)))
print("cholesky_inverse (potri): %.2f ms" % timed(lambda: torch.cholesky_inverse(L)))
print("cholesky_solve(I): %.2f ms" % timed(lambda: torch.cholesky_solve(I, L)))
Li = torch.empty_like(L)
print("tri_inv_lower (replicated part): %.2f ms" % timed(lambda: eng._tri_inv_lower(L, Li)))
ref_inv = Li.clone()
for leaf in (32, 64, 128, 256, 512, 1024):
    eng._BLOCK_INV_LEAF = leaf
    tm = timed(lambda: eng._tri_inv_lower(L, Li))
    print("   batched, leaf %4d: %.2f ms (dev %.1e)" % (leaf, tm, float((Li - ref_inv).abs().max() / ref_inv.abs().max())))
    nb = D // leaf
    Vd = L.view(nb, leaf, nb, leaf).diagonal(dim1=0, dim2=2).permute(2, 0, 1).contiguous()
    eye = torch.eye(leaf, dtype=L.dtype, device=L.device).expand(nb, leaf, leaf)
    print("        leaves alone (batched trsm): %.2f ms" % timed(lambda: torch.linalg.solve_triangular(Vd, eye, upper=False)))
eng._BLOCK_INV_BATCHED = False
eng._BLOCK_INV_LEAF = 512
print("recursive, leaf 512: %.2f ms" % timed(lambda: eng._tri_inv_lower(L, Li)))
eng._BLOCK_INV_BATCHED = True
A2 = torch.randn(4, 1024, 1024, dtype=torch.float64, device="cuda")
print("bmm 4 x 1024^3 fp64: %.2f ms" % timed(lambda: torch.bmm(A2, A2)))
A3 = torch.randn(8, 512, 512, dtype=torch.float64, device="cuda")
print("bmm 8 x 512^3 fp64: %.2f ms" % timed(lambda: torch.bmm(A3, A3)))
A4 = torch.randn(2048, 2048, dtype=torch.float64, device="cuda")
print("mm 2048^3 fp64: %.2f ms" % timed(lambda: A4 @ A4))
ws = 8
per = D // ws
for r in (0, 3, 7):
    lo, hi = r * per, (r + 1) * per
    print("rank %d of %d:" % (r, ws))
    print("   row GEMM Li[:, blk]^T Li: %.2f ms" % timed(lambda: Li[:, lo:hi].T @ Li))
    E = torch.zeros((D, per), dtype=torch.float64, device="cuda")
    E[lo:hi] = torch.eye(per, dtype=torch.float64, device="cuda")
    print("   cholesky_solve(E_blk): %.2f ms" % timed(lambda: torch.cholesky_solve(E, L)))

    def trsm_pair():
        X = torch.linalg.solve_triangular(L[lo:, lo:], E[lo:], upper=False)
        Y = torch.zeros((D, per), dtype=torch.float64, device="cuda")
        Y[lo:] = X
        return torch.linalg.solve_triangular(L.T, Y, upper=True)
    ref = torch.cholesky_solve(E, L)
    dev = float((trsm_pair() - ref).abs().max() / ref.abs().max())
    print("   trsm pair (zero structure): %.2f ms (dev %.1e)" % (timed(trsm_pair), dev))

    def blk_inv():
        # block column of L^-1 by the GEMM-rich recursion on the trailing matrix, then
        # C[:, blk] = L^-T [0; X] via the replicated... (needs all of Li: skipped)
        T = L[lo:, lo:]
        out = torch.empty_like(T)
        eng._tri_inv_lower(T, out)
        return out
    print("   tri_inv of trailing (D - lo): %.2f ms" % timed(blk_inv))
# float32 image of C from float64: conversion
Cm = torch.cholesky_inverse(L)
print("C.float(): %.2f ms" % timed(lambda: Cm.float()))
print("G / var + diag: %.2f ms" % timed(lambda: (G / 0.02).diagonal().add_(1.0)))
print("value-only solve path: %.2f ms" % timed(lambda: eng.solve_posterior(G, A[:, 0].contiguous(), 0.02, torch.ones(D, dtype=torch.float64, device='cuda'), need_C=False)))
print("full solve path: %.2f ms" % timed(lambda: eng.solve_posterior(G, A[:, 0].contiguous(), 0.02, torch.ones(D, dtype=torch.float64, device='cuda'), need_C=True)))
