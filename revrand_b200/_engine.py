"""Device engine: feature plans, buffers and launches of the C-ABI kernels.

PyTorch is used for device memory, streams, the (single GPU) dense solve and
``torch.distributed``; every data-parallel pass goes through
``librevrand_b200.so``.  Nothing here falls back to the CPU.
"""

from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _cabi
from ._cabi import RRPlan, RevrandB200Error, check

CHOLTHRESH = 1e-5   # revrand/mathfun/linalg.py:31
SVD_FLOOR = 1e-15   # revrand/mathfun/linalg.py:128 (s_tol)
TWO_PI = 2.0 * math.pi
WT_CLIP = 1e30      # bound on |W / lenscale / 2pi| sent to the device (see Wt_host)

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise RevrandB200Error(
            "revrand_b200 needs a CUDA device (B200, sm_100a); there is no CPU "
            "fallback for the feature / likelihood passes")
    return t


def device():
    t = require_cuda()
    return t.device("cuda", t.cuda.current_device())


def _stream_ptr():
    return C.c_void_p(torch().cuda.current_stream().cuda_stream)


def auto_min_rows():
    """Row count from which RR_ENGINE_AUTO uses the tensor-core engine (asked
    of the library: one source for the constant)."""
    return int(_cabi.load().rr_engine_auto_min_rows())


_FUSED16 = (_cabi.RR_ENGINE_TCGEN05_FINE, _cabi.RR_ENGINE_TCGEN05_FUSED16)

# rr_context per (thread, device): the caller-owned helper stream + events that
# let a pass overlap its feature generation with its tensor-core work.
import threading as _threading

_ctx_local = _threading.local()


class _Context(object):
    def __init__(self):
        h = C.c_void_p()
        check(_cabi.load().rr_context_create(C.byref(h)), "rr_context_create")
        self.handle = h

    def __del__(self):
        try:
            if self.handle:
                _cabi.load().rr_context_destroy(self.handle)
        except Exception:
            pass


def context():
    """The rr_context of this thread for the current device (None while a CUDA
    graph is being captured elsewhere is the caller's business)."""
    t = require_cuda()
    dev = t.cuda.current_device()
    tab = getattr(_ctx_local, "tab", None)
    if tab is None:
        tab = _ctx_local.tab = {}
    if dev not in tab:
        tab[dev] = _Context()
    return tab[dev].handle


def _ptr(tensor):
    return C.c_void_p(0 if tensor is None else tensor.data_ptr())


def to_device(a, dtype=None):
    """numpy / tensor -> contiguous CUDA tensor (float32 unless told)."""
    t = require_cuda()
    dtype = dtype or t.float32
    if isinstance(a, t.Tensor):
        return a.to(device=device(), dtype=dtype).contiguous()
    a = np.ascontiguousarray(a)
    return t.from_numpy(a).to(device=device(), dtype=dtype).contiguous()


# ---------------------------------------------------------------------------
# Feature plan
# ---------------------------------------------------------------------------

class TrigBlock(object):
    """One random trigonometric basis inside a concatenation.

    ``W`` is the raw (d_eff, K) frequency matrix, ``cols`` the input columns it
    applies to (``apply_ind``) or None, ``lenscale`` a float array of length
    1 or d_eff.
    """
    kind = "trig"

    def __init__(self, W, lenscale, cols=None, amp=None):
        self.W = np.asarray(W, dtype=np.float64)
        self.cols = None if cols is None else np.asarray(cols, dtype=np.int64)
        self.lenscale = np.atleast_1d(np.asarray(lenscale, dtype=np.float64))
        self.K = self.W.shape[1]
        self.width = 2 * self.K
        # feature amplitude; None = 1 / sqrt(K) (basis_functions.py:864)
        self.amp = amp


class ExtraBlock(object):
    """Affine columns: ``src[j] >= 0`` copies X[:, src[j]] (raised to the integer
    power ``pow[j]`` when given), else the constant ``val[j]``."""
    kind = "extra"

    def __init__(self, src, val, pow=None):
        self.src = np.asarray(src, dtype=np.int32)
        self.val = np.asarray(val, dtype=np.float32)
        self.pow = (np.ones(len(self.src), dtype=np.int32) if pow is None
                    else np.asarray(pow, dtype=np.int32))
        self.width = len(self.src)


class FeaturePlan(object):
    """Host description + device image of a concatenated feature map."""

    def __init__(self, blocks, d):
        t = require_cuda()
        self.blocks = list(blocks)
        self.d = int(d)
        self.trig = [b for b in self.blocks if b.kind == "trig"]
        self.ktot = int(sum(b.K for b in self.trig))
        self.D = int(sum(b.width for b in self.blocks))
        col_cos, col_sin, amp = [], [], []
        ext_src, ext_val, ext_col, ext_pow = [], [], [], []
        self.block_offsets = []
        self.freq_offsets = []
        off = 0
        koff = 0
        for b in self.blocks:
            self.block_offsets.append(off)
            if b.kind == "trig":
                self.freq_offsets.append(koff)
                col_cos.append(off + np.arange(b.K))
                col_sin.append(off + b.K + np.arange(b.K))
                amp.append(np.full(b.K, 1.0 / math.sqrt(b.K) if b.amp is None
                                   else float(b.amp)))
                koff += b.K
            else:
                ext_src.append(b.src)
                ext_val.append(b.val)
                ext_pow.append(b.pow)
                ext_col.append(off + np.arange(b.width))
            off += b.width
        self.next = int(sum(len(s) for s in ext_src))

        def cat(lst, dtype):
            return (np.concatenate(lst).astype(dtype) if lst
                    else np.zeros(0, dtype=dtype))
        i32, f32 = t.int32, t.float32
        self._col_cos = to_device(cat(col_cos, np.int32), i32)
        self._col_sin = to_device(cat(col_sin, np.int32), i32)
        self._amp = to_device(cat(amp, np.float32), f32)
        self._ext_src = to_device(cat(ext_src, np.int32), i32)
        self._ext_val = to_device(cat(ext_val, np.float32), f32)
        self._ext_col = to_device(cat(ext_col, np.int32), i32)
        pw = cat(ext_pow, np.int32)
        self._ext_pow = to_device(pw, i32) if (len(pw) and np.any(pw != 1)) else None
        # full-d raw frequency matrix (zeros outside each block's columns)
        Wfull = np.zeros((self.d, max(self.ktot, 1)))
        for b, ko in zip(self.trig, self.freq_offsets):
            rows = np.arange(self.d) if b.cols is None else b.cols
            Wfull[rows, ko:ko + b.K] = b.W
        self.Wfull = Wfull[:, :self.ktot] if self.ktot else Wfull[:, :0]
        self._Wfull_dev = to_device(self.Wfull, t.float64)
        self._Wt = t.zeros((self.d, max(self.ktot, 1)), dtype=f32,
                           device=device())
        self.struct = RRPlan()
        self.struct_tc = None      # extended plan (affine columns as slots), see below
        self._col_scale = None     # job-wide fixed-point scales (set_col_scale)
        self._ext_host = (cat(ext_src, np.int32), cat(ext_val, np.float32),
                          cat(ext_col, np.int32))
        self.refresh()

    def enable_tc_extras(self, col_absmax):
        """Let the fused tcgen05 value pass carry the affine (Linear / Bias)
        columns as pseudo-frequency slots (rr_plan.kind): slot j projects
        u = X[:, src] / max|X[:, src]| (|u| <= 1, as the fixed-point split of
        the kernel requires) and its amplitude undoes the scale; constant
        columns are slots of kind 2.  ``col_absmax``: (d,) max |X[:, i]| over
        ALL rows of the job (all ranks)."""
        t = torch()
        if not (self.next and self.ktot) or self._ext_pow is not None:
            return
        src, val, col = self._ext_host
        ktx = self.ktot + self.next
        scale = np.maximum(np.asarray(col_absmax, dtype=np.float64), 1e-30)
        Wx = np.zeros((self.d, self.next), dtype=np.float32)
        amp = np.zeros(self.next, dtype=np.float32)
        kind = np.zeros(ktx, dtype=np.uint8)
        for j in range(self.next):
            if src[j] >= 0:
                Wx[src[j], j] = 1.0 / scale[src[j]]
                amp[j] = scale[src[j]]
                kind[self.ktot + j] = 1
            else:
                amp[j] = val[j]
                kind[self.ktot + j] = 2
        i32, f32 = t.int32, t.float32
        self._Wt_x = t.zeros((self.d, ktx), dtype=f32, device=device())
        self._Wt_x[:, self.ktot:] = to_device(Wx, f32)
        self._amp_x = t.cat([self._amp, to_device(amp, f32)])
        self._col_cos_x = t.cat([self._col_cos, to_device(col, i32)])
        self._col_sin_x = t.cat([self._col_sin,
                                 t.full((self.next,), -1, dtype=i32, device=device())])
        self._kind_x = to_device(kind, t.uint8)
        s = RRPlan()
        s.d, s.ktot, s.next, s.D = self.d, ktx, 0, self.D
        s.Wt = self._Wt_x.data_ptr()
        s.amp = self._amp_x.data_ptr()
        s.col_cos = self._col_cos_x.data_ptr()
        s.col_sin = self._col_sin_x.data_ptr()
        s.ext_src = s.ext_val = s.ext_col = s.ext_pow = None
        s.kind = self._kind_x.data_ptr()
        self.struct_tc = s
        self.refresh()

    def set_col_scale(self, scale):
        """Fixed-point scales of the int8 value pass (rr_plan.col_scale): a (d+1,)
        float32 device tensor, max |X[:, i]| and max |y| over ALL rows of the job,
        or None (each call then finds the maxima of its own rows)."""
        self._col_scale = scale
        self.struct.col_scale = None if scale is None else scale.data_ptr()

    def set_lenscales(self, lenscales):
        """New lengthscale per trig block (scalar or (d_eff,) each)."""
        assert len(lenscales) == len(self.trig)
        for b, ls in zip(self.trig, lenscales):
            b.lenscale = np.atleast_1d(np.asarray(ls, dtype=np.float64))
        self.refresh()

    def inv_lenscale_full(self):
        """(d, ktot) array of 1/lenscale seen by each (input dim, frequency)."""
        inv = np.zeros((self.d, max(self.ktot, 1)))
        for b, ko in zip(self.trig, self.freq_offsets):
            rows = np.arange(self.d) if b.cols is None else b.cols
            ls = b.lenscale if len(b.lenscale) > 1 else np.full(len(rows),
                                                                b.lenscale[0])
            inv[rows, ko:ko + b.K] = (1.0 / ls)[:, None]
        return inv[:, :self.ktot]

    def Wt_host(self, lenscales=None):
        """Wt = W / lenscale / 2pi as a float32 host array (float64 arithmetic); for
        the blocks' current lengthscales, or for ``lenscales`` without changing them."""
        if lenscales is not None:
            keep = [b.lenscale for b in self.trig]
            for b, ls in zip(self.trig, lenscales):
                b.lenscale = np.atleast_1d(np.asarray(ls, dtype=np.float64))
        # (a random start or a line-search step to the edge of the log-warp box may
        # ask for a lengthscale so small that W / l leaves the fp32 range: keep the
        # projection u = x . Wt finite -- WT_CLIP * d * max|x| < 3.4e38 -- so that the
        # exact range reduction returns phase 0 instead of NaN; the phase is
        # meaningless there anyway, as is cos(1e100 x) in float64)
        Wt = np.clip(self.Wfull * self.inv_lenscale_full() / TWO_PI, -WT_CLIP, WT_CLIP)
        if lenscales is not None:
            for b, ls in zip(self.trig, keep):
                b.lenscale = ls
        return np.ascontiguousarray(Wt.astype(np.float32))

    def point_Wt_at(self, tensor=None):
        """Let the plan read its projection from ``tensor`` ((d, ktot) float32 on the
        device, e.g. one of a stack uploaded ahead of a pipelined batch of
        evaluations) instead of its own buffer; None restores the own buffer."""
        self.struct.Wt = (self._Wt if tensor is None else tensor).data_ptr()

    def refresh(self):
        """Recompute Wt = W / lenscale / 2pi (float64 on host) and upload."""
        t = torch()
        if self.ktot:
            self._Wt.copy_(t.from_numpy(self.Wt_host()), non_blocking=False)
        s = self.struct
        s.d, s.ktot, s.next, s.D = self.d, self.ktot, self.next, self.D
        s.Wt = self._Wt.data_ptr()
        s.amp = self._amp.data_ptr()
        s.col_cos = self._col_cos.data_ptr()
        s.col_sin = self._col_sin.data_ptr()
        s.ext_src = self._ext_src.data_ptr()
        s.ext_val = self._ext_val.data_ptr()
        s.ext_col = self._ext_col.data_ptr()
        s.kind = None
        s.ext_pow = None if self._ext_pow is None else self._ext_pow.data_ptr()
        s.col_scale = None if self._col_scale is None else self._col_scale.data_ptr()
        if self.struct_tc is not None:
            self._Wt_x[:, :self.ktot].copy_(self._Wt)

    def tcgen05_ok(self):
        return bool(_cabi.load().rr_tcgen05_supported(self.d, self.ktot,
                                                     self.next, self.D))


# ---------------------------------------------------------------------------
# Workspace cache
# ---------------------------------------------------------------------------

_workspace = {}


def workspace(nbytes):
    t = require_cuda()
    dev = t.cuda.current_device()
    buf = _workspace.get(dev)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _workspace.pop(dev, None)
        buf = t.empty(int(nbytes) + 1024, dtype=t.uint8, device=device())
        _workspace[dev] = buf
    return buf


def _ws_bytes(op, N, plan, aux0=0, aux1=0, engine=_cabi.RR_ENGINE_AUTO):
    return int(_cabi.load().rr_workspace_bytes(op, int(N), plan.d, plan.ktot,
                                               plan.D, aux0, aux1, engine))


# ---------------------------------------------------------------------------
# Kernel wrappers
# ---------------------------------------------------------------------------

def features(plan, Xd):
    t = require_cuda()
    lib = _cabi.load()
    N = Xd.shape[0]
    Phi = t.empty((N, plan.D), dtype=t.float32, device=Xd.device)
    if N == 0:
        return Phi
    check(lib.rr_features(C.byref(plan.struct), _ptr(Xd), N, _ptr(Phi), plan.D,
                          _stream_ptr()), "rr_features")
    return Phi


def trig_grad(Xd, W, lenscale, compat=True):
    """(N, 2K[, d]) gradient tensor of one trig block (reference layout)."""
    t = require_cuda()
    lib = _cabi.load()
    N, d = Xd.shape
    K = W.shape[1]
    ls = np.atleast_1d(np.asarray(lenscale, dtype=np.float64))
    P = len(ls)
    Wd = to_device(W)
    lsd = to_device(ls)
    shape = (N, 2 * K) if P == 1 else (N, 2 * K, P)
    out = t.empty(shape, dtype=t.float32, device=Xd.device)
    check(lib.rr_trig_grad(_ptr(Xd), N, d, _ptr(Wd), K, _ptr(lsd), P,
                           1 if compat else 0, _ptr(out), _stream_ptr()),
          "rr_trig_grad")
    return out


def centre_features(Xd, centres, lenscale, kind, want_grad=False):
    """Radial (kind 0) / sigmoidal (kind 1) basis and, optionally, its
    lengthscale gradient in the reference's layout."""
    t = require_cuda()
    lib = _cabi.load()
    N, d = Xd.shape
    M = centres.shape[0]
    ls = np.atleast_1d(np.asarray(lenscale, dtype=np.float64))
    P = len(ls)
    Cd, lsd = to_device(centres), to_device(ls)
    Phi = t.empty((N, M), dtype=t.float32, device=Xd.device)
    dPhi = None
    if want_grad:
        dPhi = t.empty((N, M) if P == 1 else (N, M, P), dtype=t.float32, device=Xd.device)
    check(lib.rr_centre_features(_ptr(Xd), N, d, _ptr(Cd), M, _ptr(lsd), P, int(kind),
                                 _ptr(Phi), _ptr(dPhi), _stream_ptr()),
          "rr_centre_features")
    return (Phi, dPhi) if want_grad else Phi


def gm_grad(Xd, V, mean, lenscale):
    """(d Phi / d mean, d Phi / d lenscale) of a FastFoodGM component."""
    t = require_cuda()
    lib = _cabi.load()
    N, d = Xd.shape
    n = V.shape[1]
    shape = (N, 4 * n) if d == 1 else (N, 4 * n, d)
    dm = t.empty(shape, dtype=t.float32, device=Xd.device)
    dl = t.empty(shape, dtype=t.float32, device=Xd.device)
    # (keep the uploads alive until the launch: a temporary's block would be
    # recycled by the next upload)
    Vd, md, ld = to_device(V), to_device(mean), to_device(lenscale)
    check(lib.rr_gm_grad(_ptr(Xd), N, d, _ptr(Vd), n, _ptr(md), _ptr(ld), _ptr(dm),
                         _ptr(dl), _stream_ptr()), "rr_gm_grad")
    return dm, dl


def fastfood_features(Xs_d, B, G, PI, S, want_vx=False):
    t = require_cuda()
    lib = _cabi.load()
    N, d = Xs_d.shape
    k, d2 = B.shape
    Bd, Gd, Sd = to_device(B), to_device(G), to_device(S)
    Pd = to_device(PI, t.int32)
    Phi = t.empty((N, 2 * k * d2), dtype=t.float32, device=Xs_d.device)
    VX = (t.empty((N, k * d2), dtype=t.float32, device=Xs_d.device)
          if want_vx else None)
    check(lib.rr_fastfood_features(_ptr(Xs_d), N, d, d2, k, _ptr(Bd), _ptr(Gd),
                                   _ptr(Pd), _ptr(Sd), _ptr(Phi), _ptr(VX),
                                   _stream_ptr()), "rr_fastfood_features")
    return (Phi, VX) if want_vx else Phi


class SuffStats(object):
    """Flat float64 buffer [G (D*D) | p (D) | yy (1)] so that a row-sharded
    job needs exactly one allreduce."""

    def __init__(self, D):
        t = require_cuda()
        self.D = D
        self.flat = t.zeros(D * D + D + 1, dtype=t.float64, device=device())
        self.G = self.flat[:D * D].view(D, D)
        self.p = self.flat[D * D:D * D + D]
        self.yy = self.flat[D * D + D:]

    def zero_(self):
        self.flat.zero_()


def slm_suffstats(plan, Xd, yd, stats, engine=_cabi.RR_ENGINE_AUTO,
                  want_yy=True):
    lib = _cabi.load()
    N = Xd.shape[0]
    nb = _ws_bytes(_cabi.RR_OP_SUFFSTATS, N, plan, engine=engine)
    struct = plan.struct
    if plan.struct_tc is not None and engine in _FUSED16:
        # round-1 fused kind::f16 kernel: affine columns as pseudo-frequency slots
        struct = plan.struct_tc
        nb = max(nb, plan.D * plan.D * 8 + plan.D * 4 + 8192)
    ws = workspace(nb)
    check(lib.rr_slm_suffstats(C.byref(struct), _ptr(Xd), _ptr(yd), N,
                               _ptr(stats.G), _ptr(stats.p),
                               _ptr(stats.yy) if want_yy else C.c_void_p(0),
                               _ptr(ws), ws.numel(), engine, context(),
                               _stream_ptr()),
          "rr_slm_suffstats")


def kept_features_buffer(plan, N, max_bytes):
    """Device buffer for the fp16 feature image the value pass can leave behind for
    the gradient pass of the same evaluation (rr_slm_suffstats_keep), or None when
    the plan / row count cannot keep its features or the image is larger than
    ``max_bytes`` or half of the free device memory."""
    t = require_cuda()
    need = int(_cabi.load().rr_slm_kept_features_bytes(C.byref(plan.struct), int(N)))
    if need == 0 or need > max_bytes:
        return None
    free, _ = t.cuda.mem_get_info()
    if need + 1024 > free // 2:
        return None
    raw = t.empty(need + 1024, dtype=t.uint8, device=device())
    off = (-raw.data_ptr()) % 1024
    return raw[off:off + need]


def slm_suffstats_keep(plan, Xd, yd, stats, kept, want_yy=True):
    """Value pass on the tensor-core engine that also writes the kept feature image."""
    lib = _cabi.load()
    N = Xd.shape[0]
    ws = workspace(_ws_bytes(_cabi.RR_OP_SUFFSTATS, N, plan, engine=_cabi.RR_ENGINE_TCGEN05))
    check(lib.rr_slm_suffstats_keep(C.byref(plan.struct), _ptr(Xd), _ptr(yd), N,
                                    _ptr(stats.G), _ptr(stats.p),
                                    _ptr(stats.yy) if want_yy else C.c_void_p(0),
                                    _ptr(kept), kept.numel(), _ptr(ws), ws.numel(),
                                    context(), _stream_ptr()),
          "rr_slm_suffstats_keep")


def _grad_flags():
    from . import config
    return _cabi.RR_GRAD_SPLIT_C if config.GRADIENT_SPLIT_C else 0


def slm_gradpass_kept(plan, Xd, yd, m32, C32, R, sqerr, kept):
    """Residual + gradient pass from the kept feature image of the same evaluation."""
    lib = _cabi.load()
    N = Xd.shape[0]
    flags = _grad_flags()
    ws = workspace(_ws_bytes(_cabi.RR_OP_GRADPASS_KEPT, N, plan, engine=flags))
    check(lib.rr_slm_gradpass_kept(C.byref(plan.struct), _ptr(Xd), _ptr(yd), N,
                                   _ptr(m32), _ptr(C32), _ptr(R), _ptr(sqerr),
                                   _ptr(kept), kept.numel(), _ptr(ws), ws.numel(),
                                   flags, _stream_ptr()), "rr_slm_gradpass_kept")


def slm_residual(plan, Xd, yd, m32, err=None, sqerr=None):
    t = require_cuda()
    lib = _cabi.load()
    N = Xd.shape[0]
    if sqerr is None:
        sqerr = t.zeros(1, dtype=t.float64, device=Xd.device)
    ws = workspace(_ws_bytes(_cabi.RR_OP_RESIDUAL, N, plan))
    check(lib.rr_slm_residual(C.byref(plan.struct), _ptr(Xd), _ptr(yd), N,
                              _ptr(m32), _ptr(err), _ptr(sqerr), _ptr(ws),
                              ws.numel(), _stream_ptr()), "rr_slm_residual")
    return sqerr


def slm_gradpass(plan, Xd, yd, m32, C32, R, sqerr,
                 engine=_cabi.RR_ENGINE_AUTO):
    """Residual + gradient pass: sqerr += sum (y - Phi m)^2, R += X^T Q."""
    lib = _cabi.load()
    N = Xd.shape[0]
    engine = engine | _grad_flags()
    nb = _ws_bytes(_cabi.RR_OP_GRADPASS, N, plan, engine=engine)
    ws = workspace(nb)
    check(lib.rr_slm_gradpass(C.byref(plan.struct), _ptr(Xd), _ptr(yd), N,
                              _ptr(m32), _ptr(C32), _ptr(R), _ptr(sqerr),
                              _ptr(ws), ws.numel(), engine, context(),
                              _stream_ptr()),
          "rr_slm_gradpass")


def slm_predict(plan, Xd, m32, C32=None):
    t = require_cuda()
    lib = _cabi.load()
    N = Xd.shape[0]
    Ey = t.empty(N, dtype=t.float32, device=Xd.device)
    Vf = t.empty(N, dtype=t.float32, device=Xd.device) if C32 is not None else None
    nb = _ws_bytes(_cabi.RR_OP_PREDICT, N, plan)
    ws = workspace(nb)
    check(lib.rr_slm_predict(C.byref(plan.struct), _ptr(Xd), N, _ptr(m32),
                             _ptr(C32), _ptr(Ey), _ptr(Vf), _ptr(ws),
                             ws.numel(), _stream_ptr()), "rr_slm_predict")
    return Ey, Vf


def glm_step(plan, Xd, yd, largd, mq, Cq, eps, lik, lik_param, want_ll=True,
             want_R=True):
    """Data part of one SVI step; returns device tensors."""
    t = require_cuda()
    lib = _cabi.load()
    M = Xd.shape[0]
    D, Kmix = mq.shape
    L = eps.shape[1]
    dev = Xd.device
    Edm = t.empty((D, Kmix), dtype=t.float32, device=dev)
    EdC = t.empty((D, Kmix), dtype=t.float32, device=dev)
    R = (t.zeros((plan.d, max(plan.ktot, 1)), dtype=t.float64, device=dev)
         if want_R and plan.ktot else None)
    Ell = t.zeros(Kmix, dtype=t.float64, device=dev) if want_ll else None
    dlp = t.zeros(1, dtype=t.float64, device=dev)
    nb = _ws_bytes(_cabi.RR_OP_GLM_STEP, M, plan, Kmix, L)
    ws = workspace(nb)
    check(lib.rr_glm_step(C.byref(plan.struct), _ptr(Xd), _ptr(yd), _ptr(largd),
                          M, _ptr(mq), _ptr(Cq), Kmix, _ptr(eps), L, int(lik),
                          float(lik_param), _ptr(Edm), _ptr(EdC), _ptr(R),
                          _ptr(Ell), _ptr(dlp), _ptr(ws), ws.numel(),
                          _stream_ptr()), "rr_glm_step")
    return Edm, EdC, R, Ell, dlp


def glm_predict(plan, Xd, ws_draws, lik, lik_param, largd=None, want_sq=False):
    t = require_cuda()
    lib = _cabi.load()
    N = Xd.shape[0]
    S = ws_draws.shape[0]
    Ey = t.empty(N, dtype=t.float32, device=Xd.device)
    Ey2 = t.empty(N, dtype=t.float32, device=Xd.device) if want_sq else None
    nb = _ws_bytes(_cabi.RR_OP_GLM_PREDICT, N, plan, S)
    wsb = workspace(nb)
    check(lib.rr_glm_predict(C.byref(plan.struct), _ptr(Xd), N, _ptr(ws_draws),
                             S, int(lik), float(lik_param), _ptr(largd),
                             _ptr(Ey), _ptr(Ey2), _ptr(wsb), wsb.numel(),
                             _stream_ptr()), "rr_glm_predict")
    return Ey, Ey2


def glm_cdf(Fd, lik, lik_param, quantile, largd=None):
    """mean / min / max over the draws (columns of Fd) of cdf(quantile | f)."""
    t = require_cuda()
    N, S = Fd.shape
    out = [t.empty(N, dtype=t.float32, device=Fd.device) for _ in range(3)]
    check(_cabi.load().rr_glm_cdf(_ptr(Fd), N, S, int(lik), float(lik_param), _ptr(largd),
                                  float(quantile), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]),
                                  _stream_ptr()), "rr_glm_cdf")
    return out


def glm_quantiles(Fd, lik, lik_param, lo_p, hi_p, largd=None):
    """Per-row roots of the Monte-Carlo predictive CDF at lo_p and hi_p."""
    t = require_cuda()
    N, S = Fd.shape
    ql = t.empty(N, dtype=t.float64, device=Fd.device)
    qu = t.empty(N, dtype=t.float64, device=Fd.device)
    check(_cabi.load().rr_glm_quantiles(_ptr(Fd), N, S, int(lik), float(lik_param),
                                        _ptr(largd), float(lo_p), float(hi_p), _ptr(ql),
                                        _ptr(qu), _stream_ptr()), "rr_glm_quantiles")
    return ql, qu


def tcgen05_gemm3(A, B, transa=False, transb=False, alpha=1.0, C=None, accumulate=False):
    """C (+)= alpha * op(A) op(B)^T on the tcgen05 tf32x3 kernel.  ``A``: (M, K) float32
    device tensor, or (K, M) with ``transa``; ``B``: (N, K), or (K, N) with ``transb``."""
    t = require_cuda()
    lib = _cabi.load()
    A, B = A.contiguous(), B.contiguous()
    M, K = (A.shape[1], A.shape[0]) if transa else A.shape
    N = B.shape[1] if transb else B.shape[0]
    if C is None:
        C = t.zeros((M, N), dtype=t.float32, device=A.device)
    nb = 65536 * (-(-M // 256) + -(-N // 256)) * -(-K // 32) + 8192
    ws = workspace(nb)
    check(lib.rr_tcgen05_gemm3(M, N, K, float(alpha), _ptr(A), A.shape[1], int(transa), _ptr(B),
                               B.shape[1], int(transb), _ptr(C), C.stride(0), int(accumulate),
                               _ptr(ws), ws.numel(), _stream_ptr()), "rr_tcgen05_gemm3")
    return C


def tcgen05_selftest():
    require_cuda()
    err = C.c_double(0.0)
    rc = _cabi.load().rr_tcgen05_selftest(C.byref(err))
    check(rc, "rr_tcgen05_selftest")
    return err.value


def tcgen05_i8_selftest(kblocks=9):
    """Bit-exact check of the kind::i8 GEMM kernel; returns the mismatch count."""
    require_cuda()
    bad = C.c_int64(-1)
    rc = _cabi.load().rr_tcgen05_i8_selftest(int(kblocks), C.byref(bad))
    check(rc, "rr_tcgen05_i8_selftest")
    return bad.value


# ---------------------------------------------------------------------------
# Posterior solve on one GPU (library dense linear algebra, float64)
# ---------------------------------------------------------------------------

class Posterior(object):
    """Result of the one-GPU solve.  ``C`` (float64, D x D) is formed lazily:
    an evaluation only needs diag(C), m, logdet and -- for the gradient pass --
    a float32 image of C."""

    def __init__(self, m, diagC, logdet, trgc, Linv=None, C=None, dg=None, C32=None, ok=None):
        self.m, self.diagC, self.logdet, self.trgc = m, diagC, logdet, trgc
        self._Linv, self._C, self._C32 = Linv, C, C32
        # device scalar 1.0 / 0.0: the Cholesky factorisation was stable (None: the
        # caller was told synchronously, or this is the clamped-spectrum solve)
        self.ok = ok if ok is not None else m.new_ones(())
        # (max / min of the Cholesky diagonal)^2: a cheap lower bound on cond(iC)
        self.cond_est = None if dg is None else (dg.max() / dg.min()) ** 2

    @property
    def C(self):
        if self._C is None:
            self._C = self._Linv.T @ self._Linv
        return self._C

    def C32(self):
        """float32 image of C for the gradient pass."""
        if self._C32 is None:
            self._C32 = self.C.float().contiguous()
        return self._C32


_BLOCK_INV_MIN = 1024   # below this potri is as fast
_BLOCK_INV_LEAF = 128         # diagonal blocks inverted by ONE batched trsm (cuBLAS's
                              # batched trsm is fast up to 256: 0.11 ms for 32 x 128^2,
                              # 2.0 ms for 8 x 512^2 at D = 4096)
_GRAM_LEAF = 512              # recursion leaf of L^-T L^-1
_TRI_INV_SHARD_MIN = 1024     # levels of L^-1 from this block size on are split over ranks
_BLOCK_INV_CUDA_ONLY = True   # tests flip this to exercise the blocked paths on CPU
_BLOCK_INV_BATCHED = True     # level-by-level batched triangular inverse (False: recursion)


def _use_blocked(L):
    return L.shape[0] >= _BLOCK_INV_MIN and (L.is_cuda or not _BLOCK_INV_CUDA_ONLY)


def _split(n):
    """Split point of the 2 x 2 recursion: half, rounded to a multiple of 128
    when the matrix is large enough for that to leave two non-empty blocks."""
    h = ((n // 2 + 127) // 128) * 128
    return h if 0 < h < n else n // 2


def _diag_blocks(M, nblk, c, bi, bj):
    """(nblk, c, c) strided view of sub-block (bi, bj) of every 2c x 2c diagonal
    block of M (M is (2 c nblk)^2, contiguous)."""
    V = M.view(nblk, 2, c, nblk, 2, c)[:, bi, :, :, bj, :]      # (nblk, c, nblk, c)
    return V.diagonal(dim1=0, dim2=2).permute(2, 0, 1)


def _tri_inv_lower(L, out):
    """out = L^-1 for lower-triangular L.  With L = [[L11, 0], [L21, L22]],
    L^-1 = [[L11^-1, 0], [-L22^-1 L21 L11^-1, L22^-1]]: applied bottom-up over a
    power-of-two number of diagonal blocks, every level is ONE batched triangular
    solve (the leaves) or two batched float64 GEMMs over all block pairs of that
    level -- log2(n / leaf) + 1 launches' worth of well-filled library calls instead
    of a recursion whose leaves run one small trsm at a time (3.8 ms -> see
    DESIGN.md section 3.5 at D = 4096).  Matrices whose size is not leaf * 2^k are padded
    with an identity block."""
    t = torch()
    n = L.shape[0]
    if n > _BLOCK_INV_LEAF and _BLOCK_INV_BATCHED:
        nb = 1
        while n > nb * _BLOCK_INV_LEAF:
            nb *= 2
        b = -(-n // nb)
        Dp = nb * b
        if Dp != n:
            Lp = t.zeros((Dp, Dp), dtype=L.dtype, device=L.device)
            Lp[:n, :n] = L
            Lp.diagonal()[n:] = 1.0
            Op = t.zeros_like(Lp)
        else:
            Lp = L if L.is_contiguous() else L.contiguous()
            Op = out if (out.is_contiguous() and out.shape == Lp.shape) else t.empty_like(Lp)
            Op.zero_()
        # leaves: all nb diagonal b x b blocks in one batched triangular solve
        Vd = Lp.view(nb, b, nb, b).diagonal(dim1=0, dim2=2).permute(2, 0, 1)
        eye = t.eye(b, dtype=L.dtype, device=L.device).expand(nb, b, b)
        inv = t.linalg.solve_triangular(Vd.contiguous(), eye, upper=False)
        Op.view(nb, b, nb, b).diagonal(dim1=0, dim2=2).permute(2, 0, 1).copy_(inv)
        c = b
        rank, ws = world()
        while c < Dp:
            P = Dp // (2 * c)
            L21 = _diag_blocks(Lp, P, c, 1, 0)
            A = _diag_blocks(Op, P, c, 0, 0)
            B = _diag_blocks(Op, P, c, 1, 1)
            if ws > 1 and c >= _TRI_INV_SHARD_MIN and c >= ws:
                # Row-sharded job (every rank holds the same L): the top levels are a
                # few big GEMMs -- each rank forms a column slice of every pair's block
                # and the slices are all-gathered, so the replicated (serial) part of an
                # evaluation shrinks with the number of GPUs.
                cw = -(-c // ws)
                lo, hi = min(rank * cw, c), min((rank + 1) * cw, c)
                mine = t.zeros((P, c, cw), dtype=L.dtype, device=L.device)
                if hi > lo:
                    mine[:, :, :hi - lo] = t.bmm(B, t.bmm(L21, A[:, :, lo:hi]))
                full = t.empty((ws * P, c, cw), dtype=L.dtype, device=L.device)
                t.distributed.all_gather_into_tensor(full, mine)
                X = full.view(ws, P, c, cw).permute(1, 2, 0, 3).reshape(P, c, ws * cw)[:, :, :c]
                _diag_blocks(Op, P, c, 1, 0).copy_(-X)
            else:
                X = t.bmm(B, t.bmm(L21, A))
                _diag_blocks(Op, P, c, 1, 0).copy_(X.neg_())
            c *= 2
        if Op is not out:
            out.copy_(Op[:n, :n])
        return
    if n <= max(_BLOCK_INV_LEAF, 256):
        out.copy_(t.linalg.solve_triangular(
            L, t.eye(n, dtype=L.dtype, device=L.device), upper=False))
        return
    h = _split(n)
    _tri_inv_lower(L[:h, :h], out[:h, :h])
    _tri_inv_lower(L[h:, h:], out[h:, h:])
    t.matmul(out[h:, h:], L[h:, :h] @ out[:h, :h], out=out[h:, :h])
    out[h:, :h].neg_()
    out[:h, h:].zero_()


def _gram_of_lower(Li, out):
    """out = Li^T Li for lower-triangular Li, recursively on the same 2 x 2 split
    (only half-size GEMMs; the zero block of Li is never multiplied)."""
    t = torch()
    n = Li.shape[0]
    if n <= _GRAM_LEAF:
        t.matmul(Li.T, Li, out=out)
        return
    h = _split(n)
    A, X, B = Li[:h, :h], Li[h:, :h], Li[h:, h:]
    _gram_of_lower(A, out[:h, :h])
    out[:h, :h].addmm_(X.T, X)
    t.matmul(B.T, X, out=out[h:, :h])
    out[:h, h:].copy_(out[h:, :h].T)
    _gram_of_lower(B, out[h:, h:])


def blocked_spd_inverse(L):
    """(L L^T)^-1 = L^-T L^-1 from the lower Cholesky factor, GEMM-rich."""
    t = torch()
    n = L.shape[0]
    Li = t.empty_like(L)
    _tri_inv_lower(L, Li)
    C = t.empty_like(L)
    _gram_of_lower(Li, C)
    return C


def _posterior_sharded(L, dg, p, var, lam):
    """Row-sharded job, large D: every rank holds the same factor L.  L^-1 is formed
    on every rank (batched GEMM form), from it m and diag C in float64; the O(D^3)
    product C = L^-T L^-1 is split by rows over the ranks and all-gathered as the
    float32 image the gradient pass reads (half the bytes of a float64 gather; the
    float64 C itself is only formed if somebody asks for it, from the replicated
    L^-1)."""
    t = torch()
    rank, ws = world()
    D = L.shape[0]
    Li = t.empty_like(L)
    _tri_inv_lower(L, Li)
    diagC = (Li * Li).sum(dim=0)
    m = (Li.T @ (Li @ p)) / var
    # Rows [lo, hi) of C = Li^T Li only involve rows >= lo of the lower-triangular Li:
    # C[lo:hi, :] = Li[lo:, lo:hi]^T Li[lo:, :].  The cost of a slab falls linearly with
    # lo, so the rows are cut into 2 ws slabs and rank r takes slabs r and 2 ws - 1 - r:
    # every rank does the same D^3 / (4 ws) multiply-adds, half of the dense product.
    per = (D + 2 * ws - 1) // (2 * ws)
    rows = t.zeros((2, per, D), dtype=t.float32, device=L.device)
    for j, sl in enumerate((rank, 2 * ws - 1 - rank)):
        lo = min(sl * per, D)
        hi = min(lo + per, D)
        if hi > lo:
            rows[j, :hi - lo] = (Li[lo:, lo:hi].T @ Li[lo:, :]).float()
    full = t.empty((ws, 2, per, D), dtype=t.float32, device=L.device)
    t.distributed.all_gather_into_tensor(full.view(ws * 2 * per, D), rows.view(2 * per, D))
    C32 = t.cat([full[:, 0].reshape(ws * per, D), full[:, 1].flip(0).reshape(ws * per, D)])[:D]
    logdet = 2.0 * t.log(dg).sum()
    trgc = var * (D - (diagC / lam).sum())
    return Posterior(m, diagC, logdet, trgc, Linv=Li, dg=dg, C32=C32.contiguous())


def _inverse_from_factor(L):
    """C = (L L^T)^-1.  One process: potri.  Row-sharded job (every rank holds
    the same L after the allreduce): rank r solves L L^T X = I[:, block_r] for
    its block of columns and the blocks are all-gathered as ROWS of the
    symmetric C -- the O(D^3) inverse, the serial part of an evaluation, then
    scales with the number of GPUs like the row passes do."""
    t = torch()
    rank, ws = world()
    D = L.shape[0]
    if ws == 1 or D < 2 * ws:
        if _use_blocked(L):
            return blocked_spd_inverse(L)
        return t.cholesky_inverse(L)
    per = (D + ws - 1) // ws            # equal blocks (all_gather_into_tensor)
    lo = min(rank * per, D)
    hi = min(lo + per, D)
    if _use_blocked(L):
        # L^-1 replicated (GEMM-rich, D^3/3), then this rank's rows of
        # C = L^-T L^-1 with one (per x D x D) GEMM
        Li = t.empty_like(L)
        _tri_inv_lower(L, Li)
        rows = t.zeros((per, D), dtype=L.dtype, device=L.device)
        if hi > lo:
            t.matmul(Li[:, lo:hi].T, Li, out=rows[:hi - lo])
    else:
        E = t.zeros((D, per), dtype=L.dtype, device=L.device)
        if hi > lo:
            E[lo:hi, :hi - lo] = t.eye(hi - lo, dtype=L.dtype, device=L.device)
        Xb = t.cholesky_solve(E, L)      # D x per: columns lo..hi of C (rest: zeros)
        rows = Xb.T.contiguous()         # per x D: rows lo..hi of C
    full = t.empty((ws * per, D), dtype=L.dtype, device=L.device)
    t.distributed.all_gather_into_tensor(full, rows)
    return full[:D]


def solve_posterior(G, p, var, lam, need_C=True, defer_check=False):
    """C = (diag(1/lam) + G/var)^-1, logdet(iC), m = C p / var, tr(G C).

    Semantics of ``solve_posdef`` (revrand/mathfun/linalg.py:84-125): Cholesky;
    if it fails or any diagonal of the factor is < CHOLTHRESH fall back to the
    clamped-spectrum solve (:128-179) with logdet = sum log s.
    ``G``, ``p`` float64 device tensors; ``lam`` float64 device vector.

    Because G = var (iC - diag(1/lam)),  tr(G C) = var (D - sum_j C_jj / lam_j):
    only diag(C) is needed for the value.  ``need_C=False`` (value-only
    evaluations, e.g. the random-start phase) therefore skips the explicit
    inverse: with iC = L L^T,  diag C = column sums of (L^-1)^2 and
    m = L^-T (L^-1 p) / var from one triangular solve.  The gradient pass needs
    all of C; it is then formed by potri in float64 (forming it as
    (L^-1)^T (L^-1) in reduced precision loses the small entries of C to
    cancellation and derails the optimiser).

    ``defer_check=True``: the Cholesky branch is taken WITHOUT asking the device
    whether the factorisation was stable -- no host synchronisation, so the caller
    can queue the rest of the evaluation behind the solve while the value pass is
    still running.  The verdict comes back as the device scalar ``Posterior.ok``
    (1.0 / 0.0); a caller that reads 0 repeats the solve with ``defer_check=False``.
    """
    t = torch()
    D = G.shape[0]
    iC = G / var
    iC.diagonal().add_(1.0 / lam)
    L, info = t.linalg.cholesky_ex(iC)
    dg = L.diagonal()
    okt = (info == 0) & (dg >= CHOLTHRESH).all()
    ok = True if defer_check else bool(okt.item())
    if ok and need_C and world()[1] > 1 and _use_blocked(L) and D >= 2 * world()[1]:
        post = _posterior_sharded(L, dg, p, var, lam)
        post.ok = okt.double()
        return post
    if ok and need_C:
        Cm = _inverse_from_factor(L)
        diagC = Cm.diagonal().clone()
        m = (Cm @ p) / var
        logdet = 2.0 * t.log(dg).sum()
        trgc = var * (D - (diagC / lam).sum())
        return Posterior(m, diagC, logdet, trgc, C=Cm, dg=dg, ok=okt.double())
    if ok:
        if _use_blocked(L):
            Linv = t.empty_like(L)
            _tri_inv_lower(L, Linv)
        else:
            eye = t.eye(D, dtype=G.dtype, device=G.device)
            Linv = t.linalg.solve_triangular(L, eye, upper=False)
        diagC = (Linv * Linv).sum(dim=0)
        m = (Linv.T @ (Linv @ p)) / var
        logdet = 2.0 * t.log(dg).sum()
        trgc = var * (D - (diagC / lam).sum())
        return Posterior(m, diagC, logdet, trgc, Linv=Linv, dg=dg, ok=okt.double())
    # Clamped-spectrum fallback (svd_solve, linalg.py:128-179).  iC is symmetric, so
    # its SVD is its eigendecomposition with s = |w|, U = V sign(w), Vh = V^T; syevd
    # is the robust dense routine on the device (gesvd returns NaN / exact zeros on
    # the numerically rank-one matrices a wild line-search step produces).  The
    # log-determinant uses the same clamped spectrum as the solve, so it stays
    # finite where the reference's sum(log s) relies on rounding noise in s.
    w, V = t.linalg.eigh(iC)
    w = t.nan_to_num(w, nan=0.0, posinf=1e300, neginf=-1e300)
    sc = t.clamp(w.abs(), min=SVD_FLOOR)
    sgn = t.where(w < 0, -t.ones_like(w), t.ones_like(w))
    Cm = (V * (sgn / sc)) @ V.T
    logdet = t.log(sc).sum()
    m = (Cm @ p) / var
    return Posterior(m, Cm.diagonal().clone(), logdet, (G * Cm).sum(), C=Cm)


def solve_value_scalars(G, p, var, lam, yy, slices):
    """The value-only branch of ``solve_posterior`` without any host decision, for
    pipelined batches of evaluations (random starts, sweeps): returns ONE float64
    device vector  [ok, logdet, trgc, sqerr, q_0 .. q_{S-1}]  where ``ok`` is 0 if
    the Cholesky factorisation failed or was unstable (the caller then re-runs
    that point through ``solve_posterior`` and its clamped-spectrum fallback),
    sqerr = y'y - 2 p'm + m'G m and q_s = sum over slice s of m^2 + diag C."""
    t = torch()
    D = G.shape[0]
    iC = G / var
    iC.diagonal().add_(1.0 / lam)
    L, info = t.linalg.cholesky_ex(iC)
    dg = L.diagonal()
    ok = (info == 0) & (dg >= CHOLTHRESH).all()
    if _use_blocked(L):
        Linv = t.empty_like(L)
        _tri_inv_lower(L, Linv)
    else:
        Linv = t.linalg.solve_triangular(
            L, t.eye(D, dtype=G.dtype, device=G.device), upper=False)
    diagC = (Linv * Linv).sum(dim=0)
    m = (Linv.T @ (Linv @ p)) / var
    logdet = 2.0 * t.log(dg).sum()
    trgc = var * (D - (diagC / lam).sum())
    sqerr = yy - 2.0 * p.dot(m) + m.dot(G @ m)
    mc = m * m + diagC
    q = t.stack([mc[sl].sum() for sl in slices])
    head = t.stack([ok.to(G.dtype).reshape(()), logdet, trgc, sqerr])
    return t.cat([head, q])


# ---------------------------------------------------------------------------
# Row-sharded reduction (one process per GPU)
# ---------------------------------------------------------------------------

def world():
    """(rank, world_size) of the default process group, (0, 1) if none."""
    t = torch()
    if t.distributed.is_available() and t.distributed.is_initialized():
        return t.distributed.get_rank(), t.distributed.get_world_size()
    return 0, 1


def allreduce_sum_(flat):
    """In-place sum over ranks of a flat tensor (NCCL on GPU, gloo on CPU)."""
    t = torch()
    if world()[1] > 1:
        t.distributed.all_reduce(flat, op=t.distributed.ReduceOp.SUM)
    return flat


def _coll_device():
    """Device for small collective payloads: CUDA under NCCL, CPU under gloo."""
    t = torch()
    if t.distributed.get_backend() == "nccl":
        return t.device("cuda", t.cuda.current_device())
    return t.device("cpu")


def sync_random_state(random_state):
    """Row-sharded jobs: every rank must draw the same x0 / random starts (and
    the same basis weights when ``random_state`` seeds them), or the ranks'
    optimiser iterates diverge while they all receive rank 0's objective.  Copies
    rank 0's generator state to every rank; rank 0's own stream is untouched."""
    t = torch()
    rank, ws = world()
    if ws == 1:
        return random_state
    obj = [random_state.get_state() if rank == 0 else None]
    t.distributed.broadcast_object_list(obj, src=0, device=_coll_device())
    random_state.set_state(obj[0])
    return random_state


def assert_same_on_all_ranks(value, what):
    """Raise on every rank if a float64 summary differs between ranks."""
    t = torch()
    if world()[1] == 1:
        return
    v = t.tensor([float(value), -float(value)], dtype=t.float64, device=_coll_device())
    t.distributed.all_reduce(v, op=t.distributed.ReduceOp.MAX)
    hi, lo = v[0].item(), -v[1].item()
    if hi != lo:
        raise RevrandB200Error(
            "%s differs between ranks (%r .. %r): with row sharding every rank "
            "must hold the same basis (seed `random_state`, or build the basis "
            "once and broadcast it)" % (what, lo, hi))


def shard_rows(N, rank, world_size):
    """Contiguous row range [lo, hi) owned by ``rank``."""
    base, rem = divmod(int(N), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi
