"""Device-resident SVI loop of the generalised linear model.

``GeneralizedLinearModel.fit`` composes ``structured_sgd(logtrick_sgd(sgd))``
around ``_elbo`` (glm.py:139-203, optimize/decorators.py:329-403, 406-538,
optimize/sgd.py:311-425).  Run on the host, every step costs two uploads
(variational mean / variance), one download (their gradients), a K_mix^2
entropy loop and the update rule in numpy: 4.4 ms of a 7.5 ms step at config 4.
Here the whole step lives on the device:

  flat log-space parameter vector z (float64)  ->  m, C, regulariser, lengthscales
  -> minibatch gathered from the resident training set by a resident index stream
  -> ``rr_glm_step`` (features, draws, likelihood derivatives, contractions)
  -> mixture-entropy / prior terms, chain rule of the log warp (float64)
  -> bound-aware truncation, update rule, clip

with no host synchronisation; after the first (eager) step the body is replayed as
one CUDA graph.  The arithmetic is the host path's, operation for operation
(``GeneralizedLinearModel._elbo`` stays the parity-tested restatement); minibatch
indices come from the same ``RandomState`` stream in the reference's order.
"""

from __future__ import annotations

import logging
import math

import numpy as np

from . import _engine as eng
from . import config
from .btypes import Positive
from . import optimize as _sgd      # (its names: Adam, SGDUpdater, Momentum, AdaGrad, AdaDelta)
from .optimize.structured import (EXPMAX, LOGMINPOS, Layout, _flat_bounds,
                                  _map_params, flatten_values)

log = logging.getLogger(__name__)

LOG2PI = math.log(2.0 * math.pi)
TRACE_MAX = 1 << 20     # per-step |grad| / objective kept on the device for result()


def supported(glm, largs):
    """The device loop covers the likelihoods without a learnable parameter
    (rr_glm_step takes it by value), device-drawn noise and the update rules of
    optimize/sgd.py."""
    if not config.GLM_DEVICE_LOOP or config.GLM_HOST_RNG:
        return False
    ps = glm.likelihood.params
    ps = ps if isinstance(ps, list) else [ps]
    if any(p.has_value for p in ps):
        return False
    up = glm.updater
    return up is None or type(up) in (_sgd.Adam, _sgd.SGDUpdater, _sgd.Momentum,
                                      _sgd.AdaGrad, _sgd.AdaDelta)


def glm_gen(glm):
    return glm._devgen


class _Updater(object):
    """The update rules of optimize/sgd.py on device tensors (same formulas)."""

    def __init__(self, host, n, t, dev):
        self.h = host if host is not None else _sgd.Adam()
        self.kind = type(self.h)
        z = lambda: t.zeros(n, dtype=t.float64, device=dev)  # noqa: E731
        self.a, self.b = z(), z()
        self.step = t.zeros((), dtype=t.float64, device=dev)

    def __call__(self, x, g, t):
        h = self.h
        if self.kind is _sgd.SGDUpdater:
            return x - h.eta * g
        if self.kind is _sgd.Momentum:
            self.a.mul_(h.rho).sub_(h.eta * g)
            return x + self.a
        if self.kind is _sgd.AdaGrad:
            self.a.add_(g * g)
            return x - h.eta * g / (h.epsilon + t.sqrt(self.a))
        if self.kind is _sgd.AdaDelta:
            r = h.rho
            self.a.mul_(r).add_((1 - r) * g * g)
            dx = -g * t.sqrt(self.b + h.epsilon) / t.sqrt(self.a + h.epsilon)
            self.b.mul_(r).add_((1 - r) * dx * dx)
            return x + dx
        # Adam (sgd.py:254-308)
        b1, b2 = h.beta1, h.beta2
        self.step.add_(1.0)
        self.a.mul_(b1).add_((1 - b1) * g)
        self.b.mul_(b2).add_((1 - b2) * g * g)
        mhat = self.a / (1 - t.pow(t.full_like(self.step, b1), self.step))
        vhat = self.b / (1 - t.pow(t.full_like(self.step, b2), self.step))
        return x - h.alpha * mhat / (t.sqrt(vhat) + h.epsilon)


class DeviceSVI(object):
    """State of one device-resident SVI run of ``glm`` on ``data`` (device tensors
    X, y[, larg])."""

    def __init__(self, glm, params, data, maxiter, random_state, x0=None, graph=True):
        t = eng.require_cuda()
        self.t, self.glm = t, glm
        self.X, self.y = data[0], data[1]
        self.larg = data[2] if len(data) > 2 else None
        dev = self.dev = self.X.device
        self.N, self.d = self.X.shape
        self.B = int(min(glm.batch_size, self.N))
        self.maxiter = int(maxiter)
        self.rs = random_state
        self.D, self.K, self.L = glm.D_, glm.K, glm.nsamples
        D, K = self.D, self.K
        self.layout = Layout.of_parameters(params)
        bounds = _flat_bounds(params)
        n = self.n = len(bounds)
        sizes = [len(flatten_values(_map_params(lambda p: np.zeros(p.shape), p)))
                 for p in params]
        assert sizes[0] == D * K and sizes[1] == D * K and sizes[3] == 0
        self.off = np.concatenate([[0], np.cumsum(sizes)]).astype(int)
        pos = np.array([isinstance(b, Positive) for b in bounds], dtype=bool)
        lower, upper = np.full(n, -np.inf), np.full(n, np.inf)
        for i, b in enumerate(bounds):
            if pos[i]:
                lower[i] = LOGMINPOS
                upper[i] = EXPMAX if b.upper is None else np.log(b.upper)
            else:
                lower[i] = -np.inf if b.lower is None else b.lower
                upper[i] = np.inf if b.upper is None else b.upper
        f64 = t.float64
        self.pos = eng.to_device(pos, t.bool)
        self.lower, self.upper = eng.to_device(lower, f64), eng.to_device(upper, f64)
        if x0 is None:
            x0 = flatten_values(_map_params(lambda p: p.rvs(None), params))
        self.z = t.empty(n, dtype=f64, device=dev)
        self.set_x(x0)
        # regulariser slot of every feature column
        plan = self.plan = glm._get_plan(self.d, glm.basis.params_values())
        _, slices = glm.basis.regularizer_diagonal(np.zeros((1, self.d)))
        slot = np.zeros(D, dtype=np.int64)
        if isinstance(slices, list):
            for i, s in enumerate(slices):
                slot[s] = i
            self.nreg = len(slices)
        else:
            self.nreg = 1
        assert self.nreg == sizes[2]
        self.reg_slot = eng.to_device(slot, t.int64)
        # lengthscale slot of every (input dimension, frequency)
        idx = np.zeros((self.d, max(plan.ktot, 1)), dtype=np.int64)
        o = int(self.off[4])
        self.ls_blocks = []
        for b, ko in zip(plan.trig, plan.freq_offsets):
            rows = np.arange(self.d) if b.cols is None else np.asarray(b.cols)
            nls = len(b.lenscale)
            idx[:, ko:ko + b.K] = o
            if nls > 1:
                idx[rows, ko:ko + b.K] = (o + np.arange(nls))[:, None]
            self.ls_blocks.append((o, nls, eng.to_device(rows, t.int64), ko, b.K))
            o += nls
        assert o == self.off[5], "basis hyper-parameters other than lengthscales"
        self.ls_index = eng.to_device(idx[:, :max(plan.ktot, 0)] if plan.ktot else idx[:, :0],
                                      t.int64)
        self.updater = _Updater(glm.updater, n, t, dev)
        # resident minibatch index stream (reference order: back-to-back permutations,
        # a batch may straddle two of them, sgd.py:428-459)
        self._carry = np.zeros(0, dtype=np.int64)
        self.chunk_steps = max(1, min(self.maxiter, max(1, (4 << 20) // self.B)))
        self.idx_dev = t.zeros(self.chunk_steps * self.B, dtype=t.int64, device=dev)
        self.cursor = t.zeros((), dtype=t.int64, device=dev)
        self.arangeB = t.arange(self.B, dtype=t.int64, device=dev)
        self._left_in_chunk = 0
        self.it = 0
        self.ntrace = int(max(1, min(self.maxiter, TRACE_MAX)))   # (ring beyond TRACE_MAX)
        self.norms = t.zeros(self.ntrace, dtype=f64, device=dev)
        self.objs = t.zeros(self.ntrace, dtype=f64, device=dev)
        self.itd = t.zeros(1, dtype=t.int64, device=dev)
        self.use_graph = bool(graph) and config.GLM_DEVICE_GRAPH
        self.graph = None
        self.h2d_bytes = 0

    # -- parameters ----------------------------------------------------------------
    def set_x(self, x):
        """Load raw-space parameters (flat) as the log-warped state."""
        t = self.t
        xd = eng.to_device(np.asarray(x, dtype=float), t.float64)
        self.z.copy_(t.where(self.pos, t.log(t.where(self.pos, xd, t.ones_like(xd))), xd))

    def x(self):
        t = self.t
        return t.where(self.pos, t.exp(self.z), self.z)

    def result_x(self):
        return tuple(self.layout.unflatten(self.x().cpu().numpy()))

    # -- minibatches -------------------------------------------------------------------
    def _refill(self):
        t = self.t
        steps = min(self.chunk_steps, max(self.maxiter - self.it, 1))
        need = steps * self.B
        parts, have = [self._carry], len(self._carry)
        while have < need:
            p = self.rs.permutation(self.N)
            parts.append(p)
            have += len(p)
        stream = np.concatenate(parts)
        chunk, self._carry = stream[:need], stream[need:]
        host = t.from_numpy(np.ascontiguousarray(chunk, dtype=np.int64)).pin_memory()
        self.idx_dev[:need].copy_(host, non_blocking=True)
        self._pinned = host           # keep alive until the copy has run
        self.cursor.zero_()
        self._left_in_chunk = steps
        self.h2d_bytes += need * 8

    # -- random starts ---------------------------------------------------------------
    def random_starts(self, params, nstarts):
        """Best of ``nstarts`` draws from the parameters' distributions, each scored
        on the next minibatch (decorators.py:541-583 with a data generator): the draws
        and the minibatch indices come from the host generator in the reference's
        order (batch, then candidate), go down in blocks, and the objectives are
        evaluated back to back on the device; ONE read (the arg-min) at the end.
        Loads the best candidate as the state and returns it (raw space, flat)."""
        t = self.t
        rs, N, B = self.rs, self.N, self.B
        carry = np.zeros(0, dtype=np.int64)
        cands, inds = [], []
        for _ in range(nstarts):
            while len(carry) < B:
                carry = np.concatenate((carry, rs.permutation(N)))
            inds.append(carry[:B])
            carry = carry[B:]
            cands.append(flatten_values(_map_params(lambda p: p.rvs(rs), params)))
            # (the host loop seeds the device noise generator from the same stream at
            # its first evaluation, i.e. right here)
            self.glm._device_generator(self.dev)
        objs = t.empty(nstarts, dtype=t.float64, device=self.dev)
        blk = max(1, min(self.chunk_steps, 64))
        for b0 in range(0, nstarts, blk):
            b1 = min(nstarts, b0 + blk)
            xc = t.from_numpy(np.stack(cands[b0:b1])).pin_memory().to(self.dev, non_blocking=True)
            ic = t.from_numpy(np.concatenate(inds[b0:b1])).pin_memory()
            self.idx_dev[:ic.numel()].copy_(ic, non_blocking=True)
            self.cursor.zero_()
            self.h2d_bytes += xc.numel() * 8 + ic.numel() * 8
            for i in range(b1 - b0):
                xd = xc[i]
                self.z.copy_(t.where(self.pos, t.log(t.where(self.pos, xd, t.ones_like(xd))), xd))
                objs[b0 + i] = self._body(update=False)
            t.cuda.current_stream().synchronize()      # the pinned blocks may go now
        objs = t.where(t.isnan(objs), t.full_like(objs, float("inf")), objs)
        best = int(t.argmin(objs).item())
        log.info("Best start found with objective = {}".format(float(objs[best].item())))
        self.set_x(cands[best])
        self._left_in_chunk = 0            # the SGD phase starts its own index stream
        return cands[best]

    # -- one step ----------------------------------------------------------------------
    def _body(self, update=True):
        t, glm, plan = self.t, self.glm, self.plan
        D, K, L = self.D, self.K, self.L
        off = self.off
        z = self.z
        x = t.where(self.pos, t.exp(z), z)
        m = x[off[0]:off[1]].view(D, K)
        C = x[off[1]:off[2]].view(D, K)
        reg = x[off[2]:off[3]]
        Lam = reg[self.reg_slot]
        iL = 1.0 / Lam
        if plan.ktot:
            Wt = plan._Wfull_dev / x[self.ls_index] / eng.TWO_PI
            plan._Wt.copy_(t.clamp(Wt, -eng.WT_CLIP, eng.WT_CLIP).float())
        ind = self.idx_dev[self.cursor + self.arangeB]
        self.cursor.add_(self.B)
        Xb, yb = self.X[ind], self.y[ind]
        lb = self.larg[ind] if self.larg is not None else None
        eps = glm._noise(K, L, D, self.dev)   # device generator, seeded on first use
        Edm, EdC, R, Ell, dlp = eng.glm_step(
            plan, Xb, yb, lb, m.float().contiguous(), C.float().contiguous(), eps,
            glm.likelihood._lik_id, 1.0, want_ll=True, want_R=True)
        Edm, EdC = Edm.double(), EdC.double()
        # mixture-entropy / prior terms (glm.py:222-223, 249-271)
        v = C[:, :, None] + C[:, None, :]
        dmm = m[:, :, None] - m[:, None, :]
        logNkl = -0.5 * (t.log(2 * math.pi * v) + dmm ** 2 / v).sum(dim=0)
        logzk = t.logsumexp(logNkl, dim=0)
        alpha = t.exp(logNkl - logzk[None, :]) + t.exp(logNkl - logzk[:, None])
        iCkCj = 1.0 / v
        aT = alpha.T[None, :, :]
        dm = (glm.B_ * Edm - m * iL[:, None] + (iCkCj * dmm * aT).sum(dim=2)) / K
        dC = (glm.B_ * EdC - iL[:, None]
              + ((iCkCj - (dmm * iCkCj) ** 2) * aT).sum(dim=2)) / (2 * K)
        col = ((m ** 2 + C) * (iL ** 2)[:, None]).sum(dim=1) / K - iL
        dreg = -0.5 * t.zeros(self.nreg, dtype=t.float64, device=self.dev).index_add_(
            0, self.reg_slot, col)
        parts = [-dm.reshape(-1), -dC.reshape(-1), dreg]
        if plan.ktot:
            WR = plan._Wfull_dev * R[:, :plan.ktot]
            for (o, nls, rows, ko, Kb) in self.ls_blocks:
                gi = WR[:, ko:ko + Kb].sum(dim=1)[rows]
                ls = x[o:o + nls]
                if nls > 1:
                    parts.append(gi / ls ** 2)
                elif config.REFERENCE_COMPAT:
                    parts.append((gi[0] / ls[0] ** 2).reshape(1))
                else:
                    parts.append((gi.sum() / ls[0] ** 2).reshape(1))
        g = t.cat(parts)
        ELBO = (Ell.sum() * glm.B_ - 0.5 * D * K * LOG2PI - 0.5 * K * t.log(Lam).sum()
                - 0.5 * ((m ** 2 + C) * iL[:, None]).sum() - logzk.sum() + math.log(K)) / K
        if not update:
            return -ELBO
        # log warp (decorators.py:329-403), sgd.py:380-415
        g = t.where(self.pos, g * x, g)
        slot = self.itd % self.ntrace
        self.norms.index_copy_(0, slot, t.linalg.vector_norm(g).reshape(1))
        self.objs.index_copy_(0, slot, (-ELBO).reshape(1))
        self.itd.add_(1)
        g = t.where(z <= self.lower, t.clamp(g, max=0.0), g)
        g = t.where(z >= self.upper, t.clamp(g, min=0.0), g)
        znew = self.updater(z, g, t)
        self.z.copy_(t.minimum(t.maximum(znew, self.lower), self.upper))
        return None

    def step(self):
        """One SVI iteration; False once ``maxiter`` steps have run."""
        t = self.t
        if self.it >= self.maxiter:
            return False
        if self._left_in_chunk == 0:
            self._refill()
        if not self.use_graph:
            self._body()
        elif self.graph is None and self.it < 2:
            self._body()                   # eager warm-up (lazy initialisations)
            if self.it == 1:
                side = t.cuda.Stream()
                side.wait_stream(t.cuda.current_stream())
                g = t.cuda.CUDAGraph()
                if hasattr(g, "register_generator_state"):
                    g.register_generator_state(glm_gen(self.glm))
                self._capture_pending = (g, side)
        else:
            if self.graph is None:
                g, side = self._capture_pending
                t.cuda.synchronize()
                from . import _cabi
                l0 = _cabi.load().rr_launch_count()
                with t.cuda.graph(g, stream=side):
                    self._body()
                # kernels of this library inside one replay (the library's launch
                # counter only sees the capture)
                self.graph_launches = int(_cabi.load().rr_launch_count() - l0)
                self.graph_replays = 0
                self.graph = g
                # (capture does not execute: fall through and replay it now)
            self.graph.replay()
            self.graph_replays += 1
        self._left_in_chunk -= 1
        it = self.it
        self.it += 1
        if log.isEnabledFor(logging.INFO) and (it % 500 == 0 or it == self.maxiter - 1):
            log.info("Iter {}: ELBO = {}".format(
                it, -float(self.objs[it % self.ntrace].item())))
        return True

    def run(self):
        while self.step():
            pass
        return self

    def result(self):
        from scipy.optimize import OptimizeResult
        n = min(self.it, self.ntrace)
        objs = self.objs[:n].cpu().numpy()
        return OptimizeResult(x=self.result_x(), norms=list(self.norms[:n].cpu().numpy()),
                              message='maxiter reached',
                              fun=(objs[-1] if n else None), objs=list(objs))
