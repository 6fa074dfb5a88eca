"""Run-time switches of revrand_b200 (module-level, read at call time)."""

import os

# Reproduce the reference's behaviour for a *scalar* (isotropic) lengthscale on
# d > 1 inputs, where only input dimension 0 contributes to d Phi / d lenscale
# (revrand/basis_functions.py:896-899).  Set to False for the mathematically
# complete derivative.  Posterior mean / covariance / log marginal likelihood
# are unaffected either way.
REFERENCE_COMPAT = os.environ.get("REVRAND_B200_REFERENCE_COMPAT", "1") != "0"

# Engine for the SLM passes: "auto" (tensor-core engine when the basis allows it
# and the job is large enough), "simt" (chunked CUDA-core path), "tcgen05" (error
# if unsupported); "tcgen05_fused16" / "tcgen05_fine" select the round-1 fused
# kind::f16 value pass for A/B measurements.
ENGINE = os.environ.get("REVRAND_B200_ENGINE", "auto")


def engine_code():
    from . import _cabi
    return {"auto": _cabi.RR_ENGINE_AUTO, "simt": _cabi.RR_ENGINE_SIMT,
            "tcgen05": _cabi.RR_ENGINE_TCGEN05,
            "tcgen05_fine": _cabi.RR_ENGINE_TCGEN05_FINE,
            "tcgen05_fused16": _cabi.RR_ENGINE_TCGEN05_FUSED16}[ENGINE]

# Draw the GLM reparameterisation noise on the host from the model's
# RandomState in the reference's order (slow: K_mix*L*D normals per step)
# instead of on the device.
GLM_HOST_RNG = os.environ.get("REVRAND_B200_HOST_RNG", "0") == "1"

# Only the round-1 fused kind::f16 value pass ("tcgen05_fused16" / "tcgen05_fine")
# needs this: it perturbs every trig value by ~2e-6 (zero mean), and the posterior
# moments of an ill-conditioned evaluation inherit cond * noise / sqrt(N).  When
# the conditioning estimate of the BEST evaluation exceeds this threshold, the
# reported posterior is recomputed once with the SIMT engine.  The default engine
# (24-bit fixed point on kind::i8) meets 1e-4 everywhere and never polishes.
POLISH_COND = float(os.environ.get("REVRAND_B200_POLISH_COND", "1e3"))

# Direct ``StandardLinearModel._elbo(X, y, ...)`` calls keep X and y resident on
# the device between calls while the same host arrays are passed (identity plus
# a content fingerprint).  False: upload the rows on every call.
CACHE_DEVICE_DATA = os.environ.get("REVRAND_B200_CACHE_DATA", "1") != "0"

# Value-only evaluations take sum Err^2 = y'y - 2 p'm + m'G m from the float64
# sufficient statistics when it is at least this fraction of y'y (the quadratic
# form loses y'y / sum Err^2 digits of the statistics' ~1e-7 relative accuracy to
# cancellation); nearer to interpolation they run the residual pass over the rows.
SQERR_FROM_STATS_MIN = float(os.environ.get("REVRAND_B200_SQERR_STATS_MIN", "1e-3"))

# ``fit`` evaluates its random starts (independent points) as one pipelined batch:
# the float64 solve of one start overlaps the value pass of the next, with no host
# synchronisation inside the loop.  False: one blocking evaluation per start.
PIPELINE_STARTS = os.environ.get("REVRAND_B200_PIPELINE_STARTS", "1") != "0"

# GeneralizedLinearModel.fit keeps the SVI loop on the device (parameters, update
# rule, mixture-entropy terms, minibatch gather; replayed as a CUDA graph) when the
# likelihood has no learnable parameter.  False: the host loop of the reference's
# structured_sgd(logtrick_sgd(sgd)) composition around the device step.
GLM_DEVICE_LOOP = os.environ.get("REVRAND_B200_GLM_DEVICE_LOOP", "1") != "0"
# ... replayed as a CUDA graph from its third step on (False: eager launches).
GLM_DEVICE_GRAPH = os.environ.get("REVRAND_B200_GLM_DEVICE_GRAPH", "1") != "0"
# ... and its random starts are scored back to back on the device (one read at the end);
# False: one ``_elbo`` call (host assembly) per start.
GLM_DEVICE_STARTS = os.environ.get("REVRAND_B200_GLM_DEVICE_STARTS", "1") != "0"

# An evaluation with gradients forms the feature map ONCE (as slm.py:145 does): the
# value pass leaves an fp16 image of Phi behind and the gradient pass reads it
# instead of evaluating sin/cos a second time.  The image takes 2 bytes per (row,
# feature) of device memory (8.3 GB per GPU at config 2); larger than this, or than
# half of the free memory, and the gradient pass regenerates Phi in row chunks.
KEEP_FEATURES_MAX_BYTES = int(float(os.environ.get("REVRAND_B200_KEEP_FEATURES_MAX_GB", "64"))
                              * (1 << 30))

# The tensor-core gradient pass multiplies Phi by an fp16 image of the posterior
# covariance C.  Its rounding is harmless where the quadratic form Phi C dPhi does
# not cancel (config 2: 1.6e-5 on the lengthscale gradients) but reached 1.5e-2 on a
# strongly correlated feature set (64 frequencies on 3-D inputs, cond(C) = 3e5).
# True: a second GEMM over the rounding residual of C (RR_GRAD_SPLIT_C) restores C
# to ~22 bits, at twice the tensor-core work of the gradient pass.
GRADIENT_SPLIT_C = os.environ.get("REVRAND_B200_GRADIENT_SPLIT_C", "0") == "1"
