"""Linearization methods with the protocol of parsmooth.linearization
(``method(model, MVNSqrt) -> (F, chol, b)``; reference: parsmooth/linearization/__init__.py:1-5).

Differences from the reference, by design:
* inputs may carry a leading time axis ([T, n] means, [T, n, n] factors): the drivers linearise a
  whole nominal trajectory in one call instead of ``jax.vmap``-ing a per-step function;
* functions created by ``psqrt.models`` are recognised (attribute ``_psqrt_builtin``) and
  linearised by the fused CUDA kernels of libpsqrt.so; any other callable is treated as a
  user-supplied torch function of a 1-D state and is differentiated with ``torch.func``;
* only the square-root (``MVNSqrt``) branch exists -- ``MVNStandard`` raises NotImplementedError.
"""
from ._common import get_conditional_model
from ._extended import linearize as extended
from ._cubature import linearize as cubature
from ._gh import linearize as gauss_hermite
from ._unscented import linearize as unscented

__all__ = ["extended", "cubature", "gauss_hermite", "unscented", "get_conditional_model"]
