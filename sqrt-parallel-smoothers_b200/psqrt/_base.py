"""Containers of the parsmooth API (reference: parsmooth/_base.py:5-30).  The type names, field names and field
order are the API -- code written against ``parsmooth`` builds and unpacks them positionally -- so they are kept;
the arrays inside are fp64 torch CUDA tensors here (NumPy inputs are accepted and moved by psqrt.methods)."""
from collections import namedtuple

# square-root form: (mean [..., n], lower-triangular-or-any square root of the covariance [..., n, n])
MVNSqrt = namedtuple("MVNSqrt", ["mean", "chol"])
# covariance form: accepted only to be rejected with NotImplementedError (psqrt is the square-root path)
MVNStandard = namedtuple("MVNStandard", ["mean", "cov"])
# x' = function(x) + noise, noise ~ mvn
FunctionalModel = namedtuple("FunctionalModel", ["function", "mvn"])
# E[x' | x] and a square root of Cov[x' | x] as functions of x
ConditionalMomentsModel = namedtuple("ConditionalMomentsModel",
                                     ["conditional_mean", "conditional_covariance_or_cholesky"])


def are_inputs_compatible(*y):
    """Type check of parsmooth/_base.py:25-30, lenient on purpose like upstream: it raises TypeError only when
    NO two neighbouring arguments have the same type (so also for a single argument)."""
    kinds = [type(v) for v in y]
    if not any(left is right for left, right in zip(kinds, kinds[1:])):
        raise TypeError(f"All inputs should have the same type. {y} was given")
