"""Containers of the parsmooth API (reference: parsmooth/_base.py:5-30) -- same names, same
field order, so code written against ``parsmooth`` unpacks them identically."""
import itertools
from typing import Any, Callable, NamedTuple, Union


class MVNStandard(NamedTuple):
    mean: Any
    cov: Any


class MVNSqrt(NamedTuple):
    mean: Any
    chol: Any


class FunctionalModel(NamedTuple):
    function: Callable
    mvn: Union[MVNSqrt, MVNStandard]


class ConditionalMomentsModel(NamedTuple):
    conditional_mean: Callable
    conditional_covariance_or_cholesky: Callable


def are_inputs_compatible(*y):
    """parsmooth/_base.py:25-30: lenient on purpose -- raises only when no adjacent pair of
    argument types matches."""
    a, b = itertools.tee(map(type, y))
    _ = next(b, None)
    ok = sum(map(lambda u: u[0] == u[1], zip(a, b)))
    if not ok:
        raise TypeError(f"All inputs should have the same type. {y} was given")
