"""Built-in state-space models of the reference's tests / notebooks, as torch callables that the
linearization methods recognise (analytic Jacobians; time-invariant outputs for linear models).

* lgssm       tests/_lgssm.py:5-39
* bearings    tests/bearings/bearings_utils.py:7-115, notebooks/bearing_data.py, bearing_data_pe.py:123-134
* population  notebooks/population_model.py:23-34,51-62,84-129
"""
from . import bearings, lgssm, population

__all__ = ["lgssm", "bearings", "population"]
