"""Host-side logic that needs no GPU: API error behaviour, containers, linearization constants,
analytic Jacobians of the built-in models, data generators, the parsmooth alias."""
import numpy as np
import pytest
import torch

import parsmooth_np as O


def test_containers_and_compat():
    import psqrt
    x = psqrt.MVNSqrt(np.zeros(2), np.eye(2))
    assert x._fields == ("mean", "chol") and psqrt.MVNStandard._fields == ("mean", "cov")
    assert psqrt.FunctionalModel._fields == ("function", "mvn")
    assert psqrt.ConditionalMomentsModel._fields == ("conditional_mean", "conditional_covariance_or_cholesky")
    psqrt.are_inputs_compatible(x, x)
    psqrt.are_inputs_compatible(x, x, psqrt.MVNStandard(0, 0))          # lenient: one matching pair is enough
    with pytest.raises(TypeError):
        psqrt.are_inputs_compatible(x, psqrt.MVNStandard(0, 0))


def test_unsupported_paths_raise():
    """No fallback: covariance form and sequential algorithms are refused, not emulated."""
    import psqrt
    from psqrt.models import lgssm
    tm = psqrt.FunctionalModel(lgssm.transition_function(np.eye(2)), psqrt.MVNSqrt(np.zeros(2), np.eye(2)))
    x0 = psqrt.MVNSqrt(np.zeros(2), np.eye(2))
    ys = np.zeros((3, 2))
    with pytest.raises(NotImplementedError):
        psqrt.filtering(ys, x0, tm, tm, psqrt.linearization.extended, None, parallel=False)
    with pytest.raises(NotImplementedError):
        psqrt.smoothing(tm, x0, psqrt.linearization.extended, None, parallel=False)
    with pytest.raises(NotImplementedError):
        psqrt.iterated_smoothing(ys, x0, tm, tm, psqrt.linearization.extended, None, False)
    if not torch.cuda.is_available():
        from psqrt._lib import PsqrtError
        with pytest.raises(PsqrtError):                                   # fails loudly without a GPU
            psqrt.filtering(ys, x0, tm, tm, psqrt.linearization.extended)
    with pytest.raises(NotImplementedError):
        psqrt.linearization.extended(tm, psqrt.MVNStandard(torch.zeros(2), torch.eye(2)))


def test_gh_and_cubature_tables_match_oracle():
    from psqrt.linearization._cubature import _cubature_weights
    from psqrt.linearization._gh import _gauss_hermite_weights
    for n, p in ((1, 3), (2, 3), (5, 3), (2, 5)):
        wm, wc, xi = _gauss_hermite_weights(n, p)
        owm, _, oxi = O.gauss_hermite_weights(n, p)
        np.testing.assert_array_equal(wm, owm)
        np.testing.assert_array_equal(xi, oxi.T)
        assert xi.shape == (p ** n, n)
    wm, _, xi = _cubature_weights(5)
    owm, _, oxi = O.cubature_weights(5)
    np.testing.assert_array_equal(wm, owm)
    np.testing.assert_array_equal(xi, oxi)


def test_builtin_model_jacobians_cpu():
    """analytic Jacobians (incl. the |w| < 1e-6 branch, bearings_utils.py:24-37) == forward-mode AD"""
    from psqrt.models import bearings, population
    f = bearings.make_transition_function(0.01)
    h = bearings.make_observation_function([-1.5, 0.5], [1.0, 1.0])
    xs = torch.tensor([[-1.0, -1.0, 6.0, 4.0, 2.0], [0.3, -2.0, 1.0, 0.5, 1e-8], [2.0, 1.5, -3.0, 0.2, -0.7]],
                      dtype=torch.float64)
    of, oh = O.ct_transition_function(0.01), O.bearings_observation_function([-1.5, 0.5], [1.0, 1.0])
    for fn, ofn in ((f, of), (h, oh)):
        val, jac = fn._psqrt_value_and_jac(xs)
        ad = torch.func.vmap(torch.func.jacfwd(fn))(xs)
        np.testing.assert_allclose(jac.numpy(), ad.numpy(), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(val.numpy(), ofn(xs.numpy()), rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(jac.numpy(), ofn.jac(xs.numpy()), rtol=1e-13, atol=1e-15)
    tmod, omod = population.make_parameters(10.0, np.array([[0.09]]))
    otm, oom = O.population_model(10.0, np.array([[0.09]]))
    x = torch.tensor([[0.3], [1.9]], dtype=torch.float64)
    for cm, ocm in ((tmod, otm), (omod, oom)):
        val, jac = cm.conditional_mean._psqrt_value_and_jac(x)
        np.testing.assert_allclose(val.numpy(), ocm.conditional_mean(x.numpy()), rtol=1e-14)
        np.testing.assert_allclose(jac.numpy(), ocm.conditional_mean.jac(x.numpy()), rtol=1e-14)
        np.testing.assert_allclose(cm.conditional_covariance_or_cholesky(x).numpy(),
                                   ocm.conditional_covariance_or_cholesky(x.numpy()), rtol=1e-14)


def test_extended_linearization_cpu_matches_oracle():
    """extended needs no kernel: it runs on CPU tensors too (user function via torch.func, built-ins analytic)."""
    import psqrt
    from psqrt.models import bearings
    dt = 0.01
    f = bearings.make_transition_function(dt)
    cq = np.diag([0.1, 0.1, 0.2, 0.2, 0.3])
    xs = np.random.RandomState(0).randn(7, 5)
    F, Q, b = psqrt.linearization.extended(
        psqrt.FunctionalModel(f, psqrt.MVNSqrt(torch.zeros(5, dtype=torch.float64), torch.as_tensor(cq))),
        psqrt.MVNSqrt(torch.as_tensor(xs), torch.eye(5, dtype=torch.float64).expand(7, 5, 5)))
    oF, oQ, ob = O.extended(O.FunctionalModel(O.ct_transition_function(dt), O.MVNSqrt(np.zeros(5), cq)),
                            O.MVNSqrt(xs, np.repeat(np.eye(5)[None], 7, 0)))
    np.testing.assert_allclose(F.numpy(), oF, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(b.numpy(), ob, rtol=1e-12, atol=1e-14)

    def user_f(x):                       # unknown to the library -> torch.func.jacfwd
        return torch.stack([x[0] * x[1], torch.sin(x[0])])
    F2, _, b2 = psqrt.linearization.extended(
        psqrt.FunctionalModel(user_f, psqrt.MVNSqrt(torch.zeros(2, dtype=torch.float64), torch.eye(2, dtype=torch.float64))),
        psqrt.MVNSqrt(torch.tensor([[1.0, 2.0], [0.5, -1.0]], dtype=torch.float64), None))
    np.testing.assert_allclose(F2[0].numpy(), [[2.0, 1.0], [np.cos(1.0), 0.0]], rtol=1e-14)
    np.testing.assert_allclose(b2[0].numpy(), [2.0 - 4.0, np.sin(1.0) - np.cos(1.0)], rtol=1e-14)


def test_bearings_data_generator_closed_form():
    """get_data replaces scipy.linalg.expm(F dt) of notebooks/bearing_data.py:140-146 by its closed form."""
    from scipy import linalg
    for a in (1.3, -0.4, 1e-3):
        dt = 0.01
        F = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [0, 0, 0, a], [0, 0, -a, 0]])
        sa, ca1 = np.sin(a * dt) / a, (1 - np.cos(a * dt)) / a
        c, s = np.cos(a * dt), np.sin(a * dt)
        E = np.array([[1, 0, sa, ca1], [0, 1, -ca1, sa], [0, 0, c, s], [0, 0, -s, c]])
        np.testing.assert_allclose(E, linalg.expm(F * dt), atol=1e-14)
    from psqrt.models import bearings
    ts, xs, ys = bearings.get_data(np.array([0.1, 0.2, 1.0, 0.0]), 0.01, 0.5, 50, [-1.5, 0.5], [1.0, 1.0], random_state=0)
    assert xs.shape == (51, 5) and ys.shape == (50, 2) and ys.dtype == np.float32


def test_fixed_point_iteration_count():
    """_utils.py:136-146: criterion i < N performs exactly N applications after the initial one."""
    from psqrt.methods import fixed_point
    calls = []

    def f(x):
        calls.append(x)
        return x + 1

    assert fixed_point(f, 0, lambda i, *_: i < 5) == 5 and len(calls) == 5
    assert O.fixed_point(lambda x: x + 1, 0, lambda i, *_: i < 5) == 5


def test_parsmooth_alias():
    import parsmooth
    from parsmooth.methods import iterated_smoothing, filtering            # noqa: F401
    from parsmooth.linearization import extended, cubature, gauss_hermite  # noqa: F401
    from parsmooth._base import MVNSqrt
    import psqrt
    assert MVNSqrt is psqrt.MVNSqrt and parsmooth.methods is psqrt.methods
    assert parsmooth.sampling is psqrt.methods.sampling
    from parsmooth.linearization import unscented                          # noqa: F401
