"""Oracle and host logic against golden vectors produced by EXECUTING the reference's own source
(tools/make_reference_goldens.py; fixtures in tests/golden/reference_numeric.json): collocation constants
(modeling.py:1092-1127), Latin-hypercube normal sampling (pf.py:425-447) and the resampling / roughening / moment block of
ParticleFilter.estimate (pf.py:405-421)."""
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_numeric.json')))


@pytest.mark.parametrize('method', ['radau', 'legendre'])
@pytest.mark.parametrize('deg', [1, 2, 3, 4, 5])
def test_collocation_constants_equal_the_reference(method, deg):
    from hilo_mpc_b200.transcription import collocation_coefficients
    from oracle.collocation import polynomial_basis
    g = GOLD['polynomial_basis']['values'][f'{method}{deg}']
    for fn in (polynomial_basis, collocation_coefficients):
        B, C, D, tau = fn(deg, method)
        np.testing.assert_allclose(B, g['B'], rtol=0, atol=1e-13)
        np.testing.assert_allclose(C, g['C'], rtol=0, atol=1e-11)
        np.testing.assert_allclose(D, g['D'], rtol=0, atol=1e-12)
        np.testing.assert_allclose(tau, g['T'], rtol=0, atol=1e-14)


@pytest.mark.parametrize('case', range(len(GOLD['lhsnorm']['cases'])))
def test_lhsnorm_equals_the_reference(case):
    from hilo_mpc_b200.estimators import lhsnorm as product
    from oracle.estimators import lhsnorm as oracle
    c = GOLD['lhsnorm']['cases'][case]
    for fn in (product, oracle):
        np.random.seed(c['seed'])
        x = fn(np.array(c['mu']), np.array(c['sigma']), c['n'])
        np.testing.assert_allclose(x, c['x'], rtol=0, atol=1e-12)


@pytest.mark.parametrize('case', range(len(GOLD['pf_resample_block']['cases'])))
def test_pf_resampling_roughening_moments_equal_the_reference(case):
    """pf.py:405-421 with the legacy global RNG: ``np.random.choice`` consumes ``n_s`` uniforms, the roughening draws one
    Latin-hypercube sample.  The oracle's index search, roughening covariance and moments must reproduce the block."""
    from oracle.estimators import lhsnorm, pf_resample_indices, pf_roughening_covariance
    c = GOLD['pf_resample_block']['cases'][case]
    X, q, ns, nx = np.array(c['X_in']), np.array(c['q']), c['n_s'], c['n_x']
    np.random.seed(c['seed'])
    unif = np.random.random_sample(ns)
    idx = pf_resample_indices(q, unif)
    assert np.array_equal(idx, np.array(c['ind']))          # bit-exact
    Xr = X[:, idx]
    if c['roughening']:
        Xr = Xr + lhsnorm(np.zeros(nx), pf_roughening_covariance(Xr, c['K']), ns).T
    np.testing.assert_allclose(Xr, c['X_out'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(Xr.sum(axis=1) / ns, c['x_mean'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.cov(Xr), c['P'], rtol=1e-12, atol=1e-12)
