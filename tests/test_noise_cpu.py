"""CPU: the host twin of the root-noise generator (hz_host_dirichlet_noise: Philox4x32-10 + Marsaglia-Tsang, IEEE
float64 operations only) draws Dirichlet(alpha) — checked against numpy's own sampler (what the reference calls,
core/selfplay_worker.py:279) through moments and quantiles — and is a pure function of (seed, step, root, action)."""
import numpy as np
import pytest


@pytest.mark.parametrize("alpha,A", [(0.3, 20), (0.3, 11), (1.5, 6)])
def test_host_dirichlet_matches_numpy_distribution(alpha, A):
    from hanabizero_b200.selfplay import dirichlet_noise_host
    n = 20000
    x = dirichlet_noise_host(n, A, alpha, seed=11, step=3).astype(np.float64)
    ref = np.random.default_rng(0).dirichlet([alpha] * A, n)
    np.testing.assert_allclose(x.sum(1), 1.0, atol=2e-6)
    assert abs(x.mean() - 1.0 / A) < 1e-9 + 1e-6
    # Dirichlet(alpha * 1_A): Var = (A - 1) / (A^2 (A alpha + 1)); the largest component has no closed form -> numpy
    var = (A - 1) / (A * A * (A * alpha + 1))
    assert abs(x.var() - var) / var < 0.03 and abs(ref.var() - var) / var < 0.03
    assert abs(x.max(1).mean() - ref.max(1).mean()) < 0.01
    for q in (0.25, 0.5, 0.9, 0.99):
        a, b = np.quantile(x, q), np.quantile(ref, q)
        assert abs(a - b) < 0.03 * max(b, 1e-3) + 2e-4, (q, a, b)


def test_host_dirichlet_is_a_pure_function_of_its_counters():
    from hanabizero_b200.selfplay import dirichlet_noise_host
    full = dirichlet_noise_host(64, 20, 0.3, seed=5, step=9)
    assert (full == dirichlet_noise_host(64, 20, 0.3, seed=5, step=9)).all()
    shard = dirichlet_noise_host(16, 20, 0.3, seed=5, step=9, root_offset=32)
    assert (shard == full[32:48]).all()                         # a rank's shard draws what the full batch draws
    assert not (full == dirichlet_noise_host(64, 20, 0.3, seed=5, step=10)).all()
    assert not (full == dirichlet_noise_host(64, 20, 0.3, seed=6, step=9)).all()
    legal = (np.random.default_rng(1).random((64, 20)) < 0.5).astype(np.float32)
    masked = dirichlet_noise_host(64, 20, 0.3, seed=5, step=9, legal=legal)
    assert (masked == full * legal).all()
