"""Size-independent properties of the GPU fit (no CPU oracle needed, so they also run at BASELINE.json's full config-3 size):
linearity in the labels, invariance to row order, equivariance to a permutation of the classes, first-order optimality of
the last block, and the host round trip of a fitted model."""
import os

import numpy as np
import pytest

import keystone_b200 as ks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ks.Context(0)
    yield c
    c.close()


def _params(rng, d_in, n_out, nrf, gamma):
    return [(rng.standard_normal((n_out, d_in)) * gamma, rng.random(n_out) * 2 * np.pi) for _ in range(nrf)]


def _feats(ctx, x, params):
    rfs = [ks.CosineRandomFeatures(ctx, W, b) for W, b in params]
    return ks.Pipeline.gather(rfs).andThen(ks.VectorCombiner())(x), rfs


def test_mid_size_properties(ctx):
    rng = np.random.default_rng(0)
    n, d_in, n_out, nrf, k, lam = 200_000, 64, 1024, 2, 20, 0.5
    X = rng.standard_normal((n, d_in)).astype(np.float32)
    Y = rng.standard_normal((n, k)).astype(np.float32) + (X[:, :k] > 0)
    params = _params(rng, d_in, n_out, nrf, 0.15)
    x = ctx.matrix(X); feats, _ = _feats(ctx, x, params)
    est = ks.BlockLeastSquaresEstimator(n_out, 1, lam)
    m1 = est.fit(feats, ctx.matrix(Y))
    W1 = np.concatenate(m1.xs, 0)
    scale = np.abs(W1).max()
    # (1) linearity in the labels: powers of two commute with every rounding; what is left is the run-to-run order of the
    #     fp32 reduce-adds (split-K partial tiles are combined in arrival order)
    m2 = est.fit(feats, ctx.matrix(-4.0 * Y))
    assert np.abs(np.concatenate(m2.xs, 0) + 4.0 * W1).max() < 2e-4 * 4 * scale
    assert np.abs(m2.b_opt + 4.0 * m1.b_opt).max() < 1e-5
    # (2) row order does not matter (every statistic is a sum over rows; only the atomics' order changes)
    perm = rng.permutation(n)
    xp = ctx.matrix(X[perm]); fp, _ = _feats(ctx, xp, params)
    mp_ = est.fit(fp, ctx.matrix(Y[perm]))
    assert np.abs(np.concatenate(mp_.xs, 0) - W1).max() < 2e-3 * scale   # the shift estimate (first rows) changes
    # (3) first-order optimality of the LAST block after one sweep: A_c^T (Y_c - pred_c) = lambda W_last
    pred = m1(feats).to_numpy()
    last = ks.CosineRandomFeatures(ctx, *params[-1])(x).to_numpy()        # N x n_out, generated on the device
    last_c = last - m1.feature_means[-1]
    resid = (Y - m1.b_opt) - (pred - m1.b_opt)
    g = last_c.T @ resid - lam * m1.xs[-1]
    assert np.abs(g).max() < 5e-3 * np.abs(last_c.T @ (Y - Y.mean(0))).max()
    # (4) a model rebuilt from its host arrays applies identically (BlockLinearMapper constructor, :22-27)
    m3 = ks.BlockLinearMapper.from_arrays(ctx, m1.xs, n_out, m1.b_opt, m1.feature_means)
    assert np.abs(m3(feats).to_numpy() - pred).max() < 1e-5 * max(1.0, np.abs(pred).max())


@pytest.mark.skipif(os.environ.get("KS_SKIP_FULL_SIZE") == "1", reason="full-size test disabled")
def test_full_size_config3_class_permutation_and_linearity(ctx):
    """BASELINE.json config 3 at full size (N = 1M, d_in = 440, D = 16 x 4096, k = 1000, b = 4096): inputs generated on the
    device; permuting the classes must permute the model's columns, negating the labels must negate it."""
    n, d_in, n_out, nrf, k, lam = 1_000_000, 440, 4096, 16, 1000, 1.0
    rng = np.random.default_rng(1)
    x = ctx.synthetic_normal(n, d_in, seed=3)
    cls = rng.integers(0, k, n).astype(np.int32)
    params = _params(rng, d_in, n_out, nrf, 0.0555)
    feats, _ = _feats(ctx, x, params)
    est = ks.BlockLeastSquaresEstimator(n_out, 1, lam, precision="f16")      # the mode bench.py's headline line times
    m1 = est.fit(feats, ctx.labels_from_classes(cls, k))
    stats = ctx.last_fit_stats()
    assert stats["num_blocks"] == 16 and stats["n_total"] == n and stats["mma"] == "f16"
    sigma = rng.permutation(k).astype(np.int32)                         # class c -> sigma[c]
    m2 = est.fit(feats, ctx.labels_from_classes(sigma[cls], k))
    for j in (0, 7, 15):
        W1, W2 = m1.xs[j], m2.xs[j]
        assert W1.shape == (n_out, k)
        assert np.abs(W2[:, sigma] - W1).max() < 1e-3 * np.abs(W1).max()   # class order changes the tiles each column lands in
    assert np.abs(m2.b_opt[sigma] - m1.b_opt).max() < 1e-6
    b1 = m1.b_opt
    assert np.allclose(b1, 2.0 * np.bincount(cls, minlength=k) / n - 1.0, atol=1e-6)   # intercept = label mean
