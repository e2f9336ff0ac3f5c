"""Oracle parity at BASELINE.json's shapes, row-subsampled so that the fp64 CPU oracle finishes in about a minute:
every feature block, every class and the benchmark's lambda are the real ones; only N is reduced (the per-sample arithmetic
does not depend on N).  Both operand modes that bench.py reports are checked: the parity mode (split operands) and the
fast mode (fp16 / tf32 operands, 10-bit mantissa)."""
import os

import numpy as np
import pytest

import keystone_b200 as ks
from oracle import keystone_oracle as ko

pytestmark = pytest.mark.gpu

N_SUB = int(os.environ.get("KS_TEST_SUBSAMPLE_ROWS", "32768"))


@pytest.fixture(scope="module")
def ctx():
    c = ks.Context(0)
    yield c
    c.close()


def _report(tag, Wg, Wr, pred, ref):
    rel = float(np.linalg.norm(Wg - Wr) / np.linalg.norm(Wr))
    print(f"{tag}: rel-Fro(W) = {rel:.3e}, max|pred err| = {np.abs(pred - ref).max():.3e}, "
          f"argmax agreement = {(pred.argmax(1) == ref.argmax(1)).mean():.5f}")
    return rel


def test_config3_subsampled_all_blocks(ctx):
    """C3: d_in = 440, 16 x CosineRandomFeatures(440 -> 4096), gamma = 0.0555, k = 1000, b = 4096, lambda = 1, numIter = 1."""
    n, d_in, n_out, nrf, k, lam = N_SUB, 440, 4096, 16, 1000, 1.0
    rng = np.random.default_rng(2)
    params = [(rng.standard_normal((n_out, d_in)) * 0.0555, rng.random(n_out) * 2 * np.pi) for _ in range(nrf)]
    wstar = rng.standard_normal((16, k)).astype(np.float32)
    X = rng.standard_normal((n, d_in)).astype(np.float32)
    cls = np.argmax(X[:, :16] @ wstar + 0.1 * rng.standard_normal((n, k)).astype(np.float32), axis=1)
    x = ctx.matrix(X)
    y = ctx.labels_from_classes(cls, k)
    rfs = [ks.CosineRandomFeatures(ctx, W, b) for W, b in params]
    feats = ks.Pipeline.gather(rfs).andThen(ks.VectorCombiner())(x)
    got = {}
    for prec in ("f16x2", "f16"):
        m = ks.BlockLeastSquaresEstimator(n_out, 1, lam, precision=prec).fit(feats, y)
        assert ctx.last_fit_stats()["mma"] == prec
        got[prec] = (np.concatenate(m.xs, 0).copy(), m(feats).to_numpy())
    Xd = X.astype(np.float64)
    blocks = [ko.cosine_random_features(Xd, W, b) for W, b in params]
    Y = ko.class_label_indicators(cls, k)
    xs, b0, mus = ko.block_ls_fit(None, Y, n_out, 1, lam, feature_blocks=blocks)
    Wr = np.concatenate(xs, 0)
    ref = sum((blk - mu) @ w for blk, mu, w in zip(blocks, mus, xs)) + b0
    rel2 = _report("C3 subsample, parity mode", got["f16x2"][0], Wr, got["f16x2"][1], ref)
    rel1 = _report("C3 subsample, fast mode  ", got["f16"][0], Wr, got["f16"][1], ref)
    assert rel2 < 5e-5, rel2
    assert np.abs(got["f16x2"][1] - ref).max() < 1e-4
    assert rel1 < 1.5e-3, rel1
    assert (got["f16"][1].argmax(1) == ref.argmax(1)).mean() > 0.995


def test_config2_subsampled_materialised(ctx):
    """C2: materialised features D = 16384 with the 0.1 j / D column mean of SURVEY 8d, k = 100, b = 4096, lambda = 10."""
    n, d, k, bs, lam = N_SUB, 16384, 100, 4096, 10.0
    rng = np.random.default_rng(1)
    F = rng.standard_normal((n, d), dtype=np.float32) + (0.1 * np.arange(d) / d).astype(np.float32)
    wstar = (rng.standard_normal((64, k)) / 8.0).astype(np.float32)
    cls = np.argmax(F[:, :64] @ wstar + 0.1 * rng.standard_normal((n, k)).astype(np.float32), axis=1)
    f = ctx.matrix(F)
    y = ctx.labels_from_classes(cls, k)
    got = {}
    for prec, mma in (("f16x2", "tf32x2"), ("tf32", "tf32x1")):
        m = ks.BlockLeastSquaresEstimator(bs, 1, lam, precision=prec).fit(f, y)
        assert ctx.last_fit_stats()["mma"] == mma
        got[prec] = (np.concatenate(m.xs, 0).copy(), m(f).to_numpy())
    Fd = F.astype(np.float64)
    Y = ko.class_label_indicators(cls, k)
    xs, b0, mus = ko.block_ls_fit(Fd, Y, bs, 1, lam)
    Wr = np.concatenate(xs, 0)
    ref = ko.block_linear_apply(Fd, xs, bs, b0, mus)
    rel2 = _report("C2 subsample, parity mode", got["f16x2"][0], Wr, got["f16x2"][1], ref)
    rel1 = _report("C2 subsample, tf32 mode  ", got["tf32"][0], Wr, got["tf32"][1], ref)
    assert rel2 < 5e-5, rel2
    assert np.abs(got["f16x2"][1] - ref).max() < 1e-4
    assert rel1 < 1.5e-3, rel1
