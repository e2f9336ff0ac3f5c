"""Pins the CPU oracle against the reference's own fixtures / known-answer tests.

Each test re-expresses one ScalaTest case of the reference; the cite is in the docstring.
"""
import json
import os

import numpy as np
import pytest

from oracle import keystone_oracle as ko


def _load(golden_dir, a, b):
    A = np.loadtxt(os.path.join(golden_dir, a), delimiter=",", ndmin=2)
    B = np.loadtxt(os.path.join(golden_dir, b), delimiter=",", ndmin=2)
    return A, B


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "golden.json")) as fh:
        return json.load(fh)


# ---- BlockWeightedLeastSquaresSuite ---------------------------------------------------
def test_bwls_zero_gradient(golden_dir):
    """T/nodes/learning/BlockWeightedLeastSquaresSuite.scala:142-166 (b=4, 10 iters, ||grad|| < 1e-2)."""
    A, B = _load(golden_dir, "aMat.csv", "bMat.csv")
    xs, fb = ko.bwls_fit(A, B, 4, 10, 0.1, 0.3)
    g = ko.compute_gradient(A, B, 0.1, 0.3, np.concatenate(xs, 0), fb)
    assert np.linalg.norm(g) < 1e-2


def test_bwls_matches_per_class(golden_dir):
    """:115-140  BWLS model == PerClass model to 1e-6 after 5 iters; intercept norms agree."""
    A, B = _load(golden_dir, "aMat.csv", "bMat.csv")
    xs, fb = ko.bwls_fit(A, B, 4, 5, 0.1, 0.3)
    ps, pb = ko.per_class_wls_fit(A, B, 4, 5, 0.1, 0.3)
    assert np.linalg.norm(np.concatenate(xs, 0) - np.concatenate(ps, 0)) < 1e-6
    assert abs(np.linalg.norm(fb) - np.linalg.norm(pb)) < 1e-6


def test_bwls_ragged_blocks(golden_dir):
    """:188-223  nFeatures=12 not divisible by blockSize=5; both solvers ||grad|| < 1e-1."""
    A, B = _load(golden_dir, "aMat.csv", "bMat.csv")
    xs, fb = ko.bwls_fit(A, B, 5, 10, 0.1, 0.3)
    assert [x.shape[0] for x in xs] == [5, 5, 2]
    g = ko.compute_gradient(A, B, 0.1, 0.3, np.concatenate(xs, 0), fb)
    assert np.linalg.norm(g) < 1e-1
    ps, pb = ko.per_class_wls_fit(A, B, 5, 10, 0.1, 0.3)
    g2 = ko.compute_gradient(A, B, 0.1, 0.3, np.concatenate(ps, 0), pb)
    assert np.linalg.norm(g2) < 1e-1


def test_bwls_group_by_classes(golden_dir):
    """:225-253  fitting un-grouped rows reshuffles by class and reaches the same gradient bound;
    the shuffled fixtures are a row permutation of the sorted ones."""
    A, B = _load(golden_dir, "aMat.csv", "bMat.csv")
    As, Bs = _load(golden_dir, "aMatShuffled.csv", "bMatShuffled.csv")
    # three contiguous partitions of shuffled rows violate one-class-per-partition -> reshuffle
    parts = [np.arange(0, 5), np.arange(5, 10), np.arange(10, 15)]
    xs, fb = ko.bwls_fit(As, Bs, 4, 10, 0.1, 0.3, partitions=parts)
    g = ko.compute_gradient(As, Bs, 0.1, 0.3, np.concatenate(xs, 0), fb)
    assert np.linalg.norm(g) < 1e-2
    xs0, fb0 = ko.bwls_fit(A, B, 4, 10, 0.1, 0.3)
    if sorted(map(tuple, np.round(As, 12))) == sorted(map(tuple, np.round(A, 12))):
        assert np.allclose(np.concatenate(xs, 0), np.concatenate(xs0, 0), atol=1e-9)


def test_bwls_empty_partition_and_single_class(golden_dir):
    """:72-113 (a class with no rows must not crash; its column stays 0) and :168-186 (1 class)."""
    A, B = _load(golden_dir, "aMat.csv", "bMat.csv")
    keep = np.r_[0:5, 10:15]
    parts = [np.arange(0, 5), np.arange(0), np.arange(5, 10)]
    xs, fb = ko.bwls_fit(A[keep], B[keep], 4, 10, 0.1, 0.3, partitions=parts)
    W = np.concatenate(xs, 0)
    assert np.all(np.isfinite(W)) and np.all(W[:, 1] == 0.0)
    A1, B1 = _load(golden_dir, "aMat-1class.csv", "bMat-1class.csv")
    xs1, fb1 = ko.bwls_fit(A1, B1, 4, 10, 0.1, 0.3)
    assert np.all(np.isfinite(np.concatenate(xs1, 0)))


def test_bwls_regression_pins(golden, golden_dir):
    """Committed oracle outputs on the fixtures: later oracle edits must not drift."""
    A, B = _load(golden_dir, "aMat.csv", "bMat.csv")
    for key, (b, it) in {"b4_it10": (4, 10), "b5_it10": (5, 10), "b4_it5": (4, 5)}.items():
        xs, fb = ko.bwls_fit(A, B, b, it, 0.1, 0.3)
        assert np.allclose(np.concatenate(xs, 0), np.array(golden["bwls"][key]["W"]), atol=1e-12)
        assert np.allclose(fb, np.array(golden["bwls"][key]["final_b"]), atol=1e-12)


# ---- StandardScalerSuite --------------------------------------------------------------
def test_standard_scaler_golden(golden):
    """T/nodes/stats/StandardScalerSuite.scala:21-28,58-59 -- pins the n-1 (sample) variance."""
    g = golden["standard_scaler"]
    data = np.array(g["dense_data"])
    mean, std = ko.standard_scaler_fit(data)
    out = ko.standard_scaler_apply(data, mean, std)
    assert np.allclose(out[0], g["row0"], atol=g["tol"])
    assert np.allclose(out[3], g["row3"], atol=g["tol"])
    assert np.allclose(out.mean(0), 0, atol=1e-5) and np.allclose(out.var(0, ddof=1), 1, atol=1e-5)
    # constant column -> std guard -> zeros (:13-19)
    m2, s2 = ko.standard_scaler_fit(np.full((3, 1), 2.0))
    assert np.all(ko.standard_scaler_apply(np.full((3, 1), 2.0), m2, s2) == 0.0)


# ---- VectorSplitterSuite --------------------------------------------------------------
@pytest.mark.parametrize("bs,mul", [(128, 1), (128, 2), (128, 4)])
def test_vector_splitter(bs, mul):
    """T/nodes/util/VectorSplitterSuite.scala:7-37 -- ceil(D/b) blocks, order-preserving concat, ragged tail."""
    for d in (bs * mul, bs * mul + 7, bs * mul - 3):
        v = np.arange(d, dtype=float)
        parts = ko.vector_splitter(v, bs)
        assert len(parts) == int(np.ceil(d / bs))
        assert np.array_equal(np.concatenate(parts), v)
        assert all(len(p) == bs for p in parts[:-1])
    # numFeaturesOpt override
    parts = ko.vector_splitter(np.arange(20.0), 8, num_features=12)
    assert [len(p) for p in parts] == [8, 4]


# ---- CosineRandomFeaturesSuite --------------------------------------------------------
@pytest.mark.parametrize("dist", ["gaussian", "cauchy"])
def test_cosine_random_features(dist):
    """T/nodes/stats/CosineRandomFeaturesSuite.scala:16-57 -- shapes, b in [0,2pi], formula to 1e-2."""
    rng = np.random.default_rng(0)
    n_in, n_out, gamma = 400, 1000, 1.34
    W, b = ko.cosine_random_features_params(n_in, n_out, gamma, rng, dist)
    assert W.shape == (n_out, n_in) and b.shape == (n_out,)
    assert b.max() <= 2 * np.pi and b.min() >= 0
    if dist == "gaussian":
        assert abs(W.mean()) < 1e-2 and abs(W.var() - gamma ** 2) < 2e-2
    else:
        assert abs(np.median(W)) < 2e-2
    x = rng.random(n_in)
    out = ko.cosine_random_features(x, W, b)
    assert np.allclose(out, np.cos((x @ W.T) + b), atol=1e-2)
    X = rng.random((5, n_in))
    assert np.allclose(ko.cosine_random_features(X, W, b)[2], ko.cosine_random_features(X[2], W, b))


# ---- BlockLinearMapperSuite -----------------------------------------------------------
def test_block_linear_mapper_equals_linear_mapper():
    """T/nodes/learning/BlockLinearMapperSuite.scala:18-55 -- block apply == dense apply (1e-4);
    applyAndEvaluate's last callback equals it."""
    rng = np.random.default_rng(1)
    in_dim, out_dim, bs = 1000, 100, 200
    mat = rng.standard_normal((in_dim, out_dim))
    vec = rng.standard_normal(in_dim)
    intercept = rng.standard_normal(out_dim)
    xs = [mat[s:e] for s, e in ko.block_bounds(in_dim, bs)]
    dense = ko.linear_mapper_apply(vec, mat, intercept)
    blk, partials = ko.block_linear_apply(vec[None, :], xs, bs, intercept, None, return_partials=True)
    assert np.allclose(blk[0], dense, atol=1e-4)
    assert np.allclose(partials[-1][0], dense, atol=1e-4)
    assert len(partials) == 5


# ---- LinearMapperSuite ----------------------------------------------------------------
def test_linear_map_estimator_known_answer(golden):
    """T/nodes/learning/LinearMapperSuite.scala:13-36 -- exact solve recovers x to 1e-8, mapper(point) = 5."""
    g = golden["linear_mapper"]
    rng = np.random.default_rng(42)
    A = rng.standard_normal((128, 5))
    x = np.array(g["x"])[:, None]
    b = A @ x
    xh, ymu, mu = ko.linear_map_fit(A, b)
    assert np.allclose(xh, x, atol=g["tol"])
    pt = np.array(g["point"])
    assert abs(ko.linear_mapper_apply(pt, xh, ymu, mu)[0] - g["expected"]) < 1e-8
    assert np.allclose(ko.linear_mapper_apply(A, xh, ymu, mu)[0], b[0], atol=1e-8)


# ---- BlockLeastSquaresEstimator (mlmatrix boundary: invariants only, parity unpinned) ----
def test_block_ls_single_block_is_centred_ridge():
    rng = np.random.default_rng(3)
    A = rng.standard_normal((300, 24)) + 0.5
    Y = rng.standard_normal((300, 4))
    xs, b, mus = ko.block_ls_fit(A, Y, 24, 1, 0.7)
    x2, b2, mu2 = ko.linear_map_fit(A, Y, 0.7)
    assert np.allclose(xs[0], x2, atol=1e-10) and np.allclose(b, b2) and np.allclose(mus[0], mu2)
    # LinearMapperSuite known answer through the block solver (nb = 1, lambda = 0)
    x = np.array([5.0, 4.0, 3.0, 2.0, -1.0])[:, None]
    A5 = rng.standard_normal((128, 5))
    xs5, b5, mu5 = ko.block_ls_fit(A5, A5 @ x, 5, 1, 0.0)
    assert np.allclose(xs5[0], x, atol=1e-8)


def test_block_ls_sweeps_converge_and_cost_monotone():
    rng = np.random.default_rng(4)
    A = rng.standard_normal((400, 30)) + 0.2 * rng.standard_normal((400, 1)) + 1.0
    Y = ko.class_label_indicators(rng.integers(0, 5, 400), 5)
    lam = 2.0
    exact, yb, mu = ko.linear_map_fit(A, Y, lam)
    prev = None
    for iters in (1, 2, 5, 60):
        xs, b, mus = ko.block_ls_fit(A, Y, 8, iters, lam)
        W = np.concatenate(xs, 0)
        err = np.linalg.norm(W - exact)
        Ac = A - np.concatenate(mus)
        obj = 0.5 * ((Ac @ W + b - Y) ** 2).sum() + 0.5 * lam * (W ** 2).sum()
        if prev is not None:
            assert err <= prev[0] + 1e-12 and obj <= prev[1] + 1e-9
        prev = (err, obj)
    assert prev[0] < 1e-6
    assert [x.shape[0] for x in xs] == [8, 8, 8, 6]


def test_block_ls_from_partial_sums_matches():
    """The sharded algebra (shifted blocks + reduced sums + rank-1 correction)
    equals the direct centred solve."""
    rng = np.random.default_rng(5)
    n, bsz, k, lam = 257, 12, 3, 0.4
    A = rng.standard_normal((n, bsz)) + 2.0
    Y = rng.standard_normal((n, k))
    R = Y - Y.mean(0)
    shift = A[:40].mean(0)            # any shared estimate of the mean
    shards = np.array_split(np.arange(n), 3)
    sums = [ko.block_ls_partial_sums(A[s] - shift, R[s]) for s in shards]
    G, C, sa, sr = (sum(x) for x in zip(*sums))
    dW, delta = ko.block_ls_solve_from_sums(G, C, sa, sr, n, lam)
    xs, b, mus = ko.block_ls_fit(A, Y, bsz, 1, lam)
    assert np.allclose(dW, xs[0], atol=1e-9)
    assert np.allclose(shift + delta, mus[0], atol=1e-12)


def test_compute_cost_definition():
    """K/nodes/learning/BlockLinearMapper.scala:142-187."""
    rng = np.random.default_rng(6)
    A = rng.standard_normal((50, 10)); Y = rng.standard_normal((50, 2)); W = rng.standard_normal((10, 2))
    b = rng.standard_normal(2)
    xs = [W[:4], W[4:8], W[8:]]
    c0 = ko.compute_cost(A, Y, 0.0, xs, 4, b)
    assert np.isclose(c0, ((A @ W + b - Y) ** 2).sum() / 100.0)
    c1 = ko.compute_cost(A, Y, 0.3, xs, 4, b)
    assert np.isclose(c1, c0 + 0.15 * (W ** 2).sum())


def test_multiclass_evaluator_known_answer():
    """T/evaluation/MulticlassClassifierEvaluatorSuite.scala:9-68: the suite's 9 (prediction, label) pairs, its confusion
    matrix and every metric it asserts (delta 1e-7)."""
    pairs = [(0, 0), (0, 1), (0, 0), (1, 0), (1, 1), (1, 1), (1, 1), (2, 2), (2, 0)]
    pred, act = np.array([p for p, _ in pairs]), np.array([a for _, a in pairs])
    cm = ko.confusion_matrix(pred, act, 3)
    assert np.array_equal(cm, np.array([[2, 1, 1], [1, 3, 0], [0, 0, 1]], dtype=float))   # rows = true class
    p = [2.0 / 3, 3.0 / 4, 1.0 / 2]
    r = [2.0 / 4, 3.0 / 4, 1.0]
    f1 = [2 * a * b / (a + b) for a, b in zip(p, r)]
    f2 = [5 * a * b / (4 * a + b) for a, b in zip(p, r)]
    m1, m2 = ko.multiclass_metrics(cm), ko.multiclass_metrics(cm, beta=2.0)
    d = 1e-7
    assert np.abs(m1["class_precision"] - p).max() < d and np.abs(m1["class_recall"] - r).max() < d
    assert np.abs(m1["class_fscore"] - f1).max() < d and np.abs(m2["class_fscore"] - f2).max() < d
    assert abs(m1["micro_recall"] - 6.0 / 9.0) < d
    assert abs(m1["micro_recall"] - m1["micro_precision"]) < d and abs(m1["micro_recall"] - m1["micro_fscore"]) < d
    assert abs(m1["macro_precision"] - sum(p) / 3) < d and abs(m1["macro_recall"] - sum(r) / 3) < d
    assert abs(m1["macro_fscore"] - sum(f1) / 3) < d and abs(m2["macro_fscore"] - sum(f2) / 3) < d
    assert abs(m1["total_accuracy"] - 6.0 / 9.0) < d and abs(m1["total_error"] - 3.0 / 9.0) < d


# ---- pin of block_ls_fit on in-repo reference arithmetic ---------------------------------------------------------------
@pytest.mark.parametrize("bs", [4, 5, 12])
@pytest.mark.parametrize("iters", [1, 2, 3])
def test_block_ls_pinned_by_bwls_with_zero_mixture_weight(golden_dir, bs, iters):
    """`trainWithL2` with mixtureWeight = 0 (K/nodes/learning/BlockWeightedLeastSquares.scala:216-273: jointXTX = popCov =
    A_c^T A_c / N, jointXTR = popXTR - popMean * residualMean, solve (popCov + lambda I) dW = jointXTR - lambda W) is
    algebraically the BlockLS step with lambda * N: (A_c^T A_c + lambda N I) dW = A_c^T R - lambda N W.  The reference's own
    suite pins `trainWithL2` on these fixtures, so this identity ties `block_ls_fit` (whose mlmatrix arithmetic is absent from
    /root/reference) to reference code that IS in the repo: weights per block and the folded intercept
    ybar - sum_j mu_j^T W_j (:314-319) must agree to rounding."""
    A, B = _load(golden_dir, "aMat.csv", "bMat.csv")
    n = A.shape[0]
    lam = 0.1
    xs_w, final_b = ko.bwls_fit(A, B, bs, iters, lam, 0.0)
    xs, ybar, mus = ko.block_ls_fit(A, B, bs, iters, lam * n)
    for xw, x in zip(xs_w, xs):
        assert np.abs(xw - x).max() < 1e-12
    folded = ybar - sum(mu @ x for mu, x in zip(mus, xs))
    assert np.abs(final_b - folded).max() < 1e-12
    # and on the shuffled fixture (rows regrouped by class inside trainWithL2: the BlockLS result is row-order invariant)
    As, Bs = _load(golden_dir, "aMatShuffled.csv", "bMatShuffled.csv")
    xs_s, _, _ = ko.block_ls_fit(As, Bs, bs, iters, lam * n)
    for xw, x in zip(xs_w, xs_s):
        assert np.abs(xw - x).max() < 1e-10


# ---- MNIST random-FFT featurizer nodes -------------------------------------------------------------------------------
def test_padded_fft_known_answers():
    """T/nodes/stats/PaddedFFTSuite.scala:13-36: length-100 unit impulses -> 64 real bins, agreement with R's Re(fft(.))."""
    ones = np.zeros(100); ones[0] = 1.0
    twos = np.zeros(100); twos[2] = 1.0
    out = ko.padded_fft(np.stack([twos, ones]))
    assert out.shape == (2, 64)
    assert abs(out[0, 0] - 1.0) < 1e-8 and abs(out[0, 16]) < 1e-8 and abs(out[0, 32] + 1.0) < 1e-8 and abs(out[0, 48]) < 1e-8
    assert np.abs(out[1] - 1.0).max() < 1e-8
    assert ko.next_positive_power_of_two(784) == 1024 and ko.next_positive_power_of_two(1024) == 1024


def test_random_sign_node_and_linear_rectifier():
    """T/nodes/stats/RandomSignNodeSuite.scala:11-18 and LinearRectifierSuite.scala:13-27."""
    assert np.array_equal(ko.random_sign_node(np.array([1.0, 2.0, 3.0]), np.array([1.0, -1.0, 1.0])), np.array([1.0, -2.0, 3.0]))
    x = np.random.default_rng(0).standard_normal((128, 16))
    assert (x < 0).any() and (ko.linear_rectifier(x) >= 0).all()
    assert np.array_equal(ko.linear_rectifier(np.array([-1.0, 0.5, 3.0]), 0.25, 1.0), np.array([0.25, 0.25, 2.0]))


def test_fft_featurizer_is_a_cosine_matrix_product():
    """The identity the device path uses: Re(FFT(pad(x .* s)))[f] = sum_n x[n] s[n] cos(2 pi f n / P)."""
    rng = np.random.default_rng(1)
    x = rng.random((5, 784)); s = 2.0 * rng.integers(0, 2, 784) - 1.0
    P = 1024
    M = s[None, :] * np.cos(2 * np.pi * np.outer(np.arange(P // 2), np.arange(784)) / P)
    assert np.abs(ko.padded_fft(ko.random_sign_node(x, s)) - x @ M.T).max() < 1e-9


# ---- CIFAR random-patch featurizer nodes ------------------------------------------------------------
def _gantrycrane(golden_dir):
    z = np.load(os.path.join(golden_dir, "conv_gantrycrane.npz"))
    return ko.image_from_bgr_bytes(z["rgb"]), z["expected"].astype(np.float64)


def test_convolver_matches_reference_golden_image(golden_dir):
    """T/nodes/images/ConvolverSuite.scala:100-137 ("convolutions should match scipy"): 3 x 3 x 3 filters with flipFilters = true,
    no patch normalisation; channel 0 of the result equals the reference's convolved.gantrycrane.csv EXACTLY (integers)."""
    img, expected = _gantrycrane(golden_dir)
    kimg, kimg2 = np.zeros((3, 3, 3)), np.zeros((3, 3, 3))
    i = 0
    for x in range(3):
        for y in range(3):
            for c in range(3):
                kimg[x, y, 2 - c] = float(i)        # channel order reversed to match python (:108)
                i += 1
    kimg2[0, 0, 0] = 2.0
    kimg2[2, 0, 1] = 1.0
    filt = ko.pack_filters([ko.flip_image(kimg), ko.flip_image(kimg2)])
    conv = ko.convolve(img, filt, 3, normalize=False)
    assert conv.shape == (expected.shape[0], expected.shape[1], 2)
    assert np.array_equal(conv[:, :, 0], expected)


def test_convolver_small_cases_and_patch_layout():
    """ConvolverSuite.scala:13-98: output geometry for 1 x 1 and 3 x 3 filters on the 4 x 4 / 10 x 10 ramp images, and the patch
    column order c + pox*C + poy*C*convSize of Convolver.makePatches (:152-186)."""
    w, h, ch = 10, 10, 3
    img = np.zeros((w, h, ch))
    for x in range(w):
        for y in range(h):
            for c in range(ch):
                img[x, y, c] = c + x * ch + y * w * ch
    conv1 = np.zeros(27); conv1[4] = 1.0
    conv2 = np.zeros(27); conv2[4] = conv2[13] = conv2[22] = 0.33
    out = ko.convolve(img, np.stack([conv1, conv2]), 3, normalize=True)
    assert out.shape == (8, 8, 2)
    pm = ko.make_patches(img, 3, normalize=False)
    assert pm.shape == (64, 27)
    assert pm[2 + 3 * 8, 1 + 2 * 3 + 1 * 9] == img[2 + 2, 3 + 1, 1]      # row x + y*resW, column c + pox*C + poy*C*convSize
    n = ko.normalize_rows(pm, 10.0)
    assert np.abs(n.mean(axis=1)).max() < 1e-12
    v = ((pm - pm.mean(1, keepdims=True)) ** 2).sum(1) / 26.0
    assert np.allclose(n, (pm - pm.mean(1, keepdims=True)) / np.sqrt(v + 10.0)[:, None])


def test_symmetric_rectifier_pooler_vectorizer():
    """SymmetricRectifier.scala:7-32, Pooler.scala:21-69 (CIFAR geometry: 27 x 27 -> 2 x 2 pools of 14 x 14 that overlap in
    row / column 13), Image.toArray order (Image.scala:47-65)."""
    rng = np.random.default_rng(0)
    img = rng.standard_normal((27, 27, 5))
    r = ko.symmetric_rectifier(img, 0.0, 0.25)
    assert r.shape == (27, 27, 10) and (r >= 0).all()
    assert np.array_equal(r[:, :, :5], np.maximum(0, img - 0.25)) and np.array_equal(r[:, :, 5:], np.maximum(0, -img - 0.25))
    p = ko.pooler(r, 13, 14)
    assert p.shape == (2, 2, 10)
    assert np.allclose(p[0, 0], r[0:14, 0:14].sum((0, 1))) and np.allclose(p[1, 0], r[13:27, 0:14].sum((0, 1)))
    assert np.allclose(p[1, 1], r[13:27, 13:27].sum((0, 1)))
    v = ko.image_vectorizer(p)
    assert v.shape == (40,) and v[3 + 1 * 10 + 0 * 10 * 2] == p[1, 0, 3]
