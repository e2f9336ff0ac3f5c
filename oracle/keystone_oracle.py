"""CPU oracle for the KeystoneML block least-squares hot path (numpy, IEEE fp64).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``keystone_b200/``) may import
this module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker / the timed CPU stand-in.

Every function restates one reference function and cites it (paths are relative to
``/root/reference/src/main/scala/keystoneml`` = ``K/`` and ``src/test/scala/keystoneml`` =
``T/``).  The reference itself (Scala 2.10 + Spark 2.1 + Breeze 0.12 + mlmatrix 0.2) cannot
be compiled or imported in this image (no JVM), so there is no ``oracle/_ref``.

Pinning status
--------------
* ``bwls_fit`` / ``per_class_wls_fit`` / ``compute_gradient``: PINNED by the reference's own
  fixtures ``src/test/resources/{aMat,bMat,aMat-1class,bMat-1class}.csv`` and the
  assertions of ``T/nodes/learning/BlockWeightedLeastSquaresSuite.scala`` (see
  ``tests/test_oracle_golden.py``).
* ``standard_scaler_*``: PINNED by the golden rows of ``T/nodes/stats/StandardScalerSuite.scala``.
* ``cosine_random_features``, ``vector_splitter``, ``block_linear_apply``,
  ``linear_map_fit``: PINNED by the formula / known-answer tests of the matching suites.
* ``block_ls_fit``: PINNED INDIRECTLY through in-repo reference arithmetic.  Its own arithmetic
  lives in the un-vendored dependency ``edu.berkeley.cs.amplab:mlmatrix:0.2`` (``build.sbt:44``)
  and no reference test pins its numerical output directly.  But ``trainWithL2`` with
  ``mixtureWeight = 0`` (``K/nodes/learning/BlockWeightedLeastSquares.scala:216-273``, which the
  reference's suite pins on the fixtures) is algebraically the same block coordinate descent with
  ``lambda * N``: ``tests/test_oracle_golden.py::test_block_ls_pinned_by_bwls_with_zero_mixture_weight``
  checks ``bwls_fit(A, B, b, it, lam, 0) == block_ls_fit(A, B, b, it, lam * N)`` (weights and the
  folded intercept ``ybar - sum_j mu_j^T W_j``) to 1e-12 on ``aMat/bMat`` for b = 4, 5, 12 and
  1-3 sweeps, and the GPU tests check the same identity through ``ks_blockwls_fit`` /
  ``ks_blockls_fit``.  What REMAINS UNPINNED is mlmatrix's own conventions, which no in-repo code
  shows: that ``NormalEquations`` adds ``lambda`` un-scaled by N (taken from the call site
  ``K/nodes/learning/BlockLinearMapper.scala:234-240`` passing ``Array(lambda)`` straight through),
  and the block order of multi-sweep ``solveLeastSquaresWithL2`` (sequential assumed, as in
  ``trainWithL2`` ``:179-180``).  Further invariants in ``tests/test_oracle_golden.py``: nb = 1 ==
  closed-form centred ridge == ``LinearMapEstimator``; many sweeps converge to the same;
  LinearMapperSuite known answer.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

F64 = np.float64


# --------------------------------------------------------------------------------------
# a1  VectorSplitter            K/nodes/util/VectorSplitter.scala:15-35
# --------------------------------------------------------------------------------------
def block_bounds(num_features: int, block_size: int) -> List[Tuple[int, int]]:
    """[start, end) of each feature block; nb = ceil(D / blockSize), last block ragged
    (VectorSplitter.scala:16-22)."""
    nb = int(math.ceil(num_features / float(block_size)))
    return [(j * block_size, min(num_features, (j + 1) * block_size)) for j in range(nb)]


def vector_splitter(x: np.ndarray, block_size: int, num_features: Optional[int] = None) -> List[np.ndarray]:
    """Split rows (N x D) or one vector (D) into column blocks (VectorSplitter.scala:15-35).
    ``num_features`` mirrors ``numFeaturesOpt`` (overrides D)."""
    x = np.asarray(x, dtype=F64)
    d = x.shape[-1] if num_features is None else int(num_features)
    return [np.array(x[..., s:e], dtype=F64, copy=True) for s, e in block_bounds(d, block_size)]


# --------------------------------------------------------------------------------------
# a2  StandardScaler            K/nodes/stats/StandardScaler.scala:25-59
# --------------------------------------------------------------------------------------
def standard_scaler_fit(data: np.ndarray, normalize_std_dev: bool = True, eps: float = 1e-12):
    """Column mean and (optionally) sample (n-1) std with the ``<eps -> 1.0`` guard
    (StandardScaler.scala:45-59; MLlib MultivariateOnlineSummarizer.variance is unbiased)."""
    data = np.asarray(data, dtype=F64)
    mean = data.mean(axis=0)
    if not normalize_std_dev:
        return mean, None
    n = data.shape[0]
    var = ((data - mean) ** 2).sum(axis=0) / (n - 1) if n > 1 else np.zeros_like(mean)
    std = np.sqrt(var)
    bad = np.isnan(std) | np.isinf(std) | (np.abs(std) < eps)
    std = np.where(bad, 1.0, std)
    return mean, std


def standard_scaler_apply(x: np.ndarray, mean: np.ndarray, std: Optional[np.ndarray] = None) -> np.ndarray:
    """``(in - mean) [/ std]`` (StandardScaler.scala:25-31)."""
    out = np.asarray(x, dtype=F64) - mean
    if std is not None:
        out = out / std
    return out


# --------------------------------------------------------------------------------------
# a4  CosineRandomFeatures      K/nodes/stats/CosineRandomFeatures.scala:25-60
# --------------------------------------------------------------------------------------
def cosine_random_features(x: np.ndarray, W: np.ndarray, b: np.ndarray) -> np.ndarray:
    """``cos(X W^T + b)``; W is (n_out x n_in), b is (n_out) (CosineRandomFeatures.scala:30-32, 38-43)."""
    x = np.asarray(x, dtype=F64)
    return np.cos(x @ np.asarray(W, dtype=F64).T + np.asarray(b, dtype=F64))


def cosine_random_features_params(n_in: int, n_out: int, gamma: float, rng: np.random.Generator,
                                  w_dist: str = "gaussian") -> Tuple[np.ndarray, np.ndarray]:
    """Factory (CosineRandomFeatures.scala:51-60): ``W = gamma * rand(n_out x n_in, wDist)``,
    ``b = 2*pi * uniform(n_out)``.  Breeze's RNG stream cannot be reproduced without a JVM, so
    parity runs always share the *arrays*, never a seed."""
    if w_dist == "gaussian":
        W = rng.standard_normal((n_out, n_in))
    elif w_dist == "cauchy":
        W = rng.standard_cauchy((n_out, n_in))
    else:
        raise ValueError(w_dist)
    return W * gamma, rng.random(n_out) * (2.0 * math.pi)


# --------------------------------------------------------------------------------------
# MNIST random-FFT featurizer: K/nodes/stats/RandomSignNode.scala:11-24,
# PaddedFFT.scala:13-21, LinearRectifier.scala:12-17, K/pipelines/images/mnist/MnistRandomFFT.scala:40-44
# --------------------------------------------------------------------------------------
def random_sign_node(x: np.ndarray, signs: np.ndarray) -> np.ndarray:
    """``in :* signs`` (RandomSignNode.scala:15)."""
    return np.asarray(x, dtype=F64) * np.asarray(signs, dtype=F64)


def next_positive_power_of_two(i: int) -> int:
    """PaddedFFT.scala:20: ``1 << (32 - numberOfLeadingZeros(i - 1))``."""
    return 1 << max(0, (int(i) - 1).bit_length())


def padded_fft(x: np.ndarray) -> np.ndarray:
    """Zero-pad to the next power of two P, Fourier transform, real part of bins [0, P/2) (PaddedFFT.scala:14-18)."""
    x = np.asarray(x, dtype=F64)
    p = next_positive_power_of_two(x.shape[-1])
    return np.fft.fft(x, n=p, axis=-1)[..., : p // 2].real


def linear_rectifier(x: np.ndarray, max_val: float = 0.0, alpha: float = 0.0) -> np.ndarray:
    """``max(maxVal, x - alpha)`` elementwise (LinearRectifier.scala:14-16)."""
    return np.maximum(max_val, np.asarray(x, dtype=F64) - alpha)


def mnist_random_fft_features(x: np.ndarray, signs_list: Sequence[np.ndarray]) -> np.ndarray:
    """gather(RandomSignNode andThen PaddedFFT andThen LinearRectifier(0.0)) andThen VectorCombiner
    (MnistRandomFFT.scala:40-44)."""
    return np.concatenate([linear_rectifier(padded_fft(random_sign_node(x, s)), 0.0) for s in signs_list], axis=-1)


# --------------------------------------------------------------------------------------
# label helpers (K/nodes/util/ClassLabelIndicators.scala:15-29, MaxClassifier.scala:9-11)
# --------------------------------------------------------------------------------------
def class_label_indicators(labels: np.ndarray, num_classes: int) -> np.ndarray:
    """+1 at the class index, -1 elsewhere."""
    y = -np.ones((len(labels), num_classes), dtype=F64)
    y[np.arange(len(labels)), np.asarray(labels, dtype=np.int64)] = 1.0
    return y


def max_classifier(scores: np.ndarray) -> np.ndarray:
    return np.argmax(np.asarray(scores), axis=-1)


# --------------------------------------------------------------------------------------
# a5/a6  BlockLeastSquaresEstimator.fit   K/nodes/learning/BlockLinearMapper.scala:212-243
#        (+ mlmatrix 0.2 BlockCoordinateDescent / NormalEquations, restated; parity unpinned)
# --------------------------------------------------------------------------------------
def _solve_spd(G: np.ndarray, C: np.ndarray) -> np.ndarray:
    """Breeze ``\\`` on a square system is LAPACK dgesv (LU); numpy.linalg.solve is the same routine."""
    return np.linalg.solve(G, C)


def block_ls_fit(features: np.ndarray, labels: np.ndarray, block_size: int, num_iter: int,
                 lam: float = 0.0, num_features: Optional[int] = None,
                 feature_blocks: Optional[Sequence[np.ndarray]] = None):
    """Ridge regression by block coordinate descent on mean-centred features and labels.

    BlockLinearMapper.scala:215-219  labels centred by their column mean (the intercept)
    BlockLinearMapper.scala:224-232  every feature block centred by its own column mean
    BlockLinearMapper.scala:234-240  numIter == 1 -> one Gauss-Seidel sweep from x = 0
                                     (solveOnePassL2); numIter > 1 -> cyclic sweeps
                                     (solveLeastSquaresWithL2); lambda passed un-scaled
    BlockLinearMapper.scala:242      returns (xs, blockSize, Some(labelMean), Some(featureScalers))

    Returns ``(xs, intercept, feature_means)`` with xs[j] of shape (b_j, k).
    """
    labels = np.asarray(labels, dtype=F64)
    if feature_blocks is None:
        feature_blocks = vector_splitter(features, block_size, num_features)
    y_mean = labels.mean(axis=0)
    resid = labels - y_mean                          # b  (RowPartitionedMatrix of centred labels)
    means = [blk.mean(axis=0) for blk in feature_blocks]
    k = labels.shape[1]
    xs = [np.zeros((blk.shape[1], k), dtype=F64) for blk in feature_blocks]
    grams: List[Optional[np.ndarray]] = [None] * len(feature_blocks)
    for it in range(max(1, num_iter)):
        for j, blk in enumerate(feature_blocks):     # sequential block order (see module docstring)
            A = blk - means[j]
            if grams[j] is None:
                grams[j] = A.T @ A                   # cached across sweeps
            G = grams[j]
            # A_j^T (b - sum_{i != j} A_i x_i) = A_j^T resid + G x_j
            rhs = A.T @ resid + G @ xs[j]
            x_new = _solve_spd(G + lam * np.eye(G.shape[0]), rhs)
            resid -= A @ (x_new - xs[j])
            xs[j] = x_new
    return xs, y_mean, means


# --------------------------------------------------------------------------------------
# a8  BlockLinearMapper.apply   K/nodes/learning/BlockLinearMapper.scala:40-87
# a9  LinearMapper.apply        K/nodes/learning/LinearMapper.scala:30-62
# --------------------------------------------------------------------------------------
def block_linear_apply(features: np.ndarray, xs: Sequence[np.ndarray], block_size: int,
                       intercept: Optional[np.ndarray] = None,
                       feature_means: Optional[Sequence[np.ndarray]] = None,
                       return_partials: bool = False):
    """``sum_j (x_j - mu_j) W_j + b``.  With ``return_partials`` also returns the cumulative
    sum (+ intercept) after every block, i.e. what ``applyAndEvaluate`` hands its callback
    (BlockLinearMapper.scala:117-135)."""
    features = np.asarray(features, dtype=F64)
    d = sum(x.shape[0] for x in xs)
    blocks = vector_splitter(features, block_size, d)
    out = None
    partials = []
    for j, (blk, x) in enumerate(zip(blocks, xs)):
        if feature_means is not None:
            blk = blk - feature_means[j]
        part = blk @ x
        out = part if out is None else out + part
        if return_partials:
            partials.append(out + intercept if intercept is not None else out.copy())
    if intercept is not None:
        out = out + intercept
    return (out, partials) if return_partials else out


def linear_mapper_apply(x_in: np.ndarray, x: np.ndarray, b: Optional[np.ndarray] = None,
                        mean: Optional[np.ndarray] = None) -> np.ndarray:
    """``x^T (in - mean) + b`` (LinearMapper.scala:30-37, 44-62)."""
    x_in = np.asarray(x_in, dtype=F64)
    if mean is not None:
        x_in = x_in - mean
    out = x_in @ x
    return out + b if b is not None else out


def linear_map_fit(features: np.ndarray, labels: np.ndarray, lam: Optional[float] = None):
    """LinearMapEstimator.fit (LinearMapper.scala:80-98): exact normal equations on centred data,
    intercept = label mean, returns the feature scaler mean too."""
    features = np.asarray(features, dtype=F64)
    labels = np.asarray(labels, dtype=F64)
    mu = features.mean(axis=0)
    ymu = labels.mean(axis=0)
    A = features - mu
    G = A.T @ A
    if lam is not None:
        G = G + lam * np.eye(G.shape[0])
    x = _solve_spd(G, A.T @ (labels - ymu))
    return x, ymu, mu


# --------------------------------------------------------------------------------------
# a10  computeCost              K/nodes/learning/BlockLinearMapper.scala:142-187
# --------------------------------------------------------------------------------------
def compute_cost(features: np.ndarray, labels: np.ndarray, lam: float, xs: Sequence[np.ndarray],
                 block_size: int, intercept: Optional[np.ndarray] = None) -> float:
    """``||A W + b - Y||_F^2 / (2N) + lambda/2 ||W||_F^2`` -- note: no feature centring and
    lambda un-scaled (BlockLinearMapper.scala:149-185)."""
    labels = np.asarray(labels, dtype=F64)
    axb = block_linear_apply(features, xs, block_size, intercept, None)
    cost = float(((axb - labels) ** 2).sum())
    n = labels.shape[0]
    if lam == 0:
        return cost / (2.0 * n)
    wnorm = float(sum((x ** 2).sum() for x in xs))
    return cost / (2.0 * n) + lam / 2.0 * wnorm


# --------------------------------------------------------------------------------------
# a7  BlockWeightedLeastSquaresEstimator.trainWithL2
#     K/nodes/learning/BlockWeightedLeastSquares.scala:102-321  (all arithmetic in-repo)
# --------------------------------------------------------------------------------------
def group_by_classes(labels: np.ndarray) -> List[np.ndarray]:
    """Row-index partitions, one class per partition, class c in partition c
    (groupByClasses, BlockWeightedLeastSquares.scala:333-370: HashPartitioner(nClasses) on the
    argmax class index, original order kept inside a partition).  Classes with no rows yield
    empty partitions."""
    labels = np.asarray(labels, dtype=F64)
    cls = np.argmax(labels, axis=1)
    return [np.nonzero(cls == c)[0] for c in range(labels.shape[1])]


def _needs_reshuffle(labels: np.ndarray, partitions: Sequence[np.ndarray]) -> bool:
    """BlockWeightedLeastSquares.scala:111-124."""
    class_of = []
    for idx in partitions:
        if len(idx) == 0:
            # mapPartitions on an empty iterator: distinct.length == 0 != 1 -> "not same class"
            return True
        cls = np.unique(np.argmax(labels[idx], axis=1))
        if len(cls) != 1:
            return True
        class_of.append(int(cls[0]))
    return len(set(class_of)) != len(class_of)


def bwls_fit(features: np.ndarray, labels: np.ndarray, block_size: int, num_iter: int, lam: float,
             mixture_weight: float, num_features: Optional[int] = None,
             partitions: Optional[Sequence[np.ndarray]] = None):
    """Line-by-line restatement of ``trainWithL2``.  ``partitions`` (list of row-index arrays)
    plays the role of the RDD partitioning; by default -- and whenever the one-class-per-
    partition precondition fails (:111-131) -- rows are regrouped by class.

    Returns ``(xs, final_b)``; the mapper has no feature scalers (:316-320).
    """
    features = np.asarray(features, dtype=F64)
    labels = np.asarray(labels, dtype=F64)
    w = float(mixture_weight)
    if partitions is None or _needs_reshuffle(labels, partitions):
        partitions = group_by_classes(labels)
    parts = [np.asarray(p) for p in partitions if len(p) > 0]   # rowsToMatrixIter skips empties
    class_idxs = [int(np.argmax(labels[p[0]])) for p in parts]  # :133-139
    n_train = int(sum(len(p) for p in parts))                   # labels.count  :141
    n_classes = labels.shape[1]                                  # :142
    d = features.shape[1] if num_features is None else int(num_features)
    bounds = block_bounds(d, block_size)
    nb = len(bounds)

    joint_label_mean = np.zeros(n_classes, dtype=F64)            # :148-156
    for p, c in zip(parts, class_idxs):
        joint_label_mean[c] = 2 * w + (2 * (1.0 - w) * len(p) / float(n_train)) - 1

    models = [np.zeros((e - s, n_classes), dtype=F64) for s, e in bounds]
    residual = [labels[p] - joint_label_mean for p in parts]    # :167-169
    residual_mean = np.concatenate(residual, axis=0).mean(axis=0)  # :171
    stats: List[Optional[dict]] = [None] * nb

    for it in range(num_iter):
        for blk in range(nb):                                    # sequential order :180
            s, e = bounds[blk]
            feats = [features[p, s:e] for p in parts]            # blockFeaturesMat per partition
            if it == 0:
                pop_mean = np.concatenate(feats, axis=0).mean(axis=0)            # :197
                joint_means_parts = [f.mean(axis=0) * w + pop_mean * (1.0 - w) for f in feats]  # :201-204
                joint_means = np.zeros((n_classes, e - s), dtype=F64)           # :206-210
                for jm, c in zip(joint_means_parts, class_idxs):
                    joint_means[c, :] = jm
                ata = sum(f.T @ f for f in feats)                                # :212-214
                atr = sum(f.T @ r for f, r in zip(feats, residual))
                pop_cov = ata / float(n_train) - np.outer(pop_mean, pop_mean)    # :216
                pop_xtr = atr / float(n_train)                                   # :221
                stats[blk] = dict(pop_cov=pop_cov, pop_mean=pop_mean, joint_mean=joint_means,
                                  joint_means_parts=joint_means_parts)
            else:
                atr = sum(f.T @ r for f, r in zip(feats, residual))              # :223-225
                st = stats[blk]
                pop_cov, pop_mean = st["pop_cov"], st["pop_mean"]
                joint_means_parts = st["joint_means_parts"]
                pop_xtr = atr / float(n_train)

            delta = np.zeros_like(models[blk])
            for f, r, jm, c in zip(feats, residual, joint_means_parts, class_idxs):   # :241-276
                res_local = r[:, c]
                n_pos = f.shape[0]
                class_mean = f.mean(axis=0)
                zm = f - class_mean
                class_cov = (zm.T @ zm) / float(n_pos)
                class_xtr = (f.T @ res_local) / float(n_pos)
                mean_diff = class_mean - pop_mean
                joint_xtx = (pop_cov * (1.0 - w) + class_cov * w
                             + np.outer(mean_diff, mean_diff) * (1.0 - w) * w)
                mean_mixture_wt = residual_mean[c] * (1.0 - w) + w * res_local.mean()
                joint_xtr = pop_xtr[:, c] * (1.0 - w) + class_xtr * w - jm * mean_mixture_wt
                nd = joint_xtx.shape[1]
                W = _solve_spd(joint_xtx + np.eye(nd) * lam, joint_xtr - models[blk][:, c] * lam)
                delta[:, c] = W
            models[blk] = models[blk] + delta                                    # :278-284
            residual = [r - f @ delta for f, r in zip(feats, residual)]          # :287-294
            residual_mean = np.concatenate(residual, axis=0).mean(axis=0)        # :296

    full = np.concatenate(models, axis=0)                                        # :314
    jm_comb = np.concatenate([st["joint_mean"] for st in stats], axis=1)         # :315
    final_b = joint_label_mean - (jm_comb.T * full).sum(axis=0)                  # :316
    return models, final_b


# --------------------------------------------------------------------------------------
# Cross-check solver: PerClassWeightedLeastSquaresEstimator
#   K/nodes/learning/PerClassWeightedLeastSquares.scala:65-222
#   K/nodes/learning/internal/ReWeightedLeastSquares.scala:36-141
# --------------------------------------------------------------------------------------
def _reweighted_ls(blocks: Sequence[np.ndarray], labels_zm: np.ndarray, weights: np.ndarray,
                   feature_mean: np.ndarray, bounds, num_iter: int, lam: float) -> List[np.ndarray]:
    """ReWeightedLeastSquaresSolver.trainWithL2 for one class (numClasses = 1)."""
    n = labels_zm.shape[0]
    residual = np.zeros((n, 1), dtype=F64)
    model = [np.zeros((e - s, 1), dtype=F64) for s, e in bounds]
    ata_cache: List[Optional[np.ndarray]] = [None] * len(bounds)
    wcol = weights.reshape(-1, 1)
    for it in range(num_iter):
        for blk, (s, e) in enumerate(bounds):
            a_zm = blocks[blk] - feature_mean[s:e]
            if it == 0:
                ata_cache[blk] = a_zm.T @ (a_zm * wcol)                          # :92-99
            xw_old = a_zm @ model[blk]
            res_updated = residual - xw_old * wcol                                # :114-116
            atb = a_zm.T @ (labels_zm * wcol - res_updated)                       # :118-119
            new_model = _solve_spd(ata_cache[blk] + np.eye(e - s) * lam, atb)    # :123
            residual = residual + (a_zm @ (new_model - model[blk])) * wcol       # :129-134
            model[blk] = new_model
    return model


def per_class_wls_fit(features: np.ndarray, labels: np.ndarray, block_size: int, num_iter: int,
                      lam: float, mixture_weight: float, num_features: Optional[int] = None):
    """PerClassWeightedLeastSquaresEstimator.trainWithL2 (:65-124)."""
    features = np.asarray(features, dtype=F64)
    labels = np.asarray(labels, dtype=F64)
    w = float(mixture_weight)
    n_classes = labels.shape[1]
    d = features.shape[1] if num_features is None else int(num_features)
    n = labels.shape[0]
    cls = np.argmax(labels, axis=1)
    pop_mean = features[:, :d].sum(axis=0) / float(n)                              # :78-81
    counts = np.array([(cls == c).sum() for c in range(n_classes)], dtype=np.int64)   # :136-142
    jfm = np.zeros((n_classes, d), dtype=F64)
    present = [c for c in range(n_classes) if counts[c] > 0]
    for c in present:                                                             # :150-165
        cm = features[cls == c, :d].sum(axis=0) / float(counts[c])
        jfm[c] = cm * w + pop_mean * (1.0 - w)
    neg_wt = (1.0 - w) / float(n)                                                 # :176-183
    weights = np.full((n, n_classes), neg_wt, dtype=F64)
    weights[np.arange(n), cls] += w / counts[cls].astype(F64)
    joint_label_mean = (counts / float(n)) * (2.0 * (1.0 - w)) - 1.0 + 2.0 * w   # :188-194
    labels_zm = labels - joint_label_mean
    bounds = block_bounds(d, block_size)
    blocks = [features[:, s:e] for s, e in bounds]
    # jfmMat = rows of jointFeatureMean sorted by key: only classes that occur (:85)
    xs = [np.zeros((e - s, n_classes), dtype=F64) for s, e in bounds]
    for c in present:   # classes without rows keep a zero model column
        cm = jfm[c]
        model = _reweighted_ls(blocks, labels_zm[:, [c]], weights[:, c], cm, bounds, num_iter, lam)
        for blk in range(len(bounds)):
            xs[blk][:, c] = model[blk][:, 0]
    full = np.concatenate(xs, axis=0)
    final_b = joint_label_mean - (jfm.T * full).sum(axis=0)                       # :119
    return xs, final_b


# --------------------------------------------------------------------------------------
# Gradient oracle of the weighted objective
#   T/nodes/learning/BlockWeightedLeastSquaresSuite.scala:19-61
# --------------------------------------------------------------------------------------
def compute_gradient(features: np.ndarray, labels: np.ndarray, lam: float, mixture_weight: float,
                     x: np.ndarray, b: np.ndarray, partitions: Optional[Sequence[np.ndarray]] = None) -> np.ndarray:
    """Per partition (= class): weights negWt everywhere, posWt in the class column, where the
    class is read from the partition's first label row (:32-41); grad = sum A^T((A x + b - Y) .* wts)
    + lambda x (:47-60)."""
    features = np.asarray(features, dtype=F64)
    labels = np.asarray(labels, dtype=F64)
    if partitions is None:
        partitions = group_by_classes(labels)
    parts = [np.asarray(p) for p in partitions if len(p) > 0]
    n_train = int(sum(len(p) for p in parts))
    grad = np.zeros_like(x)
    for p in parts:
        lab = labels[p]
        feats = features[p]
        c = int(np.argmax(lab[0]))
        neg = (1.0 - mixture_weight) / float(n_train)
        pos = neg + mixture_weight / float(len(p))
        wts = np.full(lab.shape, neg, dtype=F64)
        wts[:, c] = pos
        out = feats @ x + b - lab
        grad += feats.T @ (out * wts)
    return grad + x * lam


# --------------------------------------------------------------------------------------
# Multi-partition restatement used by the sharding tests: the same BlockLS fit computed from
# per-shard partial sums only (what the GPU path all-reduces).  Algebra of DESIGN.md section 4.
# --------------------------------------------------------------------------------------
def block_ls_partial_sums(A_shift: np.ndarray, R: np.ndarray):
    """Per-shard contribution for one block: ``A^T A``, ``A^T R``, column sums of A and of R,
    where ``A_shift = A - shift`` for a shift vector shared by all shards."""
    return A_shift.T @ A_shift, A_shift.T @ R, A_shift.sum(axis=0), R.sum(axis=0)


def block_ls_solve_from_sums(G, C, sa, sr, n_total: int, lam: float, w_old: Optional[np.ndarray] = None):
    """Exact centring correction from reduced sums: with delta = sa / N (mean of the shifted
    block) and rbar = sr / N,  G_c = G - N delta delta^T,  C_c = C - N delta rbar^T, then the
    BCD step  dW = (G_c + lam I)^-1 (C_c - lam W_old)."""
    delta = sa / float(n_total)
    rbar = sr / float(n_total)
    Gc = G - n_total * np.outer(delta, delta)
    Cc = C - n_total * np.outer(delta, rbar)
    if w_old is not None:
        Cc = Cc - lam * w_old
    dW = _solve_spd(Gc + lam * np.eye(G.shape[0]), Cc)
    return dW, delta


# --------------------------------------------------------------------------------------
# Evaluation ("next" row of SURVEY section 8(f): the consumer of BlockLinearMapper -> MaxClassifier output).
# Restated for the parity tests of the on-device confusion matrix that follows the solver; not on the measured path yet.
# --------------------------------------------------------------------------------------
def confusion_matrix(predictions: np.ndarray, actuals: np.ndarray, num_classes: int) -> np.ndarray:
    """K/evaluation/MulticlassClassifierEvaluator.scala:149-160: rows = true class, columns = predicted class, counts."""
    cm = np.zeros((num_classes, num_classes))
    np.add.at(cm, (np.asarray(actuals, dtype=np.int64), np.asarray(predictions, dtype=np.int64)), 1.0)
    return cm


def multiclass_metrics(cm: np.ndarray, beta: float = 1.0) -> dict:
    """MulticlassMetrics (K/evaluation/MulticlassClassifierEvaluator.scala:23-54) over BinaryClassificationMetrics
    (K/evaluation/BinaryClassifierEvaluator.scala:16-41): per-class contingency tables from the confusion matrix, macro =
    mean over classes, micro = the metric of the merged (summed) table."""
    cm = np.asarray(cm, dtype=np.float64)
    total = cm.sum()
    tp = np.diag(cm).copy()
    fp = cm.sum(axis=0) - tp                 # predictedSums - tp
    tn = total - cm.sum(axis=1) - fp         # total - actualsSums - fp
    fn = total - tp - fp - tn

    def fscore(tp_, fp_, fn_):
        return (1.0 + beta * beta) * tp_ / ((1.0 + beta * beta) * tp_ + beta * beta * fn_ + fp_)

    with np.errstate(invalid="ignore", divide="ignore"):
        precision, recall = tp / (tp + fp), tp / (tp + fn)
        accuracy = (tp + tn) / (tp + fp + tn + fn)
        f = fscore(tp, fp, fn)
        TP, FP, FN = tp.sum(), fp.sum(), fn.sum()
        out = {"class_precision": precision, "class_recall": recall, "class_fscore": f,
               "avg_accuracy": accuracy.mean(), "avg_error": 1.0 - accuracy.mean(),
               "macro_precision": precision.mean(), "macro_recall": recall.mean(), "macro_fscore": f.mean(),
               "total_accuracy": TP / (TP + FP), "total_error": FN / (FN + TP),
               "micro_precision": TP / (TP + FP), "micro_recall": TP / (TP + FN), "micro_fscore": fscore(TP, FP, FN)}
    return out


# --------------------------------------------------------------------------------------
# CIFAR random-patch featurizer.  Images are arrays img[x, y, c] with the reference's
# coordinates: x runs over image ROWS (xDim = height), y over columns (yDim = width)
# (K/utils/images/ImageConversions.scala:10-24, K/utils/images/Image.scala:140-143).
#   Convolver          K/nodes/images/Convolver.scala:20-203
#   Stats.normalizeRows K/utils/Stats.scala:112-123
#   SymmetricRectifier K/nodes/images/SymmetricRectifier.scala:7-32
#   Pooler             K/nodes/images/Pooler.scala:21-69
#   ImageVectorizer    K/nodes/images/ImageVectorizer.scala:12-16 (Image.toArray, Image.scala:47-65)
# --------------------------------------------------------------------------------------
def image_from_bgr_bytes(rgb_hwc: np.ndarray) -> np.ndarray:
    """What ImageUtils.loadImage yields for an 8-bit RGB file: Java decodes to TYPE_3BYTE_BGR, ByteArrayVectorizedImage indexes
    it as get(x = row, y = column, c) with c = 0 blue, 1 green, 2 red (ImageConversions.scala:10-24, Image.scala:152-176)."""
    return np.asarray(rgb_hwc)[:, :, ::-1].astype(F64)


def flip_image(img: np.ndarray) -> np.ndarray:
    """ImageUtils.flipImage (K/utils/images/ImageUtils.scala:376-389): all three axes reversed."""
    return np.asarray(img, dtype=F64)[::-1, ::-1, ::-1].copy()


def pack_filters(filters: Sequence[np.ndarray]) -> np.ndarray:
    """Convolver.packFilters (:104-131): row i, column c + x*C + y*C*xDim = filters[i][x, y, c]."""
    out = []
    for f in filters:
        f = np.asarray(f, dtype=F64)
        out.append(np.transpose(f, (1, 0, 2)).reshape(-1))       # y slowest, then x, then c
    return np.stack(out, axis=0)


def normalize_rows(mat: np.ndarray, alpha: float = 1.0) -> np.ndarray:
    """Stats.normalizeRows (K/utils/Stats.scala:112-123): subtract the row mean, divide by sqrt(sample variance (n-1) + alpha)."""
    mat = np.asarray(mat, dtype=F64)
    mean = np.nan_to_num(mat.mean(axis=1))
    var = ((mat - mean[:, None]) ** 2).sum(axis=1) / (mat.shape[1] - 1.0)
    sd = np.sqrt(var + alpha)
    sd = np.where(np.isnan(sd), math.sqrt(alpha), sd)
    return (mat - mean[:, None]) / sd[:, None]


def make_patches(img: np.ndarray, conv_size: int, normalize: bool = True, whitener_means: Optional[np.ndarray] = None,
                 var_constant: float = 10.0) -> np.ndarray:
    """Convolver.makePatches (:152-203): patch row py = x + y*resWidth, column px = c + pox*C + poy*C*convSize holds
    img[x + pox, y + poy, c]; rows optionally normalised (Stats.normalizeRows with varConstant) and shifted by the whitener's means."""
    img = np.asarray(img, dtype=F64)
    xd, yd, ch = img.shape
    rw, rh = xd - conv_size + 1, yd - conv_size + 1
    pm = np.empty((rw * rh, conv_size * conv_size * ch), dtype=F64)
    for poy in range(conv_size):
        for pox in range(conv_size):
            win = img[pox:pox + rw, poy:poy + rh, :]                   # [x, y, c]
            col0 = pox * ch + poy * ch * conv_size
            pm[:, col0:col0 + ch] = np.transpose(win, (1, 0, 2)).reshape(rw * rh, ch)   # row index x + y*rw
    if normalize:
        pm = normalize_rows(pm, var_constant)
    if whitener_means is not None:
        pm = pm - np.asarray(whitener_means, dtype=F64)
    return pm


def convolve(img: np.ndarray, filters: np.ndarray, conv_size: int, normalize: bool = True,
             whitener_means: Optional[np.ndarray] = None, var_constant: float = 10.0) -> np.ndarray:
    """Convolver.convolve (:128-149): patches * filters^T, returned as an image out[x, y, f] (RowMajorArrayVectorizedImage of the
    column-major product: value (x, y, f) at x + y*resWidth + f*resWidth*resHeight)."""
    img = np.asarray(img, dtype=F64)
    rw, rh = img.shape[0] - conv_size + 1, img.shape[1] - conv_size + 1
    res = make_patches(img, conv_size, normalize, whitener_means, var_constant) @ np.asarray(filters, dtype=F64).T
    return np.transpose(res.reshape(rh, rw, -1), (1, 0, 2))           # row index x + y*rw -> [y][x] -> [x, y, f]


def symmetric_rectifier(img: np.ndarray, max_val: float = 0.0, alpha: float = 0.0) -> np.ndarray:
    """SymmetricRectifier (:7-32): channels [0, C) = max(maxVal, v - alpha), channels [C, 2C) = max(maxVal, -v - alpha)."""
    img = np.asarray(img, dtype=F64)
    return np.concatenate([np.maximum(max_val, img - alpha), np.maximum(max_val, -img - alpha)], axis=2)


def pooler(img: np.ndarray, stride: int, pool_size: int, pixel_fn=None, pool_fn=np.sum) -> np.ndarray:
    """Pooler (:21-69): pools centred at strideStart = poolSize / 2, strideStart + stride, ...; a pool covers
    [x - poolSize/2, min(x + poolSize/2, xDim)) in each direction (integer division); out[px, py, c] = poolFn(pixelFn(values))."""
    img = np.asarray(img, dtype=F64)
    xd, yd, ch = img.shape
    s0 = pool_size // 2
    npx, npy = int(math.ceil((xd - s0) / float(stride))), int(math.ceil((yd - s0) / float(stride)))
    out = np.zeros((npx, npy, ch), dtype=F64)
    for x in range(s0, xd, stride):
        for y in range(s0, yd, stride):
            reg = img[x - pool_size // 2:min(x + pool_size // 2, xd), y - pool_size // 2:min(y + pool_size // 2, yd), :]
            if pixel_fn is not None:
                reg = pixel_fn(reg)
            out[(x - s0) // stride, (y - s0) // stride, :] = pool_fn(reg.reshape(-1, ch), axis=0)
    return out


def image_vectorizer(img: np.ndarray) -> np.ndarray:
    """Image.toArray (Image.scala:47-65): flat[c + x*C + y*C*xDim] = img[x, y, c]."""
    return np.transpose(np.asarray(img, dtype=F64), (1, 0, 2)).reshape(-1)


def random_patch_cifar_features(img: np.ndarray, filters: np.ndarray, whitener_means: Optional[np.ndarray], conv_size: int = 6,
                                alpha: float = 0.25, pool_stride: int = 13, pool_size: int = 14) -> np.ndarray:
    """The featurizer of K/pipelines/images/cifar/RandomPatchCifar.scala:59-63: Convolver(filters, whitener, normalizePatches = true)
    andThen SymmetricRectifier(alpha) andThen Pooler(stride, size, identity, sum) andThen ImageVectorizer."""
    conv = convolve(img, filters, conv_size, True, whitener_means, 10.0)
    return image_vectorizer(pooler(symmetric_rectifier(conv, 0.0, alpha), pool_stride, pool_size))
