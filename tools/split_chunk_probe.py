"""Parity mode: rows per accumulation chain (context option split_chunk_rows) against accuracy and time.

    python tools/split_chunk_probe.py [chunk ...]          default 2048 4096 8192

Accuracy: the C3 shape with N = 32768 rows and the first 4 feature maps (D = 16384, k = 1000, lambda = 1) against the fp64 oracle
(one oracle fit, one device fit per chunk size).  Time: config-3 fits at N = 1M, 16 maps (1 warm-up + 2 timed per chunk size)."""
import json
import os
import sys
import time

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = str(os.cpu_count() or 1)
import numpy as np

sys.path.insert(0, ".")
import keystone_b200 as ks
from oracle import keystone_oracle as ko

chunks = [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192]
ctx = ks.Context(0)

# ---- accuracy
n, d_in, n_out, nrf, k, lam = 32768, 440, 4096, 4, 1000, 1.0
rng = np.random.default_rng(2)
params = [(rng.standard_normal((n_out, d_in)) * 0.0555, rng.random(n_out) * 2 * np.pi) for _ in range(nrf)]
wstar = rng.standard_normal((16, k)).astype(np.float32)
X = rng.standard_normal((n, d_in)).astype(np.float32)
cls = np.argmax(X[:, :16] @ wstar + 0.1 * rng.standard_normal((n, k)).astype(np.float32), axis=1)
x = ctx.matrix(X)
y = ctx.labels_from_classes(cls, k)
feats = ks.Pipeline.gather([ks.CosineRandomFeatures(ctx, W, b) for W, b in params]).andThen(ks.VectorCombiner())(x)
blocks = [ko.cosine_random_features(X.astype(np.float64), W, b) for W, b in params]
xs, b0, mus = ko.block_ls_fit(None, ko.class_label_indicators(cls, k), n_out, 1, lam, feature_blocks=blocks)
Wr = np.concatenate(xs, 0)
acc = {}
for ch in chunks:
    ctx.set_option("split_chunk_rows", ch)
    m = ks.BlockLeastSquaresEstimator(n_out, 1, lam, precision="f16x2").fit(feats, y)
    acc[ch] = float(np.linalg.norm(np.concatenate(m.xs, 0) - Wr) / np.linalg.norm(Wr))
del x, y, feats, m, blocks

# ---- time
n = 1_000_000
prm = np.random.default_rng(5)
x = ctx.synthetic_normal(n, d_in, seed=3)
y = ctx.labels_from_classes(prm.integers(0, k, n).astype(np.int32), k)
feats = ks.Pipeline.gather([ks.CosineRandomFeatures.create(ctx, d_in, n_out, 0.0555, prm) for _ in range(16)]).andThen(ks.VectorCombiner())(x)
est = ks.BlockLeastSquaresEstimator(n_out, 1, lam, precision="f16x2")
for ch in chunks:
    ctx.set_option("split_chunk_rows", ch)
    ms = []
    for rep in range(3):
        ctx.synchronize()
        t0 = time.perf_counter()
        m = est.fit(feats, y)
        _ = m.xs[-1][0, 0]
        ms.append(1e3 * (time.perf_counter() - t0))
    st = ctx.last_fit_stats()
    print(json.dumps({"probe": "split_chunk", "chunk_rows": ch, "rel_fro_W_n32768_d16384": acc[ch], "fit_ms": [round(v, 1) for v in ms[1:]],
                      "gram_ms": st["gram_ms"], "update_ms": st["update_ms"]}), flush=True)
ctx.close()
