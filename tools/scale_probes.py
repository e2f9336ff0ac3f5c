"""Scale sanity runs for the two other solver configurations of BASELINE.json (run on the GPU box):
  C2-like : materialised features N = 1M x D = 16384 (generated on the device), k = 100, blockSize 4096, BlockLS
  C5-mini : BlockWeightedLeastSquares, 147 classes, d_in 440 -> 2 x 4096 cosine features, N = 200k class-sorted rows
Prints one JSON line each (timings from the library's phase timers; no CPU oracle at these sizes)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import keystone_b200 as ks


def c2(ctx, n=1_000_000, d=16384, k=100, bs=4096, lam=10.0):
    f = ctx.synthetic_normal(n, d, seed=1, mean=0.1)
    cls = np.random.default_rng(1).integers(0, k, n).astype(np.int32)
    y = ctx.labels_from_classes(cls, k)
    est = ks.BlockLeastSquaresEstimator(bs, 1, lam)
    est.fit(f, y)
    t0 = time.perf_counter()
    m = est.fit(f, y)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    st = ctx.last_fit_stats()
    W = m.xs[0]
    alg = 2.0 * n * d * (bs + 2 * k)
    print(json.dumps({"probe": "C2 materialised", "n": n, "d": d, "k": k, "wall_s": dt, "samples_per_s": n / dt,
                      "alg_tflops": alg / dt / 1e12, "finite": bool(np.isfinite(W).all()), "stats": st}), flush=True)


def c5_mini(ctx, n=200_000, d_in=440, n_out=4096, nrf=2, k=147, lam=6e-5, w=0.25):
    rng = np.random.default_rng(4)
    sizes = rng.multinomial(n - 1000 * k, np.ones(k) / k) + 1000
    cls = np.repeat(np.arange(k), sizes).astype(np.int32)          # class-sorted rows
    x = ctx.synthetic_normal(n, d_in, seed=4)
    y = ctx.labels_from_classes(cls, k)
    rfs = [ks.CosineRandomFeatures.create(ctx, d_in, n_out, 0.0555, rng) for _ in range(nrf)]
    feats = ks.Pipeline.gather(rfs).andThen(ks.VectorCombiner())(x)
    est = ks.BlockWeightedLeastSquaresEstimator(n_out, 1, lam * n, w)
    t0 = time.perf_counter()
    m = est.fit(feats, y)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    W = np.concatenate(m.xs, 0)
    pred = m.apply_argmax(feats)
    print(json.dumps({"probe": "C5-mini BWLS", "n": n, "D": nrf * n_out, "k": k, "wall_s": dt, "finite": bool(np.isfinite(W).all()),
                      "train_acc": float((pred == cls).mean()), "stats": ctx.last_fit_stats()}), flush=True)


if __name__ == "__main__":
    with ks.Context(0) as ctx:
        if "c5" in sys.argv or len(sys.argv) == 1:
            c5_mini(ctx)
        if "c2" in sys.argv or len(sys.argv) == 1:
            c2(ctx)
