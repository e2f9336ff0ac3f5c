"""One config-3 fit (fast mode by default) for profiler captures: python tools/one_fit.py [precision] [n_rows]."""
import sys

import numpy as np

sys.path.insert(0, ".")
import keystone_b200 as ks

prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
rng = np.random.default_rng(0)
with ks.Context(0) as ctx:
    x = ctx.synthetic_normal(n, 440, seed=3)
    y = ctx.labels_from_classes(rng.integers(0, 1000, n).astype(np.int32), 1000)
    rfs = [ks.CosineRandomFeatures.create(ctx, 440, 4096, 0.0555, rng) for _ in range(4)]
    feats = ks.Pipeline.gather(rfs).andThen(ks.VectorCombiner())(x)
    m = ks.BlockLeastSquaresEstimator(4096, 1, 1.0, precision=prec).fit(feats, y)
    print(ctx.last_fit_stats())
