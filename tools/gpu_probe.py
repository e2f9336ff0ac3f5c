"""First-contact probe for the GPU box: numerics of each kernel in isolation + first timings.  Prints JSON lines."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import keystone_b200 as ks
from keystone_b200._capi import check, lib
from oracle import keystone_oracle as ko


def gram(ctx, a, b, m, kc):
    G = np.zeros((m, m)); Cm = np.zeros((m, kc))
    check(ctx.handle, lib().ks_debug_gram(ctx.handle, a.handle, b.handle, G.ctypes.data_as(C.c_void_p), m, Cm.ctypes.data_as(C.c_void_p), kc))
    return G, Cm


def main():
    ctx = ks.Context(0)
    rng = np.random.default_rng(0)
    # 1. exactness on integers
    A = rng.integers(-3, 4, (777, 200)).astype(np.float64); B = rng.integers(-3, 4, (777, 70)).astype(np.float64)
    try:
        G, Cm = gram(ctx, ctx.matrix(A.astype(np.float32)), ctx.matrix(B.astype(np.float32)), 200, 70)
        print(json.dumps({"probe": "gram_int_exact", "G_maxerr": float(np.abs(G - A.T @ A).max()), "C_maxerr": float(np.abs(Cm - A.T @ B).max())}), flush=True)
    except Exception as e:
        print(json.dumps({"probe": "gram_int_exact", "error": str(e)}), flush=True)
        return
    # 2. gaussian accuracy
    for n, m, kc in [(4096, 512, 64), (65536, 1024, 100)]:
        A = rng.standard_normal((n, m)).astype(np.float32); B = rng.standard_normal((n, kc)).astype(np.float32)
        G, Cm = gram(ctx, ctx.matrix(A), ctx.matrix(B), m, kc)
        Gr = A.astype(np.float64).T @ A.astype(np.float64)
        print(json.dumps({"probe": "gram_gauss", "n": n, "m": m, "G_maxerr": float(np.abs(G - Gr).max()), "G_relfro": float(np.linalg.norm(G - Gr) / np.linalg.norm(Gr)),
                          "mean_signed_diag_relerr": float(np.mean((np.diag(G) - np.diag(Gr)) / np.diag(Gr)))}), flush=True)
    # 3. timing of the Gram kernel at the C3 block shape (b=4096, k=1000) for several chunk sizes
    n = 131072
    a = ctx.synthetic_normal(n, 4096, 1)
    b = ctx.synthetic_normal(n, 1001, 2)
    for chunk in (2048, 4096, 8192, 16384):
        ctx.set_option("gram_chunk_rows", chunk)
        ms = C.c_double(0)
        check(ctx.handle, lib().ks_debug_time_gram(ctx.handle, a.handle, b.handle, 3, C.byref(ms)))
        alg = 2.0 * n * 4096 * (4096 + 1001)
        print(json.dumps({"probe": "gram_time", "n": n, "chunk": chunk, "ms": ms.value, "alg_tflops": alg / ms.value / 1e9,
                          "real_tflops": 2.0 * n * 128 * 256 * (272 + 128) / ms.value / 1e9}), flush=True)
    ctx.set_option("gram_chunk_rows", 4096)
    del a, b
    # 4. end-to-end mini C3
    n, d_in, n_out, nrf, k = 131072, 440, 4096, 2, 1000
    X = rng.standard_normal((n, d_in)).astype(np.float32)
    cls = rng.integers(0, k, n)
    params = [ko.cosine_random_features_params(d_in, n_out, 0.0555, rng) for _ in range(nrf)]
    x = ctx.matrix(X); y = ctx.labels_from_classes(cls, k)
    rfs = [ks.CosineRandomFeatures(ctx, W, bb) for W, bb in params]
    feats = ks.Pipeline.gather(rfs).andThen(ks.VectorCombiner())(x)
    for rep in range(2):
        t0 = time.time()
        model = ks.BlockLeastSquaresEstimator(n_out, 1, 1.0).fit(feats, y)
        ctx.synchronize()
        print(json.dumps({"probe": "fit_c3_mini", "rep": rep, "wall_s": time.time() - t0, "stats": ctx.last_fit_stats()}), flush=True)
    # accuracy of that fit on a subsample-sized oracle is too slow here; covered by tests at small sizes
    ctx.close()


if __name__ == "__main__":
    main()
