"""Accuracy of the GPU BlockLS fit vs the fp64 oracle at a mid-size config-3 shape (runs on the GPU box; CPU part ~1 min)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import keystone_b200 as ks
from oracle import keystone_oracle as ko


def main(n=32768, d_in=440, n_out=4096, nrf=2, k=100, lam=1.0, iters=1, precisions=("tf32", "f16")):
    rng = np.random.default_rng(5)
    X = rng.standard_normal((n, d_in)).astype(np.float32)
    wstar = rng.standard_normal((16, k)).astype(np.float32)
    cls = np.argmax(X[:, :16] @ wstar + 0.1 * rng.standard_normal((n, k)).astype(np.float32), 1)
    params = [ko.cosine_random_features_params(d_in, n_out, 0.0555, rng) for _ in range(nrf)]
    with ks.Context(0) as ctx:
        x = ctx.matrix(X); y = ctx.labels_from_classes(cls, k)
        rfs = [ks.CosineRandomFeatures(ctx, W, b) for W, b in params]
        feats = ks.Pipeline.gather(rfs).andThen(ks.VectorCombiner())(x)
        got = {}
        for prec in precisions:
            model = ks.BlockLeastSquaresEstimator(n_out, iters, lam, precision=prec).fit(feats, y)
            got[prec] = (np.concatenate(model.xs, 0), np.concatenate(model.feature_means), model(feats).to_numpy(),
                         ctx.last_fit_stats())
    t0 = time.time()
    Xd = X.astype(np.float64)
    blocks = [ko.cosine_random_features(Xd, W, b) for W, b in params]
    Y = ko.class_label_indicators(cls, k)
    xs, b0, mus = ko.block_ls_fit(None, Y, n_out, iters, lam, feature_blocks=blocks)
    Wr = np.concatenate(xs, 0)
    ref = sum((blk - mu) @ w for blk, mu, w in zip(blocks, mus, xs)) + b0
    for prec, (Wg, mg, pred, stats) in got.items():
      print(json.dumps({"probe": "accuracy", "precision": prec, "n": n, "D": nrf * n_out, "k": k, "lambda": lam, "iters": iters,
                        "relfro_W": float(np.linalg.norm(Wg - Wr) / np.linalg.norm(Wr)),
                        "maxabs_W": float(np.abs(Wg - Wr).max()), "max_W": float(np.abs(Wr).max()),
                        "maxabs_mean": float(np.abs(mg - np.concatenate(mus)).max()),
                        "maxabs_pred": float(np.abs(pred - ref).max()), "rms_pred_err": float(np.sqrt(np.mean((pred - ref) ** 2))),
                        "argmax_agree": float((pred.argmax(1) == ref.argmax(1)).mean()), "train_acc_ref": float((ref.argmax(1) == cls).mean()),
                        "cpu_s": time.time() - t0, "gpu_ms": stats["total_ms"]}), flush=True)


if __name__ == "__main__":
    only = sys.argv[1:] or None
    main(precisions=only or ("tf32", "f16"))
    main(n=32768, k=100, lam=1e-3, iters=1, precisions=only or ("tf32", "f16"))
    main(n=16384, k=50, lam=1.0, iters=3, precisions=only or ("tf32", "f16"))
