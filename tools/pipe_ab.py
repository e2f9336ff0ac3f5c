"""A/B timing of the fit's stream arrangement (option "pipeline") and operand modes inside ONE process on the config-3
workload (GPU box).  Every configuration gets 1 warm-up + 3 timed fits; the list runs twice so drift shows up as a
difference between the passes.  With KS_TIMELINE set, the last fit of every configuration dumps its span timeline."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import keystone_b200 as ks


def main():
    tl_base = os.environ.pop("KS_TIMELINE", None)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    configs = [c.split(":") for c in sys.argv[2:]] or [["1", "f16"], ["0", "f16"], ["1", "f16x2"], ["1", "tf32"]]
    rng = np.random.default_rng(0)
    d_in, n_out, nrf, k = 440, 4096, 16, 1000
    with ks.Context(0) as ctx:
        x = ctx.synthetic_normal(n, d_in, seed=3)
        cls = rng.integers(0, k, n).astype(np.int32)
        y = ctx.labels_from_classes(cls, k)
        rfs = [ks.CosineRandomFeatures.create(ctx, d_in, n_out, 0.0555, rng) for _ in range(nrf)]
        feats = ks.Pipeline.gather(rfs).andThen(ks.VectorCombiner())(x)
        for rep in range(2):
            for cfg in configs:
                pipe, prec = int(cfg[0]), cfg[1]
                opts = {"custom_solve": 0, "dyn_tiles": 1, "gram_chunk_rows": 0, "reserve_sms": 8}     # defaults, so that configs do not leak
                opts.update({name: int(val) for name, val in (o.split("=") for o in cfg[2:])})
                ctx.set_option("pipeline", pipe)
                for name, val in opts.items():
                    ctx.set_option(name, val)
                est = ks.BlockLeastSquaresEstimator(n_out, 1, 1.0, precision=prec)
                est.fit(feats, y)
                ts = []
                for i in range(3):
                    if tl_base and i == 2:
                        os.environ["KS_TIMELINE"] = f"{tl_base}_p{pipe}_{prec}" + "".join("_" + o.replace("=", "") for o in cfg[2:])
                    t0 = time.perf_counter()
                    m = est.fit(feats, y)
                    _ = m.xs[-1][0, 0]          # the model is on the host when fit returns
                    ts.append(1e3 * (time.perf_counter() - t0))
                os.environ.pop("KS_TIMELINE", None)
                st = ctx.last_fit_stats()
                print(json.dumps({"probe": "pipe_ab", "pass": rep, "pipeline": pipe, "precision": prec, "opts": cfg[2:],
                                  "ms": [round(t, 1) for t in ts], "device_ms": round(st["total_ms"], 1),
                                  "featurize_ms": round(st["featurize_ms"], 1), "gram_ms": round(st["gram_ms"], 1),
                                  "update_ms": round(st["update_ms"], 1), "solve_ms": round(st["solve_ms"], 1),
                                  "other_ms": round(st["other_ms"], 1), "host_ms": round(st["host_ms"], 1)}), flush=True)


if __name__ == "__main__":
    main()
