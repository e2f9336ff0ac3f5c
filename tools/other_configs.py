"""The other solver configurations of BASELINE.json, timed on the GPU box with parity on a row subsample:

  C1  MNIST random-FFT shape: 60000 x 784 -> 4 x PaddedFFT (D = 2048), k = 10, b = 2048, lambda = 0, with the full-size CPU oracle
  C2  N = 1M x D = 16384 materialised features (generated on the device, column mean 0.1), k = 100, b = 4096, lambda = 10, 1 GPU
  C4  N = 50K x D = 160000 direct synthetic features max(0, N(0,1)) (39 blocks of 4096 + one of 256), k = 10, lambda = 3000
  C5  BlockWeightedLeastSquares, N = 2M (or --c5-rows), d_in 440 -> D = 132 x 4096 cosine features, k = 147 class-sorted,
      lambda = 6e-5, mixtureWeight 0.25; rows sharded BY CLASS over the ranks

    python tools/other_configs.py c2 c4 c5 [--c5-rows N] [--c5-blocks B] [--precision f16x2|tf32|f16]
    python -m torch.distributed.run --nproc-per-node 4 ... tools/other_configs.py c4 c5

One JSON line per configuration (rank 0).  The timed fit uses the requested precision; `parity` is measured with a second, small
fit of the same shape family (same D / k / b / lambda, N_sub rows) against the fp64 oracle on rank 0 -- single rank only."""
import argparse
import json
import os
import sys
import time

if int(os.environ.get("RANK", "0")) == 0:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

sys.path.insert(0, ".")
import keystone_b200 as ks

ap = argparse.ArgumentParser()
ap.add_argument("configs", nargs="*", default=["c2", "c4", "c5"])
ap.add_argument("--precision", default="f16x2")
ap.add_argument("--c5-rows", type=int, default=2_000_000)
ap.add_argument("--c5-blocks", type=int, default=132)
ap.add_argument("--c4-rows", type=int, default=50_000)
ap.add_argument("--c1-lambdas", type=float, nargs="*", default=[0.0, 1.0, 100.0])
ap.add_argument("--parity-rows", type=int, default=8192)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ks.Context.from_torch_distributed(local)
else:
    ctx = ks.Context(local)


def barrier():
    if world > 1:
        dist.barrier(device_ids=[local])
    ctx.synchronize()


def timed_fit(est, feats, y, reps=2):
    est.fit(feats, y)                       # warm-up (pool allocations, pinned mirror)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        m = est.fit(feats, y)
        _ = m.xs[-1][0, 0]
    barrier()
    return m, (time.perf_counter() - t0) / reps


def emit(rec):
    if rank == 0:
        print(json.dumps(rec), flush=True)


def c2():
    n, d, k, bs, lam = 1_000_000, 16384, 100, 4096, 10.0
    lo, hi = ks.shard_range(n, rank, world)
    f = ctx.synthetic_normal(hi - lo, d, seed=1, global_row_offset=lo, mean=0.1)
    cls = np.random.default_rng([1, lo]).integers(0, k, hi - lo).astype(np.int32)
    y = ctx.labels_from_classes(cls, k)
    m, dt = timed_fit(ks.BlockLeastSquaresEstimator(bs, 1, lam, precision=args.precision), f, y)
    st = ctx.last_fit_stats()
    rec = {"config": "C2 materialised N=1M D=16384 k=100 b=4096 lambda=10", "gpus": world, "precision": st["mma"], "fit_s": dt,
           "samples_per_s": n / dt, "alg_tflops": 1.409e14 / dt / 1e12, "finite": bool(np.isfinite(m.xs[0]).all()),
           "phase_ms": {q: st[q] for q in st if q.endswith("_ms")}}
    del f, y, m
    if world == 1 and args.parity_rows:
        from oracle import keystone_oracle as ko
        ns = args.parity_rows
        rng = np.random.default_rng(11)
        F = rng.standard_normal((ns, d), dtype=np.float32) + (0.1 * np.arange(d) / d).astype(np.float32)
        c = rng.integers(0, k, ns)
        mg = ks.BlockLeastSquaresEstimator(bs, 1, lam, precision=args.precision).fit(ctx.matrix(F), ctx.labels_from_classes(c, k))
        xs, _, _ = ko.block_ls_fit(F.astype(np.float64), ko.class_label_indicators(c, k), bs, 1, lam)
        Wg, Wr = np.concatenate(mg.xs, 0), np.concatenate(xs, 0)
        rec["parity"] = {"n_rows": ns, "rel_fro_W": float(np.linalg.norm(Wg - Wr) / np.linalg.norm(Wr))}
    emit(rec)


def c4():
    n, d, k, bs, lam = args.c4_rows, 160_000, 10, 4096, 3000.0
    lo, hi = ks.shard_range(n, rank, world)
    rng = np.random.default_rng([3, lo])
    # pooled rectified responses max(0, N(0,1)), generated on the device (32 GB in total at full size: n / world rows per rank)
    g = ctx.synthetic_normal(hi - lo, d, seed=3, global_row_offset=lo)
    f = ks.LinearRectifier(0.0, 0.0, ctx)(g)
    del g
    cls = rng.integers(0, k, hi - lo).astype(np.int32)
    y = ctx.labels_from_classes(cls, k)
    m, dt = timed_fit(ks.BlockLeastSquaresEstimator(bs, 1, lam, precision=args.precision), f, y)
    st = ctx.last_fit_stats()
    rec = {"config": f"C4 direct synthetic N={n} D=160000 (39 x 4096 + 256) k=10 lambda=3000", "gpus": world, "precision": st["mma"],
           "fit_s": dt, "samples_per_s": n / dt, "num_blocks": st["num_blocks"], "last_block_rows": int(m.xs[-1].shape[0]),
           "finite": bool(all(np.isfinite(w).all() for w in m.xs)), "phase_ms": {q: st[q] for q in st if q.endswith("_ms")}}
    if world == 1 and args.parity_rows:
        from oracle import keystone_oracle as ko
        ns, dsub = min(args.parity_rows, 4096), 3 * 4096 + 256                 # same ragged last block, 4 blocks
        F = np.maximum(np.random.default_rng(12).standard_normal((ns, dsub), dtype=np.float32), 0)
        c = np.random.default_rng(13).integers(0, k, ns)
        mg = ks.BlockLeastSquaresEstimator(bs, 1, lam, precision=args.precision).fit(ctx.matrix(F), ctx.labels_from_classes(c, k))
        xs, _, _ = ko.block_ls_fit(F.astype(np.float64), ko.class_label_indicators(c, k), bs, 1, lam)
        Wg, Wr = np.concatenate(mg.xs, 0), np.concatenate(xs, 0)
        rec["parity"] = {"n_rows": ns, "d": dsub, "rel_fro_W": float(np.linalg.norm(Wg - Wr) / np.linalg.norm(Wr))}
    emit(rec)


def c5():
    n, d_in, n_out, nrf, k, lam, w = args.c5_rows, 440, 4096, args.c5_blocks, 147, 6e-5, 0.25
    rng = np.random.default_rng(4)
    sizes = rng.multinomial(n - 4096 * k, np.ones(k) / k) + 4096 if n >= 4096 * k * 2 else rng.multinomial(n - 64 * k, np.ones(k) / k) + 64
    owner = np.arange(k) % world                                                   # class c lives on rank c % world
    mine = [c for c in range(k) if owner[c] == rank]
    cls = np.repeat(np.array(mine), sizes[mine]).astype(np.int32)                  # class-sorted rows of this rank
    x = ctx.synthetic_normal(len(cls), d_in, seed=4, global_row_offset=int(sizes[:mine[0]].sum()) if mine else 0)
    y = ctx.labels_from_classes(cls, k)
    prm = np.random.default_rng(5)
    rfs = [ks.CosineRandomFeatures.create(ctx, d_in, n_out, 0.0555, prm) for _ in range(nrf)]
    feats = ks.Pipeline.gather(rfs).andThen(ks.VectorCombiner())(x)
    est = ks.BlockWeightedLeastSquaresEstimator(n_out, 1, lam, w, precision=args.precision)
    barrier()
    t0 = time.perf_counter()
    m = est.fit(feats, y)
    _ = m.xs[-1][0, 0]
    barrier()
    dt = time.perf_counter() - t0
    st = ctx.last_fit_stats()
    emit({"config": f"C5 BlockWeightedLeastSquares N={n} D={nrf}x4096 k=147 lambda=6e-5 w=0.25, rows sharded by class", "gpus": world,
          "precision": st["mma"], "fit_s": dt, "samples_per_s": n / dt, "classes_on_rank0": st["classes_present"],
          "solve_lanes": st["solve_lanes"], "launches": st["launches"], "finite": bool(all(np.isfinite(wj).all() for wj in m.xs))})


def c1():
    """C1 (MnistRandomFFT.scala:40-47 at its CPU-runnable size): synthetic 60000 x 784 pixel-scale rows, gather(RandomSignNode ->
    PaddedFFT -> LinearRectifier(0)) x 4 -> D = 2048, BlockLeastSquaresEstimator(2048, 1, lambda), k = 10; full-size oracle parity.
    The pipeline's default lambda is 0: the 512 cosine bins of a 784-sample signal padded to 1024 are numerically dependent
    (time-bandwidth product ~392), so the Gram matrix is close to singular; the device path factorises with Cholesky and reports
    it (the reference's LU returns some solution).  Each lambda of --c1-lambdas is one record."""
    for lam in args.c1_lambdas:
        try:
            c1_one(float(lam))
        except ks.KeystoneError as e:
            emit({"config": f"C1 MnistRandomFFT-shaped 60000 x 784, numFFTs = 4 (D = 2048), k = 10, b = 2048, lambda = {lam}", "gpus": world,
                  "error": str(e)})


def c1_one(lam):
    n, d_in, nfft, k, bs = 60_000, 784, 4, 10, 2048
    rng = np.random.default_rng(0)
    X = rng.random((n, d_in), dtype=np.float32)
    cls = rng.integers(0, k, n).astype(np.int32)
    signs = [2.0 * rng.integers(0, 2, d_in) - 1.0 for _ in range(nfft)]
    x = ctx.matrix(X)
    y = ctx.labels_from_classes(cls, k)
    branches = [ks.RandomSignNode(sg, ctx).andThen(ks.PaddedFFT(ctx)).andThen(ks.LinearRectifier(0.0, ctx=ctx)) for sg in signs]
    feats = ks.Pipeline.gather(branches).andThen(ks.VectorCombiner())(x)
    m, dt = timed_fit(ks.BlockLeastSquaresEstimator(bs, 1, lam, precision=args.precision), feats, y)
    st = ctx.last_fit_stats()
    rec = {"config": f"C1 MnistRandomFFT-shaped 60000 x 784, numFFTs = 4 (D = 2048), k = 10, b = 2048, lambda = {lam}", "gpus": world,
           "precision": st["mma"], "fit_s": dt, "samples_per_s": n / dt, "alg_tflops": 5.11e11 / dt / 1e12}
    if world == 1:
        from oracle import keystone_oracle as ko
        t0 = time.perf_counter()
        F = ko.mnist_random_fft_features(X.astype(np.float64), signs)
        xs, b0, mus = ko.block_ls_fit(F, ko.class_label_indicators(cls, k), bs, 1, lam)
        t_cpu = time.perf_counter() - t0
        Wg, Wr = np.concatenate(m.xs, 0), np.concatenate(xs, 0)
        pred = m.apply_argmax(feats)
        ref = ko.block_linear_apply(F, xs, bs, b0, mus)
        rec["cpu_oracle_s"] = t_cpu
        rec["cpu_oracle_samples_per_s"] = n / t_cpu
        rec["cpu_cores"] = os.cpu_count()
        rec["parity"] = {"n_rows": n, "rel_fro_W": float(np.linalg.norm(Wg - Wr) / np.linalg.norm(Wr)),
                         "argmax_agree": float((pred == np.argmax(ref, 1)).mean())}
    emit(rec)


def c4f():
    """C4 WITH its featurizer (RandomPatchCifar.scala:59-66): synthetic CIFAR-shaped images -> Convolver(20000 random 6x6x3 filters,
    whitener means, normalised patches) -> SymmetricRectifier(0.25) -> Pooler(13, 14, sum) -> D = 160000 features on the device ->
    BlockLeastSquaresEstimator(4096, 1, lambda = 3000)."""
    n, nf, k, bs, lam = args.c4_rows, 20000, 10, 4096, 3000.0
    lo, hi = ks.shard_range(n, rank, world)
    rng = np.random.default_rng([3, lo])
    imgs = rng.integers(0, 256, (hi - lo, 3, 32, 32), dtype=np.uint8)              # CifarLoader layout
    prm = np.random.default_rng(33)
    filters = prm.standard_normal((nf, 108)) / 10.0
    wmeans = prm.standard_normal(108) * 0.05
    x = ctx.matrix(ks.cifar_bytes_to_matrix(imgs))
    conv = ks.Convolver(ctx, filters, 32, 32, 3, wmeans, True, 10.0)
    chain = conv.andThen(ks.SymmetricRectifier(alpha=0.25)).andThen(ks.Pooler(13, 14)).andThen(ks.ImageVectorizer())
    chain(ctx.matrix(ks.cifar_bytes_to_matrix(imgs[:64])))                          # warm-up
    barrier()
    t0 = time.perf_counter()
    feats = chain(x)
    barrier()
    t_feat = time.perf_counter() - t0
    cls = rng.integers(0, k, hi - lo).astype(np.int32)
    y = ctx.labels_from_classes(cls, k)
    m, dt = timed_fit(ks.BlockLeastSquaresEstimator(bs, 1, lam, precision=args.precision), feats, y, reps=1)
    st = ctx.last_fit_stats()
    rec = {"config": f"C4 with featurizer: N={n} CIFAR-shaped images, 20000 filters -> D=160000, k=10, lambda=3000", "gpus": world,
           "featurize_s": t_feat, "images_per_s": n / t_feat, "featurize_precision": "split fp16 operands (context default)",
           "fit_precision": st["mma"], "fit_s": dt, "pipeline_samples_per_s": n / (t_feat + dt), "num_blocks": st["num_blocks"],
           "finite": bool(all(np.isfinite(w).all() for w in m.xs))}
    if world == 1 and args.parity_rows:
        from oracle import keystone_oracle as ko
        ns = 32
        sub = np.transpose(imgs[:ns].astype(np.float64), (0, 2, 3, 1))              # [n][x][y][c]
        ref = np.stack([ko.random_patch_cifar_features(im, filters[:256], wmeans, 6, 0.25, 13, 14) for im in sub])
        conv_s = ks.Convolver(ctx, filters[:256], 32, 32, 3, wmeans, True, 10.0)
        got = conv_s.andThen(ks.SymmetricRectifier(alpha=0.25)).andThen(ks.Pooler(13, 14)).andThen(ks.ImageVectorizer())(
            ctx.matrix(ks.cifar_bytes_to_matrix(imgs[:ns]))).to_numpy()
        rec["parity"] = {"images": ns, "filters": 256, "max_rel_err_features": float(np.abs(got - ref).max() / np.abs(ref).max())}
    emit(rec)


for name in args.configs:
    try:
        {"c1": c1, "c2": c2, "c4": c4, "c4f": c4f, "c5": c5}[name]()
    except ks.KeystoneError as e:          # one failing configuration must not lose the others (same error on every rank)
        emit({"config": name, "gpus": world, "error": str(e)})
ctx.close()
if world > 1:
    dist.destroy_process_group()
