#!/usr/bin/env python
"""Headline benchmark: block least-squares fit throughput (samples/s) on BASELINE.json config 3
(CosineRandomFeatures 440 -> D = 16 x 4096 = 65536, BlockLeastSquaresEstimator(4096, numIter=1, lambda=1), k = 1000,
N = 1M rows sharded over the GPUs of one node; strong scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one complete fit over the whole batch.  `value` = N_total / (time per step) with the inputs resident
in HBM; `e2e` = the same fit through the public API starting from pinned HOST buffers (H2D of X and the class
labels, fit, D2H of the fitted model) -- see DESIGN.md section 7 for what each number includes.
`--impl reference` times the CPU stand-in for the reference (the numpy/OpenBLAS fp64 oracle: the reference's own
Spark/Breeze path needs a JVM that this image does not have) on a bounded row sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-rows", type=int, default=1_000_000)
    ap.add_argument("--d-in", type=int, default=440)
    ap.add_argument("--num-rf", type=int, default=16)
    ap.add_argument("--block", type=int, default=4096)
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--lam", type=float, default=1.0)
    ap.add_argument("--num-iter", type=int, default=1)
    ap.add_argument("--gamma", type=float, default=0.0555)
    ap.add_argument("--cpu-rows", type=int, default=2048, help="rows of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-tf32-ref", action="store_true", help="skip the extra tf32-operand fits reported beside an f16 run")
    ap.add_argument("--precision", default=os.environ.get("KS_BENCH_PRECISION", "f16"), choices=["tf32", "f16"],
                    help="operand type of the three big GEMMs (fp32 accumulate, fp64 solve either way)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------- workload
def make_workload(args, lo: int, hi: int, seed: int = 2):
    """Synthetic config-3 inputs for global rows [lo, hi): X ~ N(0,1) fp32, labels planted through a fixed
    16-dim linear map + noise, CosineRandomFeatures parameters shared by all ranks."""
    prm = np.random.default_rng(seed)
    params = [(prm.standard_normal((args.block, args.d_in)) * args.gamma, prm.random(args.block) * 2 * np.pi)
              for _ in range(args.num_rf)]
    wstar = prm.standard_normal((16, args.classes)).astype(np.float32)
    rng = np.random.default_rng([seed, lo])
    X = rng.standard_normal((hi - lo, args.d_in), dtype=np.float32)
    scores = X[:, :16] @ wstar + 0.1 * rng.standard_normal((hi - lo, args.classes), dtype=np.float32)
    cls = np.argmax(scores, axis=1).astype(np.int32)
    return X, cls, params


def alg_flops(n, d_in, D, b, k, nb, num_iter=1):
    """Gram 2NDb (full-GEMM convention) + A^T R 2NDk + update 2NDk + projection 2 N d_in D + solves."""
    first = 2.0 * n * D * b + 4.0 * n * D * k + 2.0 * n * d_in * D + nb * (b ** 3 / 3.0 + 2.0 * b * b * k)
    more = (num_iter - 1) * (4.0 * n * D * k + 2.0 * n * d_in * D + nb * 2.0 * b * b * k)
    return first + more


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------- CPU stand-in
def cpu_reference_fit(args, X, cls, params, rows, blocks=None):
    """One oracle fit (numpy/OpenBLAS fp64, all host threads) on the first `rows` rows.  Returns (seconds, seconds spent in
    the N-independent b x b solves), so the per-row cost can be separated from the fixed cost."""
    from oracle import keystone_oracle as ko
    Xs = X[:rows].astype(np.float64)
    Y = ko.class_label_indicators(cls[:rows], args.classes)
    use = params if blocks is None else params[:blocks]
    t_solve = [0.0]
    orig = ko._solve_spd

    def timed(G, C):
        t = time.perf_counter()
        out = orig(G, C)
        t_solve[0] += time.perf_counter() - t
        return out

    ko._solve_spd = timed
    try:
        t0 = time.perf_counter()
        blocks_f = [ko.cosine_random_features(Xs, W, b) for W, b in use]   # block i == feature map i (b_out == blockSize)
        ko.block_ls_fit(None, Y, args.block, args.num_iter, args.lam, feature_blocks=blocks_f)
        total = time.perf_counter() - t0
    finally:
        ko._solve_spd = orig
    return total, t_solve[0]


def cpu_baseline_record(args, X, cls, params, nb, D):
    cores, cpu_model = cpu_info()
    cpu_reference_fit(args, X, cls, params, min(args.cpu_rows, 512), blocks=1)       # warm the BLAS threads
    t_cpu, t_solve = cpu_reference_fit(args, X, cls, params, args.cpu_rows)
    per_row = (t_cpu - t_solve) / args.cpu_rows
    full = args.n_rows / (per_row * args.n_rows + t_solve)
    # `value` is the metric at the benchmark's N: the per-row cost measured on the sample scaled to N rows plus the measured
    # N-independent solves once (the sample alone would charge those 16 solves to 2048 rows and understate the CPU 2.6x)
    return {"value": full, "unit": "samples/s", "sample_value": args.cpu_rows / t_cpu, "cores": cores, "cpu": cpu_model, "kind": "port",
            "sample": f"first {args.cpu_rows} rows, all {nb} blocks (D={D}, k={args.classes}); numpy/OpenBLAS fp64 oracle, "
                      f"{t_cpu:.1f} s of which {t_solve:.1f} s are the N-independent {args.block}^2 solves; the Spark/Breeze "
                      f"reference itself needs a JVM (absent)",
            "seconds": t_cpu, "solve_seconds": t_solve,
            "extrapolated_full_n": {"value": full, "unit": "samples/s",
                                    "how": "N / (per_row_seconds * N + solve_seconds) with per_row from the sample"}}


def cpu_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return os.cpu_count() or 1, model


# ----------------------------------------------------------------------------------------- main
def main():
    # keep stdout to the one JSON line: NCCL prints a version banner there when NCCL_DEBUG is VERSION (or unset on some builds)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    D = args.num_rf * args.block
    nb = args.num_rf
    config = {"workload": f"C3 CosineRandomFeatures({args.d_in}->{args.block})x{args.num_rf} + BlockLeastSquaresEstimator",
              "n_rows": args.n_rows, "d_in": args.d_in, "d": D, "k": args.classes, "block_size": args.block,
              "num_iter": args.num_iter, "lambda": args.lam, "parallelism": f"rows x{world}",
              "l2": "inputs larger than L2 (X 1.76 GB, slab 8-16 GB per block)",
              "precision": f"{args.precision} operands, fp32 accumulate, fp64 solve"}

    if args.impl == "reference":
        if rank != 0:
            return
        X, cls, params = make_workload(args, 0, args.cpu_rows)
        recs = [cpu_baseline_record(args, X, cls, params, nb, D) for _ in range(max(1, args.steps))]
        secs = float(np.mean([r["seconds"] for r in recs]))
        sps = float(np.mean([r["value"] for r in recs]))
        rec = dict(recs[-1]); rec["value"] = sps
        print(json.dumps({"impl": "reference", "metric": "block-LS fit samples/sec (N=1M, D=64K, k=1K)", "value": sps,
                          "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": config, "cpu_baseline": rec,
                          "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    import keystone_b200 as ks

    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ctx = ks.Context.from_torch_distributed(local_rank)
    else:
        ctx = ks.Context(local_rank)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(local_rank)
        ctx.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lo, hi = ks.shard_range(args.n_rows, rank, world)
    X, cls, params = make_workload(args, lo, hi)
    Xp = torch.from_numpy(X).pin_memory()          # pinned host buffers for the e2e leg
    cp = torch.from_numpy(cls).pin_memory()
    rfs = [ks.CosineRandomFeatures(ctx, W, b) for W, b in params]
    est = ks.BlockLeastSquaresEstimator(args.block, args.num_iter, args.lam, precision=args.precision)

    def feats_of(x):
        return ks.Pipeline.gather(rfs).andThen(ks.VectorCombiner())(x)

    # ---- kernel-only leg: inputs resident in HBM
    x_dev = ctx.matrix(Xp.numpy())
    y_dev = ctx.labels_from_classes(cp.numpy(), args.classes)
    feats = feats_of(x_dev)
    stats = []
    for _ in range(args.warmup):
        est.fit(feats, y_dev)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    step_wall = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        model = est.fit(feats, y_dev)
        step_wall.append(1e3 * (time.perf_counter() - ts))
        stats.append(ctx.last_fit_stats())
    t_loop = time.perf_counter() - t0
    barrier()
    t_resident = max_over_ranks((time.perf_counter() - t0) / args.steps)
    launches = (ctx.launch_count() - l0) // max(args.steps, 1)
    clocks = sampler.stop()
    dev_ms = max_over_ranks(float(np.mean([s["total_ms"] for s in stats])))

    # ---- the same fit with tf32 operands (the other precision mode of the library), 1 warm-up + 2 timed fits, for reference
    tf32_ref = None
    if args.precision == "f16" and not args.no_tf32_ref:
        est32 = ks.BlockLeastSquaresEstimator(args.block, args.num_iter, args.lam, precision="tf32")
        est32.fit(feats, y_dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            est32.fit(feats, y_dev)
        barrier()
        t32 = max_over_ranks((time.perf_counter() - t0) / 2)
        tf32_ref = {"value": args.n_rows / t32, "unit": "samples/s", "ms_per_step": 1e3 * t32, "steps": 2, "warmup": 1}

    # ---- dominant kernel alone (same launch shape as inside the fit: S^T [S | R], N_loc x 4096 slab, k columns),
    #      CUDA events on the launching stream, 3 warm-up + 5 timed launches; the fit itself runs it concurrently with
    #      the residual chain on a second stream, so the in-fit span would not isolate the kernel
    import ctypes as C
    from keystone_b200._capi import check, lib
    sa = ctx.synthetic_normal(hi - lo, args.block, 11, lo)
    sb = ctx.synthetic_normal(hi - lo, args.classes, 12, lo)
    ms = C.c_double(0)
    ctx.set_option("precision", 1 if args.precision == "f16" else 0)   # the debug entry converts the operands to fp16 first
    check(ctx.handle, lib().ks_debug_time_gram(ctx.handle, sa.handle, sb.handle, 3, C.byref(ms)))
    check(ctx.handle, lib().ks_debug_time_gram(ctx.handle, sa.handle, sb.handle, 5, C.byref(ms)))
    ctx.set_option("precision", 0)
    gram_ms = max_over_ranks(ms.value)
    del sa, sb

    # ---- end-to-end leg: pinned host buffers -> public API -> fitted model on the host
    e2e = None
    if not args.no_e2e:
        del feats, x_dev, y_dev
        def e2e_step():
            xd = ctx.matrix(Xp.numpy())
            yd = ctx.labels_from_classes(cp.numpy(), args.classes)
            m = est.fit(feats_of(xd), yd)
            Wh = m.xs            # D2H of every W_j (fp64) ...
            bh = m.b_opt         # ... and the intercept
            return sum(w.nbytes for w in Wh) + bh.nbytes + sum(mu.nbytes for mu in m.feature_means)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(args.steps):
            d2h = e2e_step()
        barrier()
        t_e2e = max_over_ranks((time.perf_counter() - t0) / args.steps)
        e2e = {"value": args.n_rows / t_e2e, "unit": "samples/s", "ms_per_step": 1e3 * t_e2e,
               "h2d_bytes_per_step": int(X.nbytes + cls.nbytes), "d2h_bytes_per_step": int(d2h)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (gram_tn_kernel), timed live with CUDA events inside the fit
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    n_loc = hi - lo
    gram_launch_flops = 2.0 * n_loc * args.block * (args.block + args.classes)          # full-GEMM convention, per launch
    achieved = gram_launch_flops / (gram_ms * 1e-3) / 1e12
    # DRAM traffic of the dominant kernel from the committed ncu capture; only comparable when the capture was taken at this
    # launch shape (the round-1 capture is the reduced N_loc = 131072 run, so it is reported beside the number, not as it)
    traffic, traffic_ref = None, None
    try:
        traffic_ref = json.load(open(os.path.join(ROOT, "profiles", "gram_ncu_summary.json")))
        if traffic_ref.get("n_rows") == n_loc:
            traffic = traffic_ref.get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    kind = "kind::f16" if args.precision == "f16" else "kind::tf32"
    roofline = {"kernel": f"gram2_tn_kernel (tcgen05 cta_group::2 {kind}, S^T [S | R])", "bound": "tensor", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "note": "algorithmic flops = 2*N_loc*b*(b+k) per launch (full-GEMM convention; the kernel skips the lower "
                        "triangle: executed flops are 0.65x); peak is the measured SUSTAINED bf16 figure (35 ms launches run "
                        "under the power cap); the tf32 MMA rate is half of it, the fp16 rate equals it; kernel timed alone "
                        "with CUDA events",
                "executed_tflops": achieved * (104 * 256 * 512) / (args.block * (args.block + args.classes)) if args.block == 4096 and args.classes == 1000 else None,
                "ms_per_launch": gram_ms, "traffic_reference_capture": traffic_ref}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_record(args, X, cls, params, nb, D)

    flops = alg_flops(args.n_rows, args.d_in, D, args.block, args.classes, nb, args.num_iter)
    out = {"metric": "block-LS fit samples/sec (N=1M, D=64K, k=1K)", "value": args.n_rows / t_resident, "unit": "samples/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_resident,
           "device_ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": args.precision, "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches),
           "step_wall_ms": step_wall, "loop_ms_rank0": 1e3 * t_loop,
           "alg_tflops": flops / t_resident / 1e12, "phase_ms": {k: stats[-1][k] for k in stats[-1] if k.endswith("_ms")},
           "roofline": roofline}
    if tf32_ref:
        out["tf32_operands"] = tf32_ref
    if e2e:
        out["e2e"] = e2e
    if cpu_baseline:
        out["cpu_baseline"] = cpu_baseline
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
