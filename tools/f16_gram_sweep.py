"""Isolated timing of the Gram kernel (S^T [S | R]) for both operand types over the row-chunk size (GPU box)."""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
import keystone_b200 as ks
from keystone_b200._capi import check, lib


def main(n=262144, b=4096, k=1000, precs=(0, 1), chunks=(2048, 4096, 8192, 16384)):
    with ks.Context(0) as ctx:
        sa = ctx.synthetic_normal(n, b, 11, 0)
        sb = ctx.synthetic_normal(n, k, 12, 0)
        ms = C.c_double(0)
        for prec in precs:
            ctx.set_option("precision", prec)
            for chunk in chunks:
                ctx.set_option("gram_chunk_rows", chunk)
                check(ctx.handle, lib().ks_debug_time_gram(ctx.handle, sa.handle, sb.handle, 2, C.byref(ms)))
                check(ctx.handle, lib().ks_debug_time_gram(ctx.handle, sa.handle, sb.handle, 5, C.byref(ms)))
                print(json.dumps({"probe": "gram_sweep", "precision": ["tf32", "f16"][prec], "rows": n, "chunk_rows": chunk,
                                  "ms": ms.value, "alg_tflops": 2.0 * n * b * (b + k) / ms.value / 1e9}), flush=True)


if __name__ == "__main__":
    # no arguments: the sweep;  <rows> <tf32|f16> <chunk_rows[,chunk_rows...]>: one precision (the ncu target: warm-up launches, then 5)
    if len(sys.argv) == 4:
        main(n=int(sys.argv[1]), precs=(1 if sys.argv[2] == "f16" else 0,), chunks=tuple(int(v) for v in sys.argv[3].split(",")))
    else:
        main()
