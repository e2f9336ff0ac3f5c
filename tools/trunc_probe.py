"""Structure of the tensor core's accumulation error in the Gram kernel (kind::f16, fp32 accumulate in TMEM).

A has fp16-exact entries, so every product is exact and the only error is the accumulation.  Columns: 0..199 independent zero-mean,
200..219 = column 0 + noise (positively correlated with column 0), 220..239 = -column 0 + noise (negatively correlated), 240..255
zero-mean with mean shift +0.5 (positive products with every other shifted column).  For several chain lengths (gram_chunk_rows)
prints the signed relative error of: the diagonal, positively / negatively correlated entries, and zero-mean off-diagonal entries
(absolute, in units of the diagonal's ulp-free scale N * var)."""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import keystone_b200 as ks
from keystone_b200._capi import check, lib

n, m, kc = 65536, 256, 64
rng = np.random.default_rng(0)
A = rng.standard_normal((n, m)) * 0.7
A[:, 200:220] = A[:, :1] + 0.3 * rng.standard_normal((n, 20))
A[:, 220:240] = -A[:, :1] + 0.3 * rng.standard_normal((n, 20))
A[:, 240:256] += 0.5
A = A.astype(np.float16).astype(np.float64)
B = (rng.standard_normal((n, kc)) * 0.7).astype(np.float16).astype(np.float64)
Gx, Cx = A.T @ A, A.T @ B
ctx = ks.Context(0)
ctx.set_option("precision", 1)
a, b = ctx.matrix(A.astype(np.float32)), ctx.matrix(B.astype(np.float32))
for chunk in [int(v) for v in sys.argv[1:]] or [2048, 8192, 32768, 65536]:
    ctx.set_option("gram_chunk_rows", chunk)
    G = np.zeros((m, m)); Cm = np.zeros((m, kc))
    check(ctx.handle, lib().ks_debug_gram(ctx.handle, a.handle, b.handle, G.ctypes.data_as(C.c_void_p), m, Cm.ctypes.data_as(C.c_void_p), kc))
    E = G - Gx
    d = np.arange(m)
    scale = float(np.mean(np.diag(Gx)[:200]))
    zero = E[:200, :200][~np.eye(200, dtype=bool)]
    rec = {"probe": "trunc", "chain_rows": chunk, "mma_steps": chunk // 16,
           "diag_rel": float(np.mean(E[d, d] / Gx[d, d])),
           "pos_corr_rel": float(np.mean(E[0, 200:220] / Gx[0, 200:220])),
           "neg_corr_rel": float(np.mean(E[0, 220:240] / Gx[0, 220:240])),
           "shifted_pairs_rel": float(np.mean((E[240:256, 240:256] / Gx[240:256, 240:256])[~np.eye(16, dtype=bool)])),
           "zero_mean_offdiag_mean_over_diag": float(zero.mean() / scale), "zero_mean_offdiag_rms_over_diag": float(np.sqrt((zero ** 2).mean()) / scale),
           "C_mean_over_diag": float((Cm - Cx).mean() / scale), "C_rms_over_diag": float(np.sqrt(((Cm - Cx) ** 2).mean()) / scale)}
    print(json.dumps(rec), flush=True)
ctx.close()
