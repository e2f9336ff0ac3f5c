// Multi-right-hand-side Cholesky solve  X = (L L^T)^{-1} B  in ONE kernel (fp64, CUDA cores).
//
// Replaces cusolverDnDpotrs on the critical chain of the block solver (the reference's `\` on the driver,
// K/nodes/learning/BlockWeightedLeastSquares.scala:272, mlmatrix NormalEquations for BlockLS).  cuSOLVER runs the two
// triangular solves as ~100 small dependent kernels: 5.5 ms alone and ~11 ms when it shares the SMs with the look-ahead
// tensor work -- the dominant serial term of the strong-scaling curve (profiles/README.md).  Here every CTA owns 8
// right-hand sides and performs the complete forward and backward substitution for them; L (the lower triangle, 67 MB
// at b = 4096) is streamed from L2 once per pass per CTA and the kernel is a single launch.
//
//   forward   for each 128-row tile I:  V = B_I - L[I, 0:i0] Y[0:i0]   (bulk, all 256 threads, 32-row chunks of Y in smem)
//                                       Y_I = L[I,I]^{-1} V            (one warp per right-hand side, warp shuffles)
//   backward  for each tile I (last to first): V = Y_I - L[i1:n, I]^T X[i1:n] ; X_I = L[I,I]^{-T} V
//
// L: column-major n x n (ld = n), lower triangle valid (cusolverDnDpotrf, CUBLAS_FILL_MODE_LOWER).  B: column-major n x k.
#include "kernels.h"

namespace ks {

namespace {
constexpr int TS = 128;      // tile size (rows of the diagonal tile)
constexpr int NC = 8;        // right-hand sides per CTA == warps per CTA
constexpr int CH = 32;       // rows of Y / X staged per chunk of the bulk update
constexpr int LP = TS + 1;   // padded pitch (doubles) of transposed tiles in shared memory

struct SolveSmem {
  double diag[TS * LP];   // diagonal tile: forward  diag[q * LP + r] = L[i0 + r][i0 + q]  (r >= q)
                          //                backward diag[q * LP + r] = L[i0 + q][i0 + r]  (r <= q)
  double lt[CH * LP];     // backward bulk: lt[jj * LP + c] = L[j0 + jj][i0 + c]
  double xs[CH * NC];     // staged chunk of already-solved rows: xs[jj * NC + c]
  double v[TS * NC];      // tile right-hand side after the bulk update: v[r * NC + c]
  double invd[TS];
};
}  // namespace

__global__ void __launch_bounds__(256, 1)
chol_solve_kernel(const double* __restrict__ L, int n, double* __restrict__ B, int k) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SolveSmem& S = *reinterpret_cast<SolveSmem*>(smem_raw);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int c0 = blockIdx.x * NC;            // first right-hand side of this CTA
  const int row = t & (TS - 1);              // bulk update: tile row (forward) / tile column (backward)
  const int half = t >> 7;                   // bulk update: right-hand sides [4*half, 4*half + 4)
  const int ntiles = (n + TS - 1) / TS;
  const size_t ld = static_cast<size_t>(n);
  const int my_col = c0 + warp;              // in-tile solve: the right-hand side owned by this warp
  const bool col_ok = my_col < k;

  // =========================================================== forward: L Y = B
  for (int it = 0; it < ntiles; ++it) {
    const int i0 = it * TS;
    const int rows = min(TS, n - i0);
    // ---- stage the diagonal tile (column q of the tile contiguous in r: coalesced global reads, conflict-free smem)
    for (int idx = t; idx < TS * TS; idx += 256) {
      const int q = idx / TS, r = idx - q * TS;
      double val = 0.0;
      if (r >= q && r < rows) val = L[(i0 + r) + (i0 + q) * ld];
      S.diag[q * LP + r] = val;
    }
    // ---- bulk update  acc[row][c] = sum_{j < i0} L[i0 + row][j] * Y[j][c]
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const bool row_ok = row < rows;
    for (int j0 = 0; j0 < i0; j0 += CH) {
      __syncthreads();
      {  // stage Y[j0 .. j0+CH)[c0 .. c0+NC): 256 threads, one element each (coalesced along j for a fixed column)
        const int jj = t & (CH - 1), c = t >> 5;
        S.xs[jj * NC + c] = (c0 + c < k) ? B[(j0 + jj) + (c0 + c) * ld] : 0.0;
      }
      __syncthreads();
      if (row_ok) {
        const double* lp = L + (i0 + row) + j0 * ld;
#pragma unroll 8
        for (int jj = 0; jj < CH; ++jj) {
          const double l = lp[jj * ld];
          const double2 y01 = *reinterpret_cast<const double2*>(&S.xs[jj * NC + 4 * half]);
          const double2 y23 = *reinterpret_cast<const double2*>(&S.xs[jj * NC + 4 * half + 2]);
          acc[0] = fma(l, y01.x, acc[0]);
          acc[1] = fma(l, y01.y, acc[1]);
          acc[2] = fma(l, y23.x, acc[2]);
          acc[3] = fma(l, y23.y, acc[3]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col = c0 + 4 * half + c;
      S.v[row * NC + 4 * half + c] = (row_ok && col < k) ? B[(i0 + row) + col * ld] - acc[c] : 0.0;
    }
    if (t < TS) {
      const double d = S.diag[t * LP + t];
      S.invd[t] = (t < rows && d != 0.0) ? 1.0 / d : 0.0;
    }
    __syncthreads();
    // ---- in-tile forward substitution: warp w owns right-hand side w; lane holds rows lane, lane+32, lane+64, lane+96
    {
      double vr[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) vr[m] = S.v[(lane + 32 * m) * NC + warp];
      for (int q = 0; q < rows; ++q) {
        const int m = q >> 5, src = q & 31;
        double cand = vr[0];
        if (m == 1) cand = vr[1];
        else if (m == 2) cand = vr[2];
        else if (m == 3) cand = vr[3];
        const double yq = __shfl_sync(0xffffffffu, cand, src) * S.invd[q];
        const double* dq = &S.diag[q * LP];
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) {
          const int r = lane + 32 * mm;
          if (r == q) vr[mm] = yq;
          else if (r > q) vr[mm] = fma(-dq[r], yq, vr[mm]);
        }
      }
      if (col_ok) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int r = lane + 32 * m;
          if (r < rows) B[(i0 + r) + my_col * ld] = vr[m];
        }
      }
    }
    __syncthreads();  // Y_I is visible to the whole CTA before the next tile stages it
  }

  // =========================================================== backward: L^T X = Y
  for (int it = ntiles - 1; it >= 0; --it) {
    const int i0 = it * TS;
    const int rows = min(TS, n - i0);
    const int i1 = i0 + TS;
    // ---- stage the diagonal tile transposed: diag[q * LP + r] = L[i0 + q][i0 + r], r <= q
    for (int idx = t; idx < TS * TS; idx += 256) {
      const int r = idx / TS, q = idx - r * TS;  // q fastest: consecutive rows of L for a fixed column r -> coalesced
      double val = 0.0;
      if (r <= q && q < rows) val = L[(i0 + q) + (i0 + r) * ld];
      S.diag[q * LP + r] = val;
    }
    // ---- bulk update  acc[col][c] = sum_{j >= i1} L[j][i0 + col] * X[j][c]
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j0 = i1; j0 < n; j0 += CH) {
      __syncthreads();
      {
        const int jj = t & (CH - 1), c = t >> 5;
        S.xs[jj * NC + c] = (j0 + jj < n && c0 + c < k) ? B[(j0 + jj) + (c0 + c) * ld] : 0.0;
        // transposed chunk of L: thread reads 32 consecutive j of one tile column (256 B per warp)
#pragma unroll
        for (int m = 0; m < TS / 8; ++m) {
          const int cc = (t >> 5) + 8 * m;
          S.lt[jj * LP + cc] = (j0 + jj < n && cc < rows) ? L[(j0 + jj) + (i0 + cc) * ld] : 0.0;
        }
      }
      __syncthreads();
#pragma unroll 8
      for (int jj = 0; jj < CH; ++jj) {
        const double l = S.lt[jj * LP + row];
        const double2 x01 = *reinterpret_cast<const double2*>(&S.xs[jj * NC + 4 * half]);
        const double2 x23 = *reinterpret_cast<const double2*>(&S.xs[jj * NC + 4 * half + 2]);
        acc[0] = fma(l, x01.x, acc[0]);
        acc[1] = fma(l, x01.y, acc[1]);
        acc[2] = fma(l, x23.x, acc[2]);
        acc[3] = fma(l, x23.y, acc[3]);
      }
    }
    __syncthreads();
    const bool row_ok = row < rows;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col = c0 + 4 * half + c;
      S.v[row * NC + 4 * half + c] = (row_ok && col < k) ? B[(i0 + row) + col * ld] - acc[c] : 0.0;
    }
    if (t < TS) {
      const double d = S.diag[t * LP + t];
      S.invd[t] = (t < rows && d != 0.0) ? 1.0 / d : 0.0;
    }
    __syncthreads();
    // ---- in-tile backward substitution: x_q = v_q / L_qq ; v_r -= L[q][r] x_q for r < q
    {
      double vr[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) vr[m] = S.v[(lane + 32 * m) * NC + warp];
      for (int q = rows - 1; q >= 0; --q) {
        const int m = q >> 5, src = q & 31;
        double cand = vr[0];
        if (m == 1) cand = vr[1];
        else if (m == 2) cand = vr[2];
        else if (m == 3) cand = vr[3];
        const double xq = __shfl_sync(0xffffffffu, cand, src) * S.invd[q];
        const double* dq = &S.diag[q * LP];
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) {
          const int r = lane + 32 * mm;
          if (r == q) vr[mm] = xq;
          else if (r < q) vr[mm] = fma(-dq[r], xq, vr[mm]);
        }
      }
      if (col_ok) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int r = lane + 32 * m;
          if (r < rows) B[(i0 + r) + my_col * ld] = vr[m];
        }
      }
    }
    __syncthreads();
  }
}

cudaError_t launch_chol_solve(const double* L, int n, double* B, int k, cudaStream_t st) {
  if (n <= 0 || k <= 0) return cudaSuccess;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(sizeof(SolveSmem)));
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  chol_solve_kernel<<<(k + NC - 1) / NC, 256, sizeof(SolveSmem), st>>>(L, n, B, k);
  return cudaGetLastError();
}

}  // namespace ks
