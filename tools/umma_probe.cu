// Standalone descriptor probe (debug tool, not part of the library): one CTA loads one operand stage with TMA,
// dumps the raw shared memory, issues kind::tf32 MMAs with a given (LBO, SBO, major) setting and dumps the TMEM
// accumulator.  The host compares against A^T B / A B^T and prints which hypothesis matches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I keystone_b200/csrc tools/umma_probe.cu -o gpurun_out/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>

#include "tc_common.cuh"

using namespace ks;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_enc;

static void make_map(CUtensorMap* m, const float* base, int rows, int cols, int ld, int box_cols, int box_rows, bool atom32 = false) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

// mode 0: MN-major (A [K rows x 128], B [K rows x 256], boxes {32 floats, KR rows})
// mode 1: K-major  (A [128 x K], B [256 x K], boxes {32 floats, 128|256 rows}), K = 32 per stage
struct ProbeArgs {
  int mode, krows;          // MN-major: number of K rows in the stage (multiple of 8)
  uint32_t lbo, sbo, kstep; // descriptor byte offsets and the start-address advance per MMA
  int nmma;                 // number of MMAs (each K = 8)
  int a_major, b_major;
  uint32_t layout;
};

__global__ void __launch_bounds__(192, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ProbeArgs pa,
             float* dump_smem, float* dump_d, unsigned* dump_misc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int A_BYTES = pa.mode == 0 ? 4 * pa.krows * 128 : 128 * 128;
  const int B_BYTES = pa.mode == 0 ? 8 * pa.krows * 128 : 256 * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
  uint64_t* done_bar = full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(full_bar, 1);
      mbar_init(done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) dump_misc[0] = tmem_base;
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_BYTES;
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(full_bar, A_BYTES + B_BYTES);
    if (pa.mode == 0) {
      for (int i = 0; i < 4; ++i) tma_load_2d(sA + i * pa.krows * 128, &tmA, full_bar, 32 * i, 0);
      for (int i = 0; i < 8; ++i) tma_load_2d(sB + i * pa.krows * 128, &tmB, full_bar, 32 * i, 0);
    } else {
      tma_load_2d(sA, &tmA, full_bar, 0, 0);
      tma_load_2d(sB, &tmB, full_bar, 0, 0);
    }
  }
  mbar_wait(full_bar, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) dump_smem[i] = reinterpret_cast<float*>(smem)[i];
  __syncthreads();
  if (warp == 1 && elect_one()) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, 256, pa.a_major, pa.b_major);
    for (int kk = 0; kk < pa.nmma; ++kk) {
      const uint64_t ad = make_smem_desc(smem_u32(sA) + kk * pa.kstep, pa.lbo, pa.sbo, pa.layout);
      const uint64_t bd = make_smem_desc(smem_u32(sB) + kk * pa.kstep, pa.lbo, pa.sbo, pa.layout);
      umma_tf32(tmem_base, ad, bd, idesc, kk != 0);
    }
    umma_commit(done_bar);
    dump_misc[1] = idesc;
  }
  if (warp >= 2) {
    const int q = warp & 3;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < 256; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) dump_d[(q * 32 + lane) * 256 + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static double check(const std::vector<float>& D, const std::vector<double>& ref) {
  double e = 0;
  for (size_t i = 0; i < ref.size(); ++i) e = fmax(e, fabs((double)D[i] - ref[i]));
  return e;
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encode fn\n"); return 1; }
  g_enc = (PFN_encodeTiled)p;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  float *dsm, *dd; unsigned* dm;
  cudaMalloc(&dsm, 64 * 1024); cudaMalloc(&dd, 128 * 256 * 4); cudaMalloc(&dm, 64);
  srand(1);
  // ---------------- MN-major: A [32 rows x 128], B [32 rows x 256]
  {
    const int KR = 32;
    std::vector<float> A(KR * 128), B(KR * 256);
    for (auto& x : A) x = float(rand() % 7 - 3);
    for (auto& x : B) x = float(rand() % 7 - 3);
    float *dA, *dB;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tA, tB, tA32, tB32;
    make_map(&tA, dA, KR, 128, 128, 32, KR); make_map(&tB, dB, KR, 256, 256, 32, KR);
    make_map(&tA32, dA, KR, 128, 128, 32, KR, true); make_map(&tB32, dB, KR, 256, 256, 32, KR, true);
    struct V { const char* name; uint32_t lbo, sbo, kstep; int nmma; int am, bm; uint32_t layout; };
    V vs[] = {
        {"mn32: lbo=box sbo=512 kstep=1024 x4", (uint32_t)KR * 128, 512, 1024, 4, 1, 1, 1},
        {"mn32: lbo=box sbo=512 x1", (uint32_t)KR * 128, 512, 1024, 1, 1, 1, 1},
        {"mn32: lbo=512 sbo=box x1", 512, (uint32_t)KR * 128, 1024, 1, 1, 1, 1},
        {"mn: lbo=box sbo=1024 kstep=1024 x4", (uint32_t)KR * 128, 1024, 1024, 4, 1, 1, 2},
        {"mn: lbo=1024 sbo=box kstep=1024 x4", 1024, (uint32_t)KR * 128, 1024, 4, 1, 1, 2},
        {"mn: lbo=box sbo=1024 x1 (K=8 only)", (uint32_t)KR * 128, 1024, 1024, 1, 1, 1, 2},
        {"mn: lbo=1024 sbo=box x1 (K=8 only)", 1024, (uint32_t)KR * 128, 1024, 1, 1, 1, 2},
        {"mn-bits-off: lbo=box sbo=1024 x1", (uint32_t)KR * 128, 1024, 1024, 1, 0, 0, 2},
    };
    for (auto& v : vs) {
      ProbeArgs pa{0, KR, v.lbo, v.sbo, v.kstep, v.nmma, v.am, v.bm, v.layout};
      cudaMemset(dd, 0xff, 128 * 256 * 4);
      probe_kernel<<<1, 192, 80 * 1024>>>(v.layout == 1 ? tA32 : tA, v.layout == 1 ? tB32 : tB, pa, dsm, dd, dm);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 2; }
      std::vector<float> D(128 * 256), S(12 * KR * 32);
      unsigned misc[2];
      cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(S.data(), dsm, S.size() * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(misc, dm, 8, cudaMemcpyDeviceToHost);
      const int kuse = v.nmma * 8;
      std::vector<double> ref(128 * 256, 0.0);
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 256; ++n) { double s = 0; for (int r = 0; r < kuse; ++r) s += (double)A[r * 128 + m] * B[r * 256 + n]; ref[m * 256 + n] = s; }
      double nz = 0; for (auto x : D) nz += (x != 0.f);
      printf("%-40s maxerr=%g  nonzero=%g  D[0][0..3]=%g %g %g %g ref=%g %g %g %g tmem=%08x idesc=%08x\n", v.name, check(D, ref), nz,
             D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3], misc[0], misc[1]);
      // smem layout check: expected swizzled placement of A box 0: row r (128 B), chunk c -> chunk c ^ (r & 7)
      int bad = 0;
      for (int r = 0; r < KR; ++r) for (int c = 0; c < 8; ++c) for (int t = 0; t < 4; ++t) {
        float got = v.layout == 1 ? S[(r * 128 + (((c >> 1) ^ (r & 3)) * 32) + (c & 1) * 16) / 4 + t] : S[(r * 128 + ((c ^ (r & 7)) * 16)) / 4 + t];
        if (got != A[r * 128 + c * 4 + t]) ++bad;
      }
      printf("    smem A box0 swizzle check: %d mismatches; S[0..7]= %g %g %g %g %g %g %g %g (A row0: %g %g %g %g)\n", bad, S[0], S[1], S[2], S[3], S[4], S[5], S[6], S[7], A[0], A[1], A[2], A[3]);
    }
  }
  // ---------------- K-major: A [128 x 32], B [256 x 32]
  {
    std::vector<float> A(128 * 32), B(256 * 32);
    for (auto& x : A) x = float(rand() % 7 - 3);
    for (auto& x : B) x = float(rand() % 7 - 3);
    float *dA, *dB;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tA, tB;
    make_map(&tA, dA, 128, 32, 32, 32, 128); make_map(&tB, dB, 256, 32, 32, 32, 256);
    struct V { const char* name; uint32_t lbo, sbo, kstep; int nmma; };
    V vs[] = {{"k: lbo=16 sbo=1024 kstep=32 x4", 16, 1024, 32, 4}, {"k: lbo=16 sbo=1024 x1", 16, 1024, 32, 1}};
    for (auto& v : vs) {
      ProbeArgs pa{1, 0, v.lbo, v.sbo, v.kstep, v.nmma, 0, 0, 2};
      cudaMemset(dd, 0xff, 128 * 256 * 4);
      probe_kernel<<<1, 192, 80 * 1024>>>(tA, tB, pa, dsm, dd, dm);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 2; }
      std::vector<float> D(128 * 256);
      cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
      const int kuse = v.nmma * 8;
      std::vector<double> ref(128 * 256, 0.0);
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 256; ++n) { double s = 0; for (int r = 0; r < kuse; ++r) s += (double)A[m * 32 + r] * B[n * 32 + r]; ref[m * 256 + n] = s; }
      double nz = 0; for (auto x : D) nz += (x != 0.f);
      printf("%-40s maxerr=%g  nonzero=%g  D[0][0..3]=%g %g %g %g ref=%g %g %g %g\n", v.name, check(D, ref), nz, D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
    }
  }
  return 0;
}
