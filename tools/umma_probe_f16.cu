// fp16 twin of umma_probe.cu (debug tool, not part of the library): one CTA loads one fp16 operand stage with TMA, issues
// kind::f16 MMAs (K = 16 each) with a given (LBO, SBO, K step, major) setting and dumps the TMEM accumulator; the host compares
// against A^T B (MN-major operands, the Gram kernel) / A B^T (K-major operands, the update kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I keystone_b200/csrc tools/umma_probe_f16.cu -o build/umma_probe_f16
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

static void make_map(CUtensorMap* m, const __half* base, int rows, int cols, int ld, int box_cols, int box_rows, bool atom32 = false) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

// mode 0: MN-major (A [K rows x 128], B [K rows x 256], boxes {64 halfs, KR rows})
// mode 1: K-major  (A [128 x K], B [256 x K], boxes {64 halfs, 128|256 rows}), K = 64 per stage
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
  const int A_BYTES = pa.mode == 0 ? 2 * pa.krows * 128 : 128 * 128;
  const int B_BYTES = pa.mode == 0 ? 4 * pa.krows * 128 : 256 * 128;
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
      for (int i = 0; i < 2; ++i) tma_load_2d(sA + i * pa.krows * 128, &tmA, full_bar, 64 * i, 0);
      for (int i = 0; i < 4; ++i) tma_load_2d(sB + i * pa.krows * 128, &tmB, full_bar, 64 * i, 0);
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
    const uint32_t idesc = make_idesc_f16(128, 256, pa.a_major, pa.b_major);
    for (int kk = 0; kk < pa.nmma; ++kk) {
      const uint64_t ad = make_smem_desc(smem_u32(sA) + kk * pa.kstep, pa.lbo, pa.sbo, pa.layout);
      const uint64_t bd = make_smem_desc(smem_u32(sB) + kk * pa.kstep, pa.lbo, pa.sbo, pa.layout);
      umma_f16(tmem_base, ad, bd, idesc, kk != 0);
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
  // ---------------- MN-major: A [64 rows x 128], B [64 rows x 256]
  {
    const int KR = 64;
    std::vector<__half> A(KR * 128), B(KR * 256);
    for (auto& x : A) x = __float2half(float(rand() % 7 - 3));
    for (auto& x : B) x = __float2half(float(rand() % 7 - 3));
    __half *dA, *dB;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tA, tB;
    make_map(&tA, dA, KR, 128, 128, 64, KR); make_map(&tB, dB, KR, 256, 256, 64, KR);
    struct V { const char* name; uint32_t lbo, sbo, kstep; int nmma; int am, bm; uint32_t layout; };
    V vs[] = {
        {"mn16: lbo=box sbo=1024 kstep=2048 x4", (uint32_t)KR * 128, 1024, 2048, 4, 1, 1, 2},
        {"mn16: lbo=box sbo=1024 x1", (uint32_t)KR * 128, 1024, 2048, 1, 1, 1, 2},
        {"mn16: lbo=1024 sbo=box x1", 1024, (uint32_t)KR * 128, 2048, 1, 1, 1, 2},
        {"mn16: lbo=1024 sbo=box kstep=2048 x4", 1024, (uint32_t)KR * 128, 2048, 4, 1, 1, 2},
    };
    for (auto& v : vs) {
      ProbeArgs pa{0, KR, v.lbo, v.sbo, v.kstep, v.nmma, v.am, v.bm, v.layout};
      cudaMemset(dd, 0xff, 128 * 256 * 4);
      probe_kernel<<<1, 192, 80 * 1024>>>(tA, tB, pa, dsm, dd, dm);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 2; }
      std::vector<float> D(128 * 256);
      cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
      const int kuse = v.nmma * 16;
      std::vector<double> ref(128 * 256, 0.0);
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 256; ++n) { double s = 0; for (int r = 0; r < kuse; ++r) s += (double)__half2float(A[r * 128 + m]) * __half2float(B[r * 256 + n]); ref[m * 256 + n] = s; }
      double nz = 0; for (auto x : D) nz += (x != 0.f);
      printf("%-40s maxerr=%g  nonzero=%g  D[0][0..3]=%g %g %g %g ref=%g %g %g %g\n", v.name, check(D, ref), nz,
             D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
    }
  }
  // ---------------- K-major: A [128 x 64], B [256 x 64]
  {
    std::vector<__half> A(128 * 64), B(256 * 64);
    for (auto& x : A) x = __float2half(float(rand() % 7 - 3));
    for (auto& x : B) x = __float2half(float(rand() % 7 - 3));
    __half *dA, *dB;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tA, tB;
    make_map(&tA, dA, 128, 64, 64, 64, 128); make_map(&tB, dB, 256, 64, 64, 64, 256);
    struct V { const char* name; uint32_t lbo, sbo, kstep; int nmma; };
    V vs[] = {{"k16: lbo=16 sbo=1024 kstep=32 x4", 16, 1024, 32, 4}, {"k16: lbo=16 sbo=1024 x1", 16, 1024, 32, 1}};
    for (auto& v : vs) {
      ProbeArgs pa{1, 0, v.lbo, v.sbo, v.kstep, v.nmma, 0, 0, 2};
      cudaMemset(dd, 0xff, 128 * 256 * 4);
      probe_kernel<<<1, 192, 80 * 1024>>>(tA, tB, pa, dsm, dd, dm);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 2; }
      std::vector<float> D(128 * 256);
      cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
      const int kuse = v.nmma * 16;
      std::vector<double> ref(128 * 256, 0.0);
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 256; ++n) { double s = 0; for (int r = 0; r < kuse; ++r) s += (double)__half2float(A[m * 64 + r]) * __half2float(B[n * 64 + r]); ref[m * 256 + n] = s; }
      double nz = 0; for (auto x : D) nz += (x != 0.f);
      printf("%-40s maxerr=%g  nonzero=%g  D[0][0..3]=%g %g %g %g ref=%g %g %g %g\n", v.name, check(D, ref), nz, D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
    }
  }
  return 0;
}
