// Host-side engine state shared by engine.cu (C ABI, BlockLS, apply) and bwls.cu (weighted solver).
#pragma once
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <nccl.h>
#include <stdint.h>

#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/keystone_b200.h"
#include "kernels.h"

namespace ks {

struct KsError {
  int code;
  std::string msg;
};

#define KS_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      throw ::ks::KsError{KS_ERR_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(e__)};    \
  } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// Device memory comes from a per-process caching pool (engine.cu): a fit needs ~35 GB of workspace (slab, residuals,
// operand copies) and cudaMalloc/cudaFree of that much costs ~0.25 s per call -- more than the featurize phase.
void* pool_alloc(size_t bytes);
void pool_free(void* p, size_t bytes);
void pool_release_all();

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) pool_free(p, bytes);
    p = nullptr;
    bytes = 0;
  }
  void alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    p = pool_alloc(n);
    bytes = n;
  }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

struct Matrix {  // fp32 row-major, ld % 32 == 0, padding columns are zero
  DevBuf buf;
  float* d = nullptr;
  int64_t rows = 0, cols = 0, ld = 0;
};

struct CosRF {
  DevBuf wbuf, bbuf;
  float* W = nullptr;     // [n_out][ld] tf32-rounded, K-major GEMM operand
  float* bias = nullptr;  // [n_out]
  int64_t n_out = 0, n_in = 0, ld = 0;
};

struct Model {  // BlockLinearMapper state (K/nodes/learning/BlockLinearMapper.scala:22-33)
  int block_size = 0;
  int64_t k = 0;
  std::vector<int64_t> brows;
  std::vector<std::unique_ptr<DevBuf>> W;     // column-major (rows_j x k) fp64
  std::vector<std::unique_ptr<DevBuf>> mean;  // rows_j fp64 (if has_mean)
  DevBuf intercept;                           // k fp64
  bool has_mean = false, has_intercept = false;
};

struct SolverApi {
  void* lib = nullptr;
  cusolverStatus_t (*Create)(cusolverDnHandle_t*) = nullptr;
  cusolverStatus_t (*Destroy)(cusolverDnHandle_t) = nullptr;
  cusolverStatus_t (*SetStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
  cusolverStatus_t (*DpotrfBufferSize)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, int*) = nullptr;
  cusolverStatus_t (*Dpotrf)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, double*, int, int*) = nullptr;
  cusolverStatus_t (*Dpotrs)(cusolverDnHandle_t, cublasFillMode_t, int, int, const double*, int, double*, int, int*) = nullptr;
  cusolverStatus_t (*DpotriBufferSize)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, int*) = nullptr;
  cusolverStatus_t (*Dpotri)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, double*, int, int*) = nullptr;
};
struct BlasApi {  // cuBLAS, one plain library call: the fp64 symmetric product H^-1 * rhs
  void* lib = nullptr;
  cublasStatus_t (*Create)(cublasHandle_t*) = nullptr;
  cublasStatus_t (*Destroy)(cublasHandle_t) = nullptr;
  cublasStatus_t (*SetStream)(cublasHandle_t, cudaStream_t) = nullptr;
  cublasStatus_t (*Dsymm)(cublasHandle_t, cublasSideMode_t, cublasFillMode_t, int, int, const double*, const double*, int,
                          const double*, int, const double*, double*, int) = nullptr;
};
BlasApi& blas_api();
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;  // optional (NCCL >= 2.18)
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
SolverApi& solver_api();  // throws KsError if libcusolver cannot be loaded
NcclApi& nccl_api();      // throws KsError if libnccl cannot be loaded

enum Phase { PH_FEATURIZE = 0, PH_GRAM, PH_ALLREDUCE, PH_SOLVE, PH_UPDATE, PH_OTHER, PH_COUNT };

struct Ctx {
  int device = 0, rank = 0, world = 1;
  int num_sms = 148;
  cudaStream_t st = nullptr;   // main stream: residual-dependent chain (A^T R, triangular solves, update)
  cudaStream_t st2 = nullptr;  // prep stream: featurize + Gram of the blocks AHEAD (independent of the residual)
  cudaStream_t st3 = nullptr;  // factor stream: fp64 assembly + Cholesky of the blocks ahead
  cudaStream_t st4 = nullptr;  // broadcast stream of the owner-computed inverses (so a broadcast never blocks a rank's own factor work)
  ncclComm_t comm = nullptr;   // collectives issued on st
  ncclComm_t comm2 = nullptr;  // collectives issued on st2 (split of comm; falls back to comm)
  ncclComm_t comm3 = nullptr;  // broadcasts of the per-block inverse issued on st3
  int inv_min_world = 1 << 30; // experimental: from this world size on, block j's Cholesky + explicit inverse run on rank
                               // j % world only and are broadcast (off by default: cusolverDnDpotri is ~25 ms per 4096^2
                               // block and its fp64 work slows the tensor kernels more than the shorter solve gains)
  int shard_solve = 1;  // world > 1: every rank runs the triangular solves for its k / world right-hand sides only and the
                        // columns of dW are gathered (grouped ncclBroadcast, 32 MB at b = 4096, k = 1000) -- the solve is the
                        // serial term of the strong-scaling curve and its cost is proportional to the number of columns
  int exclusive_solve_min_world = 1 << 30;  // experimental: from this world size on the look-ahead Gram waits for the
                                            // critical chain's triangular solves (4.4 ms alone, ~11 ms contended); measured
                                            // neutral at 4 GPUs (332 vs 338 ms) because it serialises Gram and solve
  cusolverDnHandle_t solver = nullptr;   // triangular solves (main stream)
  cublasHandle_t blas = nullptr;         // H^-1 * rhs on the main stream
  cudaStream_t solver_stream = nullptr, solver2_stream = nullptr;  // streams the handles are currently bound to
  cusolverDnHandle_t solver2 = nullptr;  // factorizations (factor stream): a handle's internal cuBLAS workspace is per stream
  DevBuf solver_work;
  int solver_lwork = 0;
  DevBuf dev_info;  // int[kMaxInfo]
  std::string err;
  std::string stats_json;
  int64_t launches = 0;
  int64_t gram_chunk_rows = 0;  // rows of the contraction per Gram CTA (pair); 0 = chosen from the local row count (engine.cu)
  int gram_pair = 1;  // CTA-pair (cta_group::2) Gram kernel
  int epi_multi = 1;  // CTA-pair kernels: 8 rotating epilogue staging buffers per warp (0: one buffer, store-and-wait)
  int proj_f16 = 1;   // fp16 mode: the projection GEMM X W^T runs with fp16 operands too (0: tf32 operands, fp16 slab)
  int precision = 0;  // KS_PRECISION_TF32 (0) or KS_PRECISION_F16 (1: fp16 slab / residual / increment operands, kind::f16;
                      // BlockLeastSquares on generated (cosine) features only, everything else stays tf32)
  int reserve_sms = 8; // SMs the persistent look-ahead kernel leaves to the critical chain
  int custom_solve = 0; // experimental: 1 = chol_solve_kernel (single launch; 10.8 ms at b=4096,k=1000 vs 4.4 ms for
                        // cusolverDnDpotrs alone / ~11 ms when potrs shares the SMs), 0 = cusolverDnDpotrs
  int64_t sample_rows = 16384;
  int64_t next_id = 1;
  std::unordered_map<int64_t, std::unique_ptr<Matrix>> matrices;
  std::unordered_map<int64_t, std::unique_ptr<CosRF>> rfs;
  std::unordered_map<int64_t, std::unique_ptr<Model>> models;
  std::map<std::vector<int>, std::unique_ptr<DevBuf>> tile_cache;
  // phase timing of the current fit
  struct Span { int phase; cudaEvent_t a, b; int stream; };
  cudaEvent_t timeline_origin = nullptr;  // when set, collect_spans also renders (phase, stream, start, end) per span
  std::string timeline_json;
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
  bool timing = true;

  Matrix& matrix(int64_t h);
  CosRF& rf(int64_t h);
  Model& model(int64_t h);
  int64_t add(std::unique_ptr<Matrix> m);
  int64_t add(std::unique_ptr<Model> m);
  cudaEvent_t get_event();
  void span_begin(int phase, cudaStream_t s = nullptr);
  void span_end(cudaStream_t s = nullptr);
  void collect_spans(double out_ms[PH_COUNT]);
  void allreduce_f32(float* p, size_t n, bool prep = false);
  void allreduce_f64(double* p, size_t n, bool prep = false);
  void allreduce_max_u32(unsigned* p, size_t n);
  void ensure_solver();
  void potrf(double* H, int n, int info_slot, cudaStream_t s);
  void potrs(const double* H, int n, double* B, int nrhs, int info_slot, cudaStream_t s);
  // H (Cholesky factor, lower) -> lower triangle of H^-1, in place (cusolverDnDpotri); latency-bound, runs on the factor stream
  void potri(double* H, int n, int info_slot, cudaStream_t s);
  // out (n x nrhs) = sym(Hinv, lower) * B   (cublasDsymm): the whole solve of the critical chain is ONE fp64 GEMM
  void symm_solve(const double* Hinv, int n, const double* B, int nrhs, double* out, cudaStream_t s);
  void check_infos(int used_slots);
  void check_async(const char* what);
};

// Feature source: a materialised matrix or raw input + concatenated CosineRandomFeatures parameters.
struct FeatSrc {
  Matrix* F = nullptr;
  Matrix* X = nullptr;
  DevBuf xop;   // tf32-rounded copy of X (GEMM operand)
  // fp16 operand mode: X and the projection weights as fp16, each multiplied by a device-chosen power of two;
  // pscale[0] = 1 / (x scale * w scale) is applied to the accumulator before the cosine
  DevBuf xop16, w16, pscale;
  bool proj16 = false;
  // split-operand mode (KS_PRECISION_F16X2): K-concatenated fp16 operands [x_hi | x_lo | x_hi] and [w_hi | w_hi | w_lo], so one
  // GEMM of depth 3 d_in accumulates x_hi w_hi + x_lo w_hi + x_hi w_lo
  DevBuf x3, w3;
  int64_t ldx3 = 0, ldw3 = 0;
  bool proj_x2 = false;
  DevBuf wcat, bcat;
  float* Wall = nullptr;
  float* ball = nullptr;
  int64_t ldw = 0, d_in = 0;
  int64_t D = 0, n_rows = 0;
  DevBuf zeros;  // max(D-block, d_in) zero floats
};
void make_feat_src(Ctx& c, int64_t features, int64_t x_in, const int64_t* rfs, int32_t n_rfs, FeatSrc& out,
                   int precision = 0);  // KS_PRECISION_*: which operand copies of X / W to prepare
// slab[rows x lds] = round_tf32(features[row_begin : row_begin+rows, c0 : c0+cols] - shift)   (shift may be the zero vector)
// colsum (optional, fp32[cols], must be zeroed): receives the column sums of the stored slab
// out16: the slab is fp16 (lds in fp16 elements), generated features only
void produce_slab(Ctx& c, FeatSrc& src, int64_t c0, int64_t cols, const float* shift, void* slab, int64_t lds,
                  int64_t row_begin, int64_t rows, bool round_out = true, float* colsum = nullptr, cudaStream_t st = nullptr,
                  bool out16 = false, bool x2 = false);  // x2: unrounded fp32 slab from the K-concatenated split operands
const GramTile* gram_tiles(Ctx& c, int b, int kcols, bool with_g, bool with_c, bool pair, int* num_tiles);
// f16: slab and R are fp16 matrices (leading dimensions in elements); G / C stay fp32
void launch_gram_block(Ctx& c, const void* slab, int64_t lds, int64_t rows, int b, const void* R, int64_t ldr, int kcols,
                       float* G, int ldg, float* C, int ldc, bool with_g, bool with_c, cudaStream_t st = nullptr,
                       bool f16 = false, int64_t chunk_rows = 0);  // chunk_rows 0: the context's choice
// out[rows x k] (+)= (epi == EPI_UPDATE ? -1 : +1) * slab[rows x b] * bop[k x b]^T + cbias   (reduce: add into out)
// f16: slab and bop are fp16; the product is multiplied by *acc_scale_ptr (device scalar, may be null) before the epilogue
void launch_update(Ctx& c, const void* slab, int64_t lds, int64_t rows, int b, const void* bop, int64_t ldb, int k,
                   float* out, int64_t ldo, const float* cbias, int epi, bool reduce, cudaStream_t st = nullptr,
                   bool f16 = false, const float* acc_scale_ptr = nullptr);

int64_t fit_bwls(Ctx& c, FeatSrc& src, Matrix& Y, int bs, int num_iter, double lam, double w, int64_t nf_opt);

}  // namespace ks
