// tcgen05 / TMA tensor-core contraction path (sm_100a): bf16 operands staged in shared memory by
// TMA (128B swizzle), tcgen05.mma issued by one elected thread, fp32 accumulators in TMEM,
// tcgen05.ld back to registers for the fused epilogues.
//
//   forward   Y_ext[z] = P[z]^T S_ext[z]        A = P  (MN-major), B = S_ext (MN-major), split over cells
//   row-dot   r_i = <S_ext_i, (P dY_ext)_i>     A = P  (K-major),  B = dY_ext (MN-major), split over voxels
//   backward  dP = S_ext dY_ext^T  -> softmax-Jacobian + Adam in the epilogue (A, B K-major)
//
// Warp roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
#pragma once
#include <cuda.h>
#include <cstdio>
#include "common.cuh"
#include "gemm_simt.cuh"

namespace tgb {

constexpr int TC_BM = 128;        // UMMA M (TMEM lanes)
constexpr int TC_BK = 64;         // one 128-byte swizzle row of bf16 per k-block
constexpr int TC_UMMA_K = 16;
constexpr int TC_THREADS = 192;
constexpr int TC_RDOT_BN = 256;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> f32
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread t = lane base + t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors ----------------------------------------------------------------------------
// Shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1,
// a_major [15], b_major [16] (1 = MN-major), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Operand tile in shared memory, one pipeline stage.
//   K-major  (rows = MN index, 64 k-elements = 128 B per row):   ROWS x 128 B, SBO = 1024 B (8 rows)
//   MN-major (rows = k index,  64 mn-elements = 128 B per row):  ROWS/64 boxes of [64 k][128 B];
//            LBO = 8192 B (next 64-wide MN atom), SBO = 1024 B (next 8 k-rows)
template <bool KMAJOR, int ROWS>
struct OperandTile {
  static constexpr int kBytes = ROWS * TC_BK * 2;
  static __device__ __forceinline__ void load(const CUtensorMap* map, uint64_t* bar, uint8_t* dst, int mn0, int k0) {
    if (KMAJOR) {
      tma_load_2d(map, bar, dst, k0, mn0);                       // box {64 k, ROWS mn}
    } else {
#pragma unroll
      for (int b = 0; b < ROWS / 64; ++b) tma_load_2d(map, bar, dst + b * 8192, mn0 + b * 64, k0);  // box {64 mn, 64 k}
    }
  }
  static __device__ __forceinline__ uint64_t desc(uint32_t saddr, int k_step /* 0..3 */) {
    if (KMAJOR) return make_smem_desc(saddr + k_step * (TC_UMMA_K * 2), 16, 1024);
    return make_smem_desc(saddr + k_step * (TC_UMMA_K * 128), 8192, 1024);
  }
};

// ---- epilogues (thread = one TMEM lane = one output row; 16 consecutive columns per call) ------
struct TcEpiStore {
  float* C; int ldc; size_t split_stride; int M, N;   // N = ldc extent guard
  __device__ __forceinline__ void begin(int) {}
  __device__ __forceinline__ void chunk(int row, int col, const float (&v)[16]) {
    if (row >= M || col >= ldc) return;
    float4* dst = reinterpret_cast<float4*>(C + (size_t)blockIdx.z * split_stride + (size_t)row * ldc + col);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  __device__ __forceinline__ void end(int) {}
};

struct TcEpiRowDot {
  const __nv_bfloat16* S; int lds; float* rpart; int M;
  float acc;
  __device__ __forceinline__ void begin(int) { acc = 0.f; }
  __device__ __forceinline__ void chunk(int row, int col, const float (&v)[16]) {
    if (row >= M || col >= lds) return;
    const uint4* src = reinterpret_cast<const uint4*>(S + (size_t)row * lds + col);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const uint4 u = src[q];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
        acc = fmaf(v[q * 8 + 2 * e], __low2float(b), acc);
        acc = fmaf(v[q * 8 + 2 * e + 1], __high2float(b), acc);
      }
    }
  }
  __device__ __forceinline__ void end(int row) {
    if (row < M) rpart[((size_t)blockIdx.z * gridDim.x + blockIdx.x) * M + row] = acc;
  }
};

struct TcAdamArgs {
  float* Mp; float* mp; float* vp; int ld; int V;
  const RowStat* stats; const float* rdot;
  float lam_r, lam_l1, lam_l2;
  AdamScalars a;
};
struct TcEpiAdam {
  TcAdamArgs p; int M;
  RowStat st; float r;
  __device__ __forceinline__ void begin(int row) {
    if (row < M) { st = p.stats[row]; r = p.rdot[row]; }
  }
  __device__ __forceinline__ float one(float x, float dp, float& m, float& v) const {
    const float pr = softmax_prob(x, st);
    float g = dp - r;
    if (p.lam_r != 0.f) g -= p.lam_r * (((x - st.mx) - st.log_z) - st.h);
    g *= pr;
    if (p.lam_l1 != 0.f) g += p.lam_l1 * (float)((x > 0.f) - (x < 0.f));
    if (p.lam_l2 != 0.f) g += 2.f * p.lam_l2 * x;
    return adam_update(x, g, m, v, p.a);
  }
  __device__ __forceinline__ void chunk(int row, int col, const float (&acc)[16]) {
    if (row >= M || col >= p.V) return;
    const size_t o = (size_t)row * p.ld + col;
    float4* Mq = reinterpret_cast<float4*>(p.Mp + o);
    float4* mq = reinterpret_cast<float4*>(p.mp + o);
    float4* vq = reinterpret_cast<float4*>(p.vp + o);
    float4 x[4], m[4], v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { x[q] = Mq[q]; m[q] = mq[q]; v[q] = vq[q]; }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = col + 4 * q;
      if (c + 0 < p.V) x[q].x = one(x[q].x, acc[4 * q + 0], m[q].x, v[q].x);
      if (c + 1 < p.V) x[q].y = one(x[q].y, acc[4 * q + 1], m[q].y, v[q].y);
      if (c + 2 < p.V) x[q].z = one(x[q].z, acc[4 * q + 2], m[q].z, v[q].z);
      if (c + 3 < p.V) x[q].w = one(x[q].w, acc[4 * q + 3], m[q].w, v[q].w);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { Mq[q] = x[q]; mq[q] = m[q]; vq[q] = v[q]; }
  }
  __device__ __forceinline__ void end(int) {}
};

// ---- the kernel -------------------------------------------------------------------------------
// grid: x = N tiles, y = M tiles, z = k splits.  One output tile (128 x BN) per CTA.
template <bool A_KMAJOR, bool B_KMAJOR, int BN, int STAGES, int MIN_CTAS, class Epi>
__global__ void __launch_bounds__(TC_THREADS, MIN_CTAS)
k_gemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
          int k_total, int k_per_split, Epi epi) {
  using TileA = OperandTile<A_KMAJOR, TC_BM>;
  using TileB = OperandTile<B_KMAJOR, BN>;
  constexpr int kStageBytes = TileA::kBytes + TileB::kBytes;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;   // power of two (BN is 128 or 256)
  constexpr uint32_t kIdesc = make_idesc(TC_BM, BN, !A_KMAJOR, !B_KMAJOR);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TC_BM;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(k_total, k_begin + k_per_split);
  const int num_kb = (k_end > k_begin) ? (k_end - k_begin + TC_BK - 1) / TC_BK : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * kStageBytes;
        uint8_t* sb = sa + TileA::kBytes;
        mbar_expect_tx(&full_bar[s], kStageBytes);
        const int k0 = k_begin + kb * TC_BK;
        TileA::load(&map_a, &full_bar[s], sa, m0, k0);
        TileB::load(&map_b, &full_bar[s], sb, n0, k0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * kStageBytes);
        const uint32_t sb = sa + TileA::kBytes;
#pragma unroll
        for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
          umma_bf16(tmem_base, TileA::desc(sa, k), TileB::desc(sb, k), kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[s]);                       // frees the smem stage when these MMAs retire
        if (kb == num_kb - 1) umma_commit(&tmem_full_bar);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> fused epilogue =====
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    epi.begin(row);
    if (num_kb > 0) {
      mbar_wait(&tmem_full_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      float v[16];
      if (num_kb > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      epi.chunk(row, n0 + c, v);
    }
    epi.end(row);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcContext {
  PFN_encodeTiled encode = nullptr;
};

static inline int tc_init(TcContext& tc, char* err, size_t n) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    snprintf(err, n, "cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return -2;
  }
  tc.encode = (PFN_encodeTiled)fn;
  return 0;
}

// 2D bf16 row-major [rows][cols] (row pitch ld elements): inner dim = cols.  box = {box_inner, box_outer}.
static inline int tc_make_map(TcContext& tc, CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows,
                              uint64_t ld, uint32_t box_inner, uint32_t box_outer, char* err, size_t n) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = tc.encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, n, "cuTensorMapEncodeTiled failed (%d) cols=%llu rows=%llu ld=%llu box=%ux%u", (int)r,
             (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld, box_inner, box_outer);
    return -2;
  }
  return 0;
}

template <class Kern>
static inline int tc_set_smem(Kern kern, int bytes, char* err, size_t n) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { snprintf(err, n, "cudaFuncSetAttribute(smem=%d): %s", bytes, cudaGetErrorString(e)); return -2; }
  return 0;
}
static inline int tc_check_launch(const char* name, char* err, size_t n) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(err, n, "launch %s: %s", name, cudaGetErrorString(e)); return -2; }
  return 0;
}

constexpr int TC_FWD_BN = 256, TC_FWD_STAGES = 4;
constexpr int TC_BWD_BN = 128, TC_BWD_STAGES = 3;
constexpr int TC_RD_STAGES = 4;

static inline int tc_splits(long long tiles, long long k_total, int min_k) {
  long long s = (2 * 148 + tiles - 1) / tiles;
  const long long max_s = (k_total + min_k - 1) / min_k;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  // make every split non-empty
  const long long kps = ((k_total + s - 1) / s + TC_BK - 1) / TC_BK * TC_BK;
  return (int)((k_total + kps - 1) / kps);
}
static inline int tc_forward_splits(int N, int V, int Ke) {
  return tc_splits((long long)ceil_div(V, TC_BM) * ceil_div(Ke, TC_FWD_BN), N, 512);
}
static inline int tc_rowdot_splits(int N, int V, int Ke) {
  return tc_splits((long long)ceil_div(N, TC_BM) * ceil_div(Ke, TC_RDOT_BN), V, 1024);
}
static inline int tc_kps(int k_total, int splits) {
  return (int)(round_up(ceil_div(k_total, splits), TC_BK));
}

// Y_ext[z] (V x Ke) = P[cells of split z]^T S_ext[...]
static inline int tc_forward(TcContext& tc, const __nv_bfloat16* P, const __nv_bfloat16* Sx, float* out, int N, int V,
                             int Ke, int ld, int splits, cudaStream_t s, char* err, size_t n) {
  CUtensorMap ma, mb;
  if (tc_make_map(tc, &ma, P, V, N, ld, 64, 64, err, n)) return -2;      // A: MN-major (voxels contiguous), rows = cells
  if (tc_make_map(tc, &mb, Sx, Ke, N, Ke, 64, 64, err, n)) return -2;    // B: MN-major (genes contiguous), rows = cells
  auto kern = k_gemm_tc<false, false, TC_FWD_BN, TC_FWD_STAGES, 1, TcEpiStore>;
  const int smem = TC_FWD_STAGES * (TC_BM + TC_FWD_BN) * TC_BK * 2 + 1024;
  if (tc_set_smem(kern, smem, err, n)) return -2;
  TcEpiStore epi{out, Ke, (size_t)V * Ke, V, Ke};
  dim3 grid((unsigned)ceil_div(Ke, TC_FWD_BN), (unsigned)ceil_div(V, TC_BM), splits);
  kern<<<grid, TC_THREADS, smem, s>>>(ma, mb, N, tc_kps(N, splits), epi);
  return tc_check_launch("tc_gemm_fwd", err, n);
}

// rpart[(z * ntiles_n + tn)][i] = sum over the tile's genes of (P dY_ext)_ik S_ext_ik
static inline int tc_rowdot(TcContext& tc, const __nv_bfloat16* P, const __nv_bfloat16* dYb, const __nv_bfloat16* Sxb,
                            float* rpart, int N, int V, int Ke, int ld, int splits, cudaStream_t s, char* err, size_t n) {
  CUtensorMap ma, mb;
  if (tc_make_map(tc, &ma, P, V, N, ld, 64, TC_BM, err, n)) return -2;   // A: K-major (contraction over voxels)
  if (tc_make_map(tc, &mb, dYb, Ke, V, Ke, 64, 64, err, n)) return -2;   // B: MN-major (genes contiguous), rows = voxels
  auto kern = k_gemm_tc<true, false, TC_RDOT_BN, TC_RD_STAGES, 1, TcEpiRowDot>;
  const int smem = TC_RD_STAGES * (TC_BM + TC_RDOT_BN) * TC_BK * 2 + 1024;
  if (tc_set_smem(kern, smem, err, n)) return -2;
  TcEpiRowDot epi{Sxb, Ke, rpart, N, 0.f};
  dim3 grid((unsigned)ceil_div(Ke, TC_RDOT_BN), (unsigned)ceil_div(N, TC_BM), splits);
  kern<<<grid, TC_THREADS, smem, s>>>(ma, mb, V, tc_kps(V, splits), epi);
  return tc_check_launch("tc_gemm_rowdot", err, n);
}

// dP = S_ext dY_ext^T fused with the softmax-Jacobian and Adam
static inline int tc_backward(TcContext& tc, const __nv_bfloat16* Sxb, const __nv_bfloat16* dYb, const TcAdamArgs& a,
                              int N, int V, int Ke, cudaStream_t s, char* err, size_t n) {
  CUtensorMap ma, mb;
  if (tc_make_map(tc, &ma, Sxb, Ke, N, Ke, 64, TC_BM, err, n)) return -2;     // A: K-major, rows = cells
  if (tc_make_map(tc, &mb, dYb, Ke, V, Ke, 64, TC_BWD_BN, err, n)) return -2; // B: K-major, rows = voxels
  auto kern = k_gemm_tc<true, true, TC_BWD_BN, TC_BWD_STAGES, 2, TcEpiAdam>;
  const int smem = TC_BWD_STAGES * (TC_BM + TC_BWD_BN) * TC_BK * 2 + 1024;
  if (tc_set_smem(kern, smem, err, n)) return -2;
  TcEpiAdam epi{a, N};
  dim3 grid((unsigned)ceil_div(V, TC_BWD_BN), (unsigned)ceil_div(N, TC_BM), 1);
  kern<<<grid, TC_THREADS, smem, s>>>(ma, mb, Ke, Ke, epi);
  return tc_check_launch("tc_gemm_bwd_adam", err, n);
}

}  // namespace tgb
