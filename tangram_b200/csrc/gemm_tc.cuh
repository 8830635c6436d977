// tcgen05 / TMA tensor-core contraction path (sm_100a): bf16 operands staged in shared memory by
// TMA (128B swizzle), tcgen05.mma issued by one elected thread, fp32 accumulators in TMEM,
// tcgen05.ld back to registers for the fused epilogues.
//
//   forward   Y_ext[z] = P[z]^T S_ext[z]        A = P  (MN-major), B = S_ext (MN-major), split over cells / cell chunks
//   backward  dP = S_ext dY_ext^T  -> stored (bf16 centred / fp32) + row-dot partials in the epilogue (A, B K-major);
//             from 2048 cells up on CTA pairs (k_gemm_tc_pair: tcgen05 cta_group::2, 256-row tiles, half of B per CTA).
//             The update itself is a streaming kernel (adam_rows.cuh).
//
// Persistent, warp-specialised kernel: one CTA per SM loops over output tiles.
//   warp 0      TMA producer: keeps the operand ring full across tile boundaries
//   warp 1      TMEM allocator + MMA issuer: two accumulator buffers in TMEM, so the MMAs of tile i+1
//               run while the epilogue warps drain tile i
//   warps 2..   epilogue (TMEM lane quarter = warp_id % 4)
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// L2 policy: operand tiles are re-read by many CTAs -> evict_last (CUTLASS TMA::CacheHintSm90::EVICT_LAST)
constexpr uint64_t kPolicyEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kPolicyEvictLast = 0x14F0000000000000ull;
constexpr uint64_t kPolicyEvictFirst = 0x12F0000000000000ull;
// one full 32-byte sector per instruction (sm_100+: STG.256) -- 16-byte stores from a row-per-thread layout are partial-sector
// writes, which cost L2 fill reads from DRAM
__device__ __forceinline__ void stg256(void* dst, const uint32_t* w) {
  asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
// one full 32-byte sector per thread, read-only path (sm_100+: LDG.256)
__device__ __forceinline__ void ldg256(const void* src, uint32_t (&w)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(src));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy) : "memory");
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
  static __device__ __forceinline__ void load(const CUtensorMap* map, uint64_t* bar, uint8_t* dst, int mn0, int k0, uint64_t policy) {
    if (KMAJOR) {
      tma_load_2d(map, bar, dst, k0, mn0, policy);               // box {64 k, ROWS mn}
    } else {
#pragma unroll
      for (int b = 0; b < ROWS / 64; ++b) tma_load_2d(map, bar, dst + b * 8192, mn0 + b * 64, k0, policy);  // box {64 mn, 64 k}
    }
  }
  static __device__ __forceinline__ uint64_t desc(uint32_t saddr, int k_step /* 0..3 */) {
    if (KMAJOR) return make_smem_desc(saddr + k_step * (TC_UMMA_K * 2), 16, 1024);
    return make_smem_desc(saddr + k_step * (TC_UMMA_K * 128), 8192, 1024);
  }
};

// ---- epilogues ------------------------------------------------------------------------------
// Each epilogue warp owns TMEM lanes [32q, 32q+32) = output rows m0+32q.. of the tile.
// run() is called once per warp after the accumulator is complete; `scratch` is the (now idle)
// operand ring, 1024-byte aligned, at least 32 KB.
// Coordinates of the work item an epilogue warp is draining.
// Work item w -> (row tile m, column tile n).  Row tiles are taken in groups of `group_m`; inside a group the order is
// column-major (all rows of the group for column 0, then column 1, ...).  The CTAs running at the same time then cover
// the whole group for a few columns: the group's A rows (group_m x 128 x K, re-read for every column) are a small,
// hot L2 footprint and each B column tile is used in one burst.  group_m = 1 is plain column-fastest order.
__device__ __forceinline__ void tile_mn(int w, int tiles_m, int tiles_n, int group_m, int& m, int& n) {
  const int per_group = group_m * tiles_n;
  const int g = w / per_group, r = w - g * per_group;
  const int rows = min(group_m, tiles_m - g * group_m);
  n = r / rows;
  m = g * group_m + (r - n * rows);
}
struct TileCoord {
  int m0, n0;        // first output row / column of the tile
  int tile_n;        // column-tile index
  int tiles_n;       // number of column tiles
  int split;         // k-split index
};

struct TcEpiStore {
  float* C; int ldc; size_t split_stride; int M;
  int accumulate;        // 1: C += tile (cell chunks of the pipelined forward run one after the other: fixed summation order)
  template <int BN, int NW>
  __device__ __forceinline__ void prologue(const TileCoord&, int, int, int) const {}
  __device__ __forceinline__ void finish(int, int, int) const {}
  template <int BN, int NW>
  __device__ __forceinline__ void run(uint32_t tmem_acc, int q, int, int lane, const TileCoord& t) const {
    static_assert(NW == 4, "one warp per TMEM lane quarter");
    const int row = t.m0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      float v[16];
      tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      const int col = t.n0 + c;
      if (row < M && col < ldc) {
        float4* dst = reinterpret_cast<float4*>(C + (size_t)t.split * split_stride + (size_t)row * ldc + col);
        if (accumulate) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 o = dst[e];
            v[4 * e] += o.x; v[4 * e + 1] += o.y; v[4 * e + 2] += o.z; v[4 * e + 3] += o.w;
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
      }
    }
  }
};

// Store-only backward epilogue (bf16 throughput mode, "staged" backward): dP = S_ext dY_ext^T leaves the kernel as
// bf16, centred per row on the previous iteration's row-dot (dq_ij = bf16(dP_ij - c_i): the softmax-Jacobian only sees
// dP_ij - r_i, so the bf16 rounding is relative to the deviation from the row mean, not to dP itself), and the same
// epilogue accumulates this iteration's row-dot partials r'_i = sum_j Pt_ij dq_ij from the ROUNDED values (so that
// sum_j g_ij = 0 holds for what the streaming Adam kernel consumes).  Thread = row (TMEM lane), one 32-byte sector of
// Pt in and one of dq out per 16 columns; the Pt segment of the tile is prefetched into L2 before the wait on the MMAs.
struct TcEpiDpStore {
  __nv_bfloat16* dq; const __nv_bfloat16* Pt; int ld;     // both [rows][ld]
  const float* center;                                    // c_i (per row)
  float* rpart;                                           // [tiles_n * (NW / 4)][M]
  int M;
  // NW = 4: one warp per TMEM lane quarter takes all BN columns; NW = 8: two warps, BN / 2 columns each.  Four warps keep
  // the CTA small (192 threads) so that two CTAs of the streaming Adam kernel fit next to it on the SM.
  template <int BN, int NW>
  __device__ __forceinline__ void prologue(const TileCoord& t, int q, int ew, int lane) const {
    static_assert(NW == 4 || NW == 8, "one or two warps per TMEM lane quarter");
    constexpr int W = BN / (NW / 4);
    const int row = t.m0 + q * 32 + lane, col = t.n0 + (ew >> 2) * W;
    if (row < M) {
      const __nv_bfloat16* src = Pt + (size_t)row * ld + col;
#pragma unroll
      for (int b = 0; b < W; b += 64)           // 128-byte lines of this thread's row segment
        if (col + b < ld) prefetch_l2(src + b);
    }
  }
  __device__ __forceinline__ void finish(int, int, int) const {}
  template <int BN, int NW>
  __device__ __forceinline__ void run(uint32_t tmem_acc, int q, int ew, int lane, const TileCoord& t) const {
    constexpr int PARTS = NW / 4, W = BN / PARTS, NCH = W / 16;
    const int part = ew >> 2;
    const int row = t.m0 + q * 32 + lane;
    const int cbase = t.n0 + part * W;
    const bool live = row < M;
    const float c = live ? center[row] : 0.f;
    const __nv_bfloat16* prow = Pt + (size_t)row * ld;
    __nv_bfloat16* drow = dq + (size_t)row * ld;
    uint32_t pw[2][8];
    if (live && cbase < ld) ldg256(prow + cbase, pw[0]);
    float racc = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int col0 = cbase + i * 16;
      if (i + 1 < NCH && live && col0 + 16 < ld) ldg256(prow + col0 + 16, pw[(i + 1) & 1]);
      float v[16];
      tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 - t.n0), v);
      if (live && col0 < ld) {
        uint32_t w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const __nv_bfloat162 d2 = __floats2bfloat162_rn(v[2 * e] - c, v[2 * e + 1] - c);
          const __nv_bfloat162 p2 = *reinterpret_cast<const __nv_bfloat162*>(&pw[i & 1][e]);
          racc = fmaf(__low2float(p2), __low2float(d2), racc);
          racc = fmaf(__high2float(p2), __high2float(d2), racc);
          w[e] = *reinterpret_cast<const uint32_t*>(&d2);
        }
        stg256(drow + col0, w);
      }
    }
    if (live) rpart[((size_t)t.tile_n * PARTS + part) * M + row] = racc;
  }
};

// The same store-only backward for the parity mode (bf16x3): dP leaves the kernel in fp32 (nothing to centre) and the
// row-dot partials use P reconstructed from its three bf16 planes (hi + mid + lo = the fp32 value the row pass computed).
struct TcEpiDpStoreF32 {
  float* dp; int ld;                       // [rows][ld] fp32
  const __nv_bfloat16* P3; size_t plane;   // three planes of [rows][ld] bf16
  float* rpart;                            // [tiles_n * (NW / 4)][M]
  int M;
  template <int BN, int NW>
  __device__ __forceinline__ void prologue(const TileCoord& t, int q, int ew, int lane) const {
    static_assert(NW == 4 || NW == 8, "one or two warps per TMEM lane quarter");
    constexpr int W = BN / (NW / 4);
    const int row = t.m0 + q * 32 + lane, col = t.n0 + (ew >> 2) * W;
    if (row < M) {
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        const __nv_bfloat16* src = P3 + pl * plane + (size_t)row * ld + col;
#pragma unroll
        for (int b = 0; b < W; b += 64)
          if (col + b < ld) prefetch_l2(src + b);
      }
    }
  }
  __device__ __forceinline__ void finish(int, int, int) const {}
  template <int BN, int NW>
  __device__ __forceinline__ void run(uint32_t tmem_acc, int q, int ew, int lane, const TileCoord& t) const {
    constexpr int PARTS = NW / 4, W = BN / PARTS, NCH = W / 16;
    const int part = ew >> 2;
    const int row = t.m0 + q * 32 + lane;
    const int cbase = t.n0 + part * W;
    const bool live = row < M;
    const __nv_bfloat16* prow = P3 + (size_t)row * ld;
    float* drow = dp + (size_t)row * ld;
    float racc = 0.f;
#pragma unroll 1
    for (int i = 0; i < NCH; ++i) {
      const int col0 = cbase + i * 16;
      uint32_t ph[8], pm[8], pl[8];
      if (live && col0 < ld) { ldg256(prow + col0, ph); ldg256(prow + plane + col0, pm); ldg256(prow + 2 * plane + col0, pl); }
      float v[16];
      tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 - t.n0), v);
      if (live && col0 < ld) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&ph[e]);
          const __nv_bfloat162 m2 = *reinterpret_cast<const __nv_bfloat162*>(&pm[e]);
          const __nv_bfloat162 l2 = *reinterpret_cast<const __nv_bfloat162*>(&pl[e]);
          const float p0 = (__low2float(l2) + __low2float(m2)) + __low2float(h2);
          const float p1 = (__high2float(l2) + __high2float(m2)) + __high2float(h2);
          racc = fmaf(p0, v[2 * e], racc);
          racc = fmaf(p1, v[2 * e + 1], racc);
        }
        uint32_t w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) w[e] = __float_as_uint(v[e]);
        stg256(drow + col0, w);
#pragma unroll
        for (int e = 0; e < 8; ++e) w[e] = __float_as_uint(v[8 + e]);
        stg256(drow + col0 + 8, w);
      }
    }
    if (live) rpart[((size_t)t.tile_n * PARTS + part) * M + row] = racc;
  }
};

// Per-row constants of the backward epilogue: lse = exact log-sum-exp of the row (P_ij = exp(M_ij - lse)),
// r = row-dot, h = sum_j P log P (entropy term only).
struct __align__(16) RowConst { float lse, r, h, pad; };

__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Blackwell packed-fp32 arithmetic (fma/add/mul .f32x2): two elements per instruction on the FMA pipe.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// ---- split-precision operands -------------------------------------------------------------------
// Parity mode on tensor cores ("bf16x3"): every fp32 operand value x is stored as three bf16 planes
// hi + mid + lo (x reconstructed to ~2^-24), and a*b is accumulated in fp32 from the six partial products
// whose magnitude is >= 2^-24 relative: (l,h) (h,l) (m,m) (m,h) (h,m) (h,h), smallest first.  The kernel
// simply runs its k-loop once per pair into the same TMEM accumulator; with n_pairs == 1 only (h,h) runs
// (plain bf16 mode).
struct TcMaps { CUtensorMap m[3]; };
__device__ __constant__ int kPairA[6] = {2, 0, 1, 1, 0, 0};
__device__ __constant__ int kPairB[6] = {0, 2, 1, 0, 1, 0};

// ---- the kernel -------------------------------------------------------------------------------
// Work item w -> (split z, row tile, column tile), column tile fastest so that CTAs running at the
// same time share A rows and stream B through L2.
template <bool A_KMAJOR, bool B_KMAJOR, int BN, int STAGES, int EPI_WARPS, class Epi>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
k_gemm_tc(const __grid_constant__ TcMaps maps_a, const __grid_constant__ TcMaps maps_b, int n_pairs,
          int k_total, int k_per_split, int tiles_m, int tiles_n, int splits, int group_m, uint64_t policy_a,
          uint64_t policy_b, int tm_off, int k_off, const Epi epi) {
  using TileA = OperandTile<A_KMAJOR, TC_BM>;
  using TileB = OperandTile<B_KMAJOR, BN>;
  constexpr int kStageBytes = TileA::kBytes + TileB::kBytes;
  constexpr uint32_t kTmemCols = 2 * BN;     // two accumulator buffers (256 or 512 columns)
  static_assert(kTmemCols == 256 || kTmemCols == 512, "TMEM allocation must be a power of two <= 512");
  constexpr uint32_t kIdesc = make_idesc(TC_BM, BN, !A_KMAJOR, !B_KMAJOR);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = tiles_m * tiles_n * splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps_a.m[0]);
    tma_prefetch_desc(&maps_b.m[0]);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], EPI_WARPS); }
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
      uint32_t kbg = 0;                         // k-blocks issued so far (ring position)
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int z = w / (tiles_n * tiles_m);
        int tm_i, tn_i;
        tile_mn(w - z * tiles_n * tiles_m, tiles_m, tiles_n, group_m, tm_i, tn_i);
        const int n0 = tn_i * BN, m0 = (tm_off + tm_i) * TC_BM;
        const int k_begin = k_off + z * k_per_split;
        const int k_end = min(k_total, k_begin + k_per_split);
        const int num_kb = (k_end - k_begin + TC_BK - 1) / TC_BK;
        for (int pr = 6 - n_pairs; pr < 6; ++pr) {
          const CUtensorMap* ma = &maps_a.m[kPairA[pr]];
          const CUtensorMap* mb = &maps_b.m[kPairB[pr]];
          for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
            const int s = kbg % STAGES;
            const uint32_t ph = (kbg / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* sa = smem + s * kStageBytes;
            uint8_t* sb = sa + TileA::kBytes;
            mbar_expect_tx(&full_bar[s], kStageBytes);
            const int k0 = k_begin + kb * TC_BK;
            TileA::load(ma, &full_bar[s], sa, m0, k0, policy_a);
            TileB::load(mb, &full_bar[s], sb, n0, k0, policy_b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      uint32_t kbg = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int z = w / (tiles_n * tiles_m);
        const int k_begin = k_off + z * k_per_split;
        const int k_end = min(k_total, k_begin + k_per_split);
        const int num_kb = (k_end - k_begin + TC_BK - 1) / TC_BK;
        const int b = it & 1;
        mbar_wait(&tempty_bar[b], (((uint32_t)it >> 1) & 1) ^ 1);   // epilogue has drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(b * BN);
        const int total_kb = num_kb * n_pairs;
        for (int kb = 0; kb < total_kb; ++kb, ++kbg) {
          const int s = kbg % STAGES;
          const uint32_t ph = (kbg / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * kStageBytes);
          const uint32_t sb = sa + TileA::kBytes;
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
            umma_bf16(d_tmem, TileA::desc(sa, k), TileB::desc(sb, k), kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);           // frees the smem stage when these MMAs retire
        }
        umma_commit(&tfull_bar[b]);             // accumulator of this tile complete
      }
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> fused epilogue =====
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int ew = warp - 2;
    int it = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
      TileCoord t;
      t.split = w / (tiles_n * tiles_m);
      int tm_i;
      tile_mn(w - t.split * tiles_n * tiles_m, tiles_m, tiles_n, group_m, tm_i, t.tile_n);
      t.tiles_n = tiles_n;
      t.n0 = t.tile_n * BN;
      t.m0 = (tm_off + tm_i) * TC_BM;
      const int b = it & 1;
#ifndef TGB_SKIP_EPI
      epi.template prologue<BN, EPI_WARPS>(t, q, ew, lane);
#endif
      mbar_wait(&tfull_bar[b], ((uint32_t)it >> 1) & 1);
      tc_fence_after();
#ifndef TGB_SKIP_EPI
      epi.template run<BN, EPI_WARPS>(tmem_base + (uint32_t)(b * BN), q, ew, lane, t);
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[b]);
    }
    epi.finish(ew, lane, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- CTA-pair variant (tcgen05 cta_group::2) ----------------------------------------------------
// Two CTAs of a cluster (ranks 0/1, adjacent SMs) compute one 256 x BN tile: each stages ITS 128 rows of A and
// HALF of the B tile, rank 0's MMA thread issues M=256 instructions that read both CTAs' shared memory and write
// rows 0-127 / 128-255 of the tile into the two CTAs' TMEM.  B crosses L2->SM once per 256 output rows instead
// of once per 128, and a stage is 16 KB smaller per CTA -- room for a deeper epilogue staging pipeline.
// Protocol (cf. the cluster/2-SM notes in blackwell_cuda_programming.md):
//   full[s]   lives in rank 0: armed there with both CTAs' bytes; both producers' bulk loads complete_tx on it
//   empty[s]  one multicast tcgen05.commit per use arrives on BOTH CTAs' barriers (each producer waits locally)
//   tfull[b]  multicast commit -> each CTA's epilogue waits on its own copy
//   tempty[b] lives in rank 0: the epilogue warps of both CTAs arrive on it remotely
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// K-major A and B only (the backward contraction).  tiles_m counts 256-row pair tiles; grid = 2 x clusters.
template <int BN, int STAGES, int EPI_WARPS, class Epi>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
k_gemm_tc_pair(const __grid_constant__ TcMaps maps_a, const __grid_constant__ TcMaps maps_b, int n_pairs, int k_total, int tiles_m, int tiles_n, int group_m,
               uint64_t policy_a, uint64_t policy_b, int tm_off, const Epi epi) {
  using TileA = OperandTile<true, TC_BM>;         // this CTA's 128 rows
  using TileB = OperandTile<true, BN / 2>;        // this CTA's half of the B tile
  constexpr int kStageBytes = TileA::kBytes + TileB::kBytes;
  constexpr uint32_t kTmemCols = 2 * BN;
  static_assert(kTmemCols == 256 || kTmemCols == 512, "TMEM allocation must be a power of two <= 512");
  constexpr uint32_t kIdesc = make_idesc(2 * TC_BM, BN, false, false);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int total = tiles_m * tiles_n;
  const int num_kb = (k_total + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps_a.m[0]);
    tma_prefetch_desc(&maps_b.m[0]);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 2 * EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(&tmem_base_smem, kTmemCols);
  tc_fence_before();
  cluster_sync_all();                           // barriers of both CTAs initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A rows + own half of B, completing on rank 0's full barrier =====
#ifdef TGB_SKIP_MAINLOOP
    if (false) {
#else
    if (lane == 0) {
#endif
      uint32_t kbg = 0;
      for (int w = cid; w < total; w += ncl) {
        int tm_i, tn_i;
        tile_mn(w, tiles_m, tiles_n, group_m, tm_i, tn_i);
        const int n0 = tn_i * BN + (int)rank * (BN / 2);
        const int m0 = (tm_off + tm_i) * (2 * TC_BM) + (int)rank * TC_BM;
        for (int pr = 6 - n_pairs; pr < 6; ++pr) {
          const CUtensorMap* ma = &maps_a.m[kPairA[pr]];
          const CUtensorMap* mb = &maps_b.m[kPairB[pr]];
          for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
            const int s = kbg % STAGES;
            const uint32_t ph = (kbg / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            const uint32_t sa = smem_u32(smem + s * kStageBytes);
            const uint32_t sb = sa + TileA::kBytes;
            if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * kStageBytes);
            const uint32_t lbar = mapa_shared(smem_u32(&full_bar[s]), 0);
            tma_load_2d_pair(ma, lbar, sa, kb * TC_BK, m0, policy_a);
            tma_load_2d_pair(mb, lbar, sb, kb * TC_BK, n0, policy_b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of rank 0 drives both SMs' tensor cores =====
    if (lane == 0 && rank == 0) {
      uint32_t kbg = 0;
      int it = 0;
      for (int w = cid; w < total; w += ncl, ++it) {
        const int b = it & 1;
        mbar_wait(&tempty_bar[b], (((uint32_t)it >> 1) & 1) ^ 1);   // both epilogues have drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(b * BN);
#ifdef TGB_SKIP_MAINLOOP
        const int total_kb = 0;
#else
        const int total_kb = num_kb * n_pairs;
#endif
        for (int kb = 0; kb < total_kb; ++kb, ++kbg) {
          const int s = kbg % STAGES;
          const uint32_t ph = (kbg / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * kStageBytes);
          const uint32_t sb = sa + TileA::kBytes;
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
            umma_bf16_pair(d_tmem, TileA::desc(sa, k), TileB::desc(sb, k), kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(&empty_bar[s]);
        }
        umma_commit_pair(&tfull_bar[b]);
      }
    }
  } else {
    // ===== epilogue warps (both CTAs): each CTA owns 128 rows of the pair tile =====
    const int q = warp & 3;
    const int ew = warp - 2;
    const uint32_t tempty_remote = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    int it = 0;
    for (int w = cid; w < total; w += ncl, ++it) {
      TileCoord t;
      int tm_i;
      tile_mn(w, tiles_m, tiles_n, group_m, tm_i, t.tile_n);
      t.tiles_n = tiles_n;
      t.n0 = t.tile_n * BN;
      t.m0 = (tm_off + tm_i) * (2 * TC_BM) + (int)rank * TC_BM;
      t.split = 0;
      const int b = it & 1;
#ifndef TGB_SKIP_EPI
      epi.template prologue<BN, EPI_WARPS>(t, q, ew, lane);
#endif
      mbar_wait(&tfull_bar[b], ((uint32_t)it >> 1) & 1);
      tc_fence_after();
#ifndef TGB_SKIP_EPI
      epi.template run<BN, EPI_WARPS>(tmem_base + (uint32_t)(b * BN), q, ew, lane, t);
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(tempty_remote + (uint32_t)b * 8u);
    }
    epi.finish(ew, lane, 0);
  }
  tc_fence_before();
  cluster_sync_all();                           // the peer's shared memory and barriers stay valid until both are done
  if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcContext {
  PFN_encodeTiled encode = nullptr;
  int num_sms = 148;
  int pair_clusters = -1;   // co-resident 2-CTA clusters of the backward pair kernel (-1 = not queried yet, 0 = unavailable)
  int dp_clusters = -1;     // same for the store-only backward kernel
  const void* smem_fn[16] = {};   // kernels whose dynamic shared-memory limit this handle has already raised
  int smem_bytes[16] = {};
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
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&tc.num_sms, cudaDevAttrMultiProcessorCount, dev);
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

// cudaFuncSetAttribute once per (handle, kernel): it is a driver call, not something to repeat on every launch
template <class Kern>
static inline int tc_set_smem(TcContext& tc, Kern kern, int bytes, char* err, size_t n) {
  const void* key = reinterpret_cast<const void*>(kern);
  int slot = -1;
  for (int i = 0; i < 16; ++i) {
    if (tc.smem_fn[i] == key) { if (tc.smem_bytes[i] == bytes) return 0; slot = i; break; }
    if (tc.smem_fn[i] == nullptr && slot < 0) slot = i;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { snprintf(err, n, "cudaFuncSetAttribute(smem=%d): %s", bytes, cudaGetErrorString(e)); return -2; }
  if (slot >= 0) { tc.smem_fn[slot] = key; tc.smem_bytes[slot] = bytes; }
  return 0;
}
static inline int tc_check_launch(const char* name, char* err, size_t n) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(err, n, "launch %s: %s", name, cudaGetErrorString(e)); return -2; }
  return 0;
}

constexpr int TC_FWD_BN = 256, TC_FWD_STAGES = 4;
#ifndef TGB_BWD_BN
#define TGB_BWD_BN 256
#endif
constexpr int TC_BWD_BN = TGB_BWD_BN;
// CTA-pair backward kernel (cta_group::2): used from TC_PAIR_MIN_ROWS cells up (smaller problems don't fill 74 pairs)
#ifndef TGB_BWD_PAIR
#define TGB_BWD_PAIR 1
#endif
constexpr int TC_PAIR_MIN_ROWS = 2048;
// store-only backward (staged dP): no epilogue staging in shared memory, so the operand ring can be deeper
#ifndef TGB_DP_STAGES
#define TGB_DP_STAGES 6
#endif
#ifndef TGB_DP_EPI_WARPS
#define TGB_DP_EPI_WARPS 4
#endif
constexpr int TC_DP_STAGES = TGB_DP_STAGES, TC_DP_SINGLE_STAGES = 4, TC_DP_EPI_WARPS = TGB_DP_EPI_WARPS;
static inline int tc_dp_row_parts(int V) { return (int)ceil_div(V, TC_BWD_BN) * (TC_DP_EPI_WARPS / 4); }
// (tile_mn's row-tile groups stay at 1: more CTAs pulling the same B tile at once hot-spot L2 slices -- measured in round 1.)
static inline int tc_splits(long long tiles, long long k_total, int min_k) {
  long long s = (2 * 148 + tiles - 1) / tiles;
  const long long max_s = (k_total + min_k - 1) / min_k;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  // make every split non-empty
  const long long kps = ((k_total + s - 1) / s + TC_BK - 1) / TC_BK * TC_BK;
  return (int)((k_total + kps - 1) / kps);
}
// Parity mode: the tensor core accumulates in fp32 with truncation, a bias that grows with the length of the
// accumulation chain (measured on real data: -5.6e-6 relative at 6.6k-deep chains, -0.9e-6 at 1.6k).  Chains are
// therefore cut at `max_chain` contraction elements; the partial results are summed in fp32 (round-to-nearest) by
// the kernels that consume them.
static inline int tc_splits_for_chain(long long k_total, int max_chain) {
  long long s = (k_total + max_chain - 1) / max_chain;
  if (s < 1) s = 1;
  const long long kps = ((k_total + s - 1) / s + TC_BK - 1) / TC_BK * TC_BK;
  return (int)((k_total + kps - 1) / kps);
}
static inline int tc_forward_splits(int N, int V, int Ke) {
  return tc_splits((long long)ceil_div(V, TC_BM) * ceil_div(Ke, TC_FWD_BN), N, 512);
}
static inline int tc_kps(int k_total, int splits) {
  return (int)(round_up(ceil_div(k_total, splits), TC_BK));
}
static inline unsigned tc_grid(const TcContext& tc, long long total) {
  return (unsigned)(total < tc.num_sms ? total : tc.num_sms);
}

// Operand planes: plane p of an operand starts at base + p * plane_elems (bf16).  n_pairs = 1 uses plane 0 only.
static inline int tc_make_maps(TcContext& tc, TcMaps* maps, const __nv_bfloat16* base, size_t plane_elems, int n_planes,
                               uint64_t cols, uint64_t rows, uint64_t ld, uint32_t bi, uint32_t bo, char* err, size_t n) {
  for (int p = 0; p < 3; ++p) {
    const __nv_bfloat16* ptr = base + (size_t)(p < n_planes ? p : 0) * plane_elems;
    if (tc_make_map(tc, &maps->m[p], ptr, cols, rows, ld, bi, bo, err, n)) return -2;
  }
  return 0;
}

// The tensor maps of one contraction over fixed device buffers.  cuTensorMapEncodeTiled is a driver call: the handle
// encodes each plan once (the buffers never move) instead of on every launch.
struct TcPlan {
  TcMaps a, b;
  bool pair = false;        // b is encoded for the CTA-pair kernel (half-tile boxes)
  bool ready = false;
};

// Y_ext[z] (V x Ke) = P[cells of split z]^T S_ext[...]
static inline int tc_forward_plan(TcContext& tc, TcPlan& pl, const __nv_bfloat16* P, size_t p_plane, const __nv_bfloat16* Sx,
                                  size_t s_plane, int planes, int N, int V, int Ke, int ld, char* err, size_t n) {
  if (tc_make_maps(tc, &pl.a, P, p_plane, planes, V, N, ld, 64, 64, err, n)) return -2;      // A: MN-major (voxels contiguous), rows = cells
  if (tc_make_maps(tc, &pl.b, Sx, s_plane, planes, Ke, N, Ke, 64, 64, err, n)) return -2;    // B: MN-major (genes contiguous), rows = cells
  pl.ready = true;
  return 0;
}
static inline int tc_forward_launch(TcContext& tc, const TcPlan& pl, int n_pairs, float* out, int N, int V, int Ke, int splits,
                                    cudaStream_t s, char* err, size_t n) {
  auto kern = k_gemm_tc<false, false, TC_FWD_BN, TC_FWD_STAGES, 4, TcEpiStore>;
  const int smem = TC_FWD_STAGES * (TC_BM + TC_FWD_BN) * TC_BK * 2 + 1024;
  if (tc_set_smem(tc, kern, smem, err, n)) return -2;
  TcEpiStore epi{out, Ke, (size_t)V * Ke, V, 0};
  const int tm = (int)ceil_div(V, TC_BM), tn = (int)ceil_div(Ke, TC_FWD_BN);
  kern<<<tc_grid(tc, (long long)tm * tn * splits), 64 + 32 * 4, smem, s>>>(pl.a, pl.b, n_pairs, N, tc_kps(N, splits), tm, tn,
                                                                         splits, 1, kPolicyEvictNormal, kPolicyEvictNormal, 0, 0, epi);
  return tc_check_launch("tc_gemm_fwd", err, n);
}
// cells [row0, row1) only (row0 a multiple of 64): `out` = or += this chunk's partial sum -- the host pipelines cell chunks
// behind the streaming Adam kernel; the chunks run one after the other on one stream, so the summation order is fixed
static inline int tc_forward_launch_rows(TcContext& tc, const TcPlan& pl, float* out, int accumulate, int row0, int row1, int V, int Ke,
                                         cudaStream_t s, char* err, size_t n) {
  auto kern = k_gemm_tc<false, false, TC_FWD_BN, TC_FWD_STAGES, 4, TcEpiStore>;
  const int smem = TC_FWD_STAGES * (TC_BM + TC_FWD_BN) * TC_BK * 2 + 1024;
  if (tc_set_smem(tc, kern, smem, err, n)) return -2;
  TcEpiStore epi{out, Ke, (size_t)V * Ke, V, accumulate};
  const int tm = (int)ceil_div(V, TC_BM), tn = (int)ceil_div(Ke, TC_FWD_BN);
  kern<<<tc_grid(tc, (long long)tm * tn), 64 + 32 * 4, smem, s>>>(pl.a, pl.b, 1, row1, (int)round_up(row1 - row0, TC_BK), tm, tn,
                                                                1, 1, kPolicyEvictNormal, kPolicyEvictNormal, 0, row0, epi);
  return tc_check_launch("tc_gemm_fwd", err, n);
}
// one-shot variant for temporary operands (tgb200_project)
static inline int tc_forward(TcContext& tc, const __nv_bfloat16* P, size_t p_plane, const __nv_bfloat16* Sx, size_t s_plane,
                             int n_pairs, float* out, int N, int V, int Ke, int ld, int splits, cudaStream_t s, char* err,
                             size_t n) {
  TcPlan pl;
  if (tc_forward_plan(tc, pl, P, p_plane, Sx, s_plane, n_pairs > 1 ? 3 : 1, N, V, Ke, ld, err, n)) return -2;
  return tc_forward_launch(tc, pl, n_pairs, out, N, V, Ke, splits, s, err, n);
}

// number of co-resident 2-CTA clusters of a pair kernel (persistent: never launch more), 0 = unavailable
template <class Kern>
static inline int tc_pair_clusters(TcContext& tc, Kern pk, int threads, int smem, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cfg.gridDim = dim3(tc.num_sms & ~1);
  int mc = 0;
  const int r = cudaOccupancyMaxActiveClusters(&mc, pk, &cfg) == cudaSuccess ? mc : 0;
  (void)cudaGetLastError();
  return r;
}
template <class Kern, class... Args>
static inline cudaError_t tc_launch_pair(Kern pk, unsigned clusters, int threads, int smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cfg.gridDim = dim3(2 * clusters);
  return cudaLaunchKernelEx(&cfg, pk, args...);
}

// Staged backward: dq = bf16(S_ext dY_ext^T - centre) (bf16 mode) or dP in fp32 (bf16x3 mode, three operand planes, six
// partial products) to HBM + row-dot partials; the update itself is a streaming kernel (adam_rows.cuh).  Rows [row0, row1)
// only (row0 a multiple of 256): the host pipelines row chunks.
template <class Epi>
static inline int tc_dpstore_plan(TcContext& tc, TcPlan& pl, const __nv_bfloat16* Sxb, size_t s_plane, const __nv_bfloat16* dYb,
                                  size_t dy_plane, int planes, int N, int V, int Ke, cudaStream_t s, char* err, size_t n) {
  if (tc_make_maps(tc, &pl.a, Sxb, s_plane, planes, Ke, N, Ke, 64, TC_BM, err, n)) return -2;
  pl.pair = false;
#if TGB_BWD_PAIR
  if (N >= TC_PAIR_MIN_ROWS) {
    auto pk = k_gemm_tc_pair<TC_BWD_BN, TC_DP_STAGES, TC_DP_EPI_WARPS, Epi>;
    const int psmem = TC_DP_STAGES * (TC_BM + TC_BWD_BN / 2) * TC_BK * 2 + 1024;
    if (tc_set_smem(tc, pk, psmem, err, n)) return -2;
    if (tc.dp_clusters < 0) tc.dp_clusters = tc_pair_clusters(tc, pk, 64 + 32 * TC_DP_EPI_WARPS, psmem, s);
    pl.pair = tc.dp_clusters > 0;
  }
#endif
  if (tc_make_maps(tc, &pl.b, dYb, dy_plane, planes, Ke, V, Ke, 64, pl.pair ? TC_BWD_BN / 2 : TC_BWD_BN, err, n)) return -2;
  pl.ready = true;
  return 0;
}
template <class Epi>
static inline int tc_dpstore_launch(TcContext& tc, const TcPlan& pl, int n_pairs, const Epi& epi, int row0, int row1, int V, int Ke,
                                    cudaStream_t s, char* err, size_t n) {
  const int tn = (int)ceil_div(V, TC_BWD_BN);
  if (pl.pair) {
    auto pk = k_gemm_tc_pair<TC_BWD_BN, TC_DP_STAGES, TC_DP_EPI_WARPS, Epi>;
    const int psmem = TC_DP_STAGES * (TC_BM + TC_BWD_BN / 2) * TC_BK * 2 + 1024;
    const int tm0 = row0 / (2 * TC_BM), tmp = (int)ceil_div(row1, 2 * TC_BM) - tm0;
    const long long pair_tiles = (long long)tmp * tn;
    const unsigned clusters = (unsigned)(pair_tiles < tc.dp_clusters ? pair_tiles : tc.dp_clusters);
    cudaError_t e = tc_launch_pair(pk, clusters, 64 + 32 * TC_DP_EPI_WARPS, psmem, s, pl.a, pl.b, n_pairs, Ke, tmp, tn, 1,
                                   (uint64_t)kPolicyEvictNormal, (uint64_t)kPolicyEvictLast, tm0, epi);
    if (e != cudaSuccess) { snprintf(err, n, "launch tc_gemm_bwd_dp (pair): %s", cudaGetErrorString(e)); return -2; }
    return tc_check_launch("tc_gemm_bwd_dp", err, n);
  }
  auto kern = k_gemm_tc<true, true, TC_BWD_BN, TC_DP_SINGLE_STAGES, TC_DP_EPI_WARPS, Epi>;
  const int smem = TC_DP_SINGLE_STAGES * (TC_BM + TC_BWD_BN) * TC_BK * 2 + 1024;
  if (tc_set_smem(tc, kern, smem, err, n)) return -2;
  const int tm0 = row0 / TC_BM, tm = (int)ceil_div(row1, TC_BM) - tm0;
  kern<<<tc_grid(tc, (long long)tm * tn), 64 + 32 * TC_DP_EPI_WARPS, smem, s>>>(pl.a, pl.b, n_pairs, Ke, Ke, tm, tn, 1, 1,
                                                                                kPolicyEvictNormal, kPolicyEvictLast, tm0, 0, epi);
  return tc_check_launch("tc_gemm_bwd_dp", err, n);
}

}  // namespace tgb
