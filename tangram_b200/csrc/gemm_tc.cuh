// tcgen05 / TMA tensor-core contraction path (bf16 operands, fp32 accumulate in TMEM).
// STUB: interface only -- replaced by the real kernels in the next milestone.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"

namespace tgb {

constexpr int TC_RDOT_BN = 128;
struct TcContext { int dummy = 0; };
struct TcAdamArgs {
  float* Mp; float* mp; float* vp; int ld; int V;
  const RowStat* stats; const float* rdot;
  float lam_r, lam_l1, lam_l2;
  AdamScalars a;
};
static inline int tc_unsupported(char* err, size_t n) {
  snprintf(err, n, "precision=bf16 (tcgen05 path) is not built yet");
  return -4;
}
static inline int tc_init(TcContext&, char* err, size_t n) { return tc_unsupported(err, n); }
static inline int tc_forward_splits(int, int, int) { return 1; }
static inline int tc_rowdot_splits(int, int, int) { return 1; }
static inline int tc_forward(TcContext&, const __nv_bfloat16*, const __nv_bfloat16*, float*, int, int, int, int, int,
                             cudaStream_t, char* err, size_t n) { return tc_unsupported(err, n); }
static inline int tc_rowdot(TcContext&, const __nv_bfloat16*, const __nv_bfloat16*, const float*, float*, int, int, int,
                            int, cudaStream_t, char* err, size_t n) { return tc_unsupported(err, n); }
static inline int tc_backward(TcContext&, const __nv_bfloat16*, const __nv_bfloat16*, const TcAdamArgs&, int, int, int,
                              cudaStream_t, char* err, size_t n) { return tc_unsupported(err, n); }

}  // namespace tgb
