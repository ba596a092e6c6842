// dense.cuh -- dense-layer dispatch shared by the encoders: tcgen05 (bf16 hi/lo split) where the shape
// allows, exact-fp32 FFMA otherwise.  Also the backend switch used by the parity tests.
#pragma once
#include <atomic>
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "internal.cuh"

namespace dr4sr {

extern std::atomic<int> g_attn_backend;   // 0 = FFMA attention kernels, 1 = tcgen05 window tiles (fwd + bwd), 2 (default) = persistent tcgen05 backward on the greedy tiles (dr4sr_set_attn_backend)
extern std::atomic<int> g_gemm_backend;   // 0 = tcgen05 where the shape allows, 1 = FFMA everywhere (dr4sr_set_gemm_backend)
extern std::atomic<int> g_fused_backend;  // 0 = per-op kernels, 1 = persistent fused forward where the shape allows, 2 (default) = + fused backward of each layer's position-wise half
constexpr int kSplit = 64;                // token splits of the weight-gradient GEMMs (partials reduced in fixed order)

struct Img { uint16_t *hi, *lo; };        // bf16 hi / lo weight images (UMMA SW128 K-major), see gemm_tc.cuh
inline bool tc_enabled() { return g_gemm_backend.load(std::memory_order_relaxed) == 0; }
inline bool fused_enabled() { return tc_enabled() && g_fused_backend.load(std::memory_order_relaxed) >= 1; }
inline bool fused_bwd_enabled() { return tc_enabled() && g_fused_backend.load(std::memory_order_relaxed) == 2; }
inline bool attn_tc_enabled() { return tc_enabled() && g_attn_backend.load(std::memory_order_relaxed) == 1; }
inline bool attn_bwd_tc2_enabled() { return tc_enabled() && g_attn_backend.load(std::memory_order_relaxed) == 2; }

inline bool use_tc(const GemmArgs& g, const Img& im, bool ln) {
  return tc_enabled() && im.hi && tc::tc_supported(g.N, g.K, ln);
}
// y = LN(drop(A W^T + b) + res), rows complete inside a CTA (BN == D)
inline int gemm_ln(GemmArgs& g, int D, const Img& im, cudaStream_t st) {
  if (use_tc(g, im, true)) return tc::launch_gemm_tc<tc::TC_LN>(g, im.hi, im.lo, st);
  if (D == 128) return launch_gemm<64, 128, true, true, true>(g, st);
  return launch_gemm<64, 64, true, true, true>(g, st);
}
inline int gemm_nt(GemmArgs& g, const Img& im, cudaStream_t st) {
  if (use_tc(g, im, false)) return tc::launch_gemm_tc<tc::TC_LINEAR>(g, im.hi, im.lo, st);
  if (g.N >= 256) return launch_gemm<128, 128, true, true, false>(g, st);
  if (g.N > 64) return launch_gemm<64, 128, true, true, false>(g, st);
  return launch_gemm<64, 64, true, true, false>(g, st);
}
// backward-data: C = A W, the tensor-core path consumes the transposed image of W
inline int gemm_nn(GemmArgs& g, const Img& im, cudaStream_t st) {
  if (use_tc(g, im, false)) {
    return g.epi == EPI_GELU_BWD ? tc::launch_gemm_tc<tc::TC_GELU_BWD>(g, im.hi, im.lo, st)
                                 : tc::launch_gemm_tc<tc::TC_LINEAR>(g, im.hi, im.lo, st);
  }
  if (g.N >= 256) return launch_gemm<128, 128, true, false, false>(g, st);
  if (g.N > 64) return launch_gemm<64, 128, true, false, false>(g, st);
  return launch_gemm<64, 64, true, false, false>(g, st);
}
inline int gemm_tn(GemmArgs& g, float* partial, cudaStream_t st) {   // C partials [kSplit][M*N]
  g.C = partial; g.n_split = kSplit; g.split_stride = (int64_t)g.M * g.N; g.ldc = g.N;
  return launch_gemm<64, 64, false, false, false>(g, st);
}


inline Dropout no_dropout() { Dropout d; d.key = 0; d.thresh = 0; d.scale = 1.f; return d; }
inline size_t ws_align(size_t floats) { return (floats + 63) & ~(size_t)63; }   // 256-byte granules

}  // namespace dr4sr
