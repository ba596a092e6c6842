// common.cuh -- shared device/host helpers for libdr4sr (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/dr4sr.h"

namespace dr4sr {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- error plumbing (no exceptions cross the C ABI) -------------------------------------------
void set_cuda_error(cudaError_t e, const char* where);
void count_launch();
// Optional per-kernel timing (bench / profiling only): CUDA events on the launching stream around one
// launcher scope.  Disabled => two predictable branches, no events.
struct ProfScope {
  const char* name; cudaStream_t st; int slot;
  ProfScope(const char* name, cudaStream_t st);
  ~ProfScope();
};
#define DR4SR_LAUNCH_CHECK(where)                         \
  do {                                                    \
    ::dr4sr::count_launch();                              \
    cudaError_t e__ = cudaGetLastError();                 \
    if (e__ != cudaSuccess) {                             \
      ::dr4sr::set_cuda_error(e__, where);                \
      return DR4SR_ECUDA;                                 \
    }                                                     \
  } while (0)
#define DR4SR_TRY(expr)                                   \
  do {                                                    \
    int rc__ = (expr);                                    \
    if (rc__ != DR4SR_OK) return rc__;                    \
  } while (0)

inline cudaStream_t as_stream(dr4sr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- counter-based RNG --------------------------------------------------------------------------
// One 32-bit draw per (stream key, element index): murmur3-style finaliser over a Weyl-scrambled
// index.  Stateless, so the backward regenerates exactly the forward's dropout masks from
// (seed, step, site) without storing them.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ uint32_t stream_key(uint64_t seed, uint64_t step, uint32_t site) {
  uint32_t k = mix32((uint32_t)seed ^ 0x9E3779B9u);
  k = mix32(k ^ (uint32_t)(seed >> 32));
  k = mix32(k ^ (uint32_t)step);
  k = mix32(k ^ (uint32_t)(step >> 32) ^ (site * 0x632BE5ABu));
  return k;
}
__host__ __device__ __forceinline__ uint32_t draw32(uint32_t key, uint32_t idx) {
  return mix32(idx * 0x9E3779B1u + key) ^ mix32(key ^ (idx >> 7));
}
// dropout: one 32-bit draw serves two consecutive elements (16 bits each): element idx is kept iff its
// half-word of draw32(key, idx >> 1) is >= thresh16 = round(p * 65536)  (p == 0 -> thresh 0 keeps all;
// p is realised to within 2^-17, inverted scaling uses the nominal 1/(1-p)).
__host__ __device__ __forceinline__ uint32_t drop_thresh(float p) {
  double t = (double)p * 65536.0;
  return t >= 65535.0 ? 0xFFFFu : (uint32_t)(t + 0.5);
}
struct Dropout {
  uint32_t key, thresh;
  float scale;  // 1/(1-p)
  __device__ __forceinline__ bool keep(uint32_t idx) const {
    const uint32_t h = draw32(key, idx >> 1);
    return ((idx & 1u) ? (h >> 16) : (h & 0xFFFFu)) >= thresh;
  }
  __device__ __forceinline__ float apply(float x, uint32_t idx) const { return (thresh == 0u || keep(idx)) ? x * scale : 0.0f; }
  __device__ __forceinline__ float factor(uint32_t idx) const { return (thresh == 0u || keep(idx)) ? scale : 0.0f; }
  // factors of 4 consecutive elements, idx % 4 == 0 (two draws instead of four)
  __device__ __forceinline__ float4 factor4(uint32_t idx) const {
    if (thresh == 0u) return make_float4(scale, scale, scale, scale);
    const uint32_t h0 = draw32(key, idx >> 1), h1 = draw32(key, (idx >> 1) + 1u);
    return make_float4((h0 & 0xFFFFu) >= thresh ? scale : 0.f, (h0 >> 16) >= thresh ? scale : 0.f,
                       (h1 & 0xFFFFu) >= thresh ? scale : 0.f, (h1 >> 16) >= thresh ? scale : 0.f);
  }
};
inline Dropout make_dropout(float p, uint64_t seed, uint64_t step, uint32_t site, bool train) {
  Dropout d;
  d.key = stream_key(seed, step, site);
  d.thresh = (train && p > 0.f) ? drop_thresh(p) : 0u;
  d.scale = (train && p > 0.f) ? 1.0f / (1.0f - p) : 1.0f;
  return d;
}
// dropout sites of the SASRec step (distinct streams)
enum : uint32_t { SITE_EMBED = 1, SITE_ATTN_P = 16, SITE_ATTN_OUT = 17, SITE_FFN_H = 18, SITE_FFN_OUT = 19 };
inline uint32_t layer_site(uint32_t site, int layer) { return site + 8u * (uint32_t)layer; }

// ---- row-sharded table over peer memory (include/dr4sr.h, dr4sr_shard_map) ---------------------------------------------
// Kernel-side copy of the map; world == 1 is the single-GPU case (one compare-free path: lo[1] = N).
struct ShardView {
  const float* table[DR4SR_MAX_SHARDS];
  float* grad[DR4SR_MAX_SHARDS];
  long long lo[DR4SR_MAX_SHARDS + 1];
  int world, rank;
  // owner of a row id: ranges are few and sorted -> branch-free count of the boundaries at or below id
  __device__ __forceinline__ int owner(long long id) const {
    int o = 0;
#pragma unroll
    for (int r = 1; r < DR4SR_MAX_SHARDS; ++r) o += (r < world && id >= lo[r]) ? 1 : 0;
    return o;
  }
  __device__ __forceinline__ const float* row(long long id, int D) const {
    if (world == 1) return table[0] + (size_t)id * D;
    const int o = owner(id);
    return table[o] + (size_t)(id - lo[o]) * D;
  }
  __device__ __forceinline__ float* grad_row(long long id, int D) const {
    if (world == 1) return grad[0] + (size_t)id * D;
    const int o = owner(id);
    return grad[o] + (size_t)(id - lo[o]) * D;
  }
  __device__ __forceinline__ bool is_local(long long id) const { return world == 1 || (id >= lo[rank] && id < lo[rank + 1]); }
};
inline ShardView shard_view_local(const float* table, float* grad, long long N) {
  ShardView v{};
  v.table[0] = table; v.grad[0] = grad; v.lo[0] = 0; v.lo[1] = N; v.world = 1; v.rank = 0;
  return v;
}
inline bool shard_view_from(const dr4sr_shard_map* m, ShardView* out) {
  if (!m || m->world < 1 || m->world > DR4SR_MAX_SHARDS || m->rank < 0 || m->rank >= m->world) return false;
  ShardView v{};
  for (int r = 0; r < m->world; ++r) { v.table[r] = m->table[r]; v.grad[r] = m->grad[r]; v.lo[r] = m->lo[r]; }
  v.lo[m->world] = m->lo[m->world];
  v.world = m->world; v.rank = m->rank;
  *out = v;
  return true;
}

// ---- warp helpers -------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU and its derivative (torch 'gelu', approximate='none')
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// numerically stable log-sigmoid / softplus / sigmoid, same branches as ATen
__device__ __forceinline__ float log_sigmoid_f(float x) { return fminf(x, 0.0f) - log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace dr4sr
