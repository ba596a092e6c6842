// batch.cu -- batch preparation (packed-token index), negative sampling, fused embedding stage.
#include "internal.cuh"

namespace dr4sr {
namespace {

// tok_off = exclusive scan of clamp(seqlen, 0, L); row_seq[row] = b.  One CTA (B is a few thousand).
// counts = {live tokens, non-pad targets among item_id[0, n_tgt) (0 when item_id is null), 0, 0}.
__global__ void __launch_bounds__(1024) prep_scan_kernel(const int64_t* __restrict__ seqlen, int B, int L,
                                                         int32_t* __restrict__ tok_off, int32_t* __restrict__ row_seq,
                                                         int32_t* __restrict__ counts, const int64_t* __restrict__ item_id, int64_t n_tgt) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  __shared__ int s_before[1024], s_len[1024];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < B; base += blockDim.x) {
    const int b = base + tid;
    int len = 0;
    if (b < B) {
      const int64_t s = seqlen[b];
      len = s < 0 ? 0 : (s > L ? L : (int)s);
    }
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      warp_tot[lane] = w;   // inclusive totals of warps
    }
    __syncthreads();
    const int before = carry + (warp ? warp_tot[warp - 1] : 0) + incl - len;
    if (b < B) tok_off[b] = before;
    s_before[tid] = before; s_len[tid] = len;
    __syncthreads();
    // row_seq: one warp per sequence, lanes = positions (coalesced stores instead of a serial loop per thread)
    for (int i = warp; i < (int)blockDim.x && base + i < B; i += (int)(blockDim.x >> 5)) {
      const int o = s_before[i], n = s_len[i];
      for (int t = lane; t < n; t += 32) row_seq[o + t] = base + i;
    }
    __syncthreads();
    if (tid == blockDim.x - 1) carry = before + len;
    __syncthreads();
  }
  int c = 0;                                   // valid targets: same CTA, no second launch (integer sum: order-independent)
  if (item_id)
    for (int64_t i = tid; i < n_tgt; i += blockDim.x) c += item_id[i] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __syncthreads();
  if (lane == 0) warp_tot[warp] = c;
  __syncthreads();
  if (tid == 0) {
    int tot = 0;
    for (int w = 0; w < 32; ++w) tot += warp_tot[w];
    tok_off[B] = carry;
    counts[0] = carry;
    counts[1] = tot; counts[2] = 0; counts[3] = 0;
  }
}

__global__ void __launch_bounds__(256) neg_sample_kernel(int64_t* __restrict__ out, int64_t n, uint32_t range, uint32_t key) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t r = draw32(key, (uint32_t)i) ^ mix32((uint32_t)(i >> 32) + 0x7F4A7C15u);
    out[i] = 1 + (int64_t)__umulhi(r, range);   // uniform on {1..N-1}: floor(r * (N-1) / 2^32)
  }
}

// x0[row] = dropout(E[in_id] + P[t]); one warp per packed row, 16-byte lanes.
__global__ void __launch_bounds__(256) embed_fwd_kernel(const float* __restrict__ table, const float* __restrict__ pos,
                                                        const int64_t* __restrict__ in_ids, const int32_t* __restrict__ tok_off,
                                                        const int32_t* __restrict__ row_seq, const int32_t* __restrict__ counts,
                                                        int L, int D, Dropout drop, float* __restrict__ x0) {
  const int T = counts[0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + warp; row < T; row += gridDim.x * 8) {
    const int b = row_seq[row];
    const int t = row - tok_off[b];
    const int64_t id = in_ids[(size_t)b * L + t];
    const float* e = table + (size_t)id * D;
    const float* p = pos ? pos + (size_t)t * D : nullptr;
    for (int c = lane * 4; c < D; c += 128) {
      float4 v = *reinterpret_cast<const float4*>(e + c);
      if (p) {
        const float4 q = *reinterpret_cast<const float4*>(p + c);
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      const uint32_t idx = (uint32_t)row * (uint32_t)D + c;
      const float4 f = drop.factor4(idx);
      v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
      *reinterpret_cast<float4*>(x0 + (size_t)row * D + c) = v;
    }
  }
}

// out[r, :] = src[idx[r], :] for int64 rows of `width` elements (batch assembly from device-resident columns)
__global__ void __launch_bounds__(256) gather_i64_kernel(const int64_t* __restrict__ src, int width, const int64_t* __restrict__ idx,
                                                         int64_t m, int64_t* __restrict__ out) {
  const int64_t total = m * width;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / width;
    out[e] = src[idx[r] * width + (e - r * width)];
  }
}

}  // namespace
}  // namespace dr4sr

using namespace dr4sr;

extern "C" int dr4sr_gather_i64(const int64_t* src, int32_t width, const int64_t* idx, int64_t m, int64_t* out, dr4sr_stream_t stream) {
  if (!src || !idx || !out || width <= 0 || m < 0) return DR4SR_EINVAL;
  if (m == 0) return DR4SR_OK;
  const int64_t total = m * width;
  const int blocks = ceil_div(total, 256) < 4 * kNumSMs ? ceil_div(total, 256) : 4 * kNumSMs;
  ProfScope prof("gather_batch", as_stream(stream));
  gather_i64_kernel<<<blocks, 256, 0, as_stream(stream)>>>(src, width, idx, m, out);
  DR4SR_LAUNCH_CHECK("gather_i64_kernel");
  return DR4SR_OK;
}

extern "C" int dr4sr_prep_batch(const int64_t* seqlen, const int64_t* item_id, int32_t B, int32_t L, int32_t target_is_1d,
                                int32_t* tok_off, int32_t* row_seq, int32_t* counts, dr4sr_stream_t stream) {
  if (!seqlen || !tok_off || !row_seq || !counts || B <= 0 || L <= 0) return DR4SR_EINVAL;
  cudaStream_t st = as_stream(stream);
  ProfScope prof("prep_batch", st);
  prep_scan_kernel<<<1, 1024, 0, st>>>(seqlen, B, L, tok_off, row_seq, counts, item_id, target_is_1d ? (int64_t)B : (int64_t)B * L);
  DR4SR_LAUNCH_CHECK("prep_scan_kernel");
  return DR4SR_OK;
}

extern "C" int dr4sr_neg_sample(int64_t* out, int64_t n, int64_t num_items, uint64_t seed, uint64_t step,
                                dr4sr_stream_t stream) {
  if (!out || n < 0 || num_items < 2 || num_items > 0xFFFFFFFFll) return DR4SR_EINVAL;
  if (n == 0) return DR4SR_OK;
  const int blocks = ceil_div(n, 256 * 4) < 4 * kNumSMs ? ceil_div(n, 256 * 4) : 4 * kNumSMs;
  ProfScope prof("neg_sample", as_stream(stream));
  neg_sample_kernel<<<blocks, 256, 0, as_stream(stream)>>>(out, n, (uint32_t)(num_items - 1), stream_key(seed, step, 0xA11CEu));
  DR4SR_LAUNCH_CHECK("neg_sample_kernel");
  return DR4SR_OK;
}

extern "C" int dr4sr_embed_fwd(const float* table, const float* pos, const int64_t* in_item_id, const int32_t* tok_off,
                               const int32_t* row_seq, const int32_t* counts, int32_t B, int32_t L, int32_t D,
                               float dropout_p, uint64_t seed, uint64_t step, float* x0_packed, dr4sr_stream_t stream) {
  if (!table || !in_item_id || !tok_off || !row_seq || !counts || !x0_packed || D % 4) return DR4SR_EINVAL;
  const Dropout drop = make_dropout(dropout_p, seed, step, SITE_EMBED, dropout_p > 0.f);
  const int T_cap = B * L;
  const int blocks = ceil_div(T_cap, 8) < 4 * kNumSMs ? ceil_div(T_cap, 8) : 4 * kNumSMs;
  ProfScope prof("embed_fwd", as_stream(stream));
  embed_fwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(table, pos, in_item_id, tok_off, row_seq, counts, L, D, drop, x0_packed);
  DR4SR_LAUNCH_CHECK("embed_fwd_kernel");
  return DR4SR_OK;
}
