// rowops.cu -- HBM-bound row kernels of the backward: LayerNorm backward (+ fused bias/affine column
// partials), column sums, and the deterministic partial reducer.  One warp per packed row, 16-byte
// lanes, grid = 2 x 148 CTAs, grid-stride over the live rows only.
#include "internal.cuh"

namespace dr4sr {
namespace {

// Rows are D <= 256 floats; lane l owns columns {4l..4l+3} (+128 for the second chunk).
template <int NCHUNK>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                     const float* __restrict__ stats, const float* __restrict__ gamma,
                                                     float* __restrict__ dz, float* __restrict__ partials, int D, int T_cap,
                                                     const int32_t* __restrict__ tok_dev, Dropout bias_drop, Dropout dy_drop) {
  const int T = min(T_cap, tok_dev ? *tok_dev : T_cap);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float invD = 1.0f / (float)D;
  float4 g4[NCHUNK], acc_g[NCHUNK], acc_b[NCHUNK], acc_bias[NCHUNK];
  bool on[NCHUNK];
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    const int col = c * 128 + lane * 4;
    on[c] = col < D;
    g4[c] = on[c] ? *reinterpret_cast<const float4*>(gamma + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc_g[c] = acc_b[c] = acc_bias[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = blockIdx.x * 8 + warp; row < T; row += gridDim.x * 8) {
    const float mu = stats[2 * row], rstd = stats[2 * row + 1];
    float4 g[NCHUNK], xh[NCHUNK];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      const int col = c * 128 + lane * 4;
      if (on[c]) {
        g[c] = *reinterpret_cast<const float4*>(dy + (size_t)row * D + col);
        if (dy_drop.thresh) {   // the normalised output went through a dropout (FMLP embedding stage)
          const float4 f = dy_drop.factor4((uint32_t)row * (uint32_t)D + col);
          g[c].x *= f.x; g[c].y *= f.y; g[c].z *= f.z; g[c].w *= f.w;
        }
        const float4 zz = *reinterpret_cast<const float4*>(z + (size_t)row * D + col);
        xh[c] = make_float4((zz.x - mu) * rstd, (zz.y - mu) * rstd, (zz.z - mu) * rstd, (zz.w - mu) * rstd);
        acc_g[c].x += g[c].x * xh[c].x; acc_g[c].y += g[c].y * xh[c].y; acc_g[c].z += g[c].z * xh[c].z; acc_g[c].w += g[c].w * xh[c].w;
        acc_b[c].x += g[c].x; acc_b[c].y += g[c].y; acc_b[c].z += g[c].z; acc_b[c].w += g[c].w;
        g[c].x *= g4[c].x; g[c].y *= g4[c].y; g[c].z *= g4[c].z; g[c].w *= g4[c].w;   // dxhat
        s1 += (g[c].x + g[c].y) + (g[c].z + g[c].w);
        s2 += (g[c].x * xh[c].x + g[c].y * xh[c].y) + (g[c].z * xh[c].z + g[c].w * xh[c].w);
      }
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      const int col = c * 128 + lane * 4;
      if (on[c]) {
        float4 o;
        o.x = rstd * (g[c].x - s1 - xh[c].x * s2); o.y = rstd * (g[c].y - s1 - xh[c].y * s2);
        o.z = rstd * (g[c].z - s1 - xh[c].z * s2); o.w = rstd * (g[c].w - s1 - xh[c].w * s2);
        if (dz) *reinterpret_cast<float4*>(dz + (size_t)row * D + col) = o;   // null: column partials only
        const uint32_t idx = (uint32_t)row * (uint32_t)D + col;
        const float4 f = bias_drop.factor4(idx);
        acc_bias[c].x += o.x * f.x; acc_bias[c].y += o.y * f.y; acc_bias[c].z += o.z * f.z; acc_bias[c].w += o.w * f.w;
      }
    }
  }
  // block reduce the three column accumulators over the 8 warps (fixed order -> deterministic)
  __shared__ float red[8][3][256];
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    const int col = c * 128 + lane * 4;
    if (on[c]) {
      *reinterpret_cast<float4*>(&red[warp][0][col]) = acc_g[c];
      *reinterpret_cast<float4*>(&red[warp][1][col]) = acc_b[c];
      *reinterpret_cast<float4*>(&red[warp][2][col]) = acc_bias[c];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 3 * D; e += blockDim.x) {
    const int k = e / D, col = e % D;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][k][col];
    partials[(size_t)blockIdx.x * 3 * D + e] = s;
  }
}

// y = dropout(LayerNorm(z)) row-wise, stats[row] = {mean, rstd}; warp per row, two-pass moments in registers
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, float* __restrict__ y,
                                                     float* __restrict__ stats, int D, int T_cap, const int32_t* __restrict__ tok_dev,
                                                     Dropout out_drop) {
  const int T = min(T_cap, tok_dev ? *tok_dev : T_cap);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float invD = 1.0f / (float)D;
  for (int row = blockIdx.x * 8 + warp; row < T; row += gridDim.x * 8) {
    float4 v[2];
    float s = 0.f;
    int k = 0;
    for (int c = lane * 4; c < D; c += 128, ++k) {
      v[k] = *reinterpret_cast<const float4*>(z + (size_t)row * D + c);
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const float mu = warp_sum(s) * invD;
    float q = 0.f;
    k = 0;
    for (int c = lane * 4; c < D; c += 128, ++k) {
      const float a = v[k].x - mu, b = v[k].y - mu, cc = v[k].z - mu, d = v[k].w - mu;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * invD + eps);
    k = 0;
    for (int c = lane * 4; c < D; c += 128, ++k) {
      const float4 g = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
      const float4 f = out_drop.factor4((uint32_t)row * (uint32_t)D + c);
      float4 o;
      o.x = ((v[k].x - mu) * rstd * g.x + be.x) * f.x; o.y = ((v[k].y - mu) * rstd * g.y + be.y) * f.y;
      o.z = ((v[k].z - mu) * rstd * g.z + be.z) * f.z; o.w = ((v[k].w - mu) * rstd * g.w + be.w) * f.w;
      *reinterpret_cast<float4*>(y + (size_t)row * D + c) = o;
    }
    if (lane == 0) { stats[2 * row] = mu; stats[2 * row + 1] = rstd; }
  }
}

__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int N, int T_cap,
                                                     const int32_t* __restrict__ tok_dev, float* __restrict__ partials) {
  const int T = min(T_cap, tok_dev ? *tok_dev : T_cap);
  const int groups = N / 4;                 // float4 column groups
  const int lanes_r = 256 / groups;         // row lanes per block (>= 1 because N <= 1024)
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < lanes_r) {
    const int step = gridDim.x * lanes_r;
    int row = blockIdx.x * lanes_r + rl;
    for (; row + 3 * step < T; row += 4 * step) {          // four independent loads in flight; fixed summation order
      const float4 v0 = *reinterpret_cast<const float4*>(x + (size_t)row * N + cg * 4);
      const float4 v1 = *reinterpret_cast<const float4*>(x + (size_t)(row + step) * N + cg * 4);
      const float4 v2 = *reinterpret_cast<const float4*>(x + (size_t)(row + 2 * step) * N + cg * 4);
      const float4 v3 = *reinterpret_cast<const float4*>(x + (size_t)(row + 3 * step) * N + cg * 4);
      acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
      acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; row < T; row += step) {
      const float4 v = *reinterpret_cast<const float4*>(x + (size_t)row * N + cg * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  __shared__ __align__(16) float red[256 * 4];
  *reinterpret_cast<float4*>(&red[threadIdx.x * 4]) = acc;
  __syncthreads();
  for (int e = threadIdx.x; e < N; e += blockDim.x) {
    const int g = e / 4, k = e % 4;
    float s = 0.f;
    for (int r = 0; r < lanes_r; ++r) s += red[(r * groups + g) * 4 + k];
    partials[(size_t)blockIdx.x * N + e] = s;
  }
}

// out[e] = sum_k src[k][e] in a fixed order.  A CTA owns 32 consecutive outputs; its 8 warps each sum
// an interleaved eighth of the splits (coalesced 128-byte rows, 8 independent loads in flight per lane),
// then the eight partial sums are added in warp order -> deterministic and latency-tolerant even for
// the 296-way LayerNorm / bias partials.
__global__ void __launch_bounds__(256) reduce_segments_kernel(ReduceTable tab) {
  const ReduceSeg s = tab.seg[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float part[8][32];
  for (int e0 = blockIdx.x * 32; e0 < s.n; e0 += gridDim.x * 32) {
    const int e = e0 + lane;
    float acc = 0.f;
    if (e < s.n) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int k = warp;
      for (; k + 24 < s.n_split; k += 32) {
        a0 += s.src[(size_t)k * s.stride + e];
        a1 += s.src[(size_t)(k + 8) * s.stride + e];
        a2 += s.src[(size_t)(k + 16) * s.stride + e];
        a3 += s.src[(size_t)(k + 24) * s.stride + e];
      }
      for (; k < s.n_split; k += 8) a0 += s.src[(size_t)k * s.stride + e];
      acc = (a0 + a1) + (a2 + a3);
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && e < s.n) {
      float t = part[0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) t += part[w][lane];
      s.dst[e] = t;
    }
    __syncthreads();
  }
}

}  // namespace

int launch_ln_bwd(const float* dy, const float* z, const float* stats, const float* gamma, float* dz, float* partials,
                  int D, int T_cap, const int32_t* tok_dev, Dropout bias_drop, cudaStream_t st, Dropout dy_drop) {
  if (D % 4 || D > 256) return DR4SR_EINVAL;
  ProfScope prof("ln_bwd", st);
  if (D <= 128)
    ln_bwd_kernel<1><<<kLnBwdBlocks, 256, 0, st>>>(dy, z, stats, gamma, dz, partials, D, T_cap, tok_dev, bias_drop, dy_drop);
  else
    ln_bwd_kernel<2><<<kLnBwdBlocks, 256, 0, st>>>(dy, z, stats, gamma, dz, partials, D, T_cap, tok_dev, bias_drop, dy_drop);
  DR4SR_LAUNCH_CHECK("ln_bwd_kernel");
  return DR4SR_OK;
}

int launch_ln_fwd(const float* z, const float* gamma, const float* beta, float eps, float* y, float* stats, int D, int T_cap,
                  const int32_t* tok_dev, Dropout out_drop, cudaStream_t st) {
  if (D % 4 || D > 256) return DR4SR_EINVAL;
  ProfScope prof("ln_fwd", st);
  const int blocks = ceil_div(T_cap, 8) < 8 * kNumSMs ? ceil_div(T_cap, 8) : 8 * kNumSMs;
  ln_fwd_kernel<<<blocks, 256, 0, st>>>(z, gamma, beta, eps, y, stats, D, T_cap, tok_dev, out_drop);
  DR4SR_LAUNCH_CHECK("ln_fwd_kernel");
  return DR4SR_OK;
}

int launch_colsum(const float* x, int N, int T_cap, const int32_t* tok_dev, float* partials, cudaStream_t st) {
  if (N % 4 || N > 1024) return DR4SR_EINVAL;
  ProfScope prof("colsum", st);
  colsum_kernel<<<kColsumBlocks, 256, 0, st>>>(x, N, T_cap, tok_dev, partials);
  DR4SR_LAUNCH_CHECK("colsum_kernel");
  return DR4SR_OK;
}

int launch_reduce_segments(const ReduceTable& tab, cudaStream_t st) {
  if (tab.count <= 0) return DR4SR_OK;
  if (tab.count > kMaxSeg) return DR4SR_EINVAL;
  ProfScope prof("reduce_partials", st);
  int nmax = 0;
  for (int i = 0; i < tab.count; ++i) nmax = nmax > tab.seg[i].n ? nmax : tab.seg[i].n;
  dim3 grid(ceil_div(nmax, 32) < 2 * kNumSMs ? ceil_div(nmax, 32) : 2 * kNumSMs, tab.count);
  reduce_segments_kernel<<<grid, 256, 0, st>>>(tab);
  DR4SR_LAUNCH_CHECK("reduce_segments_kernel");
  return DR4SR_OK;
}

}  // namespace dr4sr
