// attention.cu -- causal, key-padding-masked self-attention over packed rows (SASRec, L <= 64).
//
// Replaces F.multi_head_attention_forward + scaled_dot_product_attention as dispatched from
// torch.nn.TransformerEncoderLayer at reference model/sasrec.py:65-68:
//   S = Q_h K_h^T / sqrt(d_h) + M,  M[i,j] = -inf if j > i or in_id[b,j] == 0
//   A = dropout(softmax_j S);  O_h = A V_h
// A whole (sequence, head) problem is 50 x 50 x 64 -- it lives in one CTA's shared memory, so the
// score matrix never touches HBM and the backward recomputes the probabilities instead of storing
// them.  One CTA per (sequence, head); only the t < seqlen rows exist (packed layout).
#include "internal.cuh"

namespace dr4sr {
namespace {

constexpr int kAttnThreads = 128;

struct AttnSmem {
  float *Q, *K, *V, *S;
  int* pad;
  int ldk, lds;
};

__device__ __forceinline__ AttnSmem carve_fwd(float* sm, int L, int dh) {
  AttnSmem a;
  a.ldk = dh + 1; a.lds = L + 1;
  a.Q = sm; a.K = a.Q + L * dh; a.V = a.K + L * a.ldk; a.S = a.V + L * dh;
  a.pad = reinterpret_cast<int*>(a.S + L * a.lds);
  return a;
}
inline size_t fwd_smem_bytes(int L, int dh) { return sizeof(float) * (size_t)(L * dh + L * (dh + 1) + L * dh + L * (L + 1) + L); }

// loads the head slices of Q, K, V (and optionally dO) for one sequence; K is row-padded (+1) so the
// j-strided reads of the score loop are bank-conflict free
__device__ __forceinline__ void load_head(const float* __restrict__ qkv, int off, int len, int D, int h, int dh, float* Q,
                                          float* K, int ldk, float* V, int ldv) {
  const int f4_per_row = dh / 4;
  for (int e = threadIdx.x; e < len * f4_per_row; e += blockDim.x) {
    const int r = e / f4_per_row, c = (e % f4_per_row) * 4;
    const float* src = qkv + (size_t)(off + r) * 3 * D + h * dh + c;
    const float4 q = *reinterpret_cast<const float4*>(src);
    const float4 k = *reinterpret_cast<const float4*>(src + D);
    const float4 v = *reinterpret_cast<const float4*>(src + 2 * D);
    *reinterpret_cast<float4*>(Q + r * dh + c) = q;
    float* kd = K + r * ldk + c;
    kd[0] = k.x; kd[1] = k.y; kd[2] = k.z; kd[3] = k.w;
    float* vd = V + r * ldv + c;
    vd[0] = v.x; vd[1] = v.y; vd[2] = v.z; vd[3] = v.w;
  }
}

// S[i][j] = softmax_j(scale * <Q_i, K_j> + mask); rows owned by warps, two keys per lane (L <= 64)
__device__ __forceinline__ void scores_softmax(const AttnSmem& a, int len, int dh, float scale) {
  for (int e = threadIdx.x; e < len * len; e += blockDim.x) {
    const int i = e / len, j = e % len;
    float s = -INFINITY;
    if (j <= i && !a.pad[j]) {
      float acc = 0.f;
      const float* q = a.Q + i * dh;
      const float* k = a.K + j * a.ldk;
#pragma unroll 8
      for (int d = 0; d < dh; ++d) acc = fmaf(q[d], k[d], acc);
      s = acc * scale;
    }
    a.S[i * a.lds + j] = s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int i = warp; i < len; i += nwarp) {
    float* row = a.S + i * a.lds;
    const float s0 = lane < len ? row[lane] : -INFINITY;
    const float s1 = lane + 32 < len ? row[lane + 32] : -INFINITY;
    const float mx = warp_max(fmaxf(s0, s1));
    const float e0 = expf(s0 - mx), e1 = expf(s1 - mx);   // exp(-inf - mx) = 0 at masked keys
    const float inv = 1.0f / warp_sum(e0 + e1);
    if (lane < len) row[lane] = e0 * inv;
    if (lane + 32 < len) row[lane + 32] = e1 * inv;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kAttnThreads) attn_fwd_kernel(const float* __restrict__ qkv, const int64_t* __restrict__ in_ids,
                                                                const int32_t* __restrict__ tok_off, float* __restrict__ out,
                                                                int L, int D, int n_head, float scale, Dropout drop) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x / n_head, h = blockIdx.x % n_head;
  const int off = tok_off[b], len = min(tok_off[b + 1] - off, L);
  if (len <= 0) return;
  const int dh = D / n_head;
  AttnSmem a = carve_fwd(sm, L, dh);
  load_head(qkv, off, len, D, h, dh, a.Q, a.K, a.ldk, a.V, dh);
  for (int j = threadIdx.x; j < len; j += blockDim.x) a.pad[j] = in_ids[(size_t)b * L + j] == 0;
  __syncthreads();
  scores_softmax(a, len, dh, scale);
  const uint32_t base = (uint32_t)(b * n_head + h) * (uint32_t)(L * L);
  for (int e = threadIdx.x; e < len * dh; e += blockDim.x) {
    const int i = e / dh, d = e % dh;
    const float* p = a.S + i * a.lds;
    float acc = 0.f;
    for (int j = 0; j <= i; ++j) acc = fmaf(drop.apply(p[j], base + i * L + j), a.V[j * dh + d], acc);
    out[(size_t)(off + i) * D + h * dh + d] = acc;
  }
}

__host__ __device__ inline int align4(int x) { return (x + 3) & ~3; }
inline size_t bwd_smem_bytes(int L, int dh) {
  return sizeof(float) * (size_t)(L * dh + 2 * align4(L * (dh + 1)) + L * dh + 2 * align4(L * (L + 1)) + L);
}

// dQKV from dO, recomputing the probabilities (no [B,H,L,L] tensor is ever stored)
__global__ void __launch_bounds__(kAttnThreads) attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ d_out,
                                                                const int64_t* __restrict__ in_ids,
                                                                const int32_t* __restrict__ tok_off, float* __restrict__ d_qkv,
                                                                int L, int D, int n_head, float scale, Dropout drop) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x / n_head, h = blockIdx.x % n_head;
  const int off = tok_off[b], len = min(tok_off[b + 1] - off, L);
  if (len <= 0) return;
  const int dh = D / n_head;
  AttnSmem a;
  a.ldk = dh + 1; a.lds = L + 1;
  a.Q = sm; a.K = a.Q + L * dh; a.V = a.K + align4(L * a.ldk);   // V row-padded too (j-strided reads below)
  float* dO = a.V + align4(L * a.ldk);                           // [L][dh], 16-byte aligned
  a.S = dO + L * dh;                                             // P  [L][L+1]
  float* dS = a.S + align4(L * a.lds);                           // dS [L][L+1]
  a.pad = reinterpret_cast<int*>(dS + align4(L * a.lds));
  load_head(qkv, off, len, D, h, dh, a.Q, a.K, a.ldk, a.V, a.ldk);
  for (int e = threadIdx.x; e < len * (dh / 4); e += blockDim.x) {
    const int r = e / (dh / 4), c = (e % (dh / 4)) * 4;
    *reinterpret_cast<float4*>(dO + r * dh + c) =
        *reinterpret_cast<const float4*>(d_out + (size_t)(off + r) * D + h * dh + c);
  }
  for (int j = threadIdx.x; j < len; j += blockDim.x) a.pad[j] = in_ids[(size_t)b * L + j] == 0;
  __syncthreads();
  scores_softmax(a, len, dh, scale);
  const uint32_t base = (uint32_t)(b * n_head + h) * (uint32_t)(L * L);

  // dP[i][j] = factor * <dO_i, V_j>  (gradient w.r.t. the pre-dropout probability)
  for (int e = threadIdx.x; e < len * len; e += blockDim.x) {
    const int i = e / len, j = e % len;
    float v = 0.f;
    if (j <= i && !a.pad[j]) {
      float acc = 0.f;
      const float* g = dO + i * dh;
      const float* vv = a.V + j * a.ldk;
#pragma unroll 8
      for (int d = 0; d < dh; ++d) acc = fmaf(g[d], vv[d], acc);
      v = acc * drop.factor(base + i * L + j);
    }
    dS[i * a.lds + j] = v;
  }
  __syncthreads();
  // dS = P * (dP - sum_j dP P)
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int i = warp; i < len; i += nwarp) {
      float* g = dS + i * a.lds;
      const float* p = a.S + i * a.lds;
      const float g0 = lane < len ? g[lane] : 0.f, p0 = lane < len ? p[lane] : 0.f;
      const float g1 = lane + 32 < len ? g[lane + 32] : 0.f, p1 = lane + 32 < len ? p[lane + 32] : 0.f;
      const float dot = warp_sum(fmaf(g0, p0, g1 * p1));
      if (lane < len) g[lane] = p0 * (g0 - dot);
      if (lane + 32 < len) g[lane + 32] = p1 * (g1 - dot);
    }
  }
  __syncthreads();
  // dQ_i = scale sum_{j<=i} dS_ij K_j ; dK_j = scale sum_{i>=j} dS_ij Q_i ; dV_j = sum_{i>=j} Pdrop_ij dO_i
  for (int e = threadIdx.x; e < len * dh; e += blockDim.x) {
    const int r = e / dh, d = e % dh;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j <= r; ++j) dq = fmaf(dS[r * a.lds + j], a.K[j * a.ldk + d], dq);
    for (int i = r; i < len; ++i) {
      dk = fmaf(dS[i * a.lds + r], a.Q[i * dh + d], dk);
      dv = fmaf(drop.apply(a.S[i * a.lds + r], base + i * L + r), dO[i * dh + d], dv);
    }
    float* dst = d_qkv + (size_t)(off + r) * 3 * D + h * dh + d;
    dst[0] = dq * scale;
    dst[D] = dk * scale;
    dst[2 * D] = dv;
  }
}

}  // namespace

int launch_attn_fwd(const float* qkv, const int64_t* in_ids, const int32_t* tok_off, float* out, int B, int L, int D,
                    int n_head, Dropout drop, cudaStream_t st) {
  if (L > 64 || D % n_head || (D / n_head) % 4) return DR4SR_EINVAL;
  const int dh = D / n_head;
  const size_t smem = fwd_smem_bytes(L, dh);
  ProfScope prof("attn_fwd", st);
  static thread_local size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_cuda_error(cudaGetLastError(), "attn_fwd smem attribute");
      return DR4SR_ECUDA;
    }
    configured = smem;
  }
  attn_fwd_kernel<<<B * n_head, kAttnThreads, smem, st>>>(qkv, in_ids, tok_off, out, L, D, n_head, 1.0f / sqrtf((float)dh), drop);
  DR4SR_LAUNCH_CHECK("attn_fwd_kernel");
  return DR4SR_OK;
}

int launch_attn_bwd(const float* qkv, const float* d_out, const int64_t* in_ids, const int32_t* tok_off, float* d_qkv,
                    int B, int L, int D, int n_head, Dropout drop, cudaStream_t st) {
  if (L > 64 || D % n_head || (D / n_head) % 4) return DR4SR_EINVAL;
  const int dh = D / n_head;
  const size_t smem = bwd_smem_bytes(L, dh);
  ProfScope prof("attn_bwd", st);
  static thread_local size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_cuda_error(cudaGetLastError(), "attn_bwd smem attribute");
      return DR4SR_ECUDA;
    }
    configured = smem;
  }
  attn_bwd_kernel<<<B * n_head, kAttnThreads, smem, st>>>(qkv, d_out, in_ids, tok_off, d_qkv, L, D, n_head,
                                                          1.0f / sqrtf((float)dh), drop);
  DR4SR_LAUNCH_CHECK("attn_bwd_kernel");
  return DR4SR_OK;
}

}  // namespace dr4sr
