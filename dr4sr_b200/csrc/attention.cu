// attention.cu -- causal, key-padding-masked self-attention over packed rows (SASRec, L <= 64).
//
// Replaces F.multi_head_attention_forward + scaled_dot_product_attention as dispatched from
// torch.nn.TransformerEncoderLayer at reference model/sasrec.py:65-68:
//   S = Q_h K_h^T / sqrt(d_h) + M,  M[i,j] = -inf if j > i or in_id[b,j] == 0
//   A = dropout(softmax_j S);  O_h = A V_h
// A whole (sequence, head) problem is at most 64 x 64 x 64: it lives in one CTA's shared memory, the
// score matrix never touches HBM, and the backward recomputes the probabilities instead of storing
// them.  One CTA (256 threads) per (sequence, head); only the t < seqlen rows exist (packed layout).
//
// Register tiling: the two matmul shapes of the problem are
//   abt : C[i][j] = sum_d A[i][d] B[j][d]      -- 4x4 outputs per thread, rows strided by 16 so that the
//         float4 reads of 8 neighbouring lanes fall in 8 different 16-byte bank groups (ld % 32 == 4)
//   pv  : C[r][d] = sum_x S(r,x) M[x][d]       -- 4 rows x one float4 of d per thread
// i.e. 64 (resp. 16) FMAs per 8 (resp. 5) shared loads instead of 1 FMA per 2 loads.
#include "internal.cuh"

namespace dr4sr {
namespace {

constexpr int kAttnThreads = 256;

__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }

struct Tile {            // shared-memory geometry of one (sequence, head) problem
  int rows, ld, lds;     // rows = L rounded up to 16; ld = dh + 4; lds = rows rounded up to 32, + 4 (lds % 32 == 4)
};
__host__ __device__ inline Tile make_tile(int L, int dh) {
  Tile t;
  t.rows = (L + 15) & ~15;
  t.ld = dh + 4;
  t.lds = ((t.rows + 31) / 32) * 32 + 4;
  return t;
}

// Asynchronous 16-byte global -> shared copies (LDGSTS): every thread queues all of its copies for the
// whole problem back to back and waits once, so a CTA pays one memory round trip instead of one per row.
__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// rows [0,len) of one head slice of a packed [T, stride] matrix -> smem [rows][ld]
__device__ __forceinline__ void load_rows(const float* __restrict__ src, int stride, int len, int dh, float* dst, int ld) {
  const int f4 = dh / 4;
  for (int e = threadIdx.x; e < len * f4; e += blockDim.x) {
    const int r = e / f4, c = (e % f4) * 4;
    cp_async16(dst + r * ld + c, src + (size_t)r * stride + c);
  }
}

// Q, K, V head slices of the packed [T, 3D] projection
__device__ __forceinline__ void load_qkv(const float* __restrict__ src, int D, int len, int dh, float* Q, float* K, float* V, int ld) {
  const int f4 = dh / 4;
  for (int e = threadIdx.x; e < len * f4; e += blockDim.x) {
    const int r = e / f4, c = (e % f4) * 4;
    const float* p = src + (size_t)r * 3 * D + c;
    cp_async16(Q + r * ld + c, p);
    cp_async16(K + r * ld + c, p + D);
    cp_async16(V + r * ld + c, p + 2 * D);
  }
}

// C[i][j] = sum_d A[i][d] * B[j][d] for j <= i < len; thread (ti, tj) owns i = ti + 16a, j = tj + 16b.
// Whole 16x16 blocks are skipped uniformly across the CTA: blocks beyond ceil(len/16) and the blocks
// b > a that lie strictly above the causal diagonal (their outputs are reported as `above`).
// Rows >= len of A / B hold stale shared memory; each output depends only on its own two rows, and
// outputs with i or j >= len are never reported.
// Packed fp32x2 FMA (FFMA2 on sm_100): both halves of a 64-bit register pair in one instruction.
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

template <class Out>
__device__ __forceinline__ void mm_abt(const float* __restrict__ A, const float* __restrict__ B, int ld, int len, int dh, Out out) {
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  const int nblk = (len + 15) >> 4;                       // CTA-uniform
  unsigned long long acc[4][4];                           // two partial sums (even / odd d) per output, FFMA2-accumulated
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0ull;
#pragma unroll 2
  for (int d = 0; d < dh; d += 4) {
    ulonglong2 av[4], bv[4];                               // {x,y}, {z,w} of a float4
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (a < nblk) {
        av[a] = *reinterpret_cast<const ulonglong2*>(A + (ti + 16 * a) * ld + d);
        bv[a] = *reinterpret_cast<const ulonglong2*>(B + (tj + 16 * a) * ld + d);
      }
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (a < nblk) {
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          ffma2(acc[a][b], av[a].x, bv[b].x);
          ffma2(acc[a][b], av[a].y, bv[b].y);
        }
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = ti + 16 * a, j = tj + 16 * b;
      if (i < len && j < len)
        out(i, j, __uint_as_float((uint32_t)acc[a][b]) + __uint_as_float((uint32_t)(acc[a][b] >> 32)), b > a);
    }
}

// C[r][4c..4c+3] = sum_x S(r,x) * M[x][4c..], S(r,x) = TRANS ? S[x][r] : S[r][x].
// Causal structure: !TRANS sums x <= r (keys up to the query), TRANS sums x >= r (queries from the key
// on); S holds exact zeros outside the band, so the loop bounds only have to cover it.
// A thread owns 4 CONSECUTIVE rows r = 4 tr .. 4 tr + 3 and one float4 of columns, so that its four S operands of a
// step are one 16-byte shared load (a row segment for !TRANS, a column segment of the transposed access for TRANS):
// 8 LDS.128 per 64 FMAs.  Rows >= len read stale but in-bounds shared memory and are never reported.
template <bool TRANS, class Out>
__device__ __forceinline__ void mm_pv(const float* __restrict__ S, int lds, const float* __restrict__ M, int ld, int len, int dh, Out out) {
  const int ncol = dh / 4;                       // float4 columns (16 for dh = 64)
  const int tc = threadIdx.x % ncol, tr = threadIdx.x / ncol;
  const int rows_per_pass = 4 * (kAttnThreads / ncol);
  for (int rb = 4 * tr; rb < len; rb += rows_per_pass) {
    unsigned long long acc[4][2];                 // {x,y}, {z,w} of the thread's float4 of columns, FFMA2-accumulated
#pragma unroll
    for (int a = 0; a < 4; ++a) acc[a][0] = acc[a][1] = 0ull;
    const int x_lo = TRANS ? rb : 0, x_hi = TRANS ? len : min(len, rb + 4);
    int rs[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) rs[a] = min(rb + a, len - 1);        // !TRANS: clamped, always-valid row addresses
    int x = x_lo;                                                  // multiple of 4 in both cases
    for (; x + 3 < x_hi; x += 4) {
      ulonglong2 m[4];
      float4 sv[4];                              // sv[u].{x,y,z,w} = S(rb + {0,1,2,3}, x + u)  (TRANS) ; sv[a] = S(rb + a, x .. x+3) (!TRANS)
#pragma unroll
      for (int u = 0; u < 4; ++u) m[u] = *reinterpret_cast<const ulonglong2*>(M + (x + u) * ld + tc * 4);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        sv[q] = TRANS ? *reinterpret_cast<const float4*>(S + (x + q) * lds + rb) : *reinterpret_cast<const float4*>(S + rs[q] * lds + x);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 t = TRANS ? sv[u] : sv[a];
          const int k = TRANS ? a : u;
          const float s = k == 0 ? t.x : (k == 1 ? t.y : (k == 2 ? t.z : t.w));
          unsigned long long ss;
          asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(s));
          ffma2(acc[a][0], ss, m[u].x);
          ffma2(acc[a][1], ss, m[u].y);
        }
      }
    }
    for (; x < x_hi; ++x) {
      const ulonglong2 m = *reinterpret_cast<const ulonglong2*>(M + x * ld + tc * 4);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float s = TRANS ? S[x * lds + rb + a] : S[rs[a] * lds + x];
        unsigned long long ss;
        asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(s));
        ffma2(acc[a][0], ss, m.x);
        ffma2(acc[a][1], ss, m.y);
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (rb + a < len)
        out(rb + a, tc * 4, make_float4(__uint_as_float((uint32_t)acc[a][0]), __uint_as_float((uint32_t)(acc[a][0] >> 32)),
                                        __uint_as_float((uint32_t)acc[a][1]), __uint_as_float((uint32_t)(acc[a][1] >> 32))));
  }
}

// P[i][j] = softmax_j(S[i][j]) over j < len (masked entries hold -inf -> exactly 0).  Forward: S is
// overwritten with dropout(P).  Backward (Pd != null): S <- P and Pd <- dropout(P).
// One warp per row, two keys per lane (L <= 64).
__device__ __forceinline__ void softmax_rows(float* S, float* Pd, int lds, int len, const Dropout& drop, uint32_t base, int L) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int i = warp; i < len; i += nwarp) {
    float* row = S + i * lds;
    const float s0 = lane < len ? row[lane] : -INFINITY;
    const float s1 = lane + 32 < len ? row[lane + 32] : -INFINITY;
    const float mx = warp_max(fmaxf(s0, s1));
    const float e0 = expf(s0 - mx), e1 = expf(s1 - mx);
    const float inv = 1.0f / warp_sum(e0 + e1);
    const float p0 = e0 * inv, p1 = e1 * inv;
    if (Pd) {
      if (lane < len) { row[lane] = p0; Pd[i * lds + lane] = drop.apply(p0, base + i * L + lane); }
      if (lane + 32 < len) { row[lane + 32] = p1; Pd[i * lds + lane + 32] = drop.apply(p1, base + i * L + lane + 32); }
    } else {
      if (lane < len) row[lane] = drop.apply(p0, base + i * L + lane);
      if (lane + 32 < len) row[lane + 32] = drop.apply(p1, base + i * L + lane + 32);
    }
  }
}

__global__ void __launch_bounds__(kAttnThreads, 2) attn_fwd_kernel(const float* __restrict__ qkv, const int64_t* __restrict__ in_ids,
                                                                const int32_t* __restrict__ tok_off, float* __restrict__ out,
                                                                int L, int D, int n_head, float scale, Dropout drop) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x / n_head, h = blockIdx.x % n_head;
  const int off = tok_off[b], len = min(tok_off[b + 1] - off, L);
  if (len <= 0) return;
  const int dh = D / n_head;
  const Tile t = make_tile(L, dh);
  float* Q = sm; float* K = Q + t.rows * t.ld; float* V = K + t.rows * t.ld; float* S = V + t.rows * t.ld;
  int* pad = reinterpret_cast<int*>(S + t.rows * t.lds);
  const float* base_q = qkv + (size_t)off * 3 * D + h * dh;
  load_qkv(base_q, D, len, dh, Q, K, V, t.ld);
  for (int j = threadIdx.x; j < len; j += blockDim.x) pad[j] = in_ids[(size_t)b * L + j] == 0;
  cp_async_wait_all();
  __syncthreads();
  mm_abt(Q, K, t.ld, len, dh, [&](int i, int j, float v, bool above) {
    S[i * t.lds + j] = (!above && j <= i && !pad[j]) ? v * scale : -INFINITY;
  });
  __syncthreads();
  softmax_rows(S, nullptr, t.lds, len, drop, (uint32_t)(b * n_head + h) * (uint32_t)(L * L), L);
  __syncthreads();
  float* dst = out + (size_t)off * D + h * dh;
  mm_pv<false>(S, t.lds, V, t.ld, len, dh, [&](int i, int c, float4 v) { *reinterpret_cast<float4*>(dst + (size_t)i * D + c) = v; });
}

// dQKV from dO, recomputing the probabilities (no [B,H,L,L] tensor is ever stored)
__global__ void __launch_bounds__(kAttnThreads, 2) attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ d_out,
                                                                const int64_t* __restrict__ in_ids,
                                                                const int32_t* __restrict__ tok_off, float* __restrict__ d_qkv,
                                                                int L, int D, int n_head, float scale, Dropout drop) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x / n_head, h = blockIdx.x % n_head;
  const int off = tok_off[b], len = min(tok_off[b + 1] - off, L);
  if (len <= 0) return;
  const int dh = D / n_head;
  const Tile t = make_tile(L, dh);
  const int msz = t.rows * t.ld, ssz = t.rows * t.lds;
  float* Q = sm; float* K = Q + msz; float* V = K + msz; float* dO = V + msz;
  float* P = dO + msz; float* W = P + ssz;                  // W: Pd, then dP, then dS
  int* pad = reinterpret_cast<int*>(W + ssz);
  const float* base_q = qkv + (size_t)off * 3 * D + h * dh;
  load_qkv(base_q, D, len, dh, Q, K, V, t.ld);
  load_rows(d_out + (size_t)off * D + h * dh, D, len, dh, dO, t.ld);
  for (int j = threadIdx.x; j < len; j += blockDim.x) pad[j] = in_ids[(size_t)b * L + j] == 0;
  cp_async_wait_all();
  __syncthreads();
  const uint32_t base = (uint32_t)(b * n_head + h) * (uint32_t)(L * L);
  mm_abt(Q, K, t.ld, len, dh, [&](int i, int j, float v, bool above) {
    P[i * t.lds + j] = (!above && j <= i && !pad[j]) ? v * scale : -INFINITY;
  });
  __syncthreads();
  softmax_rows(P, W, t.lds, len, drop, base, L);            // P = softmax, W = dropout(P)
  __syncthreads();
  float* dst = d_qkv + (size_t)off * 3 * D + h * dh;
  // dV_j = sum_{i>=j} Pd_ij dO_i
  mm_pv<true>(W, t.lds, dO, t.ld, len, dh, [&](int j, int c, float4 v) { *reinterpret_cast<float4*>(dst + (size_t)j * 3 * D + 2 * D + c) = v; });
  __syncthreads();
  // dP_ij = factor_ij <dO_i, V_j> on the causal, non-pad band (0 elsewhere)
  mm_abt(dO, V, t.ld, len, dh, [&](int i, int j, float v, bool above) {
    W[i * t.lds + j] = (!above && j <= i && !pad[j]) ? v * drop.factor(base + i * L + j) : 0.f;
  });
  __syncthreads();
  {  // dS = P * (dP - sum_j dP P), pre-multiplied by the 1/sqrt(d_h) of the score
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int i = warp; i < len; i += nwarp) {
      float* g = W + i * t.lds;
      const float* p = P + i * t.lds;
      const float g0 = lane < len ? g[lane] : 0.f, p0 = lane < len ? p[lane] : 0.f;
      const float g1 = lane + 32 < len ? g[lane + 32] : 0.f, p1 = lane + 32 < len ? p[lane + 32] : 0.f;
      const float dot = warp_sum(fmaf(g0, p0, g1 * p1));
      if (lane < len) g[lane] = p0 * (g0 - dot) * scale;
      if (lane + 32 < len) g[lane + 32] = p1 * (g1 - dot) * scale;
    }
  }
  __syncthreads();
  // dQ_i = sum_{j<=i} dS_ij K_j ; dK_j = sum_{i>=j} dS_ij Q_i
  mm_pv<false>(W, t.lds, K, t.ld, len, dh, [&](int i, int c, float4 v) { *reinterpret_cast<float4*>(dst + (size_t)i * 3 * D + c) = v; });
  mm_pv<true>(W, t.lds, Q, t.ld, len, dh, [&](int j, int c, float4 v) { *reinterpret_cast<float4*>(dst + (size_t)j * 3 * D + D + c) = v; });
}

inline size_t fwd_smem_bytes(int L, int dh) {
  const Tile t = make_tile(L, dh);
  return sizeof(float) * (size_t)(3 * t.rows * t.ld + t.rows * t.lds + t.rows);
}
inline size_t bwd_smem_bytes(int L, int dh) {
  const Tile t = make_tile(L, dh);
  return sizeof(float) * (size_t)(4 * t.rows * t.ld + 2 * t.rows * t.lds + t.rows);
}

int check_shape(int L, int D, int n_head) {
  if (L > 64 || L <= 0 || D % n_head) return DR4SR_EINVAL;
  const int dh = D / n_head;
  if (dh % 4 || dh > 64 || kAttnThreads % (dh / 4)) return DR4SR_EINVAL;
  return DR4SR_OK;
}

}  // namespace

int launch_attn_fwd(const float* qkv, const int64_t* in_ids, const int32_t* tok_off, float* out, int B, int L, int D,
                    int n_head, Dropout drop, cudaStream_t st) {
  DR4SR_TRY(check_shape(L, D, n_head));
  const int dh = D / n_head;
  const size_t smem = fwd_smem_bytes(L, dh);
  ProfScope prof("attn_fwd", st);
  if (cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "attn_fwd smem attribute");
    return DR4SR_ECUDA;
  }
  attn_fwd_kernel<<<B * n_head, kAttnThreads, smem, st>>>(qkv, in_ids, tok_off, out, L, D, n_head, 1.0f / sqrtf((float)dh), drop);
  DR4SR_LAUNCH_CHECK("attn_fwd_kernel");
  return DR4SR_OK;
}

int launch_attn_bwd(const float* qkv, const float* d_out, const int64_t* in_ids, const int32_t* tok_off, float* d_qkv,
                    int B, int L, int D, int n_head, Dropout drop, cudaStream_t st) {
  DR4SR_TRY(check_shape(L, D, n_head));
  const int dh = D / n_head;
  const size_t smem = bwd_smem_bytes(L, dh);
  ProfScope prof("attn_bwd", st);
  if (cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "attn_bwd smem attribute");
    return DR4SR_ECUDA;
  }
  attn_bwd_kernel<<<B * n_head, kAttnThreads, smem, st>>>(qkv, d_out, in_ids, tok_off, d_qkv, L, D, n_head,
                                                          1.0f / sqrtf((float)dh), drop);
  DR4SR_LAUNCH_CHECK("attn_bwd_kernel");
  return DR4SR_OK;
}

}  // namespace dr4sr
