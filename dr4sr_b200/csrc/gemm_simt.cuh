// gemm_simt.cuh -- fp32 FFMA tiled GEMM family used by the encoder kernels (exact-fp32 path).
//
//   C[m,n] (+epilogue) = sum_k  proA(A)[m,k] * proB(B)[k,n]
//
// One template covers the three operand arrangements of the step:
//   forward   y  = x  W^T   : A k-contiguous [M,K],  B k-contiguous [N,K]        (A_KC=1, B_KC=1)
//   backward  dx = dy W     : A k-contiguous [M,K'], B n-contiguous [K',N]       (A_KC=1, B_KC=0)
//   backward  dW = dy^T x   : A m-contiguous [R,M],  B n-contiguous [R,N], R = tokens, split over
//                             blockIdx.z with per-split partial outputs            (A_KC=0, B_KC=0)
// 256 threads as a 16x16 grid, (BM/16)x(BN/16) accumulators per thread, BK=16, two smem stages with
// register prefetch.  The token dimension is read from device memory (packed ragged batches), so the
// launch shape is static and CUDA-graph friendly.
#pragma once
#include "common.cuh"

namespace dr4sr {

enum Prologue : int { PRO_NONE = 0, PRO_GELU_DROP = 1, PRO_DROPMASK = 2 };
enum Epilogue : int { EPI_LINEAR = 0, EPI_GELU_BWD = 1 };

struct GemmArgs {
  const float* A; const float* B; float* C;
  int lda, ldb, ldc;
  int M, N, K;            // logical extents (capacity for the dynamic one)
  int b_rows;             // B_KC form: rows of B that exist (<= N; columns beyond read as 0)
  const int* tok_dev;     // device scalar: live token count (bounds M for *_KC=1 forms, K for the TN form)
  int n_split;            // TN form: number of blockIdx.z splits (partials at C + z * split_stride)
  int64_t split_stride;
  // prologues
  int proA, proB;
  Dropout dropA, dropB;   // dropout streams for the prologues (index = row * ld + col of the source)
  // epilogue
  int epi;
  const float* bias;      // [N] or null
  const float* add;       // [M, ldadd] or null
  int ldadd;
  const float* pre;       // EPI_GELU_BWD: pre-activation [M, ldc]
  const float* mul;       // EPI_GELU_BWD, tcgen05 kernel only: saved dropout-mask * gelu'(pre) [M, ldc] (replaces pre / dropE when set)
  Dropout dropE;          // EPI_GELU_BWD / LN epilogue dropout stream (index = m * N + n)
  // LN epilogue (GemmLN kernel only): z = drop(acc + bias) + add ; y = LN(z)
  const float* gamma; const float* beta; float ln_eps;
  float* Z; float* stats;  // Z [M,N] pre-norm sum, stats [M,2] = {mean, rstd}
  const char* tag;         // name under which the optional event timer files this launch
};

// idx % 4 == 0 everywhere a prologue / epilogue touches a float4 (leading dimensions are multiples of 4)
__device__ __forceinline__ float4 apply_prologue(float4 v, int pro, const Dropout& d, uint32_t idx) {
  if (pro == PRO_GELU_DROP) {
    const float4 f = d.factor4(idx);
    v.x = gelu_f(v.x) * f.x; v.y = gelu_f(v.y) * f.y; v.z = gelu_f(v.z) * f.z; v.w = gelu_f(v.w) * f.w;
  } else if (pro == PRO_DROPMASK) {
    const float4 f = d.factor4(idx);
    v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
  }
  return v;
}

template <int BM, int BN, bool A_KC, bool B_KC, bool LN_EPI>
__global__ void __launch_bounds__(256, (BM * BN <= 64 * 128) ? 2 : 1) gemm_simt_kernel(GemmArgs g) {
  constexpr int BK = 16;
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int LDA = BM + 4, LDB = BN + 4;
  constexpr int A_F4 = BM * BK / 4 / 256;  // float4 loads per thread per tile
  constexpr int B_F4 = BN * BK / 4 / 256;
  static_assert(A_F4 >= 1 && B_F4 >= 1, "tile too small for 256 threads");
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][LDB];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tok = g.tok_dev ? *g.tok_dev : (A_KC ? g.M : g.K);

  int Mlim = g.M, kbeg = 0, kend = g.K;
  if (A_KC) {
    Mlim = min(g.M, tok);
    if (m0 >= Mlim) return;
  } else {  // TN: reduction over tokens, split over blockIdx.z
    const int R = min(g.K, tok);
    int len = (R + g.n_split - 1) / g.n_split;
    len = (len + BK - 1) / BK * BK;
    kbeg = min(R, (int)blockIdx.z * len);
    kend = min(R, kbeg + len);
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_F4], rb[B_F4];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int q = 0; q < A_F4; ++q) {
      const int f = tid + q * 256;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A_KC) {
        const int m = f / (BK / 4), kq = f % (BK / 4);
        const int gm = m0 + m, gk = k0 + kq * 4;
        if (gm < Mlim && gk < kend) {
          v = *reinterpret_cast<const float4*>(g.A + (size_t)gm * g.lda + gk);
          v = apply_prologue(v, g.proA, g.dropA, (uint32_t)gm * (uint32_t)g.lda + gk);
        }
      } else {
        const int kk = f / (BM / 4), mq = f % (BM / 4);
        const int gk = k0 + kk, gm = m0 + mq * 4;
        if (gk < kend && gm < g.M) {
          v = *reinterpret_cast<const float4*>(g.A + (size_t)gk * g.lda + gm);
          v = apply_prologue(v, g.proA, g.dropA, (uint32_t)gk * (uint32_t)g.lda + gm);
        }
      }
      ra[q] = v;
    }
#pragma unroll
    for (int q = 0; q < B_F4; ++q) {
      const int f = tid + q * 256;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (B_KC) {
        const int n = f / (BK / 4), kq = f % (BK / 4);
        const int gn = n0 + n, gk = k0 + kq * 4;
        if (gn < g.b_rows && gk < kend) {
          v = *reinterpret_cast<const float4*>(g.B + (size_t)gn * g.ldb + gk);
          v = apply_prologue(v, g.proB, g.dropB, (uint32_t)gn * (uint32_t)g.ldb + gk);
        }
      } else {
        const int kk = f / (BN / 4), nq = f % (BN / 4);
        const int gk = k0 + kk, gn = n0 + nq * 4;
        if (gk < kend && gn < g.N) {
          v = *reinterpret_cast<const float4*>(g.B + (size_t)gk * g.ldb + gn);
          v = apply_prologue(v, g.proB, g.dropB, (uint32_t)gk * (uint32_t)g.ldb + gn);
        }
      }
      rb[q] = v;
    }
  };
  auto store_tiles = [&](int s) {
#pragma unroll
    for (int q = 0; q < A_F4; ++q) {
      const int f = tid + q * 256;
      if (A_KC) {
        const int m = f / (BK / 4), kq = f % (BK / 4);
        As[s][kq * 4 + 0][m] = ra[q].x; As[s][kq * 4 + 1][m] = ra[q].y;
        As[s][kq * 4 + 2][m] = ra[q].z; As[s][kq * 4 + 3][m] = ra[q].w;
      } else {
        const int kk = f / (BM / 4), mq = f % (BM / 4);
        *reinterpret_cast<float4*>(&As[s][kk][mq * 4]) = ra[q];
      }
    }
#pragma unroll
    for (int q = 0; q < B_F4; ++q) {
      const int f = tid + q * 256;
      if (B_KC) {
        const int n = f / (BK / 4), kq = f % (BK / 4);
        Bs[s][kq * 4 + 0][n] = rb[q].x; Bs[s][kq * 4 + 1][n] = rb[q].y;
        Bs[s][kq * 4 + 2][n] = rb[q].z; Bs[s][kq * 4 + 3][n] = rb[q].w;
      } else {
        const int kk = f / (BN / 4), nq = f % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[s][kk][nq * 4]) = rb[q];
      }
    }
  };

  const int ntile = (kend - kbeg + BK - 1) / BK;
  if (ntile > 0) {
    load_tiles(kbeg);
    store_tiles(0);
  }
  __syncthreads();
  for (int t = 0; t < ntile; ++t) {
    const int s = t & 1;
    if (t + 1 < ntile) load_tiles(kbeg + (t + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int h = 0; h < TM / 4; ++h) {
        const float4 v = *reinterpret_cast<const float4*>(&As[s][k][h * (BM / (TM / 4)) + ty * 4]);
        a[h * 4 + 0] = v.x; a[h * 4 + 1] = v.y; a[h * 4 + 2] = v.z; a[h * 4 + 3] = v.w;
      }
#pragma unroll
      for (int h = 0; h < TN / 4; ++h) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[s][k][h * (BN / (TN / 4)) + tx * 4]);
        b[h * 4 + 0] = v.x; b[h * 4 + 1] = v.y; b[h * 4 + 2] = v.z; b[h * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < ntile) store_tiles(s ^ 1);
    __syncthreads();
  }

  // ---------------- epilogue ----------------
  float* Cbase = g.C + (A_KC ? 0 : (size_t)blockIdx.z * g.split_stride);
  if (!LN_EPI) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + (i / 4) * (BM / (TM / 4)) + ty * 4 + (i & 3);
      if (m >= Mlim) continue;
#pragma unroll
      for (int h = 0; h < TN / 4; ++h) {
        const int n = n0 + h * (BN / (TN / 4)) + tx * 4;
        if (n >= g.N) continue;
        float4 v = make_float4(acc[i][h * 4], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
        if (g.epi == EPI_GELU_BWD) {
          const float4 p = *reinterpret_cast<const float4*>(g.pre + (size_t)m * g.ldc + n);
          const uint32_t idx = (uint32_t)m * (uint32_t)g.ldc + n;
          const float4 f = g.dropE.factor4(idx);
          v.x *= f.x * gelu_grad_f(p.x); v.y *= f.y * gelu_grad_f(p.y);
          v.z *= f.z * gelu_grad_f(p.z); v.w *= f.w * gelu_grad_f(p.w);
        } else {
          if (g.bias) {
            const float4 bb = *reinterpret_cast<const float4*>(g.bias + n);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
          }
          if (g.add) {
            const float4 aa = *reinterpret_cast<const float4*>(g.add + (size_t)m * g.ldadd + n);
            v.x += aa.x; v.y += aa.y; v.z += aa.z; v.w += aa.w;
          }
          if (g.dropE.thresh) {   // gradient through a dropout that sits on this GEMM's output tensor
            const uint32_t idx = (uint32_t)m * (uint32_t)g.ldc + n;
            const float4 f = g.dropE.factor4(idx);
            v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
          }
        }
        *reinterpret_cast<float4*>(Cbase + (size_t)m * g.ldc + n) = v;
      }
    }
  } else {
    // z = drop(acc + bias) + residual; y = LayerNorm(z).  The CTA owns complete rows (BN == N); a row
    // lives in the 16 consecutive lanes tx = 0..15, so the moments are half-warp shuffles.
    const float invN = 1.0f / (float)g.N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + (i / 4) * (BM / (TM / 4)) + ty * 4 + (i & 3);
      const bool live = m < Mlim;
      float z[TN];
      float s = 0.f;
#pragma unroll
      for (int h = 0; h < TN / 4; ++h) {
        const int n = h * (BN / (TN / 4)) + tx * 4;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f), rr = bb;
        if (live) {
          if (g.bias) bb = *reinterpret_cast<const float4*>(g.bias + n);
          rr = *reinterpret_cast<const float4*>(g.add + (size_t)m * g.ldadd + n);
        }
        const uint32_t idx = (uint32_t)m * (uint32_t)g.N + n;
        const float4 f = g.dropE.factor4(idx);
        z[h * 4 + 0] = (acc[i][h * 4 + 0] + bb.x) * f.x + rr.x;
        z[h * 4 + 1] = (acc[i][h * 4 + 1] + bb.y) * f.y + rr.y;
        z[h * 4 + 2] = (acc[i][h * 4 + 2] + bb.z) * f.z + rr.z;
        z[h * 4 + 3] = (acc[i][h * 4 + 3] + bb.w) * f.w + rr.w;
        s += (z[h * 4] + z[h * 4 + 1]) + (z[h * 4 + 2] + z[h * 4 + 3]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mu = s * invN;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < TN; ++j) { const float d = z[j] - mu; q = fmaf(d, d, q); }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q * invN + g.ln_eps);
      if (!live) continue;
#pragma unroll
      for (int h = 0; h < TN / 4; ++h) {
        const int n = h * (BN / (TN / 4)) + tx * 4;
        const float4 gg = *reinterpret_cast<const float4*>(g.gamma + n);
        const float4 be = *reinterpret_cast<const float4*>(g.beta + n);
        float4 y;
        y.x = (z[h * 4 + 0] - mu) * rstd * gg.x + be.x;
        y.y = (z[h * 4 + 1] - mu) * rstd * gg.y + be.y;
        y.z = (z[h * 4 + 2] - mu) * rstd * gg.z + be.z;
        y.w = (z[h * 4 + 3] - mu) * rstd * gg.w + be.w;
        *reinterpret_cast<float4*>(g.C + (size_t)m * g.ldc + n) = y;
        if (g.Z)
          *reinterpret_cast<float4*>(g.Z + (size_t)m * g.N + n) =
              make_float4(z[h * 4], z[h * 4 + 1], z[h * 4 + 2], z[h * 4 + 3]);
      }
      if (tx == 0 && g.stats) { g.stats[2 * m] = mu; g.stats[2 * m + 1] = rstd; }
    }
  }
}

// host-side launchers ----------------------------------------------------------------------------
template <int BM, int BN, bool A_KC, bool B_KC, bool LN_EPI>
inline int launch_gemm(const GemmArgs& g, cudaStream_t st) {
  dim3 grid(ceil_div(g.N, BN), A_KC ? ceil_div(g.M, BM) : ceil_div(g.M, BM), A_KC ? 1 : g.n_split);
  if (LN_EPI && g.N != BN) return DR4SR_EINVAL;
  ProfScope prof(g.tag ? g.tag : "gemm", st);
  gemm_simt_kernel<BM, BN, A_KC, B_KC, LN_EPI><<<grid, 256, 0, st>>>(g);
  DR4SR_LAUNCH_CHECK("gemm_simt_kernel");
  return DR4SR_OK;
}

inline GemmArgs gemm_args(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                          const int* tok_dev) {
  GemmArgs g{};
  g.A = A; g.B = B; g.C = C; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.b_rows = N;
  g.tok_dev = tok_dev; g.n_split = 1; g.split_stride = 0; g.proA = PRO_NONE; g.proB = PRO_NONE; g.epi = EPI_LINEAR;
  Dropout none; none.key = 0; none.thresh = 0; none.scale = 1.f;
  g.dropA = none; g.dropB = none; g.dropE = none;
  return g;
}

}  // namespace dr4sr
