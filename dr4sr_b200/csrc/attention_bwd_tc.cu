// attention_bwd_tc.cu -- SASRec self-attention backward on tcgen05 / TMEM, persistent (attention backend 2, the default
// for head_dim 64).
//
// Same math as attention.cu / attention_tc.cu (autograd backward of torch SDPA as called from the TransformerEncoder
// configured at reference model/sasrec.py:21-34), probabilities recomputed from the saved qkv:
//   S = Q K^T, P = softmax(mask(S / 8)), Pd = dropout(P);  dPd = dO V^T, dP = dropout'(dPd)
//   dV = Pd^T dO;  dS = P (dP - rowsum(dP P)) / 8;  dQ = dS K;  dK = dS^T Q
// Work item = (tile, head): a tile is a greedy group of whole sequences with <= 128 packed rows (the fused forward's
// tiling, fused_tiles_kernel), so the attention of a tile is ONE 128 x 128 score matrix per head with a block-diagonal
// (same sequence) + causal + key-padding mask.  148 persistent CTAs (one per SM: 12 operand images = 192 KB of shared
// memory) loop over the items; per item
//   global fp32 Q, K, V, dO head slices -> registers (every load of the item in flight at once) -> bf16 hi/lo images
//   S, dPd   : 24 UMMAs  -> TMEM [0,128), [128,256)        (each product = hi*hi + hi*lo + lo*hi, fp32 accumulate)
//   softmax  : thread = (row, 32-key quarter); quarters outside the row's key window are skipped warp-uniformly
//   dV       : Pd image (MN-major A) x dO (MN-major B) -> TMEM [256,320), while dS is computed in registers
//   dQ, dK   : dS image x K / Q -> TMEM [320,384), [384,448)
//   outputs  : TMEM -> registers -> d_qkv (row slices of 128 B per thread)
// What changed against attn_tc_bwd_kernel (attention_tc.cu, kept for the parity tests): greedy tiles (241 instead of
// 335 at cfg-2), persistent CTAs (no per-item launch / TMEM allocation), one batch of global loads per item, the lean
// UMMA issue path of gemm_tc.cuh (descriptor lower words advanced by adds), exp2-based exponentials, dropout draws
// hoisted per row (17 hashes per 32 elements instead of 64), pad mask as bit words.
#include "gemm_tc.cuh"
#include "internal.cuh"

namespace dr4sr {
namespace {

using namespace tc;

constexpr int kBT = 512;
constexpr uint32_t kImg = 128 * 64 * 2;               // one [128 x 64] bf16 image = 16 KB
constexpr uint32_t kIdescN64_KK = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t kIdescN64_K_MN = kIdescN64_KK | (1u << 16);                 // A K-major, B MN-major
constexpr uint32_t kIdescN64_MN_MN = kIdescN64_KK | (1u << 15) | (1u << 16);   // both MN-major
constexpr uint32_t kLboImg = ((kImg >> 4) << 16) - (1u << 16);   // desc_lo(addr) + kLboImg: MN-major, 64-wide MN blocks one image apart

__device__ __forceinline__ bool bar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !bar_try(bar, parity); ++spins)
    if (spins > (1u << 22)) __trap();                      // a protocol bug traps instead of hanging the device
}

__device__ __forceinline__ float exp2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 32 fp32 values of one matrix row (columns [32 qc, 32 qc + 32)) -> their 64 bytes of the row in k-block qc / 2
__device__ __forceinline__ void store_row_image(const float* v, int row, int qc, uint8_t* hi, uint8_t* lo) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 h, l;
    split_bf16x8(make_float4(v[c * 8], v[c * 8 + 1], v[c * 8 + 2], v[c * 8 + 3]),
                 make_float4(v[c * 8 + 4], v[c * 8 + 5], v[c * 8 + 6], v[c * 8 + 7]), h, l);
    const uint32_t off = (uint32_t)(qc >> 1) * kImg + sw128_offset((uint32_t)row, (uint32_t)((qc & 1) * 32 + c * 8));
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}
__device__ __forceinline__ void store_row_zero(int row, int qc, uint8_t* hi, uint8_t* lo) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t off = (uint32_t)(qc >> 1) * kImg + sw128_offset((uint32_t)row, (uint32_t)((qc & 1) * 32 + c * 8));
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// keep-mask of the 32 consecutive dropout elements idx0 .. idx0 + 31: bit j <=> element idx0 + j is kept (Dropout::keep of
// common.cuh, with the draws shared by element pairs and the slowly varying half of draw32 hoisted: 17 + 2 hashes instead
// of 64); the factor of a kept element is d.scale
__device__ __forceinline__ uint32_t drop_keepbits32(const Dropout& d, uint32_t idx0) {
  if (d.thresh == 0u) return 0xFFFFFFFFu;
  const uint32_t p0 = idx0 >> 1, odd = idx0 & 1u, hi0 = p0 >> 7;
  const uint32_t ha = mix32(d.key ^ hi0), hb = mix32(d.key ^ (hi0 + 1u));
  uint32_t bits = 0u, prev = 0u;
#pragma unroll
  for (int k = 0; k < 17; ++k) {
    const uint32_t pk = p0 + (uint32_t)k;
    const uint32_t dr = mix32(pk * 0x9E3779B1u + d.key) ^ ((pk >> 7) == hi0 ? ha : hb);
    if (k > 0) {                     // elements 2(k-1), 2(k-1)+1: half-words of (prev, dr) shifted by the parity of idx0
      const uint32_t e = odd ? __funnelshift_r(prev, dr, 16) : prev;
      bits |= ((e & 0xFFFFu) >= d.thresh ? 1u : 0u) << (2 * (k - 1));
      bits |= ((e >> 16) >= d.thresh ? 1u : 0u) << (2 * (k - 1) + 1);
    }
    prev = dr;
  }
  return bits;
}

// optional per-phase cycle timeline of the first CTAs (-DDR4SR_TRACE; tools/dbg_timeline_attn.py)
__device__ int* g_trace_attn = nullptr;
#ifdef DR4SR_TRACE
#define ATRACE(code) do { if (threadIdx.x == 0 && tr_n < 126) { s_trace[2 * tr_n] = (code); s_trace[2 * tr_n + 1] = (int)(clock64() - tr_t0); ++tr_n; } } while (0)
#else
#define ATRACE(code) do { } while (0)
#endif

struct BwdArgs {
  const float* qkv; const float* d_out; const int64_t* in_ids; const int32_t* tok_off; const int32_t* row_seq; const int32_t* tiles;
  float* d_qkv;
  int L, n_head;
  float scale_log2e, scale;
  Dropout drop;
};

__global__ void __launch_bounds__(kBT, 1) attn_bwd_tc2_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_start[128], s_seq[128];
  __shared__ uint32_t s_padbits[4];
  __shared__ float s_x[4][128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_items = a.tiles[0] * a.n_head;
  if ((int)blockIdx.x >= n_items) return;
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *q_hi = smem, *q_lo = smem + kImg, *k_hi = smem + 2 * kImg, *k_lo = smem + 3 * kImg, *v_hi = smem + 4 * kImg,
          *v_lo = smem + 5 * kImg, *g_hi = smem + 6 * kImg, *g_lo = smem + 7 * kImg,      // g = dO
          *w_hi = smem + 8 * kImg, *w_lo = smem + 10 * kImg;                               // Pd, then dS (2 k-blocks each)
  if (tid == 0) { mbar_init(&bar, 4); fence_mbar_init(); }      // four issuing threads (lane 0 of warps 0..3), one commit each per phase
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t qh = desc_lo(smem_u32(q_hi)), ql = desc_lo(smem_u32(q_lo)), kh = desc_lo(smem_u32(k_hi)), kl = desc_lo(smem_u32(k_lo)),
                 vh = desc_lo(smem_u32(v_hi)), vl = desc_lo(smem_u32(v_lo)), gh = desc_lo(smem_u32(g_hi)), gl = desc_lo(smem_u32(g_lo)),
                 wh = desc_lo(smem_u32(w_hi)), wl = desc_lo(smem_u32(w_lo));
  const int quad = warp & 3, qc = warp >> 2, row = quad * 32 + lane;      // softmax / epilogue identity: (row, 32-column quarter)
  const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
  const int chunk = tid & 7, rsub = tid >> 3;                             // staging identity: 8 floats of rows rsub, 64 + rsub
  uint32_t nbar = 0;                                                      // completed phases of `bar` (uniform)
#ifdef DR4SR_TRACE
  __shared__ int s_trace[256];
  int tr_n = 0; const long long tr_t0 = clock64();
#endif

  // Next item's row metadata (threads 0..127 own one tile row each) travels through registers one item ahead, ONE load per
  // phase of the current item, so that none of the dependent loads (tiles -> tok_off -> row_seq -> tok_off -> in_ids) is
  // ever waited for.  n_r0 / n_R: the next tile's first packed row and row count (CTA-uniform).
  int m_start = 0, m_seq = 0, m_real = 0, n_b0 = 0, n_b1 = 0, n_r0 = 0, n_R = 0, m_off = 0;
  {  // the first item: the plain dependent chain, once per CTA
    const int tl = (int)blockIdx.x / a.n_head;
    n_b0 = a.tiles[1 + tl]; n_b1 = a.tiles[2 + tl];
    n_r0 = a.tok_off[n_b0]; n_R = a.tok_off[n_b1] - n_r0;
    if (tid < 128 && tid < n_R) {
      m_seq = a.row_seq[n_r0 + tid];
      m_off = a.tok_off[m_seq];
      m_start = m_off - n_r0;
      m_real = a.in_ids[(size_t)m_seq * a.L + (n_r0 + tid - m_off)] != 0;
    }
  }
  const int m4 = tid >> 7, mrow = tid & 127;                              // L2 prefetch identity: matrix (Q, K, V, dO), tile row
#pragma unroll 1
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int h = item % a.n_head;
    const int r0 = n_r0, R = n_R;                                         // (fetched one item ahead)
    const int nitem = item + (int)gridDim.x;
    const bool more = nitem < n_items;
    const int nh = nitem % a.n_head;
    if (more) { const int tl = nitem / a.n_head; n_b0 = a.tiles[1 + tl]; n_b1 = a.tiles[2 + tl]; }   // stage 0 of the next item's metadata
    ATRACE(1);
    // ---- every global load of the item in flight at once: Q, K, V, dO head slices (2 rows x 8 floats per thread each) ----
    float4 ld[4][2][2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int r = it * 64 + rsub;
      if (r < R) {
        const float* pq = a.qkv + (size_t)(r0 + r) * 384 + h * 64 + chunk * 8;
        const float* pg = a.d_out + (size_t)(r0 + r) * 128 + h * 64 + chunk * 8;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          ld[m][it][0] = *reinterpret_cast<const float4*>(pq + m * 128);
          ld[m][it][1] = *reinterpret_cast<const float4*>(pq + m * 128 + 4);
        }
        ld[3][it][0] = *reinterpret_cast<const float4*>(pg);
        ld[3][it][1] = *reinterpret_cast<const float4*>(pg + 4);
      } else {
#pragma unroll
        for (int m = 0; m < 4; ++m) ld[m][it][0] = ld[m][it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // ---- row metadata of this item (registers -> shared) ----
    if (tid < 128) {
      s_start[tid] = m_start; s_seq[tid] = m_seq;
      const uint32_t bits = __ballot_sync(0xffffffffu, m_real);
      if (lane == 0) s_padbits[warp] = bits;
    }
    ATRACE(2);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      uint8_t* hi = smem + (uint32_t)(2 * m) * kImg;
      uint8_t* lo = hi + kImg;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        uint4 hh, ll;
        split_bf16x8(ld[m][it][0], ld[m][it][1], hh, ll);
        const uint32_t off = sw128_offset((uint32_t)(it * 64 + rsub), (uint32_t)(chunk * 8));
        *reinterpret_cast<uint4*>(hi + off) = hh;
        *reinterpret_cast<uint4*>(lo + off) = ll;
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    ATRACE(3);
    // several independent instruction streams issue the UMMAs (one thread sustains ~1 tcgen05.mma per 60-100 cycles, a
    // 128 x 64 x 16 UMMA executes in 32); each issuing thread commits, the barrier counts kIssuers arrivals per phase
    if (tid == 0) {                                             // S = Q K^T -> [0,128)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        umma_lo(tmem, qh + 2u * k, kh + 2u * k, kIdesc, k > 0);
        umma_lo<true>(tmem, qh + 2u * k, kl + 2u * k, kIdesc);
        umma_lo<true>(tmem, ql + 2u * k, kh + 2u * k, kIdesc);
      }
      umma_commit(&bar);
    } else if (tid == 32) {                                     // dPd = dO V^T -> [128,256)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        umma_lo(tmem + 128, gh + 2u * k, vh + 2u * k, kIdesc, k > 0);
        umma_lo<true>(tmem + 128, gh + 2u * k, vl + 2u * k, kIdesc);
        umma_lo<true>(tmem + 128, gl + 2u * k, vh + 2u * k, kIdesc);
      }
      umma_commit(&bar);
    } else if (tid == 64 || tid == 96) {
      umma_commit(&bar);                                        // (keeps the arrival count uniform)
    }
    // next item: its tile's row range (the tiles[] words were requested at the top of this item), and its Q, K, V, dO
    // slices pulled from HBM into L2 while this item computes: one 256-byte row slice per thread
    if (more) {
      n_r0 = a.tok_off[n_b0]; n_R = a.tok_off[n_b1] - n_r0;
      if (mrow < n_R) {
        const float* src = m4 < 3 ? a.qkv + (size_t)(n_r0 + mrow) * 384 + m4 * 128 + nh * 64 : a.d_out + (size_t)(n_r0 + mrow) * 128 + nh * 64;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(256) : "memory");
      }
    } else {
      n_R = 0;
    }
    const bool live = row < R;
    const int start = s_start[row];
    const int c0 = qc * 32;
    const bool mine = live && start <= c0 + 31 && row >= c0;     // the quarter intersects the row's key window [start, row]
    const bool any = __any_sync(0xffffffffu, mine);
    // bit j <=> key c0 + j is inside the window and real
    uint32_t okb = 0u;
    if (mine) {
      const int lo = max(start - c0, 0), hi = min(row - c0, 31);
      okb = (hi >= 31 ? 0xFFFFFFFFu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u) & s_padbits[qc];
    }
    const uint32_t dbase = (uint32_t)(s_seq[row] * a.n_head + h) * (uint32_t)(a.L * a.L) + (uint32_t)(row - start) * (uint32_t)a.L +
                           (uint32_t)(c0 - start);                // dropout index of (row, key c0): ((sequence, head), query pos, key pos)
    ATRACE(4);
    bar_wait(&bar, nbar & 1u); ++nbar;
    tc_fence_after();
    ATRACE(5);
    float p[32];
    bool have = false;
    {
      float mx_q = -INFINITY;
      if (any) {
        tmem_ld32(trow + (uint32_t)c0, p);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          p[j] = (okb >> j & 1u) ? p[j] * a.scale_log2e : -INFINITY;     // scores in log2 units: exp(s/8 - m) = 2^(s log2e/8 - m')
          mx_q = fmaxf(mx_q, p[j]);
        }
      }
      s_x[qc][row] = mx_q;
      __syncthreads();
      const float mx = fmaxf(fmaxf(s_x[0][row], s_x[1][row]), fmaxf(s_x[2][row], s_x[3][row]));
      __syncthreads();
      float sum = 0.f;
      if (okb != 0u) {
#pragma unroll
        for (int j = 0; j < 32; ++j) { p[j] = (okb >> j & 1u) ? exp2_fast(p[j] - mx) : 0.f; sum += p[j]; }
      }
      s_x[qc][row] = sum;
      __syncthreads();
      const float tot = (s_x[0][row] + s_x[1][row]) + (s_x[2][row] + s_x[3][row]);
      have = okb != 0u && tot > 0.f;
      if (have) {
        const float inv = 1.0f / tot;
#pragma unroll
        for (int j = 0; j < 32; ++j) p[j] *= inv;
      }
      __syncthreads();
    }
    ATRACE(6);
    if (more && tid < 128) { m_seq = 0; if (tid < n_R) m_seq = a.row_seq[n_r0 + tid]; }     // next item's metadata, stage 2
    uint32_t keep = 0u;                                           // dropout keep-mask of this quarter (shared by Pd and dP)
    if (have) {
      keep = drop_keepbits32(a.drop, dbase);
      float pd[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) pd[j] = (keep >> j & 1u) ? p[j] * a.drop.scale : 0.f;
      store_row_image(pd, row, qc, w_hi, w_lo);
    } else {
      store_row_zero(row, qc, w_hi, w_lo);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    ATRACE(7);
    // dV = Pd^T dO: A = Pd MN-major (M = keys), B = dO MN-major; 8 k-steps of 16 query rows (two 8-row groups = 2048 B, +128),
    // two per issuing thread into four partial accumulators [256,320), [448,512), [0,64), [64,128) (S is dead), summed in the epilogue
    if ((tid & 31) == 0 && tid < 128) {
      const int t4 = tid >> 5;
      const uint32_t acc = tmem + (t4 == 0 ? 256u : t4 == 1 ? 448u : t4 == 2 ? 0u : 64u);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const uint32_t k = (uint32_t)(2 * t4 + kk);
        umma_lo(acc, wh + kLboImg + 128u * k, gh + 128u * k, kIdescN64_MN_MN, kk > 0);
        umma_lo<true>(acc, wh + kLboImg + 128u * k, gl + 128u * k, kIdescN64_MN_MN);
        umma_lo<true>(acc, wl + kLboImg + 128u * k, gh + 128u * k, kIdescN64_MN_MN);
      }
      umma_commit(&bar);
    }
    {  // dS = P * (dP - sum_j dP P) / 8, dP = dPd * keep-factor  (overlaps the dV UMMAs; registers only)
      float dp[32];
      float dot = 0.f;
      if (any) tmem_ld32(trow + 128u + (uint32_t)c0, dp);
      if (have) {
#pragma unroll
        for (int j = 0; j < 32; ++j) { dp[j] = (keep >> j & 1u) ? dp[j] * a.drop.scale : 0.f; dot = fmaf(dp[j], p[j], dot); }
      }
      s_x[qc][row] = dot;
      __syncthreads();
      const float tot = (s_x[0][row] + s_x[1][row]) + (s_x[2][row] + s_x[3][row]);
      if (have) {
#pragma unroll
        for (int j = 0; j < 32; ++j) p[j] = p[j] * (dp[j] - tot) * a.scale;       // p now holds dS
      }
    }
    if (more && tid < 128) { m_off = 0; if (tid < n_R) m_off = a.tok_off[m_seq]; }          // next item's metadata, stage 3
    ATRACE(8);
    bar_wait(&bar, nbar & 1u); ++nbar;                            // dV UMMAs done reading the Pd image
    ATRACE(9);
    tc_fence_after();
    if (have) store_row_image(p, row, qc, w_hi, w_lo);
    else store_row_zero(row, qc, w_hi, w_lo);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();                                              // (also: every thread has read its dP quarter -> [128,256) is free)
    tc_fence_after();
    ATRACE(10);
    // dQ = dS K   : A = dS K-major (k-block k/4, 32 B per step), B = K MN-major (16 keys = 2048 B)   -> [320,384) + [128,192)
    // dK = dS^T Q : A = dS MN-major (M = keys), B = Q MN-major, 16 query rows = 2048 B               -> [384,448) + [192,256)
    // four issuing threads: (dQ, k 0..3), (dQ, k 4..7), (dK, k 0..3), (dK, k 4..7); the halves are summed in the epilogue
    if ((tid & 31) == 0 && tid < 128) {
      const int t4 = tid >> 5;
      const uint32_t acc = tmem + (t4 == 0 ? 320u : t4 == 1 ? 128u : t4 == 2 ? 384u : 192u);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t k = (uint32_t)(4 * (t4 & 1) + kk), bo = 128u * k;
        if (t4 < 2) {
          const uint32_t ao = (k >> 2) * (kImg >> 4) + (k & 3u) * 2u;
          umma_lo(acc, wh + ao, kh + bo, kIdescN64_K_MN, kk > 0);
          umma_lo<true>(acc, wh + ao, kl + bo, kIdescN64_K_MN);
          umma_lo<true>(acc, wl + ao, kh + bo, kIdescN64_K_MN);
        } else {
          umma_lo(acc, wh + kLboImg + bo, qh + bo, kIdescN64_MN_MN, kk > 0);
          umma_lo<true>(acc, wh + kLboImg + bo, ql + bo, kIdescN64_MN_MN);
          umma_lo<true>(acc, wl + kLboImg + bo, qh + bo, kIdescN64_MN_MN);
        }
      }
      umma_commit(&bar);
    }
    if (more && tid < 128) {                                      // next item's metadata, stage 4
      m_start = 0; m_real = 0;
      if (tid < n_R) { m_start = m_off - n_r0; m_real = a.in_ids[(size_t)m_seq * a.L + (n_r0 + tid - m_off)] != 0; }
    }
    ATRACE(11);
    bar_wait(&bar, nbar & 1u); ++nbar;
    tc_fence_after();
    ATRACE(12);
    // ---- outputs: six (matrix, 32-column half) units per row quadrant over its four warps; rows of dQ / dK / dV are all
    // tile rows (queries and keys are the same tokens).  tcgen05.ld is warp-collective: every lane loads, live rows store ----
#pragma unroll 1
    for (int u = qc; u < 6; u += 4) {
      const int mat = u >> 1, half = u & 1;                       // 0: dQ, 1: dK, 2: dV
      float o[32], o2[32];
      if (mat == 2) {                                             // four partial accumulators
        float o3[32];
        tmem_ld32(trow + 256u + (uint32_t)(half * 32), o);
        tmem_ld32(trow + 448u + (uint32_t)(half * 32), o2);
        tmem_ld32(trow + 0u + (uint32_t)(half * 32), o3);
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = (o[j] + o2[j]) + o3[j];
        tmem_ld32(trow + 64u + (uint32_t)(half * 32), o2);
      } else {
        tmem_ld32(trow + (mat == 0 ? 320u : 384u) + (uint32_t)(half * 32), o);
        tmem_ld32(trow + (mat == 0 ? 128u : 192u) + (uint32_t)(half * 32), o2);
      }
      if (live) {
        float* dst = a.d_qkv + (size_t)(r0 + row) * 384 + mat * 128 + h * 64 + half * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] += o2[j];
#pragma unroll
        for (int j = 0; j < 32; j += 8) st_global_v8(dst + j, o + j);
      }
    }
    tc_fence_before();
    __syncthreads();                                              // images, s_x, s_start and TMEM are free for the next item
    tc_fence_after();
    ATRACE(13);
  }
#ifdef DR4SR_TRACE
  if (tid == 0 && g_trace_attn && blockIdx.x < 8)
    for (int i = 0; i < 2 * tr_n; ++i) g_trace_attn[4096 + blockIdx.x * 256 + i] = s_trace[i];
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

int attn_bwd_set_trace(int* host_mapped) {
  return cudaMemcpyToSymbol(g_trace_attn, &host_mapped, sizeof(int*)) == cudaSuccess ? DR4SR_OK : DR4SR_ECUDA;
}

int launch_attn_bwd_tc2(const float* qkv, const float* d_out, const int64_t* in_ids, const int32_t* tok_off, const int32_t* row_seq,
                        const int32_t* tiles, int tiles_cap, float* d_qkv, int B, int L, int D, int n_head, Dropout drop, cudaStream_t st) {
  (void)B;
  if (!attn_tc_supported(L, D, n_head) || D != 128 || n_head != 2) return DR4SR_EINVAL;
  const size_t smem = 12 * kImg + 1024;
  BwdArgs a{};
  a.qkv = qkv; a.d_out = d_out; a.in_ids = in_ids; a.tok_off = tok_off; a.row_seq = row_seq; a.tiles = tiles; a.d_qkv = d_qkv;
  a.L = L; a.n_head = n_head; a.scale = 0.125f; a.scale_log2e = 0.125f * 1.4426950408889634f; a.drop = drop;
  ProfScope prof("attn_bwd_tc", st);
  if (cudaFuncSetAttribute(attn_bwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "attn_bwd_tc smem attribute");
    return DR4SR_ECUDA;
  }
  const int items = tiles_cap * n_head;
  const int grid = items < kNumSMs ? items : kNumSMs;
  attn_bwd_tc2_kernel<<<grid, kBT, smem, st>>>(a);
  DR4SR_LAUNCH_CHECK("attn_bwd_tc2_kernel");
  return DR4SR_OK;
}

}  // namespace dr4sr
