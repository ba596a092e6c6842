// attention_tc.cu -- SASRec self-attention on tcgen05 / TMEM (head_dim 64), forward and backward.
//
// Same math as attention.cu (reference model/sasrec.py:65-68 -> torch SDPA).
// The per-(sequence, head) problems are far too small for a 128-row UMMA tile, so consecutive packed
// sequences are grouped: tile k holds the sequences whose first row falls in [W k, W (k+1)),
// W = 128 - (L - 1), hence at most 128 rows, and the tile's 128 x 128 score matrix is computed in one go
// with a block-diagonal (same sequence) + causal + key-padding mask applied when it is read back from
// TMEM.  One CTA = one tile x one head.
//
// fp32-grade products as in gemm_tc.cuh: every operand is split into bf16 hi + lo and each product is
// three UMMAs.  A token-major [128 rows x 64] bf16 image (128-byte rows, SWIZZLE_128B) serves both as a
// K-major operand (rows = M/N, features = K) and as an MN-major operand (features = M/N, rows = K), so
// Q, K, V, dO are staged once and used for every product:
//   forward   S = Q K^T            (A = Q  K-major,   B = K  K-major)      TMEM cols [0,128)
//             O = P V              (A = P  K-major,   B = V  MN-major)     TMEM cols [128,192)
//   backward  S, dP = dO V^T       (A = dO K-major,   B = V  K-major)      TMEM cols [128,256)
//             dV = Pd^T dO         (A = Pd MN-major,  B = dO MN-major)     TMEM cols [256,320)
//             dQ = dS K            (A = dS K-major,   B = K  MN-major)     TMEM cols [320,384)
//             dK = dS^T Q          (A = dS MN-major,  B = Q  MN-major)     TMEM cols [384,448)
// Softmax runs on registers: thread (row, 32-column quarter) reads its scores with tcgen05.ld (quarters outside the
// row's key window [sequence start, row] are skipped warp-uniformly), the quarters
// of a row exchange max / sum through shared memory, and P (resp. Pd, dS) is written back as a bf16 hi/lo
// operand image for the next UMMA.  Probabilities are recomputed in the backward, never stored.
#include "gemm_tc.cuh"
#include "internal.cuh"

namespace dr4sr {
namespace {

using namespace tc;

constexpr int kAT = 512;                              // threads: (row, 32-column quarter) per thread in the softmax phases
constexpr uint32_t kImg = 128 * 64 * 2;               // one [128 x 64] bf16 image = 16 KB
constexpr uint32_t kIdescN64_KK = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t kIdescN64_K_MN = kIdescN64_KK | (1u << 16);                 // A K-major, B MN-major
constexpr uint32_t kIdescN64_MN_MN = kIdescN64_KK | (1u << 15) | (1u << 16);   // both MN-major

// MN-major SW128 descriptor: `lbo` = byte distance between 64-wide MN blocks, 8-row K groups 1024 B apart
__device__ __forceinline__ uint64_t mn_desc(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// tile_first[k] = first sequence whose first packed row is >= W k  (k = 0..n_tiles), W = 128 - (L - 1)
__global__ void __launch_bounds__(256) attn_tiles_kernel(const int32_t* __restrict__ tok_off, int B, int W, int n_tiles,
                                                         int32_t* __restrict__ tile_first) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b <= B; b += gridDim.x * blockDim.x) {
    if (b == B) {                                          // sentinel: every tile past the last sequence starts at B
      const int last = B > 0 ? tok_off[B - 1] / W : -1;
      for (int k = last + 1; k <= n_tiles; ++k) tile_first[k] = B;
      continue;
    }
    const int s = tok_off[b];
    const int k_lo = b == 0 ? 0 : tok_off[b - 1] / W + 1;  // tiles whose boundary W k lies in (start of b-1, start of b]
    for (int k = k_lo; k <= s / W; ++k) tile_first[k] = b;
  }
}

// fp32 head slice [R x 64] (row stride `ld`) -> bf16 hi / lo images; rows >= R are zero
__device__ __forceinline__ void stage_image(const float* __restrict__ src, int ld, int R, uint8_t* hi, uint8_t* lo) {
  const int chunk = threadIdx.x & 7, rsub = threadIdx.x >> 3;      // 64 rows per pass
  float4 v[2][2];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int row = it * 64 + rsub;
    if (row < R) {
      const float* p = src + (size_t)row * ld + chunk * 8;
      v[it][0] = *reinterpret_cast<const float4*>(p);
      v[it][1] = *reinterpret_cast<const float4*>(p + 4);
    } else {
      v[it][0] = v[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int row = it * 64 + rsub;
    uint4 h, l;
    split_bf16x8(v[it][0], v[it][1], h, l);
    const uint32_t off = sw128_offset((uint32_t)row, (uint32_t)(chunk * 8));
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
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

struct TileInfo { int r0, R, b0, b1; };
__device__ __forceinline__ TileInfo tile_info(const int32_t* tile_first, const int32_t* tok_off, int tile) {
  TileInfo t;
  t.b0 = tile_first[tile]; t.b1 = tile_first[tile + 1];
  t.r0 = tok_off[t.b0];
  t.R = t.b1 > t.b0 ? tok_off[t.b1] - t.r0 : 0;
  return t;
}

// per-row metadata of the tile: first row of the row's sequence (tile-local), (sequence, position) and pad flag
__device__ __forceinline__ void row_meta(const TileInfo& t, const int32_t* tok_off, const int32_t* row_seq, const int64_t* in_ids,
                                         int L, int* s_start, int* s_seq, int* s_pad) {
  for (int i = threadIdx.x; i < 128; i += kAT) {
    int st = 0, sq = 0, pd = 1;
    if (i < t.R) {
      sq = row_seq[t.r0 + i];
      const int off = tok_off[sq];
      st = off - t.r0;
      pd = in_ids[(size_t)sq * L + (t.r0 + i - off)] == 0;
    }
    s_start[i] = st; s_seq[i] = sq; s_pad[i] = pd;
  }
}

// 32 scores of one row (columns [32 qc, 32 qc + 32)) from TMEM -> masked, scaled; returns their maximum
__device__ __forceinline__ float load_scores(uint32_t taddr, int row, int qc, int start, const int* s_pad, float scale, float* s) {
  tmem_ld32(taddr, s);
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int key = qc * 32 + j;
    const bool ok = key >= start && key <= row && !s_pad[key];
    s[j] = ok ? s[j] * scale : -INFINITY;
    mx = fmaxf(mx, s[j]);
  }
  return mx;
}

__global__ void __launch_bounds__(kAT, 1) attn_tc_fwd_kernel(const float* __restrict__ qkv, const int64_t* __restrict__ in_ids,
                                                            const int32_t* __restrict__ tok_off, const int32_t* __restrict__ row_seq,
                                                            const int32_t* __restrict__ tile_first, float* __restrict__ out, int L,
                                                            int D, int n_head, float scale, Dropout drop) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_start[128], s_seq[128], s_pad[128];
  __shared__ float s_x[4][128];
  const int tile = blockIdx.x / n_head, h = blockIdx.x % n_head;
  const TileInfo t = tile_info(tile_first, tok_off, tile);
  if (t.R <= 0) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *q_hi = smem, *q_lo = smem + kImg, *k_hi = smem + 2 * kImg, *k_lo = smem + 3 * kImg, *v_hi = smem + 4 * kImg,
          *v_lo = smem + 5 * kImg, *p_hi = smem + 6 * kImg, *p_lo = smem + 8 * kImg;   // P images: 2 k-blocks each
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  const float* base = qkv + (size_t)t.r0 * 3 * D + h * 64;
  stage_image(base, 3 * D, t.R, q_hi, q_lo);
  stage_image(base + D, 3 * D, t.R, k_hi, k_lo);
  stage_image(base + 2 * D, 3 * D, t.R, v_hi, v_lo);
  row_meta(t, tok_off, row_seq, in_ids, L, s_start, s_seq, s_pad);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {                                             // S = Q K^T
    const uint32_t ah = smem_u32(q_hi), al = smem_u32(q_lo), bh = smem_u32(k_hi), bl = smem_u32(k_lo);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t ko = (uint32_t)k * 32u;
      umma_bf16(tmem, sw128_desc(ah + ko), sw128_desc(bh + ko), kIdesc, k > 0 ? 1u : 0u);
      umma_bf16(tmem, sw128_desc(ah + ko), sw128_desc(bl + ko), kIdesc, 1u);
      umma_bf16(tmem, sw128_desc(al + ko), sw128_desc(bh + ko), kIdesc, 1u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int quad = warp & 3, qc = warp >> 2, row = quad * 32 + lane;      // thread = (row, 32-column quarter)
  const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
  {
    const bool live = row < t.R;
    const int start = s_start[row];
    // the row's keys are [start, row]: a quarter outside that window is all zeros (warp-uniform skip of the TMEM load)
    const bool mine = live && start <= qc * 32 + 31 && row >= qc * 32;
    const bool any = __any_sync(0xffffffffu, mine);
    float s[32];
    float mx_q = -INFINITY;
    if (any) mx_q = load_scores(trow + (uint32_t)(qc * 32), row, qc, start, s_pad, scale, s);
    s_x[qc][row] = mine ? mx_q : -INFINITY;
    __syncthreads();
    const float mx = fmaxf(fmaxf(s_x[0][row], s_x[1][row]), fmaxf(s_x[2][row], s_x[3][row]));
    __syncthreads();
    float sum = 0.f;
    if (mine && mx > -INFINITY) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
    }
    s_x[qc][row] = sum;
    __syncthreads();
    const float tot = (s_x[0][row] + s_x[1][row]) + (s_x[2][row] + s_x[3][row]);
    if (mine && tot > 0.f) {
      const float inv = 1.0f / tot;
      // dropout index: ((sequence, head), query position, key position), as in the FFMA kernels
      const uint32_t dbase = (uint32_t)(s_seq[row] * n_head + h) * (uint32_t)(L * L) + (uint32_t)(row - start) * (uint32_t)L;
#pragma unroll
      for (int j = 0; j < 32; ++j) s[j] = s[j] > 0.f ? drop.apply(s[j] * inv, dbase + (uint32_t)(qc * 32 + j - start)) : 0.f;
      store_row_image(s, row, qc, p_hi, p_lo);
    } else {
      store_row_zero(row, qc, p_hi, p_lo);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {                                             // O = P V : A = P (K-major, 128 keys), B = V (MN-major)
    const uint32_t ah = smem_u32(p_hi), al = smem_u32(p_lo), bh = smem_u32(v_hi), bl = smem_u32(v_lo);
#pragma unroll
    for (int k = 0; k < 8; ++k) {                             // 16 keys per UMMA
      const uint32_t ao = (uint32_t)(k >> 2) * kImg + (uint32_t)(k & 3) * 32u, bo = (uint32_t)k * 2048u;
      umma_bf16(tmem + 128, sw128_desc(ah + ao), mn_desc(bh + bo, 16), kIdescN64_K_MN, k > 0 ? 1u : 0u);
      umma_bf16(tmem + 128, sw128_desc(ah + ao), mn_desc(bl + bo, 16), kIdescN64_K_MN, 1u);
      umma_bf16(tmem + 128, sw128_desc(al + ao), mn_desc(bh + bo, 16), kIdescN64_K_MN, 1u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 1);
  tc_fence_after();
  if (qc < 2) {                                               // warps 0..7: (row, 32-column half) of the 64-wide output
    float o[32];
    tmem_ld32(trow + 128u + (uint32_t)(qc * 32), o);
    if (row < t.R) {
      float* dst = out + (size_t)(t.r0 + row) * D + h * 64 + qc * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

__global__ void __launch_bounds__(kAT, 1) attn_tc_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ d_out,
                                                            const int64_t* __restrict__ in_ids, const int32_t* __restrict__ tok_off,
                                                            const int32_t* __restrict__ row_seq, const int32_t* __restrict__ tile_first,
                                                            float* __restrict__ d_qkv, int L, int D, int n_head, float scale,
                                                            Dropout drop) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_start[128], s_seq[128], s_pad[128];
  __shared__ float s_x[4][128];
  const int tile = blockIdx.x / n_head, h = blockIdx.x % n_head;
  const TileInfo t = tile_info(tile_first, tok_off, tile);
  if (t.R <= 0) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *q_hi = smem, *q_lo = smem + kImg, *k_hi = smem + 2 * kImg, *k_lo = smem + 3 * kImg, *v_hi = smem + 4 * kImg,
          *v_lo = smem + 5 * kImg, *g_hi = smem + 6 * kImg, *g_lo = smem + 7 * kImg,      // g = dO
          *w_hi = smem + 8 * kImg, *w_lo = smem + 10 * kImg;                               // Pd, then dS (2 k-blocks each)
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  const float* base = qkv + (size_t)t.r0 * 3 * D + h * 64;
  stage_image(base, 3 * D, t.R, q_hi, q_lo);
  stage_image(base + D, 3 * D, t.R, k_hi, k_lo);
  stage_image(base + 2 * D, 3 * D, t.R, v_hi, v_lo);
  stage_image(d_out + (size_t)t.r0 * D + h * 64, D, t.R, g_hi, g_lo);
  row_meta(t, tok_off, row_seq, in_ids, L, s_start, s_seq, s_pad);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t qh = smem_u32(q_hi), ql = smem_u32(q_lo), kh = smem_u32(k_hi), kl = smem_u32(k_lo), vh = smem_u32(v_hi),
                 vl = smem_u32(v_lo), gh = smem_u32(g_hi), gl = smem_u32(g_lo), wh = smem_u32(w_hi), wl = smem_u32(w_lo);
  if (tid == 0) {                                             // S = Q K^T -> [0,128) ; dP = dO V^T -> [128,256)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t ko = (uint32_t)k * 32u;
      umma_bf16(tmem, sw128_desc(qh + ko), sw128_desc(kh + ko), kIdesc, k > 0 ? 1u : 0u);
      umma_bf16(tmem, sw128_desc(qh + ko), sw128_desc(kl + ko), kIdesc, 1u);
      umma_bf16(tmem, sw128_desc(ql + ko), sw128_desc(kh + ko), kIdesc, 1u);
      umma_bf16(tmem + 128, sw128_desc(gh + ko), sw128_desc(vh + ko), kIdesc, k > 0 ? 1u : 0u);
      umma_bf16(tmem + 128, sw128_desc(gh + ko), sw128_desc(vl + ko), kIdesc, 1u);
      umma_bf16(tmem + 128, sw128_desc(gl + ko), sw128_desc(vh + ko), kIdesc, 1u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int quad = warp & 3, qc = warp >> 2, row = quad * 32 + lane;      // thread = (row, 32-column quarter)
  const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
  const bool live = row < t.R;
  const int start = s_start[row];
  const uint32_t dbase = (uint32_t)(s_seq[row] * n_head + h) * (uint32_t)(L * L) + (uint32_t)(row - start) * (uint32_t)L;
  const bool mine = live && start <= qc * 32 + 31 && row >= qc * 32;       // quarter intersects the row's key window
  const bool any = __any_sync(0xffffffffu, mine);
  float p[32];
  bool have = false;
  {
    float mx_q = -INFINITY;
    if (any) mx_q = load_scores(trow + (uint32_t)(qc * 32), row, qc, start, s_pad, scale, p);
    s_x[qc][row] = mine ? mx_q : -INFINITY;
    __syncthreads();
    const float mx = fmaxf(fmaxf(s_x[0][row], s_x[1][row]), fmaxf(s_x[2][row], s_x[3][row]));
    __syncthreads();
    float sum = 0.f;
    if (mine && mx > -INFINITY) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { p[j] = expf(p[j] - mx); sum += p[j]; }
    }
    s_x[qc][row] = sum;
    __syncthreads();
    const float tot = (s_x[0][row] + s_x[1][row]) + (s_x[2][row] + s_x[3][row]);
    have = mine && tot > 0.f;
    if (have) {
      const float inv = 1.0f / tot;
#pragma unroll
      for (int j = 0; j < 32; ++j) p[j] *= inv;
    }
    __syncthreads();
  }
  if (have) {  // Pd image (dropout applied) for dV = Pd^T dO
    float pd[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) pd[j] = p[j] > 0.f ? drop.apply(p[j], dbase + (uint32_t)(qc * 32 + j - start)) : 0.f;
    store_row_image(pd, row, qc, w_hi, w_lo);
  } else {
    store_row_zero(row, qc, w_hi, w_lo);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {                                             // dV = Pd^T dO -> [256,320): A = Pd MN-major (M = keys), B = dO MN-major
#pragma unroll
    for (int k = 0; k < 8; ++k) {                             // 16 query rows per UMMA = two 8-row groups = 2048 B
      const uint32_t ko = (uint32_t)k * 2048u;
      umma_bf16(tmem + 256, mn_desc(wh + ko, kImg), mn_desc(gh + ko, 16), kIdescN64_MN_MN, k > 0 ? 1u : 0u);
      umma_bf16(tmem + 256, mn_desc(wh + ko, kImg), mn_desc(gl + ko, 16), kIdescN64_MN_MN, 1u);
      umma_bf16(tmem + 256, mn_desc(wl + ko, kImg), mn_desc(gh + ko, 16), kIdescN64_MN_MN, 1u);
    }
    umma_commit(&bar);
  }
  {  // dS = P * (dP * mask - sum_j dP * mask * P) * scale  (overlaps the dV UMMAs; registers only)
    float dp[32];
    float dot = 0.f;
    if (any) tmem_ld32(trow + 128u + (uint32_t)(qc * 32), dp);
    if (have) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        dp[j] = p[j] > 0.f ? dp[j] * drop.factor(dbase + (uint32_t)(qc * 32 + j - start)) : 0.f;
        dot = fmaf(dp[j], p[j], dot);
      }
    }
    s_x[qc][row] = dot;
    __syncthreads();
    const float tot = (s_x[0][row] + s_x[1][row]) + (s_x[2][row] + s_x[3][row]);
    if (have) {
#pragma unroll
      for (int j = 0; j < 32; ++j) p[j] = p[j] * (dp[j] - tot) * scale;       // p now holds dS
    }
  }
  mbar_wait(&bar, 1);                                         // dV UMMAs done reading the Pd image
  tc_fence_after();
  if (have) store_row_image(p, row, qc, w_hi, w_lo);
  else store_row_zero(row, qc, w_hi, w_lo);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      // dQ = dS K -> [320,384): A = dS K-major (k-block k/4, 32 B per step), B = K MN-major (16 keys = 2048 B)
      const uint32_t ao = (uint32_t)(k >> 2) * kImg + (uint32_t)(k & 3) * 32u, bo = (uint32_t)k * 2048u;
      umma_bf16(tmem + 320, sw128_desc(wh + ao), mn_desc(kh + bo, 16), kIdescN64_K_MN, k > 0 ? 1u : 0u);
      umma_bf16(tmem + 320, sw128_desc(wh + ao), mn_desc(kl + bo, 16), kIdescN64_K_MN, 1u);
      umma_bf16(tmem + 320, sw128_desc(wl + ao), mn_desc(kh + bo, 16), kIdescN64_K_MN, 1u);
      // dK = dS^T Q -> [384,448): A = dS MN-major (M = keys), B = Q MN-major, 16 query rows = 2048 B
      umma_bf16(tmem + 384, mn_desc(wh + bo, kImg), mn_desc(qh + bo, 16), kIdescN64_MN_MN, k > 0 ? 1u : 0u);
      umma_bf16(tmem + 384, mn_desc(wh + bo, kImg), mn_desc(ql + bo, 16), kIdescN64_MN_MN, 1u);
      umma_bf16(tmem + 384, mn_desc(wl + bo, kImg), mn_desc(qh + bo, 16), kIdescN64_MN_MN, 1u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (qc < 2) {  // warps 0..7: rows of dQ / dK / dV are all tile rows (queries and keys are the same tokens); tcgen05.ld is
                 // warp-collective, so every lane loads and only live rows store
    float* dst = d_qkv + (size_t)(t.r0 + row) * 3 * D + h * 64 + qc * 32;
    float o[32];
    tmem_ld32(trow + 320u + (uint32_t)(qc * 32), o);
    if (live) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    }
    tmem_ld32(trow + 384u + (uint32_t)(qc * 32), o);
    if (live) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + D + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    }
    tmem_ld32(trow + 256u + (uint32_t)(qc * 32), o);
    if (live) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + 2 * D + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

bool attn_tc_supported(int L, int D, int n_head) { return D % n_head == 0 && D / n_head == 64 && L >= 1 && L <= 64; }
int attn_tc_num_tiles(int B, int L) { const int W = 128 - (L - 1); return (B * L + W - 1) / W + 1; }

int launch_attn_tiles(const int32_t* tok_off, int B, int L, int32_t* tile_first, cudaStream_t st) {
  const int W = 128 - (L - 1), n_tiles = attn_tc_num_tiles(B, L);
  ProfScope prof("attn_tiles", st);
  attn_tiles_kernel<<<ceil_div(B + 1, 256), 256, 0, st>>>(tok_off, B, W, n_tiles, tile_first);
  DR4SR_LAUNCH_CHECK("attn_tiles_kernel");
  return DR4SR_OK;
}

int launch_attn_tc_fwd(const float* qkv, const int64_t* in_ids, const int32_t* tok_off, const int32_t* row_seq,
                       const int32_t* tile_first, float* out, int B, int L, int D, int n_head, Dropout drop, cudaStream_t st) {
  if (!attn_tc_supported(L, D, n_head)) return DR4SR_EINVAL;
  const size_t smem = 10 * kImg + 1024;
  ProfScope prof("attn_tc_fwd", st);
  if (cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "attn_tc_fwd smem attribute");
    return DR4SR_ECUDA;
  }
  attn_tc_fwd_kernel<<<attn_tc_num_tiles(B, L) * n_head, kAT, smem, st>>>(qkv, in_ids, tok_off, row_seq, tile_first, out, L, D, n_head,
                                                                       0.125f, drop);
  DR4SR_LAUNCH_CHECK("attn_tc_fwd_kernel");
  return DR4SR_OK;
}

int launch_attn_tc_bwd(const float* qkv, const float* d_out, const int64_t* in_ids, const int32_t* tok_off, const int32_t* row_seq,
                       const int32_t* tile_first, float* d_qkv, int B, int L, int D, int n_head, Dropout drop, cudaStream_t st) {
  if (!attn_tc_supported(L, D, n_head)) return DR4SR_EINVAL;
  const size_t smem = 12 * kImg + 1024;
  ProfScope prof("attn_tc_bwd", st);
  if (cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "attn_tc_bwd smem attribute");
    return DR4SR_ECUDA;
  }
  attn_tc_bwd_kernel<<<attn_tc_num_tiles(B, L) * n_head, kAT, smem, st>>>(qkv, d_out, in_ids, tok_off, row_seq, tile_first, d_qkv, L, D,
                                                                       n_head, 0.125f, drop);
  DR4SR_LAUNCH_CHECK("attn_tc_bwd_kernel");
  return DR4SR_OK;
}

}  // namespace dr4sr
