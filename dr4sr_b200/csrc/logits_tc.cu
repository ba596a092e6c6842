// logits_tc.cu -- K3s: full-catalog scoring q @ E^T on tcgen05 with the top-k selection fused into the epilogue; the
// B x N logits never reach HBM.  Replaces BaseModel.topk (reference model/basemodel.py:354-365) for D in {64, 128}.
//
// Exactness.  The reference scores in fp32 and returns ids; SURVEY.md section 7 measured that anything short of fp32-grade
// products changes the returned ids.  Tensor-core products here use the bf16 hi/lo split (3 UMMAs per product, error
// <= 3 * 2^-18 |q|.|e| per score), which is enough to FIND the top-k but not to order near-ties, so the result is produced
// in two steps: the tensor-core pass keeps, per row, every item whose approximate score lies within `delta` (twice the
// error bound) of the row's running k-th best, and the final pass re-scores those few candidates with the exact fp32 FFMA
// dot product (same k-ascending fma chain as the FFMA GEMM path, hence bit-identical scores and ids) and sorts them.
//
// Pipeline (all on `stream`, no host sync):
//   table_image_kernel : E fp32 -> bf16 hi / lo UMMA images (SW128 K-major, 16 KB per 128 items x 64 dims) + max row norm
//   logits_topk_kernel : (row tile, item split) CTAs: 1 producer thread streams E tiles through a 2-stage ring with
//                        cp.async.bulk, 1 thread issues the UMMAs into a double-buffered TMEM accumulator, 128 epilogue
//                        threads (thread = row) read the scores, mask dead items and append the survivors (score >= thr)
//                        to the row's candidate list of this split (predicated stores, no divergent branches).
//       seed phase     : the first 2 tiles of every split, unfiltered
//   select_kernel<MID> : per row, k'-th best over the seeds of ALL splits (k' = k + |history|: the history mask is applied
//                        at the very end) -> thr = that - delta, lists compacted
//       main phase     : the remaining tiles filtered by the row threshold (a lower bound of the final k'-th best, so no
//                        candidate is lost; about 4 % of the scores survive it).  Overflow path: a list that could fill
//                        during the next tile is refreshed by its warp (exact k'-th best of the list, compaction).
//   select_kernel<FINAL>: per row: k'-th best over the candidates, exact re-scoring of the band above it, history ids
//                        dropped, bitonic sort by (score desc, id asc) -> top-k scores and ids
#include "gemm_tc.cuh"
#include "internal.cuh"

namespace dr4sr {
namespace {

using namespace tc;

constexpr int kLT = 256;                       // threads of the scoring kernel
constexpr uint32_t kPiece = 128 * 64 * 2;      // one [128 x 64] bf16 image = 16 KB
constexpr int kCap = 1024;                     // candidate slots per (row, item split); refreshed when fewer than 128 (one tile) are free
constexpr int kFinalCap = 1024;                // candidates re-scored exactly per row
constexpr float kSlack = 3.0e-5f;              // delta = kSlack * |q| * max|e|: twice the 3 * 2^-18 split-product bound, x 1.3

__device__ __forceinline__ uint32_t order_key(float f) {            // monotone float -> uint32
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
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
    if (spins > (1u << 24)) __trap();                      // a protocol bug traps instead of hanging the device
}
__device__ __forceinline__ void bar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- table -> bf16 hi / lo images, rows >= N zero; emax2 = max squared row norm (as uint bits of a non-negative float) ----
__global__ void __launch_bounds__(256) table_image_kernel(const float* __restrict__ table, int64_t N, int64_t n_pad, int D,
                                                          uint8_t* __restrict__ hi, uint8_t* __restrict__ lo, uint32_t* __restrict__ emax2) {
  const int cpr = D / 8;                                  // 16-byte bf16 chunks per row (8 or 16: a power of two <= 16 lanes)
  const int64_t chunks = n_pad * cpr;                     // a multiple of 1024: every thread of a block iteration is in range
  for (int64_t c0 = (int64_t)blockIdx.x * blockDim.x; c0 < chunks; c0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = c0 + threadIdx.x;
    const int64_t n = c / cpr;
    const int k0 = (int)(c % cpr) * 8;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (n < N) {
      a = *reinterpret_cast<const float4*>(table + (size_t)n * D + k0);
      b = *reinterpret_cast<const float4*>(table + (size_t)n * D + k0 + 4);
    }
    float ss = (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w) + (b.x * b.x + b.y * b.y) + (b.z * b.z + b.w * b.w);
    for (int o = cpr >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);   // the cpr lanes of a row are adjacent and aligned
    if ((threadIdx.x & (cpr - 1)) == 0 && ss > 0.f) atomicMax(emax2, __float_as_uint(ss));
    uint4 h, l;
    split_bf16x8(a, b, h, l);
    const size_t off = (size_t)(k0 / 64) * n_pad * 128 + (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 +
                       (size_t)(((((k0 % 64) >> 3) ^ (int)(n & 7)) & 7) << 4);
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

struct ScoreArgs {
  const float* q;            // [B, D]
  const uint8_t *img_hi, *img_lo;
  const uint8_t* dead;       // [N] or null
  const float* thr_in;       // [B] per-row threshold (already minus delta), or null: -inf
  const uint32_t* emax2;
  float* cand_s; int32_t* cand_i; int32_t* cand_cnt;     // [B, S, kCap], [B, S, kCap], [B, S]
  int64_t N, n_pad;
  int B, D, S, item_tiles, k;   // k = k' = requested k + history length
  int rel_lo, rel_hi;        // this launch scores tiles [t_beg + rel_lo, min(t_beg + rel_hi, t_end)) of every split
  int resume;                // 1: the candidate lists continue from cand_cnt
};

// Warp-cooperative refresh of ONE row's candidate list (row `r` of the calling warp; every lane passes its own cnt / thr /
// delta / list pointers and receives the updated cnt / thr if it owns the row).  k-th best of the list by binary search on
// the upper key bits, new threshold = that lower bound - delta, stable compaction (arrival = item order is preserved).  If
// the error band itself is wider than the list (thousands of items tied with the k-th best to within delta: duplicate
// embeddings), the band keeps its lowest ids -- for exactly tied scores that is the reference's answer.  This is the
// OVERFLOW path: with the seed phase's per-row threshold a list normally never fills.
__device__ __noinline__ void refresh_row(int r, int lane, int k, float* cs_mine, int32_t* ci_mine, int& cnt, float& thr, float delta) {
  const int n = __shfl_sync(0xffffffffu, cnt, r);
  const float dl = __shfl_sync(0xffffffffu, delta, r);
  float* cs = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, (unsigned long long)cs_mine, r));
  int32_t* ci = reinterpret_cast<int32_t*>(__shfl_sync(0xffffffffu, (unsigned long long)ci_mine, r));
  constexpr int PER = kCap / 32;
  uint32_t kk[PER];
  int32_t idv[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int e = j * 32 + lane;
    kk[j] = e < n ? order_key(cs[e]) : 0u;                  // key 0 sorts below every real score
    idv[j] = e < n ? ci[e] : 0;
  }
  // k-th best of the list, from below, to the upper 16 bits of its key (sign, exponent, 7 mantissa bits): any lower bound
  // of the k-th best is a valid threshold, and 2^-7 relative only lets a handful of extra candidates through
  uint32_t prefix = 0u;
#pragma unroll 1
  for (int bit = 31; bit >= 16; --bit) {
    const uint32_t cand = prefix | (1u << bit);
    int c = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) c += kk[j] >= cand;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) prefix = cand;
  }
  const float kth = key_to_float(prefix), cut = kth - dl;
  int n_sure = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) n_sure += (j * 32 + lane < n) && key_to_float(kk[j]) > kth + dl;
  n_sure = __reduce_add_sync(0xffffffffu, n_sure);
  const int band_room = max(kCap / 2 - n_sure, k);
  int w = 0, band = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const bool valid = j * 32 + lane < n;
    const float sc = key_to_float(kk[j]);
    const bool sure = valid && sc > kth + dl, inband = valid && !sure && sc >= cut;
    const uint32_t bb = __ballot_sync(0xffffffffu, inband);
    const bool keep = sure || (inband && band + __popc(bb & ((1u << lane) - 1u)) < band_room);
    const uint32_t bk = __ballot_sync(0xffffffffu, keep);
    if (keep) { const int pos = w + __popc(bk & ((1u << lane) - 1u)); cs[pos] = sc; ci[pos] = idv[j]; }
    w += __popc(bk);
    band += __popc(bb);
  }
  __syncwarp();
  if (lane == r) { cnt = w; thr = fmaxf(thr, cut); }
}

__global__ void __launch_bounds__(kLT, 1) logits_topk_kernel(const __grid_constant__ ScoreArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_accf[2], bar_acce[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk = a.D / 64;                                  // k-blocks (1 or 2)
  const int row_tile = blockIdx.x / a.S, split = blockIdx.x % a.S;
  const int per = (a.item_tiles + a.S - 1) / a.S;
  const int s_beg = min(a.item_tiles, split * per), s_end = min(a.item_tiles, s_beg + per);
  const int t_beg = min(s_end, s_beg + a.rel_lo), t_end = min(s_end, s_beg + a.rel_hi);
  const int m0 = row_tile * 128;
  const int n_my = t_end - t_beg;
  if (n_my <= 0 && a.resume) return;                        // nothing to add to the lists of this split (CTA-uniform)
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_img = smem;                                    // [hi kb0 .. hi kb(nk-1)][lo kb0 ..]: 2 nk pieces
  uint8_t* ring = smem + (uint32_t)(2 * nk) * kPiece;       // 2 stages x (2 nk pieces): [hi kb0][lo kb0][hi kb1][lo kb1]
  const uint32_t stage_bytes = (uint32_t)(2 * nk) * kPiece;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); mbar_init(&bar_accf[i], 1); mbar_init(&bar_acce[i], 128); }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  // ---- A operand: the tile's query rows -> bf16 hi/lo images (rows >= B zero) ----
  {
    const int cpr = a.D / 8;
    for (int c = tid; c < 128 * cpr; c += kLT) {
      const int r = c / cpr, k0 = (c % cpr) * 8;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
      if (m0 + r < a.B) {
        x = *reinterpret_cast<const float4*>(a.q + (size_t)(m0 + r) * a.D + k0);
        y = *reinterpret_cast<const float4*>(a.q + (size_t)(m0 + r) * a.D + k0 + 4);
      }
      uint4 h, l;
      split_bf16x8(x, y, h, l);
      const uint32_t off = (uint32_t)(k0 / 64) * kPiece + sw128_offset((uint32_t)r, (uint32_t)(k0 % 64));
      *reinterpret_cast<uint4*>(a_img + off) = h;
      *reinterpret_cast<uint4*>(a_img + (uint32_t)nk * kPiece + off) = l;
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (tid == 128) {
    // ---- producer: E tiles through the ring ----
    for (int it = 0; it < n_my; ++it) {
      const int st = it & 1;
      if (it >= 2) bar_wait(&bar_empty[st], (uint32_t)(((it >> 1) - 1) & 1));
      mbar_expect_tx(&bar_full[st], stage_bytes);
      const int t = t_beg + it;
      uint8_t* dst = ring + (uint32_t)st * stage_bytes;
      for (int kb = 0; kb < nk; ++kb) {
        const size_t src = ((size_t)kb * a.n_pad + (size_t)t * 128) * 128;
        bulk_g2s(dst + (uint32_t)(2 * kb) * kPiece, a.img_hi + src, kPiece, &bar_full[st]);
        bulk_g2s(dst + (uint32_t)(2 * kb + 1) * kPiece, a.img_lo + src, kPiece, &bar_full[st]);
      }
    }
  } else if (tid == 160) {
    // ---- UMMA issuer: scores of tile `it` -> TMEM columns [128 (it & 1), +128) ----
    const uint32_t ah0 = desc_lo(smem_u32(a_img)), al0 = desc_lo(smem_u32(a_img + (uint32_t)nk * kPiece));
    for (int it = 0; it < n_my; ++it) {
      const int st = it & 1;
      bar_wait(&bar_full[st], (uint32_t)((it >> 1) & 1));
      if (it >= 2) bar_wait(&bar_acce[st], (uint32_t)(((it >> 1) - 1) & 1));
      tc_fence_after();
      const uint32_t acc = tmem + (uint32_t)(st * 128);
      const uint32_t e0 = desc_lo(smem_u32(ring + (uint32_t)st * stage_bytes));
      for (int kb = 0; kb < nk; ++kb) {
        const uint32_t ah = ah0 + (uint32_t)kb * (kPiece >> 4), al = al0 + (uint32_t)kb * (kPiece >> 4);
        const uint32_t eh = e0 + (uint32_t)(2 * kb) * (kPiece >> 4), el = eh + (kPiece >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_lo(acc, ah + 2u * k, eh + 2u * k, kIdesc, kb > 0 || k > 0);
          umma_lo<true>(acc, ah + 2u * k, el + 2u * k, kIdesc);
          umma_lo<true>(acc, al + 2u * k, eh + 2u * k, kIdesc);
        }
      }
      umma_commit(&bar_empty[st]);                          // ring stage free once these UMMAs have read it
      umma_commit(&bar_accf[st]);                           // accumulator ready
    }
  } else if (tid < 128) {
    // ---- epilogue: thread = query row; the four warps never wait for each other ----
    const int m = m0 + tid;
    const bool rowlive = m < a.B;
    float qn2 = 0.f;
    if (rowlive)
      for (int d = 0; d < a.D; d += 4) {
        const float4 v = *reinterpret_cast<const float4*>(a.q + (size_t)m * a.D + d);
        qn2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      }
    const float delta = kSlack * sqrtf(qn2 * __uint_as_float(*a.emax2));
    float thr = (a.thr_in && rowlive) ? a.thr_in[m] : -INFINITY;
    float* cs = a.cand_s + ((size_t)(rowlive ? m : 0) * a.S + split) * kCap;
    int32_t* ci = a.cand_i + ((size_t)(rowlive ? m : 0) * a.S + split) * kCap;
    int cnt = (a.resume && rowlive) ? a.cand_cnt[(size_t)m * a.S + split] : 0;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int it = 0; it < n_my; ++it) {
      const int st = it & 1;
      const int n0 = (t_beg + it) * 128;
      // live-item words of the tile (bit set = scorable item): every warp ballots the 4 x 32 items itself
      uint32_t okw[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int64_t id = (int64_t)n0 + w * 32 + lane;
        const bool live = id < a.N && !(a.dead && a.dead[id]);
        okw[w] = __ballot_sync(0xffffffffu, live);
      }
      bar_wait(&bar_accf[st], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_ld32(trow + (uint32_t)(st * 128 + c * 32), v);
        uint32_t surv = 0u;
#pragma unroll
        for (int j = 0; j < 32; ++j) surv |= (v[j] >= thr ? 1u : 0u) << j;
        surv &= rowlive ? okw[c] : 0u;
        if (__any_sync(0xffffffffu, surv != 0u)) {            // (the list has at least 128 free slots at the start of every tile)
          // predicated stores, no divergent branches: the 32 rows of a warp survive at unrelated columns
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const uint32_t bit = surv >> j & 1u;
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.u32 p, %0, 0;\n\t"
                "@p st.global.f32 [%1], %2;\n\t"
                "@p st.global.s32 [%3], %4;\n\t"
                "}" ::"r"(bit), "l"(cs + cnt), "f"(v[j]), "l"(ci + cnt), "r"(n0 + c * 32 + j) : "memory");
            cnt += (int)bit;
          }
        }
      }
      tc_fence_before();
      bar_arrive(&bar_acce[st]);                              // the accumulator is free: the UMMAs of tile it + 2 may start
      // overflow path: lists that could fill during the next tile are refreshed now, one row at a time, by the whole warp
      uint32_t need = __ballot_sync(0xffffffffu, rowlive && cnt > kCap - 128);
      while (need) {
        const int r = __ffs(need) - 1;
        need &= need - 1;
        refresh_row(r, lane, a.k, cs, ci, cnt, thr, delta);
      }
    }
    if (rowlive) a.cand_cnt[(size_t)m * a.S + split] = cnt;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---- per row, over the candidate lists of all its splits: k-th best (radix select on the key bytes), then either
//   MID   : thr[row] = k-th best - delta and every list compacted to the entries >= thr (stable), or
//   FINAL : exact fp32 re-scoring of the entries >= k-th best - delta, bitonic sort by (score desc, id asc), top-k out ----
struct SelectArgs {
  const float* q; const float* table; const uint32_t* emax2;
  const int64_t* hist; int H;                    // FINAL: the row's history ids (dropped before the top-k is taken)
  float* cand_s; int32_t* cand_i; int32_t* cand_cnt;
  float* thr;                                    // MID out
  float* out_scores; int64_t* out_ids;           // FINAL out
  int D, S, k, kq;                               // kq = k' = k + H (selection rank), k = ranks returned
};

constexpr int kFT2 = 256;
template <bool FINAL>
__global__ void __launch_bounds__(kFT2) select_kernel(const __grid_constant__ SelectArgs a) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_mask;
  __shared__ int s_need, s_m, s_warp[kFT2 / 32];
  __shared__ float s_q[128];
  __shared__ uint32_t s_key[FINAL ? kFinalCap : 1];       // exact-score keys of the re-scored candidates
  __shared__ int32_t s_id[FINAL ? kFinalCap : 1];
  __shared__ int32_t s_cnt[160];
  const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* cs = a.cand_s + (size_t)m * a.S * kCap;
  int32_t* ci = a.cand_i + (size_t)m * a.S * kCap;
  for (int s = tid; s < a.S; s += kFT2) s_cnt[s] = min(a.cand_cnt[(size_t)m * a.S + s], kCap);
  float qq = 0.f;
  if (tid < a.D) { s_q[tid] = a.q[(size_t)m * a.D + tid]; qq = s_q[tid] * s_q[tid]; }
  qq = warp_sum(qq);
  if (lane == 0) s_warp[warp] = __float_as_int(qq);
  if (tid == 0) { s_prefix = 0u; s_mask = 0u; s_need = a.kq; s_m = 0; }
  __syncthreads();
  float qn2 = 0.f;
  for (int wv = 0; wv < kFT2 / 32; ++wv) qn2 += __int_as_float(s_warp[wv]);
  const float delta = kSlack * sqrtf(qn2 * __uint_as_float(*a.emax2));
  // ---- k'-th best approximate score over the row's candidates: radix select on the key bytes ----
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[tid] = 0u;
    __syncthreads();
    const uint32_t prefix = s_prefix, mask = s_mask;
    for (int s = 0; s < a.S; ++s) {
      const float* ls = cs + (size_t)s * kCap;
      for (int e = tid; e < s_cnt[s]; e += kFT2) {
        const uint32_t key = order_key(ls[e]);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
      }
    }
    __syncthreads();
    if (tid == 0) {
      int need = s_need, d = 255;
      for (; d > 0; --d) {
        const int c = (int)hist[d];
        if (c >= need) break;
        need -= c;
      }
      s_need = need;
      s_prefix = prefix | ((uint32_t)d << shift);
      s_mask = mask | (255u << shift);
    }
    __syncthreads();
  }
  // fewer than k' candidates in total: the search ends on key 0 (below every real key) and everything is kept
  const bool all = s_prefix == 0u;
  const float cut = all ? -INFINITY : key_to_float(s_prefix) - delta;
  __syncthreads();
  if (!FINAL) {
    if (tid == 0) a.thr[m] = cut;
    // stable in-place compaction of every list: chunks of kFT2 entries, reads of a chunk complete before its writes
    for (int s = 0; s < a.S; ++s) {
      const int n = s_cnt[s];
      float* ls = cs + (size_t)s * kCap;
      int32_t* li = ci + (size_t)s * kCap;
      int w = 0;
      for (int start = 0; start < n; start += kFT2) {
        const int e = start + tid;
        float sc = 0.f; int32_t id = 0;
        bool keep = false;
        if (e < n) { sc = ls[e]; id = li[e]; keep = sc >= cut; }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = w;
        for (int wv = 0; wv < warp; ++wv) off += s_warp[wv];
        off += __popc(bal & ((1u << lane) - 1u));
        if (keep) { ls[off] = sc; li[off] = id; }
        for (int wv = 0; wv < kFT2 / 32; ++wv) w += s_warp[wv];
        __syncthreads();
      }
      if (tid == 0) a.cand_cnt[(size_t)m * a.S + s] = w;
    }
    return;
  }
  // ---- ordered gather (candidate order = id order: splits are ascending item ranges, lists are filled in item order) ----
  __shared__ int32_t s_hist[64];
  if (FINAL) {
    for (int i = tid; i < 64; i += kFT2) s_hist[i] = (a.hist && i < a.H) ? (int32_t)min((long long)a.hist[(size_t)m * a.H + i], 0x7fffffffLL) : -1;
    __syncthreads();
  }
  for (int sp = 0; sp < a.S; ++sp)
  for (int start = 0; start < s_cnt[sp]; start += kFT2) {
    const int e = start + tid, i = sp * kCap + e;
    bool keep = false;
    if (e < s_cnt[sp] && cs[i] >= cut) {
      keep = true;
      const int32_t id = ci[i];
      for (int hh = 0; hh < a.H; ++hh) keep = keep && s_hist[hh] != id;       // history items never rank (basemodel.py:361)
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_m;
    for (int wv = 0; wv < warp; ++wv) off += s_warp[wv];
    off += __popc(bal & ((1u << lane) - 1u));
    if (keep && off < kFinalCap) s_id[FINAL ? off : 0] = ci[i];          // a band wider than kFinalCap keeps its lowest ids
    __syncthreads();
    if (tid == 0) { int t = 0; for (int wv = 0; wv < kFT2 / 32; ++wv) t += s_warp[wv]; s_m = min(s_m + t, kFinalCap); }
    __syncthreads();
  }
  const int mcand = s_m;
  // ---- exact fp32 re-scoring: the k-ascending fma chain of the FFMA GEMM path ----
  for (int j = tid; j < kFinalCap; j += kFT2) {
    if (j < mcand) {
      const float* e = a.table + (size_t)s_id[FINAL ? j : 0] * a.D;
      float acc = 0.f;
      for (int d = 0; d < a.D; d += 4) {
        const float4 v = *reinterpret_cast<const float4*>(e + d);
        acc = fmaf(s_q[d], v.x, acc); acc = fmaf(s_q[d + 1], v.y, acc); acc = fmaf(s_q[d + 2], v.z, acc); acc = fmaf(s_q[d + 3], v.w, acc);
      }
      s_key[FINAL ? j : 0] = order_key(acc);
    } else if (FINAL) {
      s_key[j] = 0u; s_id[j] = 0x7fffffff;                   // padding sinks to the end
    }
  }
  __syncthreads();
  int kpad = 32;
  while (kpad < mcand) kpad <<= 1;
  for (int size = 2; size <= kpad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < kpad; i += kFT2) {
        const int j = i ^ stride;
        if (j > i && FINAL) {
          const bool up = (i & size) == 0;
          const uint32_t ki = s_key[i], kj = s_key[j];
          const int32_t ii = s_id[i], ij = s_id[j];
          const bool i_better = ki > kj || (ki == kj && ii < ij);
          if (up ? !i_better : i_better) { s_key[i] = kj; s_key[j] = ki; s_id[i] = ij; s_id[j] = ii; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < a.k; i += kFT2) {
    const bool real = i < mcand;
    a.out_ids[(size_t)m * a.k + i] = real ? (int64_t)s_id[FINAL ? i : 0] : (int64_t)i;       // fewer than k live items: -inf padding
    a.out_scores[(size_t)m * a.k + i] = real ? key_to_float(s_key[FINAL ? i : 0]) : -INFINITY;
  }
}

struct Plan {
  int row_tiles, S, item_tiles;
  int64_t n_pad;
  size_t off_hi, off_lo, off_emax, off_hist, off_thr, off_cs, off_ci, off_cnt, bytes;
};
inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
Plan make_plan(int B, int64_t N, int D, int H) {
  Plan p{};
  p.n_pad = (N + 127) & ~(int64_t)127;
  p.item_tiles = (int)(p.n_pad / 128);
  p.row_tiles = (B + 127) / 128;
  int S = kNumSMs / p.row_tiles;
  if (S < 1) S = 1;
  if (S > p.item_tiles) S = p.item_tiles;
  if (S > 160) S = 160;
  p.S = S;
  size_t o = 0;
  p.off_hi = o; o += al256((size_t)p.n_pad * D * 2);
  p.off_lo = o; o += al256((size_t)p.n_pad * D * 2);
  p.off_emax = o; o += 256;
  p.off_hist = o; o += al256((size_t)B * (H > 0 ? H : 1) * 4);
  p.off_thr = o; o += al256((size_t)B * 4);
  p.off_cs = o; o += al256((size_t)B * S * kCap * 4);
  p.off_ci = o; o += al256((size_t)B * S * kCap * 4);
  p.off_cnt = o; o += al256((size_t)B * S * 4);
  p.bytes = o;
  return p;
}

}  // namespace

bool logits_tc_supported(int D, int32_t k) { return (D == 64 || D == 128) && k >= 1 && k + 64 <= kCap / 2; }
size_t logits_tc_workspace_bytes(int B, int64_t N, int D, int H) { return make_plan(B, N, D, H).bytes; }

int launch_logits_topk_tc(const float* q, const float* table, const uint8_t* item_dead, const int64_t* user_hist, int B, int D, int64_t N,
                          int H, int k, float* out_scores, int64_t* out_ids, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!logits_tc_supported(D, k)) return DR4SR_EINVAL;
  const Plan p = make_plan(B, N, D, H);
  if (ws_bytes < p.bytes) return DR4SR_EWORKSPACE;
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  uint8_t *hi = base + p.off_hi, *lo = base + p.off_lo;
  uint32_t* emax2 = reinterpret_cast<uint32_t*>(base + p.off_emax);
  float* thr = reinterpret_cast<float*>(base + p.off_thr);
  float* cs = reinterpret_cast<float*>(base + p.off_cs);
  int32_t* ci = reinterpret_cast<int32_t*>(base + p.off_ci);
  int32_t* cnt = reinterpret_cast<int32_t*>(base + p.off_cnt);
  if (cudaMemsetAsync(emax2, 0, 4, st) != cudaSuccess) { set_cuda_error(cudaGetLastError(), "topk memset"); return DR4SR_ECUDA; }
  {
    ProfScope prof("topk_table_images", st);
    const int64_t chunks = p.n_pad * (D / 8);
    const int blocks = (int)((chunks + 255) / 256 < 8 * kNumSMs ? (chunks + 255) / 256 : 8 * kNumSMs);
    table_image_kernel<<<blocks, 256, 0, st>>>(table, N, p.n_pad, D, hi, lo, emax2);
    DR4SR_LAUNCH_CHECK("table_image_kernel");
  }
  const bool with_hist = user_hist && H > 0;
  const size_t smem = (size_t)(2 * (D / 64) + 2 * 2 * (D / 64)) * kPiece + 1024;
  if (cudaFuncSetAttribute(logits_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "logits_topk smem attribute");
    return DR4SR_ECUDA;
  }
  ScoreArgs a{};
  const int kq = k + (with_hist ? H : 0);          // the history mask is applied after the selection: rank k + H is always enough
  a.q = q; a.img_hi = hi; a.img_lo = lo; a.dead = item_dead; a.emax2 = emax2;
  a.N = N; a.n_pad = p.n_pad; a.B = B; a.D = D; a.k = kq;
  a.cand_s = cs; a.cand_i = ci; a.cand_cnt = cnt; a.S = p.S; a.item_tiles = p.item_tiles;
  SelectArgs f{};
  f.q = q; f.table = table; f.emax2 = emax2; f.cand_s = cs; f.cand_i = ci; f.cand_cnt = cnt; f.thr = thr;
  f.out_scores = out_scores; f.out_ids = out_ids; f.D = D; f.S = p.S; f.k = k; f.kq = kq;
  f.hist = with_hist ? user_hist : nullptr; f.H = with_hist ? H : 0;
  const int per = (p.item_tiles + p.S - 1) / p.S;
  const int seed_tiles = 2;                                // (unfiltered: every item of the seed tiles goes into the lists)
  {  // seed phase: the first tiles of every split, unfiltered
    a.rel_lo = 0; a.rel_hi = seed_tiles; a.resume = 0; a.thr_in = nullptr;
    ProfScope prof("topk_logits_seed", st);
    logits_topk_kernel<<<p.row_tiles * p.S, kLT, smem, st>>>(a);
    DR4SR_LAUNCH_CHECK("logits_topk_kernel(seed)");
  }
  if (per > seed_tiles) {
    {  // per-row threshold = k-th best over the seeds of ALL splits (S x 896 items) - delta; lists compacted to it
      ProfScope prof("topk_threshold", st);
      select_kernel<false><<<B, kFT2, 0, st>>>(f);
      DR4SR_LAUNCH_CHECK("select_kernel(mid)");
    }
    {  // main phase: the remaining tiles, filtered by the row threshold
      a.rel_lo = seed_tiles; a.rel_hi = per; a.resume = 1; a.thr_in = thr;
      ProfScope prof("topk_logits_tc", st);
      logits_topk_kernel<<<p.row_tiles * p.S, kLT, smem, st>>>(a);
      DR4SR_LAUNCH_CHECK("logits_topk_kernel");
    }
  }
  {
    ProfScope prof("topk_final", st);
    select_kernel<true><<<B, kFT2, 0, st>>>(f);
    DR4SR_LAUNCH_CHECK("select_kernel(final)");
  }
  return DR4SR_OK;
}

}  // namespace dr4sr
