// fused_fwd.cu -- the whole SASRec encoder forward as ONE persistent kernel (D = F = 128, 2 heads of 64).
//
// Replaces SASRecQueryEncoder.forward (reference model/sasrec.py:39-75): embedding gather + learned
// positions + dropout, then n_layer post-norm TransformerEncoderLayers (model/sasrec.py:21-34).
//
// Why one kernel: at the headline shape (26 K live tokens x 128) every activation is a few MB and every
// per-op kernel is a 10-30 us latency chain over 1-2 waves of CTAs.  Attention never crosses a sequence,
// so a group of whole sequences with <= 128 packed rows (a "tile", built greedily by fused_tiles_kernel)
// can be carried through ALL layers by one CTA without any grid-wide dependency:
//
//   tile -> [gather+pos+dropout] -> for each layer:
//             QKV GEMM (3 x 128 columns) -> per head: S = Q K^T, softmax+dropout, O = P V
//             -> out-proj + dropout + residual + LayerNorm -> FFN up (+bias, GELU, dropout)
//             -> FFN down + dropout + residual + LayerNorm
//
// Every product runs on tcgen05 with the bf16 hi/lo split of gemm_tc.cuh (3 UMMAs per product, fp32
// TMEM accumulator), 256 threads, ~97 KB of shared memory and 256 TMEM columns per CTA => 2 CTAs per SM
// whose phases overlap (one CTA's epilogue under the other's UMMAs).  Weight images stream from L2 through
// a 2-slot ring of 16 KB cp.async.bulk pieces.  Epilogues write the activations the backward needs to
// HBM and, where the next product consumes them, stage the next A operand straight into shared memory.
#include "gemm_tc.cuh"
#include "internal.cuh"

namespace dr4sr {
namespace {

using namespace tc;

constexpr int kFT = 256;                       // threads per CTA
constexpr uint32_t kImg = 128 * 64 * 2;        // one [128 x 64] bf16 image = 16 KB
constexpr uint32_t kFusedSmem = 6 * kImg + 1024;
constexpr uint32_t kIdescN64_K_MN = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                                    ((uint32_t)(128 >> 4) << 24);   // A K-major, B MN-major, N = 64

__device__ __forceinline__ uint64_t mn_desc16(uint32_t smem_addr) {   // MN-major SW128, one 64-wide MN block
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(16 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ---- bounded mbarrier waits + optional progress trace ------------------------------------------------------
// A wait that would spin forever (a protocol bug) traps instead of hanging the GPU; when a host-mapped trace
// buffer is installed (dr4sr_debug_trace), thread 0 of every CTA also records the last phase it reached.
__device__ int* g_trace = nullptr;
__device__ __forceinline__ void trace(int code) {
  if (g_trace && threadIdx.x == 0) { volatile int* t = g_trace; t[blockIdx.x * 4] = code; }
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity, int tag) {
  for (uint32_t spins = 0; !mbar_try(bar, parity); ++spins) {
    if (spins > (1u << 22)) {
      if (g_trace) {
        volatile int* t = g_trace;
        t[blockIdx.x * 4 + 1] = tag; t[blockIdx.x * 4 + 2] = (int)threadIdx.x; t[blockIdx.x * 4 + 3] = (int)parity;
        __threadfence_system();
      }
      __trap();
    }
  }
}

struct FusedLayer {
  float *qkv, *attn, *z1, *st1, *x1, *pre, *z2, *st2, *x2;                        // saved activations (packed rows)
  const uint16_t *in_hi, *in_lo, *out_hi, *out_lo, *w1_hi, *w1_lo, *w2_hi, *w2_lo;   // weight images
  const float *in_b, *out_b, *b1, *b2, *g1, *be1, *g2, *be2;
  Dropout d_attn_p, d_attn_out, d_ffn_h, d_ffn_out;
};
struct FusedFwdArgs {
  const float* table; const float* pos;
  const int64_t* in_ids; const int32_t* tok_off; const int32_t* row_seq; const int32_t* tiles;
  float* x0;
  int n_layer, L;
  float ln_eps, scale;
  Dropout d_embed;
  FusedLayer layer[8];
};

// ---- tiles: greedy groups of whole sequences with <= 128 packed rows ------------------------------
// tiles[0] = n_tiles, tiles[1 + k] = first sequence of tile k, tiles[1 + n_tiles] = B.
__global__ void __launch_bounds__(1024) fused_tiles_kernel(const int32_t* __restrict__ tok_off, int B, int32_t* __restrict__ tiles) {
  extern __shared__ int32_t s_off[];
  for (int i = threadIdx.x; i <= B; i += blockDim.x) s_off[i] = tok_off[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    int k = 0, r0 = 0;
    tiles[1] = 0;
    for (int b = 0; b < B; ++b) {
      if (s_off[b + 1] - r0 > 128) {          // sequence b does not fit: it opens the next tile
        ++k;
        tiles[1 + k] = b;
        r0 = s_off[b];
      }
    }
    if (s_off[B] - r0 > 0 || k == 0) ++k;      // the open tile (or a single empty one)
    tiles[1 + k] = B;
    tiles[0] = k;
  }
}

// ---- shared state of one CTA -------------------------------------------------------------------------
struct Ctx {
  uint8_t* smem;            // 1 KB aligned dynamic smem: [0,64K) A / Q,K / P images, [64K,96K) weight ring / V images
  uint64_t *full, *empty;   // [2] weight-ring barriers
  uint64_t* acc;            // accumulator-ready barrier
  uint32_t tmem;            // TMEM base (256 columns)
  uint32_t fetched, used;   // weight-ring counters (meaningful in thread 0 only)
  uint32_t n_acc;           // completed phases of `acc` (uniform across the CTA)
  int r0, R;                // packed rows [r0, r0 + R) of the tile
};

__device__ __forceinline__ uint8_t* a_hi(const Ctx& c, int kb) { return c.smem + (uint32_t)kb * kImg; }
__device__ __forceinline__ uint8_t* a_lo(const Ctx& c, int kb) { return c.smem + (uint32_t)(2 + kb) * kImg; }
__device__ __forceinline__ uint8_t* ring(const Ctx& c, uint32_t slot) { return c.smem + (4 + slot) * kImg; }

// thread 0: queue the bulk copy of one 16 KB weight piece into the next ring slot
__device__ __forceinline__ void ring_fetch(Ctx& c, const uint8_t* src) {
  const uint32_t slot = c.fetched & 1u, use = c.fetched >> 1;
  if (use >= 1) {                                   // the UMMAs that read the slot's previous piece are done
    mbar_wait_b(&c.empty[slot], (use - 1) & 1u, 100 + (int)slot);
  }
  mbar_expect_tx(&c.full[slot], kImg);
  bulk_g2s(ring(c, slot), src, kImg, &c.full[slot]);
  ++c.fetched;
}
// piece p of a 128-column chunk n0 of a logical [N_total, 128] weight: p = 2 * kb + (lo ? 1 : 0)
__device__ __forceinline__ const uint8_t* piece_src(const uint16_t* hi, const uint16_t* lo, int N_total, int n0, int p) {
  const uint8_t* base = reinterpret_cast<const uint8_t*>((p & 1) ? lo : hi);
  return base + ((size_t)(p >> 1) * N_total + n0) * 128;
}
__device__ __forceinline__ void prefetch_chunk(Ctx& c, const uint16_t* hi, const uint16_t* lo, int N_total, int n0) {
  ring_fetch(c, piece_src(hi, lo, N_total, n0, 0));
  ring_fetch(c, piece_src(hi, lo, N_total, n0, 1));
}

// thread 0: acc[128 x 128] (TMEM columns acc_col..+127) = A (hi/lo images in smem, K = 128) x W[n0..n0+127, :]^T.
// The first two pieces of the chunk must already be in flight (prefetch_chunk).  `nhi/nlo != null`: the first two
// pieces of the NEXT chunk are queued as soon as ring slots free up.
__device__ __forceinline__ void issue_chunk(Ctx& c, const uint16_t* hi, const uint16_t* lo, int N_total, int n0, uint32_t acc_col,
                                            const uint16_t* nhi, const uint16_t* nlo, int nN_total, int nn0) {
  const uint32_t acc = c.tmem + acc_col;
#pragma unroll 1
  for (int p = 0; p < 4; ++p) {
    const uint32_t slot = c.used & 1u, use = c.used >> 1;
    mbar_wait_b(&c.full[slot], use & 1u, 200 + (int)slot);
    tc_fence_after();
    const int kb = p >> 1;
    const uint32_t b = smem_u32(ring(c, slot)), ah = smem_u32(a_hi(c, kb)), al = smem_u32(a_lo(c, kb));
    if ((p & 1) == 0) {                               // W_hi piece: A_hi W_hi + A_lo W_hi
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t ko = (uint32_t)k * 32u;
        umma_bf16(acc, sw128_desc(ah + ko), sw128_desc(b + ko), kIdesc, (p > 0 || k > 0) ? 1u : 0u);
        umma_bf16(acc, sw128_desc(al + ko), sw128_desc(b + ko), kIdesc, 1u);
      }
    } else {                                          // W_lo piece: A_hi W_lo
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t ko = (uint32_t)k * 32u;
        umma_bf16(acc, sw128_desc(ah + ko), sw128_desc(b + ko), kIdesc, 1u);
      }
    }
    umma_commit(&c.empty[slot]);
    ++c.used;
    if (p + 2 < 4) ring_fetch(c, piece_src(hi, lo, N_total, n0, p + 2));
  }
  umma_commit(c.acc);
  if (nhi) prefetch_chunk(c, nhi, nlo, nN_total, nn0);
}

// all threads: wait for the accumulator committed by the most recent issue
__device__ __forceinline__ void wait_acc(Ctx& c) {
  mbar_wait_b(c.acc, c.n_acc & 1u, 300);
  ++c.n_acc;
  tc_fence_after();
}

// 32 fp32 values of tile row `row`, columns [32 qc, 32 qc + 32) of a 128-wide operand -> hi / lo images
// laid out as [hi kb0, hi kb1, lo kb0, lo kb1] from `base`
__device__ __forceinline__ void store_row_image(const float* v, int row, int qc, uint8_t* base) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 h, l;
    split_bf16x8(make_float4(v[q * 8], v[q * 8 + 1], v[q * 8 + 2], v[q * 8 + 3]),
                 make_float4(v[q * 8 + 4], v[q * 8 + 5], v[q * 8 + 6], v[q * 8 + 7]), h, l);
    const uint32_t off = (uint32_t)(qc >> 1) * kImg + sw128_offset((uint32_t)row, (uint32_t)((qc & 1) * 32 + q * 8));
    *reinterpret_cast<uint4*>(base + off) = h;
    *reinterpret_cast<uint4*>(base + 2 * kImg + off) = l;
  }
}
__device__ __forceinline__ void store_row_zero(int row, int qc, uint8_t* base) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t off = (uint32_t)(qc >> 1) * kImg + sw128_offset((uint32_t)row, (uint32_t)((qc & 1) * 32 + q * 8));
    *reinterpret_cast<uint4*>(base + off) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(base + 2 * kImg + off) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// fp32 [R x 128] rows (row stride ld) from global -> A images (rows >= R zero)
__device__ __forceinline__ void stage_a_global(const Ctx& c, const float* src, int ld) {
  const int chunk = threadIdx.x & 15, rsub = threadIdx.x >> 4;     // 16 chunks of 8 floats, 16 rows per pass
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float4 v[4][2];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = (half * 4 + it) * 16 + rsub;
      if (row < c.R) {
        const float* p = src + (size_t)(c.r0 + row) * ld + chunk * 8;
        v[it][0] = *reinterpret_cast<const float4*>(p);
        v[it][1] = *reinterpret_cast<const float4*>(p + 4);
      } else {
        v[it][0] = v[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = (half * 4 + it) * 16 + rsub;
      uint4 h, l;
      split_bf16x8(v[it][0], v[it][1], h, l);
      const uint32_t off = (uint32_t)(chunk >> 3) * kImg + sw128_offset((uint32_t)row, (uint32_t)((chunk & 7) * 8));
      *reinterpret_cast<uint4*>(c.smem + off) = h;
      *reinterpret_cast<uint4*>(c.smem + 2 * kImg + off) = l;
    }
  }
}

// fp32 head slice [R x 64] (row stride ld) -> one hi / lo image pair (rows >= R zero)
__device__ __forceinline__ void stage_head(const Ctx& c, const float* src, int ld, uint8_t* hi, uint8_t* lo) {
  const int chunk = threadIdx.x & 7, rsub = threadIdx.x >> 3;      // 32 rows per pass
  float4 v[4][2];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 32 + rsub;
    if (row < c.R) {
      const float* p = src + (size_t)row * ld + chunk * 8;
      v[it][0] = *reinterpret_cast<const float4*>(p);
      v[it][1] = *reinterpret_cast<const float4*>(p + 4);
    } else {
      v[it][0] = v[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 32 + rsub;
    uint4 h, l;
    split_bf16x8(v[it][0], v[it][1], h, l);
    const uint32_t off = sw128_offset((uint32_t)row, (uint32_t)(chunk * 8));
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

// make generic-proxy smem writes visible to the tensor core, order TMEM reads before later UMMAs, CTA barrier
__device__ __forceinline__ void sync_for_mma() {
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// ---- epilogues (thread = tile row `quad*32+lane`, 64-column half `warp>>2`) -----------------------------
// out[row, n] = acc + bias[n]   (optionally: next A operand = dropout(gelu(out)))
template <bool GELU_STAGE>
__device__ __forceinline__ void epi_linear(const Ctx& c, uint32_t acc_col, const float* bias, float* out, int ldo, const Dropout& dh) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane, m = c.r0 + row;
  const bool live = row < c.R;
  const uint32_t trow = c.tmem + ((uint32_t)(quad * 32) << 16) + acc_col + (uint32_t)(half * 64);
#pragma unroll 1
  for (int q = 0; q < 2; ++q) {
    float v[32];
    tmem_ld32(trow + (uint32_t)(q * 32), v);
    const int nb = half * 64 + q * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bb = *reinterpret_cast<const float4*>(bias + nb + j);
      v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
      if (live) *reinterpret_cast<float4*>(out + (size_t)m * ldo + nb + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    if (GELU_STAGE) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 f = dh.factor4((uint32_t)m * 128u + (uint32_t)(nb + j));
        v[j] = gelu_f(v[j]) * f.x; v[j + 1] = gelu_f(v[j + 1]) * f.y;
        v[j + 2] = gelu_f(v[j + 2]) * f.z; v[j + 3] = gelu_f(v[j + 3]) * f.w;
      }
      if (live) store_row_image(v, row, half * 2 + q, c.smem);
      else store_row_zero(row, half * 2 + q, c.smem);
    }
  }
}

// z = dropout(acc + bias) + res ; y = LayerNorm(z) -> Z, stats, Y (global) and, if STAGE, the next A operand
template <bool STAGE>
__device__ __forceinline__ void epi_ln(const Ctx& c, uint32_t acc_col, const float* bias, const float* res, const Dropout& de,
                                       const float* gamma, const float* beta, float eps, float* Z, float* stats, float* Y,
                                       float (*ln_part)[128]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane, m = c.r0 + row;
  const bool live = row < c.R;
  const uint32_t trow = c.tmem + ((uint32_t)(quad * 32) << 16) + acc_col + (uint32_t)(half * 64);
  float sum = 0.f;
#pragma unroll 1
  for (int q = 0; q < 2; ++q) {
    float v[32];
    tmem_ld32(trow + (uint32_t)(q * 32), v);
    const int nb = half * 64 + q * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const int n = nb + j;
      const float4 bb = *reinterpret_cast<const float4*>(bias + n);
      float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) rr = *reinterpret_cast<const float4*>(res + (size_t)m * 128 + n);
      const float4 f = de.factor4((uint32_t)m * 128u + (uint32_t)n);
      v[j] = (v[j] + bb.x) * f.x + rr.x;
      v[j + 1] = (v[j + 1] + bb.y) * f.y + rr.y;
      v[j + 2] = (v[j + 2] + bb.z) * f.z + rr.z;
      v[j + 3] = (v[j + 3] + bb.w) * f.w + rr.w;
      sum += (v[j] + v[j + 1]) + (v[j + 2] + v[j + 3]);
      if (live) *reinterpret_cast<float4*>(Z + (size_t)m * 128 + n) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    tmem_st32(trow + (uint32_t)(q * 32), v);
  }
  ln_part[half][row] = sum;
  __syncthreads();
  const float mu = (ln_part[0][row] + ln_part[1][row]) * (1.0f / 128.0f);
  __syncthreads();
  float var = 0.f;
#pragma unroll 1
  for (int q = 0; q < 2; ++q) {
    float v[32];
    tmem_ld32(trow + (uint32_t)(q * 32), v);
#pragma unroll
    for (int j = 0; j < 32; ++j) { const float d = v[j] - mu; var = fmaf(d, d, var); }
  }
  ln_part[half][row] = var;
  __syncthreads();
  const float rstd = rsqrtf((ln_part[0][row] + ln_part[1][row]) * (1.0f / 128.0f) + eps);
#pragma unroll 1
  for (int q = 0; q < 2; ++q) {
    float v[32];
    tmem_ld32(trow + (uint32_t)(q * 32), v);
    const int nb = half * 64 + q * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const int n = nb + j;
      const float4 gg = *reinterpret_cast<const float4*>(gamma + n);
      const float4 be = *reinterpret_cast<const float4*>(beta + n);
      v[j] = (v[j] - mu) * rstd * gg.x + be.x; v[j + 1] = (v[j + 1] - mu) * rstd * gg.y + be.y;
      v[j + 2] = (v[j + 2] - mu) * rstd * gg.z + be.z; v[j + 3] = (v[j + 3] - mu) * rstd * gg.w + be.w;
      if (live) *reinterpret_cast<float4*>(Y + (size_t)m * 128 + n) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    if (STAGE) {
      if (live) store_row_image(v, row, half * 2 + q, c.smem);
      else store_row_zero(row, half * 2 + q, c.smem);
    }
  }
  if (live && half == 0) { stats[2 * m] = mu; stats[2 * m + 1] = rstd; }
}

// ---- attention of one head over the tile (block-diagonal causal mask) -----------------------------------
__device__ __forceinline__ void attention_head(Ctx& c, const FusedFwdArgs& a, const FusedLayer& y, int h, const int* s_start,
                                               const int* s_seq, const int* s_pad, float (*s_x)[128]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;
  const bool live = row < c.R;
  uint8_t *q_hi = c.smem, *q_lo = c.smem + kImg, *k_hi = c.smem + 2 * kImg, *k_lo = c.smem + 3 * kImg;
  uint8_t *v_hi = c.smem + 4 * kImg, *v_lo = c.smem + 5 * kImg;
  const float* base = y.qkv + (size_t)c.r0 * 384 + h * 64;
  stage_head(c, base, 384, q_hi, q_lo);
  stage_head(c, base + 128, 384, k_hi, k_lo);
  stage_head(c, base + 256, 384, v_hi, v_lo);
  sync_for_mma();
  if (tid == 0) {                                             // S = Q K^T -> TMEM columns [0,128)
    const uint32_t ah = smem_u32(q_hi), al = smem_u32(q_lo), bh = smem_u32(k_hi), bl = smem_u32(k_lo);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t ko = (uint32_t)k * 32u;
      umma_bf16(c.tmem, sw128_desc(ah + ko), sw128_desc(bh + ko), kIdesc, k > 0 ? 1u : 0u);
      umma_bf16(c.tmem, sw128_desc(ah + ko), sw128_desc(bl + ko), kIdesc, 1u);
      umma_bf16(c.tmem, sw128_desc(al + ko), sw128_desc(bh + ko), kIdesc, 1u);
    }
    umma_commit(c.acc);
  }
  wait_acc(c);
  const uint32_t trow = c.tmem + ((uint32_t)(quad * 32) << 16);
  const int start = s_start[row];
  // keys of this row: [start, row] minus pads.  Quarter qc (32 keys) is live for the row when it intersects that window.
  bool mine[2], any[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int qc = half * 2 + q;
    mine[q] = live && start <= qc * 32 + 31 && row >= qc * 32;
    any[q] = __any_sync(0xffffffffu, mine[q]);
  }
  float mx = -INFINITY;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (!any[q]) continue;
    const int qc = half * 2 + q;
    float s[32];
    tmem_ld32(trow + (uint32_t)(qc * 32), s);
    if (mine[q]) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int key = qc * 32 + j;
        const bool ok = key >= start && key <= row && !s_pad[key];
        mx = fmaxf(mx, ok ? s[j] * a.scale : -INFINITY);
      }
    }
  }
  s_x[half][row] = mx;
  __syncthreads();
  mx = fmaxf(s_x[0][row], s_x[1][row]);
  __syncthreads();
  // P' = dropout(exp(s - max)) un-normalised (the 1/sum is applied to the output rows), written over the dead Q / K images
  float sum = 0.f;
  const uint32_t dbase = (uint32_t)(s_seq[row] * 2 + h) * (uint32_t)(a.L * a.L) + (uint32_t)(row - start) * (uint32_t)a.L;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int qc = half * 2 + q;
    if (any[q]) {
      float s[32];
      tmem_ld32(trow + (uint32_t)(qc * 32), s);
      if (mine[q] && mx > -INFINITY) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int key = qc * 32 + j;
          const bool ok = key >= start && key <= row && !s_pad[key];
          const float e = ok ? expf(s[j] * a.scale - mx) : 0.f;
          sum += e;
          s[j] = ok ? y.d_attn_p.apply(e, dbase + (uint32_t)(key - start)) : 0.f;
        }
        store_row_image(s, row, qc, c.smem);
      } else {
        store_row_zero(row, qc, c.smem);
      }
    } else {
      store_row_zero(row, qc, c.smem);
    }
  }
  s_x[half][row] = sum;
  sync_for_mma();
  if (tid == 0) {                                             // O = P' V -> TMEM columns [128,192): A = P' K-major, B = V MN-major
    const uint32_t ah = smem_u32(c.smem), al = smem_u32(c.smem + 2 * kImg), bh = smem_u32(v_hi), bl = smem_u32(v_lo);
#pragma unroll
    for (int k = 0; k < 8; ++k) {                             // 16 keys per UMMA
      const uint32_t ao = (uint32_t)(k >> 2) * kImg + (uint32_t)(k & 3) * 32u, bo = (uint32_t)k * 2048u;
      umma_bf16(c.tmem + 128, sw128_desc(ah + ao), mn_desc16(bh + bo), kIdescN64_K_MN, k > 0 ? 1u : 0u);
      umma_bf16(c.tmem + 128, sw128_desc(ah + ao), mn_desc16(bl + bo), kIdescN64_K_MN, 1u);
      umma_bf16(c.tmem + 128, sw128_desc(al + ao), mn_desc16(bh + bo), kIdescN64_K_MN, 1u);
    }
    umma_commit(c.acc);
  }
  wait_acc(c);
  {
    const float tot = s_x[0][row] + s_x[1][row];
    const float inv = tot > 0.f ? 1.0f / tot : 0.f;
    float o[32];
    tmem_ld32(trow + 128u + (uint32_t)(half * 32), o);
    if (live) {
      float* dst = y.attn + (size_t)(c.r0 + row) * 128 + h * 64 + half * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + j) = make_float4(o[j] * inv, o[j + 1] * inv, o[j + 2] * inv, o[j + 3] * inv);
    }
  }
  tc_fence_before();
  __syncthreads();                                            // smem images, s_x and TMEM are free again
}

__global__ void __launch_bounds__(kFT, 2) sasrec_fwd_fused_kernel(const __grid_constant__ FusedFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[5];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_start[128], s_seq[128], s_pad[128], s_id[128];
  __shared__ float s_x[2][128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = a.tiles[0];
  if ((int)blockIdx.x >= n_tiles) return;

  Ctx c;
  c.smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  c.full = &bars[0]; c.empty = &bars[2]; c.acc = &bars[4];
  c.fetched = 0; c.used = 0; c.n_acc = 0;
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = tmem_slot;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b0 = a.tiles[1 + tile], b1 = a.tiles[2 + tile];
    c.r0 = a.tok_off[b0];
    c.R = a.tok_off[b1] - c.r0;
    if (c.R <= 0) continue;                                   // CTA-uniform
    // ---- row metadata (same for every layer and head) ----
    if (tid < 128) {
      int st = 0, sq = 0, pd = 1, id = 0;
      if (tid < c.R) {
        sq = a.row_seq[c.r0 + tid];
        const int off = a.tok_off[sq];
        st = off - c.r0;
        id = (int)a.in_ids[(size_t)sq * a.L + (c.r0 + tid - off)];
        pd = id == 0;
      }
      s_start[tid] = st; s_seq[tid] = sq; s_pad[tid] = pd; s_id[tid] = id;
    }
    trace(1);
    if (tid == 0) prefetch_chunk(c, a.layer[0].in_hi, a.layer[0].in_lo, 384, 0);
    __syncthreads();
    trace(2);
    // ---- x0 = dropout(E[id] + P[t]) -> global (backward needs it) and the first A operand ----
    {
      const int chunk = tid & 15, rsub = tid >> 4;
#pragma unroll 2
      for (int it = 0; it < 8; ++it) {
        const int row = it * 16 + rsub;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (row < c.R) {
          const int m = c.r0 + row;
          const float* e = a.table + (size_t)s_id[row] * 128 + chunk * 8;
          const float* p = a.pos + (size_t)(row - s_start[row]) * 128 + chunk * 8;
          v0 = *reinterpret_cast<const float4*>(e); v1 = *reinterpret_cast<const float4*>(e + 4);
          const float4 p0 = *reinterpret_cast<const float4*>(p), p1 = *reinterpret_cast<const float4*>(p + 4);
          const uint32_t idx = (uint32_t)m * 128u + (uint32_t)(chunk * 8);
          const float4 f0 = a.d_embed.factor4(idx), f1 = a.d_embed.factor4(idx + 4);
          v0.x = (v0.x + p0.x) * f0.x; v0.y = (v0.y + p0.y) * f0.y; v0.z = (v0.z + p0.z) * f0.z; v0.w = (v0.w + p0.w) * f0.w;
          v1.x = (v1.x + p1.x) * f1.x; v1.y = (v1.y + p1.y) * f1.y; v1.z = (v1.z + p1.z) * f1.z; v1.w = (v1.w + p1.w) * f1.w;
          float* dst = a.x0 + (size_t)m * 128 + chunk * 8;
          *reinterpret_cast<float4*>(dst) = v0; *reinterpret_cast<float4*>(dst + 4) = v1;
        }
        uint4 hh, ll;
        split_bf16x8(v0, v1, hh, ll);
        const uint32_t off = (uint32_t)(chunk >> 3) * kImg + sw128_offset((uint32_t)row, (uint32_t)((chunk & 7) * 8));
        *reinterpret_cast<uint4*>(c.smem + off) = hh;
        *reinterpret_cast<uint4*>(c.smem + 2 * kImg + off) = ll;
      }
    }
    const float* xin = a.x0;
    for (int l = 0; l < a.n_layer; ++l) {
      const FusedLayer& y = a.layer[l];
      // ---- QKV projection: A operand already staged (embedding stage or the previous layer's LN2 epilogue) ----
      sync_for_mma();
#pragma unroll 1
      for (int ch = 0; ch < 3; ++ch) {
        trace(100 * l + 10 + ch);
        if (tid == 0) {
          if (ch < 2) issue_chunk(c, y.in_hi, y.in_lo, 384, ch * 128, 0, y.in_hi, y.in_lo, 384, (ch + 1) * 128);
          else issue_chunk(c, y.in_hi, y.in_lo, 384, ch * 128, 0, nullptr, nullptr, 0, 0);
        }
        wait_acc(c);
        epi_linear<false>(c, 0, y.in_b + ch * 128, y.qkv + ch * 128, 384, y.d_ffn_h);
        tc_fence_before();
        __syncthreads();                                      // accumulator free; qkv rows visible to the whole CTA
        tc_fence_after();
      }
      // ---- attention (per head; Q/K/V slices re-read from the rows this CTA just wrote) ----
      trace(100 * l + 20);
      attention_head(c, a, y, 0, s_start, s_seq, s_pad, s_x);
      trace(100 * l + 21);
      attention_head(c, a, y, 1, s_start, s_seq, s_pad, s_x);
      trace(100 * l + 30);
      // ---- out-proj + dropout + residual + LN1 (stages x1 as the FFN-up operand) ----
      if (tid == 0) prefetch_chunk(c, y.out_hi, y.out_lo, 128, 0);
      stage_a_global(c, y.attn, 128);
      sync_for_mma();
      if (tid == 0) issue_chunk(c, y.out_hi, y.out_lo, 128, 0, 0, y.w1_hi, y.w1_lo, 128, 0);
      wait_acc(c);
      epi_ln<true>(c, 0, y.out_b, xin, y.d_attn_out, y.g1, y.be1, a.ln_eps, y.z1, y.st1, y.x1, s_x);
      sync_for_mma();
      trace(100 * l + 40);
      // ---- FFN up (+bias -> pre) ; stages dropout(gelu(pre)) as the FFN-down operand ----
      if (tid == 0) issue_chunk(c, y.w1_hi, y.w1_lo, 128, 0, 128, y.w2_hi, y.w2_lo, 128, 0);
      wait_acc(c);
      epi_linear<true>(c, 128, y.b1, y.pre, 128, y.d_ffn_h);
      sync_for_mma();
      trace(100 * l + 50);
      // ---- FFN down + dropout + residual + LN2 (stages x2 as the next layer's QKV operand) ----
      const bool more = l + 1 < a.n_layer;
      if (tid == 0) {
        if (more) issue_chunk(c, y.w2_hi, y.w2_lo, 128, 0, 0, a.layer[l + 1].in_hi, a.layer[l + 1].in_lo, 384, 0);
        else issue_chunk(c, y.w2_hi, y.w2_lo, 128, 0, 0, nullptr, nullptr, 0, 0);
      }
      wait_acc(c);
      if (more) epi_ln<true>(c, 0, y.b2, y.x1, y.d_ffn_out, y.g2, y.be2, a.ln_eps, y.z2, y.st2, y.x2, s_x);
      else epi_ln<false>(c, 0, y.b2, y.x1, y.d_ffn_out, y.g2, y.be2, a.ln_eps, y.z2, y.st2, y.x2, s_x);
      xin = y.x2;
    }
    trace(9000);
    tc_fence_before();
    __syncthreads();                                          // before the next tile reuses smem / TMEM / metadata
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(c.tmem, 256);
}

}  // namespace

// debug only: install (or clear with null) a host-mapped int buffer of 4 ints per CTA for the progress trace
int fused_fwd_set_trace(int* host_mapped) {
  if (cudaMemcpyToSymbol(g_trace, &host_mapped, sizeof(int*)) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "fused trace");
    return DR4SR_ECUDA;
  }
  return DR4SR_OK;
}

bool fused_fwd_supported(int L, int D, int F, int n_head) { return D == 128 && F == 128 && n_head == 2 && L >= 1 && L <= 64; }
int fused_tiles_cap(int B, int L) { return (B * L) / 64 + 4; }   // two consecutive greedy tiles hold > 128 rows

int launch_fused_tiles(const int32_t* tok_off, int B, int32_t* tiles, cudaStream_t st) {
  const size_t smem = (size_t)(B + 1) * sizeof(int32_t);
  if (smem > 200 * 1024) return DR4SR_EINVAL;
  ProfScope prof("fused_tiles", st);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(fused_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "fused_tiles smem attribute");
    return DR4SR_ECUDA;
  }
  fused_tiles_kernel<<<1, 1024, smem, st>>>(tok_off, B, tiles);
  DR4SR_LAUNCH_CHECK("fused_tiles_kernel");
  return DR4SR_OK;
}

int launch_sasrec_fwd_fused(const FusedFwdHost& h, cudaStream_t st) {
  if (!fused_fwd_supported(h.L, 128, 128, 2) || h.n_layer < 1 || h.n_layer > 8) return DR4SR_EINVAL;
  FusedFwdArgs a{};
  a.table = h.table; a.pos = h.pos; a.in_ids = h.in_ids; a.tok_off = h.tok_off; a.row_seq = h.row_seq; a.tiles = h.tiles;
  a.x0 = h.x0; a.n_layer = h.n_layer; a.L = h.L; a.ln_eps = h.ln_eps; a.scale = 0.125f; a.d_embed = h.d_embed;
  for (int l = 0; l < h.n_layer; ++l) {
    const FusedLayerHost& s = h.layer[l];
    FusedLayer& d = a.layer[l];
    d.qkv = s.qkv; d.attn = s.attn; d.z1 = s.z1; d.st1 = s.st1; d.x1 = s.x1; d.pre = s.pre; d.z2 = s.z2; d.st2 = s.st2; d.x2 = s.x2;
    d.in_hi = s.img[0]; d.in_lo = s.img[1]; d.out_hi = s.img[2]; d.out_lo = s.img[3];
    d.w1_hi = s.img[4]; d.w1_lo = s.img[5]; d.w2_hi = s.img[6]; d.w2_lo = s.img[7];
    d.in_b = s.in_b; d.out_b = s.out_b; d.b1 = s.b1; d.b2 = s.b2; d.g1 = s.g1; d.be1 = s.be1; d.g2 = s.g2; d.be2 = s.be2;
    d.d_attn_p = s.d_attn_p; d.d_attn_out = s.d_attn_out; d.d_ffn_h = s.d_ffn_h; d.d_ffn_out = s.d_ffn_out;
  }
  ProfScope prof("sasrec_fwd_fused", st);
  if (cudaFuncSetAttribute(sasrec_fwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "sasrec_fwd_fused smem attribute");
    return DR4SR_ECUDA;
  }
  const int cap = fused_tiles_cap(h.B, h.L);
  const int grid = cap < 2 * kNumSMs ? cap : 2 * kNumSMs;
  sasrec_fwd_fused_kernel<<<grid, kFT, kFusedSmem, st>>>(a);
  DR4SR_LAUNCH_CHECK("sasrec_fwd_fused_kernel");
  return DR4SR_OK;
}

}  // namespace dr4sr
