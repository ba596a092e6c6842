// fused_fwd.cu -- the whole SASRec encoder forward as ONE persistent kernel (D = F = 128, 2 heads of 64), and (further
// down, opt-in schedule 2) the backward of a layer's position-wise half as one persistent kernel per layer.
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
#ifdef DR4SR_TRACE   // timeline of the first 8 CTAs: (code, cycles since CTA start) pairs kept in shared memory, dumped at exit
#define TRACE(code) do { if (threadIdx.x == 0 && c.tr_n < 126) { c.tr_buf[2 * c.tr_n] = (code); \
    c.tr_buf[2 * c.tr_n + 1] = (int)(clock64() - c.tr_t0); ++c.tr_n; } } while (0)
#else
#define TRACE(code) do { } while (0)
#endif
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#ifdef DR4SR_TESTWAIT   // experiment: non-suspending poll
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#else
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
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
  float *qkv, *attn, *z1, *st1, *x1, *hm, *gp, *z2, *st2, *x2;   // saved activations (packed rows); hm = mask * gelu(pre), gp = mask * gelu'(pre)
  const uint16_t *in_hi, *in_lo, *out_hi, *out_lo, *w1_hi, *w1_lo, *w2_hi, *w2_lo;   // weight images
  const float *in_b, *out_b, *b1, *b2, *g1, *be1, *g2, *be2;
  Dropout d_attn_p, d_attn_out, d_ffn_h, d_ffn_out;
};
struct FusedFwdArgs {
  ShardView table; const float* pos;
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
  extern __shared__ int32_t s_off[];            // [B + 1] offsets, then [B] next-tile-start of a tile opened at sequence b
  int32_t* s_nxt = s_off + (B + 1);
  for (int i = threadIdx.x; i <= B; i += blockDim.x) s_off[i] = tok_off[i];
  __syncthreads();
  // greedy packing: a tile opened at sequence s closes before the first b > s with off[b + 1] - off[s] > 128.  Every b
  // finds its successor by binary search in parallel; one thread then follows the ~T/110 links.
  for (int s0 = threadIdx.x; s0 < B; s0 += blockDim.x) {
    const int lim = s_off[s0] + 128;
    int lo = s0 + 1, hi = B + 1;                 // smallest j in [s0 + 1, B] with off[j] > lim, or B + 1
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_off[mid] > lim) hi = mid; else lo = mid + 1;
    }
    s_nxt[s0] = lo > B ? B : lo - 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int k = 0, s0 = 0;
    tiles[1] = 0;
    while (s0 < B) {
      s0 = s_nxt[s0];
      tiles[1 + ++k] = s0;
    }
    tiles[0] = k;
  }
}

// ---- shared state of one CTA -------------------------------------------------------------------------
// TMEM (256 columns): [0,128) accumulator (GEMMs, scores S, attention output over the dead S), [128,256) the fp32
// residual stream of the tile ("park": x0 / x1 / x2 rows, read back by the LayerNorm epilogues instead of HBM).
struct Ctx {
  uint8_t* smem;            // 1 KB aligned dynamic smem: [0,64K) A / Q,K / P images, [64K,96K) weight ring / V images
  uint64_t *full;           // [2] weight-ring slot filled (two completions per chunk and slot: parity 0 then 1)
  uint64_t *free01;         // both slots released by the first half of a chunk (one completion per chunk)
  uint64_t* acc;            // accumulator-ready barrier (also: the ring is idle again)
  uint32_t tmem;            // TMEM base
  uint32_t pchunk, ppieces; // weight ring, producer side (thread kProducer only): chunks streamed; pieces of the current chunk queued
  uint32_t n_acc;           // completed phases of `acc` (uniform across the CTA)
  int r0, R;                // packed rows [r0, r0 + R) of the tile
  // this thread's epilogue identity: tile row quad*32+lane, 64-column half warp>>2
  int row, half, m;
  bool live;
  uint32_t trow;            // TMEM address of (row's lane group, column half*64)
#ifdef DR4SR_TRACE
  int tr_n; long long tr_t0; int* tr_buf;
#endif
};
constexpr uint32_t kPark = 128;

__device__ __forceinline__ uint8_t* a_hi(const Ctx& c, int kb) { return c.smem + (uint32_t)kb * kImg; }
__device__ __forceinline__ uint8_t* a_lo(const Ctx& c, int kb) { return c.smem + (uint32_t)(2 + kb) * kImg; }
__device__ __forceinline__ uint8_t* ring(const Ctx& c, uint32_t slot) { return c.smem + (4 + slot) * kImg; }

// Weight ring: two 16 KB slots.  A GEMM chunk (128 output columns, K = 128) streams four pieces through it:
//   0: W_hi k-block 0 -> slot 0, 1: W_lo k-block 0 -> slot 1, 2: W_hi k-block 1 -> slot 0, 3: W_lo k-block 1 -> slot 1.
// The ring is fed by its own thread (lane 0 of warp 1): refills must wait for the UMMAs that read a slot's previous piece,
// and with both roles in one thread those waits sat between UMMA issues (~9 K cycles per chunk, profiles/r2_fused_fwd_timeline.md).
// Thread 0 only waits for `full`, issues, and commits twice per chunk (`free01` after the first half, `acc` at the end).
constexpr int kProducer = 32;
__device__ __forceinline__ void ring_put(Ctx& c, uint32_t slot, const uint8_t* src) {
  mbar_expect_tx(&c.full[slot], kImg);
  bulk_g2s(ring(c, slot), src, kImg, &c.full[slot]);
}
// piece p of a 128-column chunk n0 of a logical [N_total, 128] weight: p = 2 * kb + (lo ? 1 : 0)
__device__ __forceinline__ const uint8_t* piece_src(const uint16_t* hi, const uint16_t* lo, int N_total, int n0, int p) {
  const uint8_t* base = reinterpret_cast<const uint8_t*>((p & 1) ? lo : hi);
  return base + ((size_t)(p >> 1) * N_total + n0) * 128;
}
// producer: pieces 0, 1 of the next chunk.  The caller guarantees that no UMMA still reads the ring (every thread has
// passed wait_acc of the previous chunk, or produce_chunk has just waited for it).  Idempotent.
__device__ __forceinline__ void prefetch_chunk(Ctx& c, const uint16_t* hi, const uint16_t* lo, int N_total, int n0) {
  if (c.ppieces == 0u) {
    ring_put(c, 0u, piece_src(hi, lo, N_total, n0, 0));
    ring_put(c, 1u, piece_src(hi, lo, N_total, n0, 1));
    c.ppieces = 2u;
  }
}
// producer: the rest of the current chunk, then (nhi != null) the first two pieces of the next one once the ring is idle
__device__ __noinline__ void produce_chunk(Ctx& c, const uint16_t* hi, const uint16_t* lo, int N_total, int n0,
                                           const uint16_t* nhi, const uint16_t* nlo, int nN_total, int nn0) {
  prefetch_chunk(c, hi, lo, N_total, n0);
  mbar_wait_b(c.free01, c.pchunk & 1u, 100);
  ring_put(c, 0u, piece_src(hi, lo, N_total, n0, 2));
  ring_put(c, 1u, piece_src(hi, lo, N_total, n0, 3));
  ++c.pchunk;
  c.ppieces = 0u;
  if (nhi) {
    mbar_wait_b(c.acc, c.n_acc & 1u, 101);          // peek: the chunk's last UMMAs are done (wait_acc consumes the phase later)
    prefetch_chunk(c, nhi, nlo, nN_total, nn0);
  }
}

// thread 0: acc[128 x 128] (TMEM columns [0,128)) = A (hi/lo images in smem, K = 128) x the chunk's four weight pieces.
// Operand descriptors advance by integer adds on their lower words (gemm_tc.cuh, desc_lo / umma_lo).
__device__ __noinline__ void issue_chunk(Ctx& c) {
  const uint32_t acc = c.tmem;
  const uint32_t w0 = desc_lo(smem_u32(ring(c, 0u))), w1 = desc_lo(smem_u32(ring(c, 1u)));
  TRACE(1999);
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
    const uint32_t ah = desc_lo(smem_u32(a_hi(c, kb))), al = desc_lo(smem_u32(a_lo(c, kb)));
    mbar_wait_b(&c.full[0], (uint32_t)kb, 200);
    TRACE(2000 + 2 * kb);
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 4; ++k) {                     // W_hi piece: A_hi W_hi + A_lo W_hi
      if (kb == 0 && k == 0) umma_lo<false>(acc, ah, w0, kIdesc); else umma_lo<true>(acc, ah + 2u * k, w0 + 2u * k, kIdesc);
      umma_lo<true>(acc, al + 2u * k, w0 + 2u * k, kIdesc);
    }
    mbar_wait_b(&c.full[1], (uint32_t)kb, 201);
    TRACE(2001 + 2 * kb);
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_lo<true>(acc, ah + 2u * k, w1 + 2u * k, kIdesc);      // W_lo piece: A_hi W_lo
    TRACE(2010 + kb);
    umma_commit(kb == 0 ? c.free01 : c.acc);
  }
  TRACE(2020);
}
// one GEMM chunk: thread 0 issues, the producer streams, everybody else goes straight to wait_acc
__device__ __forceinline__ void run_chunk(Ctx& c, const uint16_t* hi, const uint16_t* lo, int N_total, int n0,
                                          const uint16_t* nhi, const uint16_t* nlo, int nN_total, int nn0) {
  if (threadIdx.x == 0) issue_chunk(c);
  else if (threadIdx.x == kProducer) produce_chunk(c, hi, lo, N_total, n0, nhi, nlo, nN_total, nn0);
}

// all threads: wait for the accumulator committed by the most recent issue
__device__ __forceinline__ void wait_acc(Ctx& c) {
  __syncwarp();                                   // lanes 1..31 of warp 0 park here while lane 0 issues (no spinning beside it)
  mbar_wait_b(c.acc, c.n_acc & 1u, 300);
  ++c.n_acc;
  tc_fence_after();
}

// ---- TMEM 16-column accessors (thread = its own lane's row) ---------------------------------------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// two 16-column loads in flight, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, float* va, uint32_t tb, float* vb) {
  uint32_t* r = reinterpret_cast<uint32_t*>(va);
  uint32_t* q = reinterpret_cast<uint32_t*>(vb);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(ta));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]),
        "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(tb));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---- bf16 hi/lo split with packed conversions (bit-identical to split_bf16x8) -------------------------------------
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {   // a -> low half, b -> high half
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hb), "f"(a - ha));
}
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  split2(x[0], x[1], hi.x, lo.x); split2(x[2], x[3], hi.y, lo.y);
  split2(x[4], x[5], hi.z, lo.z); split2(x[6], x[7], hi.w, lo.w);
}
// 16 fp32 values of tile row `row`, columns [n0, n0 + 16) (n0 % 16 == 0) of a 128-wide operand -> hi / lo images
// laid out as [hi kb0, hi kb1, lo kb0, lo kb1] from `base`
__device__ __forceinline__ void store_image16(const float* v, int row, int n0, uint8_t* base) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint4 h, l;
    split8(v + q * 8, h, l);
    const int n = n0 + q * 8;
    const uint32_t off = (uint32_t)(n >> 6) * kImg + sw128_offset((uint32_t)row, (uint32_t)(n & 63));
    *reinterpret_cast<uint4*>(base + off) = h;
    *reinterpret_cast<uint4*>(base + 2 * kImg + off) = l;
  }
}
__device__ __forceinline__ void store_image16_zero(int row, int n0, uint8_t* base) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int n = n0 + q * 8;
    const uint32_t off = (uint32_t)(n >> 6) * kImg + sw128_offset((uint32_t)row, (uint32_t)(n & 63));
    *reinterpret_cast<uint4*>(base + off) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(base + 2 * kImg + off) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// ---- dropout factors of one 128-wide row (idx = m * 128 + n): same draws as Dropout::factor4, with the row part
// of draw32 -- mix32(key ^ (pair_index >> 7)) = mix32(key ^ (m >> 1)) -- hoisted out of the element loop ----------
struct RowDrop {
  uint32_t key, rk, thresh, base;   // base = pair index of the row's first element (m * 64)
  float scale;
  __device__ __forceinline__ void init(const Dropout& d, uint32_t m) {
    key = d.key; thresh = d.thresh; scale = d.scale; base = m * 64u; rk = mix32(d.key ^ (m >> 1));
  }
  // f[j] = factor of column n0 + j, j < 16, n0 % 2 == 0
  __device__ __forceinline__ void factors16(int n0, float* f) const {
    if (thresh == 0u) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = scale;
      return;
    }
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const uint32_t h = mix32((base + (uint32_t)((n0 + j) >> 1)) * 0x9E3779B1u + key) ^ rk;
      f[j] = (h & 0xFFFFu) >= thresh ? scale : 0.f;
      f[j + 1] = (h >> 16) >= thresh ? scale : 0.f;
    }
  }
};

// fp32 [R x 128] rows (row stride ld) from global -> A images (rows >= R zero)
__device__ __forceinline__ void stage_a_global(const Ctx& c, const float* src, int ld) {
  const int chunk = threadIdx.x & 15, rsub = threadIdx.x >> 4;     // 16 chunks of 8 floats, 16 rows per pass
#pragma unroll 1
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
      const float x[8] = {v[it][0].x, v[it][0].y, v[it][0].z, v[it][0].w, v[it][1].x, v[it][1].y, v[it][1].z, v[it][1].w};
      uint4 h, l;
      split8(x, h, l);
      const uint32_t off = (uint32_t)(chunk >> 3) * kImg + sw128_offset((uint32_t)row, (uint32_t)((chunk & 7) * 8));
      *reinterpret_cast<uint4*>(c.smem + off) = h;
      *reinterpret_cast<uint4*>(c.smem + 2 * kImg + off) = l;
    }
  }
}

// fp32 head slices Q, K, V [R x 64] of the packed qkv rows -> their hi / lo image pairs (rows >= R zero)
__device__ __forceinline__ void stage_heads(const Ctx& c, const float* qkv_head) {
  const int chunk = threadIdx.x & 7, rsub = threadIdx.x >> 3;      // 32 rows per pass
#pragma unroll 1
  for (int which = 0; which < 3; ++which) {                        // Q, K, V
    const float* src = qkv_head + which * 128;
    uint8_t* hi = c.smem + (uint32_t)(2 * which) * kImg;           // q_hi, q_lo, k_hi, k_lo, v_hi, v_lo
    uint8_t* lo = hi + kImg;
    float4 v[4][2];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = it * 32 + rsub;
      if (row < c.R) {
        const float* p = src + (size_t)row * 384 + chunk * 8;
        v[it][0] = *reinterpret_cast<const float4*>(p);
        v[it][1] = *reinterpret_cast<const float4*>(p + 4);
      } else {
        v[it][0] = v[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = it * 32 + rsub;
      const float x[8] = {v[it][0].x, v[it][0].y, v[it][0].z, v[it][0].w, v[it][1].x, v[it][1].y, v[it][1].z, v[it][1].w};
      uint4 h, l;
      split8(x, h, l);
      const uint32_t off = sw128_offset((uint32_t)row, (uint32_t)(chunk * 8));
      *reinterpret_cast<uint4*>(hi + off) = h;
      *reinterpret_cast<uint4*>(lo + off) = l;
    }
  }
}

// make generic-proxy smem writes visible to the tensor core, order TMEM accesses before later UMMAs, CTA barrier
__device__ __forceinline__ void sync_for_mma() {
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// 8 consecutive floats (32-byte aligned) as one 256-bit store: a full 32 B sector per thread per instruction
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
               "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

// ---- epilogues (thread = tile row me.row, columns me.half*64 .. +63) --------------------------------------------
struct Me { int row, half, m; bool live; uint32_t trow; uint8_t* smem; };   // passed by value: stays in registers
__device__ __forceinline__ Me me_of(const Ctx& c) { return Me{c.row, c.half, c.m, c.live, c.trow, c.smem}; }

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// wait for every outstanding tcgen05.ld; the "+r" ties keep the consumers of v[0..16) behind the wait
__device__ __forceinline__ void tmem_ld_fence(float* v, bool wait) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  if (wait) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
  } else {
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
  }
}

// out[row, n] = acc + bias[n] (QKV projection), or -- `hm` != null, the FFN up-projection -- pre = acc + bias is NOT stored:
// the next A operand hm = mask * gelu(pre) goes to shared memory and to `hm` (the FFN-down weight gradient reads it), and
// gp = mask * gelu'(pre) to `gp` (the backward multiplies by it instead of recomputing erf / exp per element).
__device__ __noinline__ void epi_linear(Me me, const float* s_bias, float* out, int ldo, Dropout dh, float* hm, float* gp) {
  RowDrop rd;
  rd.init(dh, (uint32_t)me.m);
#pragma unroll 1
  for (int g = 0; g < 2; ++g) {
    const int n0 = me.half * 64 + g * 32;
    float v[32];
    tmem_ld16_nowait(me.trow + (uint32_t)(g * 32), v);
    tmem_ld16_nowait(me.trow + (uint32_t)(g * 32 + 16), v + 16);
    tmem_ld_fence(v, true);
    tmem_ld_fence(v + 16, false);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bb = *reinterpret_cast<const float4*>(s_bias + n0 + j);
      v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
    }
    if (!hm) {
      if (me.live) {
        float* orow = out + (size_t)me.m * ldo + me.half * 64 + g * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8) st_global_v8(orow + j, v + j);
      }
      continue;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float f[16], d[16];
      rd.factors16(n0 + q * 16, f);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x = v[q * 16 + j];
        const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
        const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
        v[q * 16 + j] = x * cdf * f[j];                       // mask * gelu(pre)
        d[j] = (cdf + x * pdf) * f[j];                        // mask * gelu'(pre)
      }
      store_image16(v + q * 16, me.row, n0 + q * 16, me.smem);
      if (me.live) {
        float* grow = gp + (size_t)me.m * ldo + n0 + q * 16;
        st_global_v8(grow, d);
        st_global_v8(grow + 8, d + 8);
      }
    }
    if (me.live) {
      float* hrow = hm + (size_t)me.m * ldo + n0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) st_global_v8(hrow + j, v + j);
    }
  }
}

// z = dropout(acc + bias) + residual(park) ; y = LayerNorm(z) -> Z, stats, Y (global), park <- y and, if stage, the
// next A operand.  The thread keeps its 64 z values in registers: one pass over TMEM.  s_b / s_g / s_be: bias, gamma,
// beta in shared memory.
__device__ __noinline__ void epi_ln(Me me, const float* s_b, const float* s_g, const float* s_be, Dropout de, float eps, float* Z,
                                    float* stats, float* Y, bool stage, float (*ln_part)[128]) {
  RowDrop rd;
  rd.init(de, (uint32_t)me.m);
  float z[64];
  float sum = 0.f;
  float* zrow = Z + (size_t)me.m * 128 + me.half * 64;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int n0 = me.half * 64 + g * 16;
    float r[16], f[16];
    tmem_ld16_nowait(me.trow + (uint32_t)(g * 16), z + g * 16);
    tmem_ld16_nowait(me.trow + kPark + (uint32_t)(g * 16), r);
    rd.factors16(n0, f);
    tmem_ld_fence(z + g * 16, true);
    tmem_ld_fence(r, false);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      z[g * 16 + j] = (z[g * 16 + j] + s_b[n0 + j]) * f[j] + r[j];
      sum += z[g * 16 + j];
    }
    if (me.live) {
      st_global_v8(zrow + g * 16, z + g * 16);
      st_global_v8(zrow + g * 16 + 8, z + g * 16 + 8);
    }
  }
  ln_part[me.half][me.row] = sum;
  __syncthreads();
  const float mu = (ln_part[0][me.row] + ln_part[1][me.row]) * (1.0f / 128.0f);
  __syncthreads();
  float var = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) { const float d = z[j] - mu; var = fmaf(d, d, var); }
  ln_part[me.half][me.row] = var;
  __syncthreads();
  const float rstd = rsqrtf((ln_part[0][me.row] + ln_part[1][me.row]) * (1.0f / 128.0f) + eps);
  float* yrow = Y + (size_t)me.m * 128 + me.half * 64;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int n0 = me.half * 64 + g * 16;
    float* v = z + g * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (v[j] - mu) * rstd * s_g[n0 + j] + s_be[n0 + j];
    if (me.live) {
      st_global_v8(yrow + g * 16, v);
      st_global_v8(yrow + g * 16 + 8, v + 8);
    }
    tmem_st16(me.trow + kPark + (uint32_t)(g * 16), v);
    if (stage) store_image16(v, me.row, n0, me.smem);
  }
  if (me.live && me.half == 0) { stats[2 * me.m] = mu; stats[2 * me.m + 1] = rstd; }
}

// bit j set <=> key c0 + j lies in the row's causal window [start, row] (and the row exists)
__device__ __forceinline__ uint32_t key_bits(bool live, int start, int row, int c0) {
  const int lo = max(start - c0, 0), hi = min(row - c0, 15);
  if (!live || hi < lo) return 0u;
  return ((2u << hi) - 1u) & ~((1u << lo) - 1u);
}

// ---- attention of one head over the tile (block-diagonal causal mask) -----------------------------------
// s_padbits[g]: bit j set <=> key 16 g + j is a real (non-pad) item
__device__ __forceinline__ void attention_head(Ctx& c, const FusedFwdArgs& a, const FusedLayer& y, int h, const int* s_start,
                                               const int* s_seq, const uint32_t* s_padbits, float (*s_x)[128]) {
  const int tid = threadIdx.x;
  uint8_t *q_hi = c.smem, *q_lo = c.smem + kImg, *k_hi = c.smem + 2 * kImg, *k_lo = c.smem + 3 * kImg;
  uint8_t *v_hi = c.smem + 4 * kImg, *v_lo = c.smem + 5 * kImg;
  stage_heads(c, y.qkv + (size_t)c.r0 * 384 + h * 64);
  sync_for_mma();
  TRACE(1000 + 10 * h + 1);
  if (tid == 0) {                                             // S = Q K^T -> TMEM columns [0,128)
    const uint32_t ah = desc_lo(smem_u32(q_hi)), al = desc_lo(smem_u32(q_lo)), bh = desc_lo(smem_u32(k_hi)), bl = desc_lo(smem_u32(k_lo));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k == 0) umma_lo<false>(c.tmem, ah, bh, kIdesc); else umma_lo<true>(c.tmem, ah + 2u * k, bh + 2u * k, kIdesc);
      umma_lo<true>(c.tmem, ah + 2u * k, bl + 2u * k, kIdesc);
      umma_lo<true>(c.tmem, al + 2u * k, bh + 2u * k, kIdesc);
    }
    umma_commit(c.acc);
  }
  wait_acc(c);
  TRACE(1000 + 10 * h + 2);
  const int row = c.row, start = s_start[row];
  // keys of this row: [start, row] minus pads.  A group of 16 keys is visited when it intersects the window of some
  // row of the warp (warp-uniform, tcgen05.ld is warp-collective).
  uint32_t live_groups = 0;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c0 = c.half * 64 + g * 16;
    const bool mine = c.live && start <= c0 + 15 && row >= c0;
    if (__any_sync(0xffffffffu, mine)) live_groups |= 1u << g;
  }
  float mx = -INFINITY;
#pragma unroll 1
  for (int g = 0; g < 4; ++g) {
    if (!(live_groups >> g & 1u)) continue;
    const int c0 = c.half * 64 + g * 16;
    float s[16];
    tmem_ld16(c.trow + (uint32_t)(g * 16), s);
    const uint32_t okb = key_bits(c.live, start, row, c0) & s_padbits[c0 >> 4];
#pragma unroll
    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, (okb >> j & 1u) ? s[j] * a.scale : -INFINITY);
  }
  s_x[c.half][row] = mx;
  __syncthreads();
  mx = fmaxf(s_x[0][row], s_x[1][row]);
  __syncthreads();
  // P' = dropout(exp(s - max)) un-normalised (the 1/sum is applied to the output rows), written over the dead Q / K images
  float sum = 0.f;
  const Dropout& dp = y.d_attn_p;
  const uint32_t dbase = (uint32_t)(s_seq[row] * 2 + h) * (uint32_t)(a.L * a.L) + (uint32_t)(row - start) * (uint32_t)a.L;
#pragma unroll 1
  for (int g = 0; g < 4; ++g) {
    const int c0 = c.half * 64 + g * 16;
    if (!(live_groups >> g & 1u)) { store_image16_zero(row, c0, c.smem); continue; }
    float s[16];
    tmem_ld16(c.trow + (uint32_t)(g * 16), s);
    // dropout draws of elements idx0 .. idx0 + 15: element i uses half-word (i & 1) of draw32(key, i >> 1)
    const uint32_t idx0 = dbase + (uint32_t)(c0 - start);
    uint32_t e16[8];
    if (dp.thresh != 0u) {
      // draw32(key, p) = mix32(p * C + key) ^ mix32(key ^ (p >> 7)); the nine pair indices p0 .. p0+8 span at most two values of p >> 7
      const uint32_t p0 = idx0 >> 1, odd = idx0 & 1u, hi0 = p0 >> 7;
      const uint32_t ha = mix32(dp.key ^ hi0), hb = mix32(dp.key ^ (hi0 + 1u));
      uint32_t d[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const uint32_t pk = p0 + (uint32_t)k;
        d[k] = mix32(pk * 0x9E3779B1u + dp.key) ^ ((pk >> 7) == hi0 ? ha : hb);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) e16[k] = odd ? __funnelshift_r(d[k], d[k + 1], 16) : d[k];
    }
    const uint32_t okb = mx > -INFINITY ? key_bits(c.live, start, row, c0) & s_padbits[c0 >> 4] : 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const bool ok = okb >> j & 1u;
      const float e = ok ? expf(s[j] * a.scale - mx) : 0.f;
      sum += e;
      float f = dp.scale;
      if (dp.thresh != 0u) {
        const uint32_t hw = (j & 1) ? (e16[j >> 1] >> 16) : (e16[j >> 1] & 0xFFFFu);
        f = hw >= dp.thresh ? dp.scale : 0.f;
      }
      s[j] = e * f;
    }
    store_image16(s, row, c0, c.smem);
  }
  s_x[c.half][row] = sum;
  sync_for_mma();
  TRACE(1000 + 10 * h + 3);
  if (tid == 0) {                                             // O = P' V -> TMEM columns [0,64) (over the dead scores)
    const uint32_t ah = desc_lo(smem_u32(c.smem)), al = desc_lo(smem_u32(c.smem + 2 * kImg)), bh = desc_lo(smem_u32(v_hi)),
                   bl = desc_lo(smem_u32(v_lo));
#pragma unroll
    for (int k = 0; k < 8; ++k) {                             // 16 keys per UMMA: A step 32 B inside the k-block image, B step 2048 B
      const uint32_t ao = (uint32_t)(k >> 2) * (kImg >> 4) + (uint32_t)(k & 3) * 2u, bo = (uint32_t)k * 128u;
      if (k == 0) umma_lo<false>(c.tmem, ah, bh, kIdescN64_K_MN); else umma_lo<true>(c.tmem, ah + ao, bh + bo, kIdescN64_K_MN);
      umma_lo<true>(c.tmem, ah + ao, bl + bo, kIdescN64_K_MN);
      umma_lo<true>(c.tmem, al + ao, bh + bo, kIdescN64_K_MN);
    }
    umma_commit(c.acc);
  }
  wait_acc(c);
  TRACE(1000 + 10 * h + 4);
  {
    const float tot = s_x[0][row] + s_x[1][row];
    const float inv = tot > 0.f ? 1.0f / tot : 0.f;
    const uint32_t to = c.trow - (uint32_t)(c.half * 64) + (uint32_t)(c.half * 32);   // this thread's 32 of the 64 output columns
    float* dst = y.attn + (size_t)c.m * 128 + h * 64 + c.half * 32;
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
      float o[16];
      tmem_ld16(to + (uint32_t)(g * 16), o);
      if (c.live) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] *= inv;
        st_global_v8(dst + g * 16, o);
        st_global_v8(dst + g * 16 + 8, o + 8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();                                            // smem images, s_x and TMEM are free again
}

// per-layer small parameters staged in shared memory (floats): in_b[384] out_b b1 b2 g1 be1 g2 be2 [128 each]
constexpr int kParIn = 0, kParOut = 384, kParB1 = 512, kParB2 = 640, kParG1 = 768, kParBe1 = 896, kParG2 = 1024, kParBe2 = 1152,
              kParTotal = 1280;

__global__ void __launch_bounds__(kFT, 2) sasrec_fwd_fused_kernel(const __grid_constant__ FusedFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[5];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_start[128], s_seq[128], s_id[128];
  __shared__ uint32_t s_padbits[8];
  __shared__ float s_x[2][128];
  __shared__ __align__(16) float s_par[kParTotal];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = a.tiles[0];
  if ((int)blockIdx.x >= n_tiles) return;

  Ctx c;
  c.smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  c.full = &bars[0]; c.free01 = &bars[2]; c.acc = &bars[3];
  c.pchunk = 0; c.ppieces = 0; c.n_acc = 0;
#ifdef DR4SR_TRACE
  __shared__ int s_trace[256];
  c.tr_n = 0; c.tr_t0 = clock64(); c.tr_buf = s_trace;
#endif
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = tmem_slot;
  c.row = (warp & 3) * 32 + lane;
  c.half = warp >> 2;
  c.trow = c.tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(c.half * 64);

#pragma unroll 1
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b0 = a.tiles[1 + tile], b1 = a.tiles[2 + tile];
    c.r0 = a.tok_off[b0];
    c.R = a.tok_off[b1] - c.r0;
    if (c.R <= 0) continue;                                   // CTA-uniform
    c.live = c.row < c.R;
    c.m = c.r0 + c.row;
    // ---- row metadata (same for every layer and head) ----
    if (tid < 128) {
      int st = 0, sq = 0, pd = 1, id = 0;
      if (tid < c.R) {
        sq = a.row_seq[c.r0 + tid];
        const int off = a.tok_off[sq];
        st = off - c.r0;
        id = (int)a.in_ids[(size_t)sq * a.L + (c.r0 + tid - off)];
        pd = id == 0;
        // pull the whole 512 B table row into L2 as ONE contiguous request: the row-per-thread gather below would
        // otherwise reach DRAM as scattered 32 B sectors
        if (a.table.is_local(id))       // (rows of other ranks come over NVLink: nothing to prefetch into this GPU's L2)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.table.row(id, 128)), "r"(512) : "memory");
      }
      s_start[tid] = st; s_seq[tid] = sq; s_id[tid] = id;
      const uint32_t real = __ballot_sync(0xffffffffu, !pd);        // warps 0..3 are complete: 32 keys each
      if (lane == 0) { s_padbits[2 * warp] = real & 0xFFFFu; s_padbits[2 * warp + 1] = real >> 16; }
    }
    TRACE(1);
    if (tid == kProducer) prefetch_chunk(c, a.layer[0].in_hi, a.layer[0].in_lo, 384, 0);
    __syncthreads();
    TRACE(2);
    // ---- x0 = dropout(E[id] + P[t]) -> global (the backward needs it), park (residual) and the first A operand ----
    {
      RowDrop rd;
      rd.init(a.d_embed, (uint32_t)c.m);
      const float* e = a.table.row(s_id[c.row], 128) + c.half * 64;     // local HBM, or the owner's HBM through peer memory
      const float* p = a.pos + (size_t)(c.row - s_start[c.row]) * 128 + c.half * 64;
      float ev[64];                                             // the thread's 64 table floats: every (256-bit) load in flight at once
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c.live) ld_global_v8(e + 8 * j, ev + 8 * j);
        else {
#pragma unroll
          for (int q = 0; q < 8; ++q) ev[8 * j + q] = 0.f;
        }
      }
      if (ev[0] == 12345.678f) TRACE(3);           // (forces the first load to land before the timestamp below)
      TRACE(4);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int n0 = c.half * 64 + g * 16;
        float v[16], f[16];
        if (c.live) {
          rd.factors16(n0, f);
          float pv[16];
          ld_global_v8(p + g * 16, pv);
          ld_global_v8(p + g * 16 + 8, pv + 8);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = (ev[g * 16 + j] + pv[j]) * f[j];
          st_global_v8(a.x0 + (size_t)c.m * 128 + n0, v);
          st_global_v8(a.x0 + (size_t)c.m * 128 + n0 + 8, v + 8);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        tmem_st16(c.trow + kPark + (uint32_t)(g * 16), v);
        store_image16(v, c.row, n0, c.smem);
      }
      TRACE(5);
    }
#pragma unroll 1
    for (int l = 0; l < a.n_layer; ++l) {
      const FusedLayer& y = a.layer[l];
      // small parameters of the layer -> shared memory (the previous layer's epilogues are done: barrier at its end)
      for (int i = tid; i < kParTotal; i += kFT) {
        const float* src = i < kParOut ? y.in_b + i
                         : i < kParB1 ? y.out_b + (i - kParOut)
                         : i < kParB2 ? y.b1 + (i - kParB1)
                         : i < kParG1 ? y.b2 + (i - kParB2)
                         : i < kParBe1 ? y.g1 + (i - kParG1)
                         : i < kParG2 ? y.be1 + (i - kParBe1)
                         : i < kParBe2 ? y.g2 + (i - kParG2) : y.be2 + (i - kParBe2);
        s_par[i] = *src;
      }
      TRACE(6);
      // ---- QKV projection: A operand already staged (embedding stage or the previous layer's LN2 epilogue) ----
      sync_for_mma();
#pragma unroll 1
      for (int ch = 0; ch < 3; ++ch) {
        TRACE(100 * l + 10 + ch);
        run_chunk(c, y.in_hi, y.in_lo, 384, ch * 128, ch < 2 ? y.in_hi : nullptr, y.in_lo, 384, (ch + 1) * 128);
        TRACE(100 * l + 13 + ch);
        wait_acc(c);
        TRACE(100 * l + 16 + ch);
        epi_linear(me_of(c), s_par + kParIn + ch * 128, y.qkv + ch * 128, 384, y.d_ffn_h, nullptr, nullptr);
        tc_fence_before();
        __syncthreads();                                      // accumulator free; qkv rows visible to the whole CTA
        tc_fence_after();
      }
      // ---- attention (per head; Q/K/V slices re-read from the rows this CTA just wrote) ----
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        TRACE(100 * l + 20 + h);
        attention_head(c, a, y, h, s_start, s_seq, s_padbits, s_x);
      }
      TRACE(100 * l + 30);
      // ---- out-proj + dropout + residual + LN1 (x1 -> park and the FFN-up operand) ----
      if (tid == kProducer) prefetch_chunk(c, y.out_hi, y.out_lo, 128, 0);
      stage_a_global(c, y.attn, 128);
      sync_for_mma();
      TRACE(100 * l + 31);
      run_chunk(c, y.out_hi, y.out_lo, 128, 0, y.w1_hi, y.w1_lo, 128, 0);
      TRACE(100 * l + 32);
      wait_acc(c);
      TRACE(100 * l + 33);
      epi_ln(me_of(c), s_par + kParOut, s_par + kParG1, s_par + kParBe1, y.d_attn_out, a.ln_eps, y.z1, y.st1, y.x1, true, s_x);
      sync_for_mma();
      TRACE(100 * l + 40);
      // ---- FFN up (+bias -> pre) ; dropout(gelu(pre)) -> the FFN-down operand ----
      run_chunk(c, y.w1_hi, y.w1_lo, 128, 0, y.w2_hi, y.w2_lo, 128, 0);
      TRACE(100 * l + 41);
      wait_acc(c);
      TRACE(100 * l + 42);
      epi_linear(me_of(c), s_par + kParB1, nullptr, 128, y.d_ffn_h, y.hm, y.gp);
      sync_for_mma();
      TRACE(100 * l + 50);
      // ---- FFN down + dropout + residual + LN2 (x2 -> park and the next layer's QKV operand) ----
      const bool more = l + 1 < a.n_layer;
      run_chunk(c, y.w2_hi, y.w2_lo, 128, 0, more ? a.layer[more ? l + 1 : l].in_hi : nullptr, a.layer[more ? l + 1 : l].in_lo, 384, 0);
      TRACE(100 * l + 51);
      wait_acc(c);
      TRACE(100 * l + 52);
      epi_ln(me_of(c), s_par + kParB2, s_par + kParG2, s_par + kParBe2, y.d_ffn_out, a.ln_eps, y.z2, y.st2, y.x2, more, s_x);
      tc_fence_before();
      __syncthreads();                                        // s_par / images / TMEM handed to the next layer (or tile)
      tc_fence_after();
    }
    TRACE(9000);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(c.tmem, 256);
#ifdef DR4SR_TRACE
  if (tid == 0 && g_trace && blockIdx.x < 8)
    for (int i = 0; i < 2 * c.tr_n; ++i) g_trace[blockIdx.x * 256 + i] = c.tr_buf[i];
#endif
}


// =====================================================================================================================
// Backward of the position-wise half of a layer as one persistent kernel over 128-row tiles (no sequence alignment needed):
//   g3   = LN2'(gin; z2)                                     (dz2)
//   dpre = ((g3 . mask_ffn_out) W2) . gp                     (gp = mask_ffn_h . gelu'(pre), saved by the fused forward)
//   dx1  = g3 + dpre W1
//   g1   = LN1'(dx1; z1)                                     (dz1)
//   g2   = (g1 . mask_attn_out) Wo                           (gradient w.r.t. the attention output)
// replacing ln_bwd + gemm_bwd_dpre + gemm_bwd_dx1 + ln_bwd + gemm_bwd_dattn on the critical path.  g3, dpre, g1 also go to
// HBM (the weight-gradient GEMMs consume them on the side stream), and the LayerNorm column partials (d gamma, d beta, bias
// gradients under the two dropouts) are accumulated here, so no separate LayerNorm-backward launch remains.
//
// Two thread mappings.  Everything that touches the TMEM accumulator runs row-per-thread (TMEM lane = tile row): operands of
// those epilogues (gp, g3) are fetched with 256-bit row loads issued BEFORE the accumulator wait, so their latency hides under
// the UMMAs.  The two LayerNorm backwards do not touch TMEM at all and run in the COALESCED mapping -- 16 lanes per row, 32
// bytes per lane, row statistics by shuffles inside the half-warp -- which is what the first version of this kernel lacked
// (its row-per-thread loads of gin / z2 alone were 49 K of 182 K cycles per tile).
struct FusedBwdFfnArgs {
  const float *gin, *z2, *st2, *gp, *z1, *st1, *gamma2, *gamma1;
  const uint16_t *w2_hi, *w2_lo, *w1_hi, *w1_lo, *out_hi, *out_lo;   // images of W2^T, W1^T, Wo^T (backward-data operands)
  float *g3, *dpre, *dx1, *g1, *g2;
  float *part_ln2, *part_ln1;                                       // [gridDim.x][3 * 128] column partials of the two LayerNorms
  const int32_t* counts;
  int T_cap;
  Dropout d_ffn_out, d_attn_out;
};

// dz = LN'(dy; z, stats, gamma) over the tile in the coalesced mapping (chunk = tid & 15: 8 columns, rsub = tid >> 4: one of 16
// rows per pass).  Writes dz (global, 32 B per lane), the next A operand dz . mask(dm) as bf16 hi/lo images, and adds this
// tile's column partials {sum dy xhat, sum dy, sum dz mask} to acc.
// (dy may have been written by this very CTA -- dx1 -- so no pointer here is __restrict__ / read-only-path qualified)
__device__ __forceinline__ void ln_bwd_pass(const Ctx& c, const float* dy, const float* z, const float* stats, const float* s_gamma,
                                            const Dropout& dm, float* dz_out, float (&acc)[3][8]) {
  const int chunk = threadIdx.x & 15, rsub = threadIdx.x >> 4, col0 = chunk * 8;
  float gam[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gam[j] = s_gamma[col0 + j];
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    float4 dv[4][2], zv[4][2];
    float mu[4], rs[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {                            // 4 rows x (dy, z) per lane in flight
      const int row = (half * 4 + it) * 16 + rsub;
      if (row < c.R) {
        const size_t o = (size_t)(c.r0 + row) * 128 + col0;
        dv[it][0] = *reinterpret_cast<const float4*>(dy + o); dv[it][1] = *reinterpret_cast<const float4*>(dy + o + 4);
        zv[it][0] = *reinterpret_cast<const float4*>(z + o);  zv[it][1] = *reinterpret_cast<const float4*>(z + o + 4);
        mu[it] = stats[2 * (c.r0 + row)]; rs[it] = stats[2 * (c.r0 + row) + 1];
      } else {
        dv[it][0] = dv[it][1] = zv[it][0] = zv[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        mu[it] = 0.f; rs[it] = 0.f;
      }
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = (half * 4 + it) * 16 + rsub;
      const uint32_t m = (uint32_t)(c.r0 + row);
      const float d8[8] = {dv[it][0].x, dv[it][0].y, dv[it][0].z, dv[it][0].w, dv[it][1].x, dv[it][1].y, dv[it][1].z, dv[it][1].w};
      const float z8[8] = {zv[it][0].x, zv[it][0].y, zv[it][0].z, zv[it][0].w, zv[it][1].x, zv[it][1].y, zv[it][1].z, zv[it][1].w};
      float xh[8], gd[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[j] = (z8[j] - mu[it]) * rs[it];
        acc[0][j] = fmaf(d8[j], xh[j], acc[0][j]);
        acc[1][j] += d8[j];
        gd[j] = d8[j] * gam[j];                                  // d xhat
        s1 += gd[j];
        s2 = fmaf(gd[j], xh[j], s2);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {                          // the 16 lanes of a row are one half-warp
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      s1 *= (1.0f / 128.0f); s2 *= (1.0f / 128.0f);
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = rs[it] * (gd[j] - s1 - xh[j] * s2);
      if (row < c.R) {
        float* o = dz_out + (size_t)m * 128 + col0;
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
      // dropout factors of elements m * 128 + col0 .. + 7: pairs m * 64 + col0 / 2 .. + 3 (draw32 with its row half hoisted)
      if (dm.thresh != 0u) {
        const uint32_t rk = mix32(dm.key ^ (m >> 1)), pb = m * 64u + (uint32_t)(col0 >> 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t hsh = mix32((pb + (uint32_t)q) * 0x9E3779B1u + dm.key) ^ rk;
          v[2 * q] = (hsh & 0xFFFFu) >= dm.thresh ? v[2 * q] * dm.scale : 0.f;
          v[2 * q + 1] = (hsh >> 16) >= dm.thresh ? v[2 * q + 1] * dm.scale : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[2][j] += v[j];              // rows >= R contribute exact zeros (rs = 0)
      uint4 hi, lo;
      split8(v, hi, lo);
      const uint32_t off = (uint32_t)(chunk >> 3) * kImg + sw128_offset((uint32_t)row, (uint32_t)((chunk & 7) * 8));
      *reinterpret_cast<uint4*>(c.smem + off) = hi;
      *reinterpret_cast<uint4*>(c.smem + 2 * kImg + off) = lo;
    }
  }
}

// acc (this thread's 8 columns over its rows) summed over the 16 row groups of the CTA in a fixed order and ADDED to
// s_part[3][128] (single writer per element)
__device__ __forceinline__ void fold_partials(float (&acc)[3][8], float (*s_red)[3][128], float (*s_part)[128]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, chunk = tid & 15;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = acc[k][j] + __shfl_xor_sync(0xffffffffu, acc[k][j], 16);     // the two row groups of the warp
      if (lane < 16) s_red[warp][k][chunk * 8 + j] = v;
      acc[k][j] = 0.f;
    }
  __syncthreads();
  for (int e = tid; e < 384; e += kFT) {
    const int k = e >> 7, col = e & 127;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += s_red[w][k][col];
    s_part[k][col] += sum;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kFT, 2) sasrec_bwd_ffn_fused_kernel(const __grid_constant__ FusedBwdFfnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[5];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_gam[256];               // gamma2, gamma1
  __shared__ float s_part2[3][128], s_part1[3][128];       // this CTA's column partials (all its tiles)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = min(a.T_cap, *a.counts);
  const int n_tiles = (T + 127) / 128;
  if ((int)blockIdx.x >= n_tiles) {                        // no tile: this CTA's partials are zero
    for (int e = tid; e < 384; e += kFT) { a.part_ln2[(size_t)blockIdx.x * 384 + e] = 0.f; a.part_ln1[(size_t)blockIdx.x * 384 + e] = 0.f; }
    return;
  }

  Ctx c;
  c.smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  c.full = &bars[0]; c.free01 = &bars[2]; c.acc = &bars[3];
  c.pchunk = 0; c.ppieces = 0; c.n_acc = 0;
#ifdef DR4SR_TRACE
  __shared__ int s_trace[256];
  c.tr_n = 0; c.tr_t0 = clock64(); c.tr_buf = s_trace;
#endif
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  s_gam[tid] = tid < 128 ? a.gamma2[tid] : a.gamma1[tid - 128];
  for (int e = tid; e < 384; e += kFT) { (&s_part2[0][0])[e] = 0.f; (&s_part1[0][0])[e] = 0.f; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = tmem_slot;
  c.row = (warp & 3) * 32 + lane;
  c.half = warp >> 2;
  c.trow = c.tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(c.half * 64);
  float acc[3][8];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;

#pragma unroll 1
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    c.r0 = tile * 128;
    c.R = min(128, T - c.r0);
    c.live = c.row < c.R;
    c.m = c.r0 + c.row;
    TRACE(1);
    if (tid == kProducer) prefetch_chunk(c, a.w2_hi, a.w2_lo, 128, 0);   // no-op when the previous tile already queued them
    // ---- g3 = LN2'(gin) -> HBM and the A operand (g3 . mask_ffn_out) ----
    ln_bwd_pass(c, a.gin, a.z2, a.st2, s_gam, a.d_ffn_out, a.g3, acc);
    TRACE(2);
    sync_for_mma();
    TRACE(3);
    // ---- dpre = (A W2) . gp -> HBM and the next A operand ----
    {
      float gpv[64];                                        // this thread's 64 gp values: in flight under the UMMAs
      const float* prow = a.gp + (size_t)c.m * 128 + c.half * 64;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c.live) ld_global_v8(prow + 8 * j, gpv + 8 * j);
        else {
#pragma unroll
          for (int q = 0; q < 8; ++q) gpv[8 * j + q] = 0.f;
        }
      }
      run_chunk(c, a.w2_hi, a.w2_lo, 128, 0, a.w1_hi, a.w1_lo, 128, 0);
      TRACE(4);
      wait_acc(c);
      TRACE(5);
      // LN2 column partials; the staging buffer is the A-image region, dead between the accumulator wait and the stores below
      fold_partials(acc, reinterpret_cast<float (*)[3][128]>(c.smem), s_part2);
      float* orow = a.dpre + (size_t)c.m * 128 + c.half * 64;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v[16];
        tmem_ld16(c.trow + (uint32_t)(g * 16), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= gpv[g * 16 + j];
        if (c.live) {
          st_global_v8(orow + g * 16, v);
          st_global_v8(orow + g * 16 + 8, v + 8);
        }
        store_image16(v, c.row, c.half * 64 + g * 16, c.smem);
      }
    }
    TRACE(6);
    sync_for_mma();
    TRACE(7);
    // ---- dx1 = g3 + A W1 -> HBM ----
    {
      float rv[64];                                         // g3 rows (written above by this CTA; L2 hits), in flight under the UMMAs
      const float* rrow = a.g3 + (size_t)c.m * 128 + c.half * 64;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c.live) ld_global_v8(rrow + 8 * j, rv + 8 * j);
        else {
#pragma unroll
          for (int q = 0; q < 8; ++q) rv[8 * j + q] = 0.f;
        }
      }
      run_chunk(c, a.w1_hi, a.w1_lo, 128, 0, a.out_hi, a.out_lo, 128, 0);
      TRACE(8);
      wait_acc(c);
      TRACE(9);
      float* xrow = a.dx1 + (size_t)c.m * 128 + c.half * 64;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v[16];
        tmem_ld16(c.trow + (uint32_t)(g * 16), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += rv[g * 16 + j];
        if (c.live) {
          st_global_v8(xrow + g * 16, v);
          st_global_v8(xrow + g * 16 + 8, v + 8);
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                        // dx1 rows of the tile are visible to the whole CTA; A images are free
    tc_fence_after();
    TRACE(10);
    // ---- g1 = LN1'(dx1) -> HBM and the next A operand (g1 . mask_attn_out) ----
    ln_bwd_pass(c, a.dx1, a.z1, a.st1, s_gam + 128, a.d_attn_out, a.g1, acc);
    TRACE(11);
    sync_for_mma();
    TRACE(12);
    // ---- g2 = A Wo -> HBM ----
    const bool more = tile + (int)gridDim.x < n_tiles;
    run_chunk(c, a.out_hi, a.out_lo, 128, 0, more ? a.w2_hi : nullptr, a.w2_lo, 128, 0);
    TRACE(13);
    wait_acc(c);
    TRACE(14);
    fold_partials(acc, reinterpret_cast<float (*)[3][128]>(c.smem), s_part1);     // LN1 column partials (A images are dead again)
    {
      float* orow = a.g2 + (size_t)c.m * 128 + c.half * 64;
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        float v[32];
        tmem_ld16_nowait(c.trow + (uint32_t)(g * 32), v);
        tmem_ld16_nowait(c.trow + (uint32_t)(g * 32 + 16), v + 16);
        tmem_ld_fence(v, true);
        tmem_ld_fence(v + 16, false);
        if (c.live) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) st_global_v8(orow + g * 32 + j, v + j);
        }
      }
    }
    TRACE(15);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  for (int e = tid; e < 384; e += kFT) {
    a.part_ln2[(size_t)blockIdx.x * 384 + e] = (&s_part2[0][0])[e];
    a.part_ln1[(size_t)blockIdx.x * 384 + e] = (&s_part1[0][0])[e];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(c.tmem, 256);
#ifdef DR4SR_TRACE
  if (tid == 0 && g_trace && blockIdx.x < 8)
    for (int i = 0; i < 2 * c.tr_n; ++i) g_trace[2048 + blockIdx.x * 256 + i] = c.tr_buf[i];
#endif
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
  const size_t smem = (size_t)(2 * B + 1) * sizeof(int32_t);
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

}  // namespace dr4sr

extern "C" int dr4sr_fused_tiles(const int32_t* tok_off, int32_t B, int32_t L, int32_t* tiles, size_t tiles_len, dr4sr_stream_t stream) {
  if (!tok_off || !tiles || B <= 0 || L <= 0 || L > 128) return DR4SR_EINVAL;
  if (tiles_len < (size_t)dr4sr::fused_tiles_cap(B, L) + 2) return DR4SR_EWORKSPACE;
  return dr4sr::launch_fused_tiles(tok_off, B, tiles, dr4sr::as_stream(stream));
}

namespace dr4sr {

int launch_sasrec_fwd_fused(const FusedFwdHost& h, cudaStream_t st) {
  if (!fused_fwd_supported(h.L, 128, 128, 2) || h.n_layer < 1 || h.n_layer > 8) return DR4SR_EINVAL;
  FusedFwdArgs a{};
  a.table = h.table; a.pos = h.pos; a.in_ids = h.in_ids; a.tok_off = h.tok_off; a.row_seq = h.row_seq; a.tiles = h.tiles;
  a.x0 = h.x0; a.n_layer = h.n_layer; a.L = h.L; a.ln_eps = h.ln_eps; a.scale = 0.125f; a.d_embed = h.d_embed;
  for (int l = 0; l < h.n_layer; ++l) {
    const FusedLayerHost& s = h.layer[l];
    FusedLayer& d = a.layer[l];
    d.qkv = s.qkv; d.attn = s.attn; d.z1 = s.z1; d.st1 = s.st1; d.x1 = s.x1; d.hm = s.hm; d.gp = s.gp; d.z2 = s.z2; d.st2 = s.st2; d.x2 = s.x2;
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


int launch_sasrec_bwd_ffn_fused(const FusedBwdFfnHost& h, cudaStream_t st) {
  FusedBwdFfnArgs a{};
  a.gin = h.gin; a.z2 = h.z2; a.st2 = h.st2; a.gp = h.gp; a.z1 = h.z1; a.st1 = h.st1; a.gamma2 = h.gamma2; a.gamma1 = h.gamma1;
  a.w2_hi = h.img[0]; a.w2_lo = h.img[1]; a.w1_hi = h.img[2]; a.w1_lo = h.img[3]; a.out_hi = h.img[4]; a.out_lo = h.img[5];
  a.g3 = h.g3; a.dpre = h.dpre; a.dx1 = h.dx1; a.g1 = h.g1; a.g2 = h.g2; a.part_ln2 = h.part_ln2; a.part_ln1 = h.part_ln1;
  a.counts = h.counts; a.T_cap = h.T_cap;
  a.d_ffn_out = h.d_ffn_out; a.d_attn_out = h.d_attn_out;
  ProfScope prof("sasrec_bwd_ffn_fused", st);
  if (cudaFuncSetAttribute(sasrec_bwd_ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "sasrec_bwd_ffn_fused smem attribute");
    return DR4SR_ECUDA;
  }
  // the grid is fixed (kLnBwdBlocks = 2 CTAs per SM): every CTA writes its slice of the LayerNorm partials, tiles are strided
  sasrec_bwd_ffn_fused_kernel<<<kLnBwdBlocks, kFT, kFusedSmem, st>>>(a);
  DR4SR_LAUNCH_CHECK("sasrec_bwd_ffn_fused_kernel");
  return DR4SR_OK;
}

}  // namespace dr4sr
