// gemm_tc.cuh -- tcgen05 / TMEM GEMM for the encoder's dense layers with fp32-grade products.
//
//   C[m,n] (+epilogue) = sum_k proA(A)[m,k] * W[n,k]       A: fp32 activations in HBM (k contiguous)
//
// The reference runs these layers as true-fp32 cuBLAS SGEMMs (allow_tf32 is off); parity is 1e-4
// relative, so a plain bf16 tensor-core product is not admissible.  Each operand is split on the fly
// into bf16 hi + bf16 lo (16 mantissa bits) and the product is issued as three kind::f16 UMMAs
//   hi*hi + hi*lo + lo*hi          (the dropped lo*lo term is below 2^-16 relative)
// accumulating in one fp32 TMEM tile.  These GEMMs are skinny (K = 128..384, N = 128..384): they are
// bound by streaming A in and C out of HBM, and the tensor pipe (3 x 8 UMMAs of 128x128x16 per tile
// and 64-deep k-stage) hides under that traffic.
//
// Structure of one CTA (256 threads = 8 warps, one output tile 128 x 128, ~64 KB smem => 3 CTAs/SM):
//   * weights are pre-split once per step by `weight_image_kernel` into bf16 hi/lo images already in
//     the UMMA K-major SWIZZLE_128B layout, so a 128 x 64 weight stage is one contiguous 16 KB block
//     that a single thread fetches with cp.async.bulk (TMA bulk copy, mbarrier complete_tx);
//   * all threads load the fp32 A stage (coalesced 16-byte loads), apply the prologue (GELU + dropout
//     or a dropout mask), split to bf16 hi/lo and write the swizzled smem operand;
//   * one thread issues the 12 UMMAs of the stage and commits to an mbarrier;
//   * epilogue: tcgen05.ld 32x32b -- warp w reads TMEM lanes 32*(w%4).. (tile rows) and the column half
//     w/4, one row per thread; LayerNorm statistics are in-thread sums combined across the two column
//     halves through shared memory (three passes over TMEM: z (stored back with tcgen05.st), variance,
//     normalise).  The A loads of the next k-stage are issued before the current stage is converted.
#pragma once
#include <cuda_bf16.h>
#include "gemm_simt.cuh"

namespace dr4sr {
namespace tc {

constexpr int kBM = 128, kBN = 128, kBK = 64;       // CTA tile; k-stage
constexpr int kThreads = 256;                      // 8 warps: enough ALU parallelism for split / prologue / epilogue
constexpr uint32_t kStageBytes = kBM * kBK * 2;     // one bf16 operand stage = 16 KB
constexpr uint32_t kSmemBytes = 4 * kStageBytes + 1024;   // A_hi, A_lo, B_hi, B_lo (+ alignment slack)

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (a failed launch the caller sees) instead of spinning forever on the device.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 22)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 256-bit global accesses (32-byte aligned): a thread that owns a row moves a full 32 B sector per instruction
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
               "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void ld_global_v8(const float* p, float* v) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]),
               "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p) : "memory");
}

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 64 bf16 (128 B), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);      // start address, 16-byte units, bits [0,14)
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset = 1024 B, bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B
  return d;
}
// Lean issue path.  Every SWIZZLE_128B descriptor used here (K-major, and MN-major with a 16-byte leading offset) has the
// same upper word and the lower word (smem_addr >> 4) | 1 << 16, so a thread that issues a long UMMA sequence keeps the
// 32-bit lower words of its operand bases and advances them with one integer add per step (+2 per 32 bytes) instead of
// rebuilding two 64-bit descriptors per instruction: the single issuing thread is a latency-bound scalar stream, and the
// descriptor arithmetic was most of its ~150 cycles per UMMA (profiles/r2_fused_fwd_timeline.md).
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
template <bool ACC>
__device__ __forceinline__ void umma_lo(uint32_t tmem_c, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .pred p;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "}" ::"r"(tmem_c), "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(ACC ? 1u : 0u) : "memory");
}
__device__ __forceinline__ void umma_lo(uint32_t tmem_c, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, bool acc) {
  if (acc) umma_lo<true>(tmem_c, a_lo, b_lo, idesc); else umma_lo<false>(tmem_c, a_lo, b_lo, idesc);
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

// byte offset of element (row, k) inside a [rows x 64] bf16 SW128 K-major block whose base is 1024-byte aligned
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t k) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((((k >> 3) ^ (row & 7u)) & 7u) << 4) + (k & 7u) * 2u;
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi); two elements per packed conversion (cvt.rn.bf16x2.f32: first source ->
// upper half), the low element of each pair in the low half-word
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hb), "f"(a - ha));
}
__device__ __forceinline__ void split_bf16x8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
  split_bf16x2(a.x, a.y, hi.x, lo.x); split_bf16x2(a.z, a.w, hi.y, lo.y);
  split_bf16x2(b.x, b.y, hi.z, lo.z); split_bf16x2(b.z, b.w, hi.w, lo.w);
}

// ---- weight images --------------------------------------------------------------------------------
// Image of a logical [N, K] operand (N % 128 == 0, K % 64 == 0): K/64 blocks, each N rows x 128 bytes in
// SW128 K-major order; the rows n0..n0+127 of block kb are the contiguous 16 KB at (kb * N + n0) * 128.
// transpose != 0: logical W'[n][k] = src[k * ld + n] (the backward-data operand W^T).
struct ImageJob { const float* src; int ld; int N, K; int transpose; uint16_t* hi; uint16_t* lo; };
constexpr int kMaxImageJobs = 16;
struct ImageTable { ImageJob job[kMaxImageJobs]; int count; };

static __global__ void __launch_bounds__(256) weight_image_kernel(ImageTable tab) {
  const ImageJob j = tab.job[blockIdx.y];
  const int chunks = j.N * (j.K / 8);                 // 16-byte chunks (8 consecutive k of one row)
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += gridDim.x * blockDim.x) {
    const int n = c / (j.K / 8), k0 = (c % (j.K / 8)) * 8;
    float x[8];
    if (!j.transpose) {
      const float4 a = *reinterpret_cast<const float4*>(j.src + (size_t)n * j.ld + k0);
      const float4 b = *reinterpret_cast<const float4*>(j.src + (size_t)n * j.ld + k0 + 4);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = j.src[(size_t)(k0 + i) * j.ld + n];
    }
    uint4 hi, lo;
    split_bf16x8(make_float4(x[0], x[1], x[2], x[3]), make_float4(x[4], x[5], x[6], x[7]), hi, lo);
    const size_t off = (size_t)(k0 / 64) * j.N * 128 + sw128_offset((uint32_t)n, (uint32_t)(k0 % 64));
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(j.hi) + off) = hi;
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(j.lo) + off) = lo;
  }
}

inline int launch_weight_images(const ImageTable& tab, cudaStream_t st) {
  if (tab.count <= 0) return DR4SR_OK;
  if (tab.count > kMaxImageJobs) return DR4SR_EINVAL;
  for (int i = 0; i < tab.count; ++i)
    if (tab.job[i].N % 128 || tab.job[i].K % 64) return DR4SR_EINVAL;
  ProfScope prof("weight_images", st);
  weight_image_kernel<<<dim3(24, tab.count), 256, 0, st>>>(tab);
  DR4SR_LAUNCH_CHECK("weight_image_kernel");
  return DR4SR_OK;
}
inline size_t image_elems(int N, int K) { return (size_t)N * K; }   // per image (hi or lo), in bf16 elements

// ---- the GEMM ---------------------------------------------------------------------------------------
struct TcArgs {
  GemmArgs g;                 // A, C, lda, ldc, M, N, K, tok_dev, prologue A, epilogue fields
  const uint16_t* b_hi;       // weight images of the logical [N, K] operand
  const uint16_t* b_lo;
};

enum TcEpi : int { TC_LINEAR = 0, TC_GELU_BWD = 1, TC_LN = 2 };

constexpr int kRowIters = kBM * (kBK / 8) / kThreads;   // 16-byte bf16 chunks per thread per operand stage (= 4)

// fp32 A stage (128 rows x 64 k) -> registers; rows >= Mlim read as zero
__device__ __forceinline__ void load_a_stage(const GemmArgs& g, int m0, int Mlim, int s, float4 (&v)[kRowIters][2]) {
  const int chunk = threadIdx.x & 7, rsub = threadIdx.x >> 3;
#pragma unroll
  for (int it = 0; it < kRowIters; ++it) {
    const int gm = m0 + it * (kThreads / 8) + rsub, gk = s * kBK + chunk * 8;
    if (gm < Mlim) {
      const float* p = g.A + (size_t)gm * g.lda + gk;
      v[it][0] = *reinterpret_cast<const float4*>(p);
      v[it][1] = *reinterpret_cast<const float4*>(p + 4);
    } else {
      v[it][0] = v[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
// registers -> prologue -> bf16 hi/lo -> swizzled smem operand
__device__ __forceinline__ void store_a_stage(const GemmArgs& g, int m0, int Mlim, int s, float4 (&v)[kRowIters][2], uint8_t* a_hi,
                                              uint8_t* a_lo) {
  const int chunk = threadIdx.x & 7, rsub = threadIdx.x >> 3;
#pragma unroll
  for (int it = 0; it < kRowIters; ++it) {
    const int row = it * (kThreads / 8) + rsub, gm = m0 + row, gk = s * kBK + chunk * 8;
    if (g.proA != PRO_NONE && gm < Mlim) {
      const uint32_t idx = (uint32_t)gm * (uint32_t)g.lda + gk;
      v[it][0] = apply_prologue(v[it][0], g.proA, g.dropA, idx);
      v[it][1] = apply_prologue(v[it][1], g.proA, g.dropA, idx + 4);
    }
    uint4 hi, lo;
    split_bf16x8(v[it][0], v[it][1], hi, lo);
    const uint32_t off = sw128_offset((uint32_t)row, (uint32_t)(chunk * 8));
    *reinterpret_cast<uint4*>(a_hi + off) = hi;
    *reinterpret_cast<uint4*>(a_lo + off) = lo;
  }
}

template <int EPI>
__global__ void __launch_bounds__(kThreads) gemm_tc_kernel(TcArgs t) {
  const GemmArgs& g = t.g;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_slot;
  __shared__ float ln_part[2][kBM];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const int Mlim = min(g.M, g.tok_dev ? *g.tok_dev : g.M);
  if (m0 >= Mlim) return;                                   // uniform: the whole CTA leaves before any allocation

  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SW128 atoms need 1 KB alignment
  uint8_t* a_hi = smem;
  uint8_t* a_lo = smem + kStageBytes;
  uint8_t* b_hi = smem + 2 * kStageBytes;
  uint8_t* b_lo = smem + 3 * kStageBytes;

  if (tid == 0) {
    mbar_init(&bar_b, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, kBN);               // 128 fp32 accumulator columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  const int nstage = g.K / kBK;
  float4 va[kRowIters][2], vb[kRowIters][2];
  load_a_stage(g, m0, Mlim, 0, va);
  for (int s = 0; s < nstage; ++s) {
    if (s + 1 < nstage) load_a_stage(g, m0, Mlim, s + 1, vb);    // in flight while this stage is converted and multiplied
    if (s > 0) {                                            // the previous stage's UMMAs have consumed the smem operands
      mbar_wait(&bar_mma, (uint32_t)((s - 1) & 1));
      tc_fence_after();
    }
    if (tid == 0) {
      mbar_expect_tx(&bar_b, 2 * kStageBytes);
      const size_t off = ((size_t)s * g.N + n0) * 128;      // rows n0..n0+127 of k-block s: contiguous 16 KB
      bulk_g2s(b_hi, reinterpret_cast<const uint8_t*>(t.b_hi) + off, kStageBytes, &bar_b);
      bulk_g2s(b_lo, reinterpret_cast<const uint8_t*>(t.b_lo) + off, kStageBytes, &bar_b);
    }
    store_a_stage(g, m0, Mlim, s, va, a_hi, a_lo);
    fence_async_smem();                                     // generic-proxy writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      mbar_wait(&bar_b, (uint32_t)(s & 1));                 // weight stage landed
      tc_fence_after();
      const uint32_t ah = desc_lo(smem_u32(a_hi)), al = desc_lo(smem_u32(a_lo)), bh = desc_lo(smem_u32(b_hi)), bl = desc_lo(smem_u32(b_lo));
#pragma unroll
      for (int k = 0; k < kBK / 16; ++k) {                  // UMMA K = 16 bf16 = 32 bytes inside the 128-byte swizzle row (+2 in 16 B units)
        umma_lo(tmem, ah + 2u * k, bh + 2u * k, kIdesc, s > 0 || k > 0);
        umma_lo<true>(tmem, ah + 2u * k, bl + 2u * k, kIdesc);
        umma_lo<true>(tmem, al + 2u * k, bh + 2u * k, kIdesc);
      }
      umma_commit(&bar_mma);                                // arrives when every UMMA above has finished
    }
    if (s + 1 < nstage) {
#pragma unroll
      for (int it = 0; it < kRowIters; ++it) { va[it][0] = vb[it][0]; va[it][1] = vb[it][1]; }
    }
  }
  mbar_wait(&bar_mma, (uint32_t)((nstage - 1) & 1));
  tc_fence_after();

  // ---------------- epilogue: warp w -> TMEM lanes 32*(w%4)..+31 (tile rows), column half w/4 ----------------
  const int lane = tid & 31, quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane, m = m0 + row;
  const bool live = m < Mlim;
  const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * (kBN / 2));
  const int ncol0 = half * (kBN / 2);
  if (EPI != TC_LN) {
#pragma unroll 1
    for (int c = 0; c < kBN / 64; ++c) {
      float v[32];
      tmem_ld32(trow + (uint32_t)(c * 32), v);
      if (!live) continue;
      const int nb = n0 + ncol0 + c * 32;
      const bool wide = ((g.ldc | g.ldadd) & 7) == 0;      // 32-byte aligned rows: 256-bit row accesses
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        const int n = nb + j;
        float o[8] = {v[j], v[j + 1], v[j + 2], v[j + 3], v[j + 4], v[j + 5], v[j + 6], v[j + 7]};
        if (EPI == TC_GELU_BWD) {
          float p[8];
          const float* src = g.mul ? g.mul : g.pre;
          if (wide) ld_global_v8(src + (size_t)m * g.ldc + n, p);
          else {
            const float4 p0 = *reinterpret_cast<const float4*>(src + (size_t)m * g.ldc + n), p1 = *reinterpret_cast<const float4*>(src + (size_t)m * g.ldc + n + 4);
            p[0] = p0.x; p[1] = p0.y; p[2] = p0.z; p[3] = p0.w; p[4] = p1.x; p[5] = p1.y; p[6] = p1.z; p[7] = p1.w;
          }
          if (g.mul) {        // the forward saved mask * gelu'(pre)
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] *= p[q];
          } else {
            const float4 f0 = g.dropE.factor4((uint32_t)m * (uint32_t)g.ldc + n), f1 = g.dropE.factor4((uint32_t)m * (uint32_t)g.ldc + n + 4);
            o[0] *= f0.x * gelu_grad_f(p[0]); o[1] *= f0.y * gelu_grad_f(p[1]); o[2] *= f0.z * gelu_grad_f(p[2]); o[3] *= f0.w * gelu_grad_f(p[3]);
            o[4] *= f1.x * gelu_grad_f(p[4]); o[5] *= f1.y * gelu_grad_f(p[5]); o[6] *= f1.z * gelu_grad_f(p[6]); o[7] *= f1.w * gelu_grad_f(p[7]);
          }
        } else {
          if (g.bias) {
            const float4 b0 = *reinterpret_cast<const float4*>(g.bias + n), b1 = *reinterpret_cast<const float4*>(g.bias + n + 4);
            o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w; o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
          }
          if (g.add) {
            float a[8];
            if (wide) ld_global_v8(g.add + (size_t)m * g.ldadd + n, a);
            else {
              const float4 a0 = *reinterpret_cast<const float4*>(g.add + (size_t)m * g.ldadd + n), a1 = *reinterpret_cast<const float4*>(g.add + (size_t)m * g.ldadd + n + 4);
              a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] += a[q];
          }
          if (g.dropE.thresh) {
            const float4 f0 = g.dropE.factor4((uint32_t)m * (uint32_t)g.ldc + n), f1 = g.dropE.factor4((uint32_t)m * (uint32_t)g.ldc + n + 4);
            o[0] *= f0.x; o[1] *= f0.y; o[2] *= f0.z; o[3] *= f0.w; o[4] *= f1.x; o[5] *= f1.y; o[6] *= f1.z; o[7] *= f1.w;
          }
        }
        if (wide) st_global_v8(g.C + (size_t)m * g.ldc + n, o);
        else {
          *reinterpret_cast<float4*>(g.C + (size_t)m * g.ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(g.C + (size_t)m * g.ldc + n + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    }
  } else {
    // z = drop(acc + bias) + residual (written back to TMEM and to Z), then two-pass LayerNorm
    const float invN = 1.0f / (float)kBN;
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < kBN / 64; ++c) {
      float v[32];
      tmem_ld32(trow + (uint32_t)(c * 32), v);
      const int nb = ncol0 + c * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int n = nb + j;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f), rr = bb;
        if (g.bias) bb = *reinterpret_cast<const float4*>(g.bias + n);
        if (live) rr = *reinterpret_cast<const float4*>(g.add + (size_t)m * g.ldadd + n);
        const float4 f = g.dropE.factor4((uint32_t)m * (uint32_t)kBN + n);
        v[j] = (v[j] + bb.x) * f.x + rr.x;
        v[j + 1] = (v[j + 1] + bb.y) * f.y + rr.y;
        v[j + 2] = (v[j + 2] + bb.z) * f.z + rr.z;
        v[j + 3] = (v[j + 3] + bb.w) * f.w + rr.w;
        sum += (v[j] + v[j + 1]) + (v[j + 2] + v[j + 3]);
        if (live && g.Z) *reinterpret_cast<float4*>(g.Z + (size_t)m * kBN + n) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      tmem_st32(trow + (uint32_t)(c * 32), v);
    }
    ln_part[half][row] = sum;
    __syncthreads();
    const float mu = (ln_part[0][row] + ln_part[1][row]) * invN;
    __syncthreads();
    float var = 0.f;
#pragma unroll 1
    for (int c = 0; c < kBN / 64; ++c) {
      float v[32];
      tmem_ld32(trow + (uint32_t)(c * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) { const float d = v[j] - mu; var = fmaf(d, d, var); }
    }
    ln_part[half][row] = var;
    __syncthreads();
    const float rstd = rsqrtf((ln_part[0][row] + ln_part[1][row]) * invN + g.ln_eps);
#pragma unroll 1
    for (int c = 0; c < kBN / 64; ++c) {
      float v[32];
      tmem_ld32(trow + (uint32_t)(c * 32), v);
      if (!live) continue;
      const int nb = ncol0 + c * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int n = nb + j;
        const float4 gg = *reinterpret_cast<const float4*>(g.gamma + n);
        const float4 be = *reinterpret_cast<const float4*>(g.beta + n);
        float4 y;
        y.x = (v[j] - mu) * rstd * gg.x + be.x; y.y = (v[j + 1] - mu) * rstd * gg.y + be.y;
        y.z = (v[j + 2] - mu) * rstd * gg.z + be.z; y.w = (v[j + 3] - mu) * rstd * gg.w + be.w;
        *reinterpret_cast<float4*>(g.C + (size_t)m * g.ldc + n) = y;
      }
    }
    if (live && half == 0 && g.stats) { g.stats[2 * m] = mu; g.stats[2 * m + 1] = rstd; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kBN);
}

// true when the tensor-core path supports the shape (otherwise the caller uses the FFMA kernel)
inline bool tc_supported(int N, int K, bool ln) { return N % kBN == 0 && K % kBK == 0 && (!ln || N == kBN); }

template <int EPI>
inline int launch_gemm_tc(const GemmArgs& g, const uint16_t* b_hi, const uint16_t* b_lo, cudaStream_t st) {
  if (!tc_supported(g.N, g.K, EPI == TC_LN) || !b_hi || !b_lo) return DR4SR_EINVAL;
  if (g.lda % 4 || g.ldc % 4) return DR4SR_EINVAL;
  TcArgs t{g, b_hi, b_lo};
  ProfScope prof(g.tag ? g.tag : "gemm_tc", st);
  if (cudaFuncSetAttribute(gemm_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "gemm_tc smem attribute");
    return DR4SR_ECUDA;
  }
  dim3 grid(g.N / kBN, ceil_div(g.M, kBM));
  gemm_tc_kernel<EPI><<<grid, kThreads, kSmemBytes, st>>>(t);
  DR4SR_LAUNCH_CHECK("gemm_tc_kernel");
  return DR4SR_OK;
}

// ---- weight-gradient GEMM: dW[i,j] = sum_m A[m,i] * B[m,j] over the live tokens -----------------------
// Both operands are token-major activations, i.e. MN-major for this product (the reduction index m is
// the slow one), so they are staged in the UMMA MN-major SWIZZLE_128B layout (64 features x 8 tokens per
// 1 KB atom) and multiplied with a_major = b_major = MN.  All weight gradients of a layer go in ONE
// launch (job table); blockIdx.y splits the tokens, each split writes its partial tile, and the
// fixed-order reducer (launch_reduce_segments) sums them -> deterministic.
struct WgradJob {
  const float* A; int lda; int proA; Dropout dropA;    // [tokens, M_out] (dY)
  const float* B; int ldb; int proB; Dropout dropB;    // [tokens, N_out] (X)
  int M_out, N_out;
  float* partial;                                      // [n_split][M_out * N_out]
  int tile0;                                           // first linear tile index of this job
};
constexpr int kMaxWgradJobs = 4;
struct WgradTable { WgradJob job[kMaxWgradJobs]; int count; int total_tiles; int T_cap; const int* tok_dev; int n_split; };

// instruction descriptor with both operands MN-major
constexpr uint32_t kIdescMN = kIdesc | (1u << 15) | (1u << 16);
// A weight-gradient stage holds kWTok tokens x 128 features per operand in the MN-major SW128 layout: atoms of 64 features x
// 8 tokens (1 KB); atom(feature block ib, token group mb) at ib * kWFeatStride + mb * 1024.
constexpr int kWTok = 32;                                   // tokens per stage (two UMMA k-steps)
constexpr uint32_t kWFeatStride = (kWTok / 8) * 1024;       // 4 KB between the two 64-feature blocks
constexpr uint32_t kWImg = 2 * kWFeatStride;                // one operand image of a stage = 8 KB
constexpr uint32_t kWBuf = 4 * kWImg;                       // a_hi, a_lo, b_hi, b_lo = 32 KB; two buffers = kSmemBytes - 1 KB
static_assert(2 * kWBuf + 1024 == kSmemBytes, "two weight-gradient stage buffers fill the GEMM shared-memory budget");
__device__ __forceinline__ uint64_t sw128_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(kWFeatStride >> 4) << 16;          // leading byte offset: next 64-feature block
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: next 8-token group
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t sw128_offset_mn(uint32_t tok, uint32_t feat) {   // tok < kWTok, feat < 128, feat % 8 == 0
  return (feat >> 6) * kWFeatStride + (tok >> 3) * 1024u + (tok & 7u) * 128u + (((((feat & 63u) >> 3) ^ (tok & 7u)) & 7u) << 4);
}

// One CTA = one 128 x 128 output tile of one job x one token split.  Stages of kWTok tokens are software-pipelined:
// the fp32 rows of stage s+1 are in flight in registers while stage s is converted (prologue, bf16 hi/lo split) into one
// of two shared-memory buffers, and the UMMAs of a buffer run asynchronously while the other one is being filled.
static __global__ void __launch_bounds__(kThreads) wgrad_tc_kernel(WgradTable tab) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_buf[2], bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  int ji = 0;
  while (ji + 1 < tab.count && (int)blockIdx.x >= tab.job[ji + 1].tile0) ++ji;
  const WgradJob& job = tab.job[ji];
  const int tile = blockIdx.x - job.tile0, ntn = job.N_out / kBN;
  const int i0 = (tile / ntn) * kBM, j0 = (tile % ntn) * kBN;

  const int R = min(tab.T_cap, tab.tok_dev ? *tab.tok_dev : tab.T_cap);
  int len = (R + tab.n_split - 1) / tab.n_split;
  len = (len + kWTok - 1) / kWTok * kWTok;
  const int kbeg = min(R, (int)blockIdx.y * len), kend = min(R, kbeg + len);
  const int nstage = (kend - kbeg + kWTok - 1) / kWTok;
  float* out = job.partial + (size_t)blockIdx.y * job.M_out * job.N_out;

  if (nstage == 0) {                                   // no tokens in this split: the partial is zero
    for (int e = tid; e < kBM * kBN / 4; e += kThreads) {
      const int r = e / (kBN / 4), c = (e % (kBN / 4)) * 4;
      *reinterpret_cast<float4*>(out + (size_t)(i0 + r) * job.N_out + j0 + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  if (tid == 0) {
    mbar_init(&bar_buf[0], 1);
    mbar_init(&bar_buf[1], 1);
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, kBN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  // a stage = kWTok tokens x 128 features per operand = 512 16-byte chunks -> 2 per thread per operand;
  // 16 consecutive lanes read one token's 512 contiguous bytes
  const int fchunk = tid & 15, tsub = tid >> 4;        // feature chunk (8 floats), token within a group of 16
  float4 ra[2][2], rb[2][2], na[2][2], nb[2][2];
  auto load_stage = [&](int s, float4 (&xa)[2][2], float4 (&xb)[2][2]) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int gm = kbeg + s * kWTok + it * 16 + tsub;
      if (gm < kend) {
        const float* pa = job.A + (size_t)gm * job.lda + i0 + fchunk * 8;
        const float* pb = job.B + (size_t)gm * job.ldb + j0 + fchunk * 8;
        xa[it][0] = *reinterpret_cast<const float4*>(pa); xa[it][1] = *reinterpret_cast<const float4*>(pa + 4);
        xb[it][0] = *reinterpret_cast<const float4*>(pb); xb[it][1] = *reinterpret_cast<const float4*>(pb + 4);
      } else {
        xa[it][0] = xa[it][1] = xb[it][0] = xb[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  load_stage(0, ra, rb);
  for (int s = 0; s < nstage; ++s) {
    uint8_t* buf = smem + (uint32_t)(s & 1) * kWBuf;
    uint8_t *a_hi = buf, *a_lo = buf + kWImg, *b_hi = buf + 2 * kWImg, *b_lo = buf + 3 * kWImg;
    if (s + 1 < nstage) load_stage(s + 1, na, nb);      // in flight while this stage is converted
    if (s >= 2) {                                      // the UMMAs that read this buffer two stages ago are done
      mbar_wait(&bar_buf[s & 1], (uint32_t)(((s >> 1) - 1) & 1));
      tc_fence_after();
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int tok = it * 16 + tsub, gm = kbeg + s * kWTok + tok;
      if (gm < kend) {
        if (job.proA != PRO_NONE) {
          const uint32_t idx = (uint32_t)gm * (uint32_t)job.lda + i0 + fchunk * 8;
          ra[it][0] = apply_prologue(ra[it][0], job.proA, job.dropA, idx);
          ra[it][1] = apply_prologue(ra[it][1], job.proA, job.dropA, idx + 4);
        }
        if (job.proB != PRO_NONE) {
          const uint32_t idx = (uint32_t)gm * (uint32_t)job.ldb + j0 + fchunk * 8;
          rb[it][0] = apply_prologue(rb[it][0], job.proB, job.dropB, idx);
          rb[it][1] = apply_prologue(rb[it][1], job.proB, job.dropB, idx + 4);
        }
      }
      uint4 hi, lo;
      const uint32_t off = sw128_offset_mn((uint32_t)tok, (uint32_t)(fchunk * 8));
      split_bf16x8(ra[it][0], ra[it][1], hi, lo);
      *reinterpret_cast<uint4*>(a_hi + off) = hi;
      *reinterpret_cast<uint4*>(a_lo + off) = lo;
      split_bf16x8(rb[it][0], rb[it][1], hi, lo);
      *reinterpret_cast<uint4*>(b_hi + off) = hi;
      *reinterpret_cast<uint4*>(b_lo + off) = lo;
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // MN-major descriptors: same upper word as desc_lo's, leading byte offset kWFeatStride in the lower word
      constexpr uint32_t kLbo = ((kWFeatStride >> 4) << 16) - (1u << 16);
      const uint32_t ah = desc_lo(smem_u32(a_hi)) + kLbo, al = desc_lo(smem_u32(a_lo)) + kLbo, bh = desc_lo(smem_u32(b_hi)) + kLbo,
                     bl = desc_lo(smem_u32(b_lo)) + kLbo;
#pragma unroll
      for (int k = 0; k < kWTok / 16; ++k) {           // 16 tokens = two 8-token groups = 2048 bytes (+128 in 16 B units)
        umma_lo(tmem, ah + 128u * k, bh + 128u * k, kIdescMN, s > 0 || k > 0);
        umma_lo<true>(tmem, ah + 128u * k, bl + 128u * k, kIdescMN);
        umma_lo<true>(tmem, al + 128u * k, bh + 128u * k, kIdescMN);
      }
      umma_commit(&bar_buf[s & 1]);
      if (s + 1 == nstage) umma_commit(&bar_done);
    }
    if (s + 1 < nstage) {
#pragma unroll
      for (int it = 0; it < 2; ++it) { ra[it][0] = na[it][0]; ra[it][1] = na[it][1]; rb[it][0] = nb[it][0]; rb[it][1] = nb[it][1]; }
    }
  }
  mbar_wait(&bar_done, 0u);
  tc_fence_after();
  const int lane = tid & 31, quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;
  const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * (kBN / 2));
#pragma unroll 1
  for (int c = 0; c < kBN / 64; ++c) {
    float v[32];
    tmem_ld32(trow + (uint32_t)(c * 32), v);
    float* dst = out + (size_t)(i0 + row) * job.N_out + j0 + half * (kBN / 2) + c * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kBN);
}

inline bool wgrad_supported(int M_out, int N_out) { return M_out % kBM == 0 && N_out % kBN == 0; }

inline int launch_wgrad_tc(WgradTable& tab, cudaStream_t st) {
  if (tab.count <= 0 || tab.count > kMaxWgradJobs) return DR4SR_EINVAL;
  int tiles = 0;
  for (int i = 0; i < tab.count; ++i) {
    if (!wgrad_supported(tab.job[i].M_out, tab.job[i].N_out) || tab.job[i].lda % 4 || tab.job[i].ldb % 4) return DR4SR_EINVAL;
    tab.job[i].tile0 = tiles;
    tiles += (tab.job[i].M_out / kBM) * (tab.job[i].N_out / kBN);
  }
  tab.total_tiles = tiles;
  ProfScope prof("wgrad_tc", st);
  if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "wgrad_tc smem attribute");
    return DR4SR_ECUDA;
  }
  wgrad_tc_kernel<<<dim3(tiles, tab.n_split), kThreads, kSmemBytes, st>>>(tab);
  DR4SR_LAUNCH_CHECK("wgrad_tc_kernel");
  return DR4SR_OK;
}

}  // namespace tc
}  // namespace dr4sr
