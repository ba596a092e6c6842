// gru_tc.cu -- the GRU recurrence (forward) on tcgen05 for hidden size 256 (reference module/layers.py:117-136: nn.GRU as
// configured by model/gru4rec.py:12-22; gate order r, z, n; h_t = (1 - z) n + z h_{t-1}, n = tanh(gi_n + r (W_hn h_{t-1}))).
//
// The recurrence is 50 dependent steps of gh = h_{t-1} W_hh^T: latency, not throughput.  The FFMA kernel (gru.cu) spends a step
// re-reading its W_hh slice and h from shared memory (LDS-bound: ~15 us per step).  Here a cluster of 8 CTAs owns a tile of 128
// length-sorted sequences (UMMA M = 128, TMEM lane = sequence); CTA `rank` owns hidden units [32 rank, 32 rank + 32) of all three
// gates, i.e. 96 rows of W_hh, resident in shared memory for the whole launch as bf16 hi / lo images (96 KB, copied once with
// cp.async.bulk straight out of the standard SWIZZLE_128B [3H x H] weight image: 32-row segments keep their swizzle phase).
// h_{t-1} of the tile lives in every CTA as the A operand, bf16 hi / lo [128 x 256] (128 KB), in the NON-swizzled K-major
// canonical layout (8-row x 16-byte core matrices; 16-byte K-columns 2 KB apart, 8-row groups 128 B apart) so that the 32 units
// a CTA produces are ONE contiguous 8 KB block per precision.  Per step and CTA
//   48 UMMAs (M128 N96 K16: 16 k-steps x {hi*hi, hi*lo, lo*hi}) -> fp32 accumulator in TMEM (96 columns)
//   epilogue: thread = (sequence, 16-unit half): gates from TMEM + the prefetched input projections gi -> new h of its units
//   exchange: the threads store their bf16 hi / lo split into the CTA's own slice of its A images (coalesced 16-byte stores), one
//             thread then sends the two 8 KB blocks to the 7 other CTAs with cp.async.bulk (shared::cta -> shared::cluster),
//             completing on the DESTINATION's mbarrier; a CTA issues the next step's UMMAs when its 7 x 16 KB have landed
//   the outputs the backward reads (h, h_prev, gates) go to global memory after the exchange, under the next step's UMMAs
// One cluster barrier per step (split arrive / wait): "every CTA's UMMAs of step t have finished reading the A images" before
// anybody overwrites them.  fp32 state: every thread keeps the exact fp32 h of its 16 units in registers (the z h_{t-1} term never
// sees the bf16 split).  Measured with generic st.shared::cluster stores in SWIZZLE_128B images (first version): 16.6 us per step,
// 12 K cycles of it in the scattered 16-byte remote stores -- hence the bulk copies.
#include <cstdio>

#include "gemm_tc.cuh"
#include "internal.cuh"

namespace dr4sr {
namespace {

using namespace tc;

constexpr int kH = 256;
constexpr int kClusterTc = 8;
constexpr int kUPC = kH / kClusterTc;           // 32 hidden units per CTA
constexpr int kNcol = 3 * kUPC;                 // 96 accumulator columns (gate-major: r | z | n)
constexpr int kTileM = 128;                     // sequences per cluster
constexpr int kGT = 288;                        // 8 epilogue warps + 1 issuing warp
constexpr uint32_t kAImg = 128 * 128;           // 64 units of h for 128 sequences: 16 KB (8 K-columns of 2 KB)
constexpr uint32_t kKCol = 2048;                // one 16-byte K-column (8 units) of all 128 rows
constexpr uint32_t kSlice = 4 * kKCol;          // the 32 units of one CTA: 8 KB per precision
constexpr uint32_t kADescHi = (128u >> 4) | (1u << 14);          // SBO = 128 B (8-row groups), version 1, no swizzle
constexpr uint32_t kALbo = (kKCol >> 4) << 16;                   // LBO = 2 KB between the two K-columns of a k-step
constexpr uint32_t kBImg = kNcol * 128;         // one [96 x 64] bf16 block of the W_hh slice: 12 KB
constexpr uint32_t kOffAhi = 0, kOffAlo = 4 * kAImg, kOffBhi = 8 * kAImg, kOffBlo = 8 * kAImg + 4 * kBImg;
constexpr uint32_t kGruTcSmem = 8 * kAImg + 8 * kBImg + 1024;      // + alignment slack
constexpr uint32_t kIdescN96 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kNcol >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void umma_ab(uint32_t tmem_c, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .pred p;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_c), "r"(a_lo), "r"(b_lo), "r"(kADescHi), "r"(kDescHi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void bulk_s2cluster(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(mbar_cluster) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float sigmoid_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __cluster_dims__(kClusterTc, 1, 1) __launch_bounds__(kGT, 1)
gru_fwd_tc_kernel(const float* __restrict__ gi, const uint16_t* __restrict__ whh_hi, const uint16_t* __restrict__ whh_lo,
                  const int32_t* __restrict__ tok_off, const int32_t* __restrict__ order, int B, float* __restrict__ h_out,
                  float* __restrict__ hprev_out, float* __restrict__ gates) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_w, bar_acc, bar_h;
  __shared__ uint32_t s_tmem;
  __shared__ int s_maxlen;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int tile = blockIdx.x / kClusterTc;
  const bool epi = warp < 8;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_acc, 1);
    mbar_init(&bar_h, 1);
    fence_mbar_init();
    s_maxlen = 0;
  }
  if (warp == 0) tmem_alloc(&s_tmem, 128);
  // h_0 = 0: clear both A images (generic proxy), made visible to the tensor core's async proxy below
  for (uint32_t e = tid; e < 8 * kAImg / 16; e += kGT) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  // this thread's sequence (TMEM lane) and unit half
  const int m = (warp & 3) * 32 + lane, uh = (warp >> 2) & 1;
  int off = 0, len = 0;
  if (epi) {
    const int bi = tile * kTileM + m;
    if (bi < B) { const int b = order[bi]; off = tok_off[b]; len = tok_off[b + 1] - off; }
    if (uh == 0) atomicMax(&s_maxlen, len);
  }
  // resident W_hh slice: 3 gates x 4 k-blocks x {hi, lo} segments of 32 rows (4 KB each) out of the [3H x H] images
  if (warp == 8 && lane == 0) {
    mbar_expect_tx(&bar_w, 8 * kBImg);
    for (int p = 0; p < 2; ++p) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p == 0 ? whh_hi : whh_lo);
      uint8_t* dst = smem + (p == 0 ? kOffBhi : kOffBlo);
      for (int kb = 0; kb < 4; ++kb)
        for (int g = 0; g < 3; ++g)
          bulk_g2s(dst + kb * kBImg + g * 4096, src + ((size_t)kb * 3 * kH + g * kH + rank * kUPC) * 128, 4096, &bar_w);
    }
  }
  __syncthreads();
  const int maxlen = s_maxlen;                  // the same in every CTA of the cluster (same tile)
  if (warp == 8 && lane == 0) mbar_wait(&bar_w, 0);
  __syncwarp();
  cluster_arrive();                             // every CTA's images are cleared before any CTA pushes into them
  cluster_wait();

  float hp[16];                                 // exact fp32 state of this thread's 16 units
#pragma unroll
  for (int j = 0; j < 16; ++j) hp[j] = 0.f;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_hi = (((smem_base + kOffAhi) & 0x3FFFFu) >> 4) | kALbo, a_lo = (((smem_base + kOffAlo) & 0x3FFFFu) >> 4) | kALbo;
  const uint32_t b_hi = desc_lo(smem_base + kOffBhi), b_lo = desc_lo(smem_base + kOffBlo);
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(uh * 16);
  const int unit0 = (int)rank * kUPC + uh * 16;             // first global hidden unit of this thread
  // this thread's two K-columns (8 units each) of row m inside an A image
  const uint32_t dst0 = (uint32_t)(unit0 / 8) * kKCol + (uint32_t)(m >> 3) * 128u + (uint32_t)(m & 7) * 16u, dst1 = dst0 + kKCol;
  const uint32_t bar_h_addr = smem_u32(&bar_h);

#ifdef DR4SR_TRACE
  long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tr_t = clock64();
#define GTR(i) do { const long long now_ = clock64(); tr_acc[i] += now_ - tr_t; tr_t = now_; } while (0)
#else
#define GTR(i) do { } while (0)
#endif
  for (int t = 0; t < maxlen; ++t) {
    const bool more = t + 1 < maxlen;
    if (warp == 8) {
      if (lane == 0) {
        if (t > 0) mbar_wait(&bar_h, (uint32_t)((t - 1) & 1));      // the 7 remote slices of h_{t-1} have landed
        tc_fence_after();
#pragma unroll 1
        for (int ks = 0; ks < 16; ++ks) {
          const uint32_t ao = (uint32_t)ks * (2u * kKCol >> 4);
          const uint32_t bo = (uint32_t)(ks >> 2) * (kBImg >> 4) + 2u * (uint32_t)(ks & 3);
          umma_ab(tmem, a_hi + ao, b_hi + bo, kIdescN96, ks != 0);
          umma_ab(tmem, a_hi + ao, b_lo + bo, kIdescN96, 1u);
          umma_ab(tmem, a_lo + ao, b_hi + bo, kIdescN96, 1u);
        }
        umma_commit(&bar_acc);
        if (more) mbar_expect_tx(&bar_h, 7u * 2u * kSlice);          // arm the arrival of h_t's remote slices
        mbar_wait(&bar_acc, (uint32_t)(t & 1));
      }
      __syncwarp();
      cluster_arrive();                         // (A) this CTA's UMMAs are done with the A images
      cluster_wait();
      __syncthreads();                          // (S) the CTA's own slice of h_t is written
      if (lane == 0 && more) {
#pragma unroll 1
        for (uint32_t r = 0; r < (uint32_t)kClusterTc; ++r) {
          if (r == rank) continue;
          const uint32_t mb = map_rank(bar_h_addr, r);
          bulk_s2cluster(map_rank(smem_base + kOffAhi + rank * kSlice, r), smem_base + kOffAhi + rank * kSlice, kSlice, mb);
          bulk_s2cluster(map_rank(smem_base + kOffAlo + rank * kSlice, r), smem_base + kOffAlo + rank * kSlice, kSlice, mb);
        }
      }
      __syncwarp();
      continue;
    }
    const bool live = t < len;
    const size_t row = (size_t)(off + t);
    float gr[16], gz[16], gn[16];
    if (live) {                                 // input projections of this step: in flight under the UMMAs
      const float* g0 = gi + row * 3 * kH + unit0;
      ld_global_v8(g0, gr); ld_global_v8(g0 + 8, gr + 8);
      ld_global_v8(g0 + kH, gz); ld_global_v8(g0 + kH + 8, gz + 8);
      ld_global_v8(g0 + 2 * kH, gn); ld_global_v8(g0 + 2 * kH + 8, gn + 8);
    }
    GTR(0);                                     // gi loads issued
    mbar_wait(&bar_acc, (uint32_t)(t & 1));
    GTR(1);                                     // accumulator ready
    tc_fence_after();
    float ar[16], az[16], an[16];
    tmem_ld16_nowait(trow, ar);
    tmem_ld16_nowait(trow + kUPC, az);
    tmem_ld16_nowait(trow + 2 * kUPC, an);
    tmem_wait_ld();
    tc_fence_before();
    __syncwarp();
    GTR(2);                                     // TMEM read
    cluster_arrive();                           // (A)
    float hold[16];
    if (live) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        hold[j] = hp[j];
        ar[j] = sigmoid_(gr[j] + ar[j]);                      // r
        az[j] = sigmoid_(gz[j] + az[j]);                      // z
        gn[j] = tanhf(gn[j] + ar[j] * an[j]);                 // n   (an stays W_hn h_{t-1}: the backward needs it)
        hp[j] = (1.0f - az[j]) * gn[j] + az[j] * hp[j];
      }
    }
    __syncwarp();
    GTR(3);                                     // gates
    cluster_wait();                             // (A) nobody reads h_{t-1} any more
    GTR(4);                                     // barrier A wait
    if (live && more) {
      uint4 hi0, lo0, hi1, lo1;
      split_bf16x8(make_float4(hp[0], hp[1], hp[2], hp[3]), make_float4(hp[4], hp[5], hp[6], hp[7]), hi0, lo0);
      split_bf16x8(make_float4(hp[8], hp[9], hp[10], hp[11]), make_float4(hp[12], hp[13], hp[14], hp[15]), hi1, lo1);
      st_shared_v4(smem_base + kOffAhi + dst0, hi0);
      st_shared_v4(smem_base + kOffAhi + dst1, hi1);
      st_shared_v4(smem_base + kOffAlo + dst0, lo0);
      st_shared_v4(smem_base + kOffAlo + dst1, lo1);
      fence_async_smem();                       // generic-proxy stores -> the bulk copies and the next UMMAs (async proxy)
    }
    __syncthreads();                            // (S)
    GTR(5);                                     // slice written
    if (live) {                                 // what the backward reads; off the critical path (under the copies / next UMMAs)
      // (measured: these row-per-thread 32-byte accesses -- 3 K store + 1.5 K load requests per CTA and step -- are what bounds the
      //  step now, ~4 cycles per request; 16-byte streaming stores were slower, ex2-based gate math changed nothing)
      float* ho = h_out + row * kH + unit0;
      st_global_v8(ho, hp); st_global_v8(ho + 8, hp + 8);
      float* po = hprev_out + row * kH + unit0;
      st_global_v8(po, hold); st_global_v8(po + 8, hold + 8);
      float* go = gates + row * 4 * kH + unit0;
      st_global_v8(go, ar); st_global_v8(go + 8, ar + 8);
      st_global_v8(go + kH, az); st_global_v8(go + kH + 8, az + 8);
      st_global_v8(go + 2 * kH, gn); st_global_v8(go + 2 * kH + 8, gn + 8);
      st_global_v8(go + 3 * kH, an); st_global_v8(go + 3 * kH + 8, an + 8);
    }
    GTR(6);                                     // global stores issued
  }
#ifdef DR4SR_TRACE
  if (blockIdx.x == 0 && (tid == 0 || tid == 128) && maxlen > 0)
    printf("gru_fwd_tc trace tid %d steps %d: cycles/step gi-issue %lld, acc-wait %lld, tmem-ld %lld, gates %lld, barA-wait %lld, slice+sync %lld, stores %lld\n",
           tid, maxlen, tr_acc[0] / maxlen, tr_acc[1] / maxlen, tr_acc[2] / maxlen, tr_acc[3] / maxlen, tr_acc[4] / maxlen, tr_acc[5] / maxlen,
           tr_acc[6] / maxlen);
#endif
  cluster_arrive();                             // no CTA leaves while a bulk copy into it may still be in flight
  cluster_wait();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace

bool gru_tc_supported(int H) { return H == kH; }

int launch_gru_fwd_tc(const float* gi, const uint16_t* whh_hi, const uint16_t* whh_lo, const int32_t* tok_off, const int32_t* order, int B,
                      float* h, float* hprev, float* gates, cudaStream_t st) {
  ProfScope prof("gru_recurrence_fwd_tc", st);
  if (cudaFuncSetAttribute(gru_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGruTcSmem) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "gru_fwd_tc smem attribute");
    return DR4SR_ECUDA;
  }
  const int tiles = ceil_div(B, kTileM);
  gru_fwd_tc_kernel<<<tiles * kClusterTc, kGT, kGruTcSmem, st>>>(gi, whh_hi, whh_lo, tok_off, order, B, h, hprev, gates);
  DR4SR_LAUNCH_CHECK("gru_fwd_tc_kernel");
  return DR4SR_OK;
}

}  // namespace dr4sr
