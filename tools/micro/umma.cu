// Microbenchmark: completion time of N back-to-back tcgen05.mma (M128 N128 K16 bf16, same accumulator) from issue to the
// tcgen05.commit mbarrier, SS mode (A, B in SW128 smem) and TS mode (A in TMEM), alone and with every SM busy.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t a) {
  uint64_t d = 0; d |= (uint64_t)((a & 0x3FFFFu) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
__global__ void __launch_bounds__(128, 2) k(int n_mma, int mode, int spin_others, long long* out, int noise, float* sink, const uint8_t* gsrc, float* gdst) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t nb[4];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&nb[i]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (noise && (blockIdx.x & 1)) {          // co-resident CTA: an epilogue-like load on ONE resource for ~100K cycles
    long long t0 = clock64();
    float acc = 0.f;
    uint32_t r[16];
    int nphase = 0;
    while (clock64() - t0 < 150000) {
      if (noise == 1) {                     // TMEM loads
        for (int i = 0; i < 8; ++i) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15])
            : "r"(tmem + ((threadIdx.x >> 5) * 32 << 16) + (i * 16 & 127)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          acc += __uint_as_float(r[3]);
        }
      } else if (noise == 2) {              // shared-memory stores (16 B per thread)
        for (int i = 0; i < 32; ++i) *reinterpret_cast<uint4*>(sm + ((threadIdx.x * 16 + i * 2048) & 65535)) = make_uint4(i, i, i, i);
      } else if (noise == 3) {              // ALU
        for (int i = 0; i < 256; ++i) acc = fmaf(acc, 1.0001f, 0.5f);
      } else if (noise == 5) {              // smem stores + async-proxy fences (what sync_for_mma does)
        for (int i = 0; i < 8; ++i) {
          *reinterpret_cast<uint4*>(sm + ((threadIdx.x * 16 + i * 2048) & 65535)) = make_uint4(i, i, i, i);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
      } else if (noise == 6) {              // bulk copies global -> smem, 4 x 16 KB in flight
        if (threadIdx.x == 0) {
          for (int i = 0; i < 4; ++i) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&nb[i])), "r"(16384) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + i * 16384)),
                         "l"(gsrc + (size_t)i * 16384), "r"(16384), "r"(s32(&nb[i])) : "memory");
          }
          for (int i = 0; i < 4; ++i) while (!try_wait(&nb[i], nphase & 1)) {}
          ++nphase;
        }
      } else if (noise == 7) {              // streaming global stores: 32 B per thread per row, like the epilogues
        for (int i = 0; i < 16; ++i) reinterpret_cast<float4*>(gdst)[((size_t)blockIdx.x * 128 + threadIdx.x) * 64 + ((i + nphase) & 63)] = make_float4(1.f, 2.f, 3.f, 4.f);
        ++nphase;
      } else if (noise == 8) {              // global loads (L2 hits), dependent use
        for (int i = 0; i < 16; ++i) acc += reinterpret_cast<const float4*>(gdst)[((size_t)blockIdx.x * 128 + threadIdx.x) * 64 + ((i + nphase) & 63)].x;
        ++nphase;
      } else if (noise == 4) {              // TMEM stores
        for (int i = 0; i < 8; ++i) {
          asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" :: "r"(tmem + ((threadIdx.x >> 5) * 32 << 16) + (i * 16 & 127)), "r"(i) : "memory");
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
      }
    }
    if (acc == 123.f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    return;
  }
  if (noise) { long long t0 = clock64(); while (clock64() - t0 < 20000) {} }   // let the noise CTA get going
  long long t_issue = 0, t_done = 0;
  for (int rep = 0; rep < 3; ++rep) {
    if (threadIdx.x == 0) {
      const uint32_t a = s32(sm), b = s32(sm + 32768);
      long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t ko = (uint32_t)(i & 3) * 32u;
        const uint32_t acc = i > 0;
        if (mode == 0)
          asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem), "l"(sw128_desc(a + ko)),
                       "l"(sw128_desc(b + ko)), "r"(kIdesc), "r"(acc) : "memory");
        else
          asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }" ::"r"(tmem), "r"(tmem + 128 + (i & 3) * 8),
                       "l"(sw128_desc(b + ko)), "r"(kIdesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      long long t1 = clock64();
      while (!try_wait(&bar, rep & 1)) {}
      long long t2 = clock64();
      t_issue = t1 - t0; t_done = t2 - t0;
    } else if (spin_others) {
      while (!try_wait(&bar, rep & 1)) {}
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = t_issue; out[2 * blockIdx.x + 1] = t_done; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}
int main() {
  long long* out; cudaMalloc(&out, 4096 * 8);
  float* sink; cudaMalloc(&sink, 64);
  long long h[4096];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  const char* names[] = {"none", "tmem-ld", "smem-st", "alu", "tmem-st", "proxyfence", "bulkcopy", "gstore", "gload"};
  uint8_t* gsrc; cudaMalloc(&gsrc, 1 << 20); cudaMemset(gsrc, 0, 1 << 20);
  float* gdst; cudaMalloc(&gdst, (size_t)296 * 128 * 64 * 16);
  for (int noise : {0, 6, 7, 8})
    for (int mode : {0, 1})
      for (int n : {4, 24}) {
        const int grid = 296;
        k<<<grid, 128, 66 * 1024 + 1024>>>(n, mode, 0, out, noise, sink, gsrc, gdst);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, out, grid * 16, cudaMemcpyDeviceToHost);
        double ai = 0, ad = 0; int cnt = 0;
        for (int i = 0; i < grid; i += (noise ? 2 : 1)) { ai += h[2 * i]; ad += h[2 * i + 1]; ++cnt; }
        printf("noise %-8s %s n=%2d: issue %.0f cyc, done %.0f cyc (%.0f / mma)  %s\n", names[noise], mode ? "TS" : "SS", n, ai / cnt, ad / cnt, ad / cnt / n, cudaGetErrorString(e));
      }
  return 0;
}
