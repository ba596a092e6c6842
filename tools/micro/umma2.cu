// Does a burst of tcgen05.mma run slower after the tensor pipe has been idle?  One CTA per SM issues bursts of n MMAs
// separated by `gap` idle cycles and reports the issue->commit-complete time of each burst.
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
__global__ void __launch_bounds__(128, 1) k(int n_mma, int gap, int data, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = data ? (0x3c003c00u + (i * 2654435761u >> 20)) : 0u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t a = s32(sm), b = s32(sm + 32768);
    for (int rep = 0; rep < 8; ++rep) {
      long long tg = clock64();
      while (clock64() - tg < gap) {}
      long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t ko = (uint32_t)(i & 3) * 32u;
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem), "l"(sw128_desc(a + ko)),
                     "l"(sw128_desc(b + ko)), "r"(kIdesc), "r"((uint32_t)(i > 0)) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      while (!try_wait(&bar, rep & 1)) {}
      out[blockIdx.x * 8 + rep] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}
int main() {
  long long* out; cudaMalloc(&out, 4096 * 8);
  long long h[4096];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  for (int grid : {1, 148})
    for (int data : {0, 1})
      for (int gap : {0, 2000, 20000, 100000}) {
        k<<<grid, 128, 66 * 1024 + 1024>>>(24, gap, data, out);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, out, grid * 64, cudaMemcpyDeviceToHost);
        printf("grid %3d data %d gap %6d: bursts of 24 MMAs:", grid, data, gap);
        for (int r = 0; r < 8; ++r) { double a = 0; for (int i = 0; i < grid; ++i) a += h[i * 8 + r]; printf(" %.0f", a / grid); }
        printf("  %s\n", cudaGetErrorString(e));
      }
  return 0;
}
