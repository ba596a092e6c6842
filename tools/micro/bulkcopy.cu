// Microbenchmark: latency / throughput of cp.async.bulk global->shared pieces under the access patterns of the
// fused encoder kernels (every CTA streaming the same weight image from L2).  nvcc -arch=sm_100a -o bulkcopy bulkcopy.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
  return ok;
}
// mode 0: all CTAs read the same region sequence; 1: each CTA its own region; depth = copies in flight; piece bytes
__global__ void __launch_bounds__(256, 2) k(const uint8_t* src, size_t region, int mode, int depth, int piece, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* base = src + (mode == 1 ? (size_t)blockIdx.x * region : 0);
    const int npieces = (int)(region / piece);
    long long t0 = clock64();
    int issued = 0, done = 0;
    while (done < iters) {
      while (issued < iters && issued - done < depth) {
        const int slot = issued % depth;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[slot])), "r"(piece) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + (size_t)slot * piece)),
                     "l"(base + (size_t)(issued % npieces) * piece), "r"(piece), "r"(s32(&bar[slot])) : "memory");
        ++issued;
      }
      const int slot = done % depth;
      while (!try_wait(&bar[slot], (done / depth) & 1)) {}
      ++done;
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  const size_t total = 64u << 20;
  uint8_t* src; cudaMalloc(&src, total); cudaMemset(src, 1, total);
  long long* out; cudaMalloc(&out, 1024 * 8);
  long long h[1024];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 192;
  struct Cfg { int grid, mode, depth, piece; size_t region; } cfgs[] = {
    {1, 0, 1, 16384, 786432}, {1, 0, 2, 16384, 786432}, {1, 0, 4, 16384, 786432},
    {148, 0, 1, 16384, 786432}, {148, 0, 2, 16384, 786432},
    {296, 0, 1, 16384, 786432}, {296, 0, 2, 16384, 786432}, {296, 0, 4, 16384, 786432}, {296, 0, 6, 16384, 786432},
    {296, 1, 2, 16384, 196608}, {296, 1, 4, 16384, 196608},
    {296, 0, 2, 32768, 786432}, {296, 0, 4, 8192, 786432}, {296, 0, 8, 8192, 786432}, {296, 0, 2, 16384, 65536},
  };
  for (auto& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) k<<<c.grid, 256, 98304>>>(src, c.region, c.mode, c.depth, c.piece, iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, c.grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0, mx = 0;
    for (int i = 0; i < c.grid; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; }
    avg /= c.grid;
    printf("grid %3d mode %d depth %d piece %5d region %7zu: avg %.0f cyc/copy (max CTA %.0f)  -> %.1f B/cyc/CTA  %s\n", c.grid, c.mode, c.depth, c.piece,
           c.region, avg / iters, mx / iters, (double)c.piece * iters / avg, cudaGetErrorString(e));
  }
  return 0;
}
