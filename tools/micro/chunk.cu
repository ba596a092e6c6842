// Stripped model of the fused kernel's GEMM chunk loop, with switches to bisect what delays UMMA completion.
//   bit 0: weights arrive by cp.async.bulk into a 6-slot ring (else: B stays resident in smem)
//   bit 1: commit per piece (else: one commit per chunk)
//   bit 2: epilogue reads the accumulator with tcgen05.ld and stores rows to global
//   bit 3: A operand rewritten with tcgen05.st before every chunk
//   bit 4: other threads poll the acc barrier while thread 0 issues (else they wait at __syncthreads only)
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
__global__ void __launch_bounds__(256, 2) k(int flags, int chunks, const uint8_t* w, float* gout, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  __shared__ __align__(8) uint64_t full[6], empty[6], accb;
  __shared__ uint32_t slot_s;
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 98304 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u;
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[i]))); }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&accb)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot_s)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot_s;
  const uint32_t trow = tmem + (((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t fetched = 0, used = 0, nacc = 0;
  long long t_issue = 0, t_done = 0, t_epi = 0;
  auto fetch = [&](int piece) {
    const uint32_t s = fetched % 6, use = fetched / 6;
    if (use >= 1) while (!try_wait(&empty[s], (use - 1) & 1)) {}
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(16384) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + s * 16384)),
                 "l"(w + (size_t)(piece % 48) * 16384), "r"(16384), "r"(s32(&full[s])) : "memory");
    ++fetched;
  };
  int next_piece = 0;
  if ((flags & 1) && tid == 0) for (int i = 0; i < 6; ++i) fetch(next_piece++);
  for (int c = 0; c < chunks; ++c) {
    if (flags & 8) {                       // A operand via tcgen05.st (128 columns: hi + lo)
      for (int g = 0; g < 4; ++g) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(trow - (warp >> 2) * 64 + 128 + (warp >> 2) * 32 + g * 8), "r"(0x3c003c00u) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(trow - (warp >> 2) * 64 + 192 + (warp >> 2) * 32 + g * 8), "r"(0x3c003c00u) : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t0 = clock64();
    if (tid == 0) {
      for (int p = 0; p < 4; ++p) {
        uint32_t b;
        if (flags & 1) {
          const uint32_t s = used % 6, use = used / 6;
          while (!try_wait(&full[s], use & 1)) {}
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          b = s32(sm + s * 16384);
        } else {
          b = s32(sm + p * 16384);
        }
        const int nm = (p & 1) ? 4 : 8;
        for (int i = 0; i < nm; ++i) {
          const uint32_t kk = (uint32_t)(i & 3);
          asm volatile("{ .reg .pred q; setp.ne.b32 q, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, q; }" ::"r"(tmem), "r"(tmem + 128 + (p >> 1) * 32 + kk * 8),
                       "l"(sw128_desc(b + kk * 32)), "r"(kIdesc), "r"((uint32_t)(p > 0 || i > 0)) : "memory");
        }
        if (flags & 1) {
          if (flags & 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&empty[used % 6])) : "memory");
          ++used;
        }
      }
      if ((flags & 1) && !(flags & 2))      // one commit covers the chunk: every slot it used becomes free together
        for (int p = 0; p < 4; ++p) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&empty[(used - 4 + p) % 6])) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&accb)) : "memory");
      t_issue += clock64() - t0;
      if (flags & 1) while (fetched - used < 6) fetch(next_piece++);
    }
    if (!(flags & 16)) __syncthreads();
    __syncwarp();
    while (!try_wait(&accb, nacc & 1)) {}
    ++nacc;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t1 = clock64();
    t_done += t1 - t0;
    if (flags & 4) {
      for (int g = 0; g < 4; ++g) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15])
          : "r"(trow + g * 16));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float4* dst = reinterpret_cast<float4*>(gout + ((size_t)blockIdx.x * 128 + (warp & 3) * 32 + lane) * 128 + (warp >> 2) * 64 + g * 16);
        if (flags & 32) {
          for (int j = 0; j < 2; ++j)
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 2 * j), "r"(r[8 * j]), "r"(r[8 * j + 1]), "r"(r[8 * j + 2]), "r"(r[8 * j + 3]),
                         "r"(r[8 * j + 4]), "r"(r[8 * j + 5]), "r"(r[8 * j + 6]), "r"(r[8 * j + 7]) : "memory");
        } else {
          for (int j = 0; j < 4; ++j) dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
      }
    }
    t_epi += clock64() - t1;
  }
  if (tid == 0) { out[3 * blockIdx.x] = t_issue / chunks; out[3 * blockIdx.x + 1] = t_done / chunks; out[3 * blockIdx.x + 2] = t_epi / chunks; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}
int main() {
  long long* out; cudaMalloc(&out, 4096 * 8);
  uint8_t* w; cudaMalloc(&w, 48 * 16384); cudaMemset(w, 0x3c, 48 * 16384);
  float* gout; cudaMalloc(&gout, (size_t)296 * 128 * 128 * 4);
  long long h[4096];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int grid : {148, 296})
    for (int flags : {4, 36, 31, 63}) {
      k<<<grid, 256, 98304 + 1024>>>(flags, 40, w, gout, out);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, out, grid * 24, cudaMemcpyDeviceToHost);
      double a = 0, b = 0, c = 0;
      for (int i = 0; i < grid; ++i) { a += h[3 * i]; b += h[3 * i + 1]; c += h[3 * i + 2]; }
      printf("grid %3d flags %2d (ring %d, commit/piece %d, epi %d, A-st %d, poll %d): issue %.0f  issue->acc %.0f  epilogue %.0f cyc per chunk  %s\n", grid, flags, flags & 1,
             (flags >> 1) & 1, (flags >> 2) & 1, (flags >> 3) & 1, (flags >> 4) & 1, a / grid, b / grid, c / grid, cudaGetErrorString(e));
    }
  return 0;
}
