#include <cuda_runtime.h>
__global__ void k(const float2* a, const float2* b, float2* c, int n) {
  float2 acc = make_float2(0.f, 0.f);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float2 x = a[i], y = b[i];
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(*reinterpret_cast<unsigned long long*>(&acc)) : "l"(*reinterpret_cast<unsigned long long*>(&x)), "l"(*reinterpret_cast<unsigned long long*>(&y)));
  }
  c[threadIdx.x] = acc;
}
int main() { return 0; }
