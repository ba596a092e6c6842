#include <cuda_runtime.h>
__global__ void k(float* p) {
  float a = threadIdx.x;
  asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + threadIdx.x * 8), "f"(a) : "memory");
}
int main() { return 0; }
