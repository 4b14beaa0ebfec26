#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void k_tf32(float* out, int iters) {
  float c[4][4] = {};
  uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f000000u, 0x3f800000u}, b0 = 0x3f800000u, b1 = 0x3f000000u + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c[0][0] + c[1][1] + c[2][2] + c[3][3];
}
__global__ void k_bf16(float* out, int iters) {
  float c[4][4] = {};
  uint32_t a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f003f00u, 0x3f803f80u}, b0 = 0x3f803f80u, b1 = 0x3f003f00u + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c[0][0] + c[1][1] + c[2][2] + c[3][3];
}
int main() {
  float* out; cudaMalloc(&out, 148 * 512 * 4);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  int iters = 20000;
  for (int which = 0; which < 2; ++which) for (int warps = 4; warps <= 16; warps *= 2) {
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(s);
      if (which == 0) k_tf32<<<148, warps * 32>>>(out, iters); else k_bf16<<<148, warps * 32>>>(out, iters);
      cudaEventRecord(e); cudaEventSynchronize(e); cudaEventElapsedTime(&ms, s, e);
    }
    double mmas = 148.0 * warps * iters * 4;
    double macs = mmas * (which == 0 ? 16 * 8 * 8 : 16 * 8 * 16);
    printf("%s warps/SM=%2d: %.3f ms  %.1f cycles/MMA/SM(at 1.965GHz)  %.1f TMAC/s\n", which == 0 ? "tf32 m16n8k8 " : "bf16 m16n8k16", warps, ms,
           ms * 1e-3 * 1.965e9 / (warps * iters * 4.0), macs / (ms * 1e-3) / 1e12);
  }
  return 0;
}
