// Per-SM write throughput from shared memory to global memory (L2): coalesced 128-bit st.global from all threads of a CTA vs
// one thread issuing cp.async.bulk (shared::cta -> global).  Question behind it: the GEMM epilogues move ~147 KB per CTA at
// ~25 B/clk through st.global - would bulk stores be faster?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/store_probe profiles/dbg/store_probe.cu && /tmp/store_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int BYTES = 144 * 1024;     // per CTA and repetition
constexpr int THREADS = 512;

__global__ void __launch_bounds__(THREADS) store_kernel(uint8_t* out, int mode, int reps, long long* cycles) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(i, 1, 2, 3);
  __syncthreads();
  uint8_t* dst = out + (size_t)blockIdx.x * BYTES;
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (mode == 0) {
      for (int i = threadIdx.x; i < BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(sm)[i];
    } else {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int c = 0; c < BYTES; c += 16384) {
          const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm + c);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c), "r"(s), "r"(16384) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncthreads();
    }
  }
  if (mode == 1 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

int main() {
  uint8_t* out;
  long long* cyc;
  cudaMalloc(&out, (size_t)148 * BYTES);
  cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES);
  const int reps = 50;
  for (int ctas : {1, 8, 32, 128, 148}) {
    for (int mode = 0; mode < 2; ++mode) {
      store_kernel<<<ctas, THREADS, BYTES>>>(out, mode, 2, cyc);       // warm-up
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a);
      store_kernel<<<ctas, THREADS, BYTES>>>(out, mode, reps, cyc);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("ctas %3d  %-14s  %7.1f us  %6.1f B/clk per SM (CTA 0: %lld cycles)  aggregate %6.0f GB/s  %s\n", ctas,
             mode ? "cp.async.bulk" : "st.global.v4", ms * 1e3, (double)BYTES * reps / (double)c, c,
             (double)ctas * BYTES * reps / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
