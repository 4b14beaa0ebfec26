// How does a 4-D TMA box whose inner dimension is 64 B land in shared memory under SWIZZLE_128B, and are negative /
// unaligned inner coordinates allowed?   nvcc -arch=sm_100a -I../../<pkg>/csrc -I../../include tma_box_probe.cu -o tma_box_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "umma.cuh"
using namespace flowk::tc;

__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int bytes, int c0, int c1, int c2, int c3) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ int failed;
  if (threadIdx.x == 0) {
    failed = 0;
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = -1.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    fence_proxy_async();
    mbar_expect_tx(&bar, bytes);
    tma_load_4d(smem, &map, &bar, c0, c1, c2, c3);
  }
  mbar_wait(&bar, 0, &failed);
  __syncthreads();
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
  if (threadIdx.x == 0) out[bytes / 4] = (float)failed;
}

int main(int argc, char** argv) {
  const int B = 2, C = 16, H = 8, W = 16;
  std::vector<float> h(B * C * H * W);
  for (int b = 0; b < B; ++b) for (int c = 0; c < C; ++c) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x)
    h[((b * C + c) * H + y) * W + x] = b * 100000 + c * 1000 + y * 16 + x;     // c*1000 + pixel
  float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int rows = 8, bytes = rows * 128;
  CUtensorMap map;
  // flattened-pixel view: dims (P = H*W, C, B, 1), box (32 pixels, rows, 1, 1)
  cuuint64_t dims[4] = {H * W, C, B, 1};
  cuuint64_t strides[3] = {H * W * 4, C * H * W * 4, B * C * H * W * 4};
  cuuint32_t box[4] = {32, rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  float* out; cudaMalloc(&out, bytes + 4);
  std::vector<float> o(bytes / 4 + 1);
  int coords[1][4] = {{atoi(argv[1]), 4, 1, 0}};
  for (auto& c : coords) {
    probe<<<1, 128, bytes + 1024>>>(map, out, bytes, c[0], c[1], c[2], c[3]);
    cudaError_t e = cudaDeviceSynchronize();
    printf("coords (%d,%d,%d,%d): %s\n", c[0], c[1], c[2], c[3], cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(o.data(), out, bytes + 4, cudaMemcpyDeviceToHost);
    printf("failed=%g\n", o[bytes / 4]);
    for (int rrow = 0; rrow < rows; ++rrow) {
      printf("row %d:", rrow);
      for (int j = 0; j < 32; ++j) printf(" %g", o[rrow * 32 + j]);
      printf("\n");
    }
  }
  return 0;
}
