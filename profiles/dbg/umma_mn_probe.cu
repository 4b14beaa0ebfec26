// Do MN-major (transposed) tf32 operands work with tcgen05.mma, and what do LBO / SBO mean for them?
// A_g [K=32][128] and B_g [K=32][64]: the contraction index is the ROW (pixel) index, channels are contiguous - the
// layout of NHWC activations for a weight-gradient GEMM.  TMA boxes of [32 rows x 32 channels] (128 B, SWIZZLE_128B).
//   usage: umma_mn_probe LBO_bytes SBO_bytes
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "umma.cuh"
using namespace flowk::tc;

__constant__ int g_layout = 2;      // UMMA layout type: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)g_layout << 61;
  return d;
}

__global__ void probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out,
                      uint32_t lbo, uint32_t sbo, int kmajor) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;               // 4 channel blocks x [32 rows x 128 B]
  uint8_t* sb = smem + 4 * 4096;    // 2 channel blocks
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_slot;
  __shared__ int failed;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    failed = 0;
    mbar_init(&full, 1);
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full, 6 * 4096);
    if (kmajor) {                      // control: K-major tiles [rows = channels][32 k], the layout conv_gemm uses
      tma_load_2d(sa, &map_a, &full, 0, 0);
      tma_load_2d(sb, &map_b, &full, 0, 0);
    } else {
      for (int j = 0; j < 4; ++j) tma_load_2d(sa + j * 4096, &map_a, &full, j * 32, 0);
      for (int j = 0; j < 2; ++j) tma_load_2d(sb + j * 4096, &map_b, &full, j * 32, 0);
    }
    mbar_wait(&full, 0, &failed);
    tc_fence_after();
    if (kmajor) {
      const uint32_t idk = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      for (int kk = 0; kk < 4; ++kk)
        umma_tf32(tmem, make_smem_desc(smem_u32(sa)) + (uint64_t)(kk * 2), make_smem_desc(smem_u32(sb)) + (uint64_t)(kk * 2), idk, kk > 0);
      umma_commit(&done);
    } else {
    // idesc: c=F32, a=b=TF32, a_major (bit 15) = b_major (bit 16) = 1 (MN-major), N=64, M=128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int kk = 0; kk < 4; ++kk)
      umma_tf32(tmem, desc_mn(smem_u32(sa) + kk * 1024, lbo, sbo), desc_mn(smem_u32(sb) + kk * 1024, lbo, sbo), idesc, kk > 0);
    umma_commit(&done);
    }
  }
  mbar_wait(&done, 0, &failed);
  tc_fence_after();
  if (warp < 4) {
    for (int col = 0; col < 64; col += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + col, v);
      for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + col + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
  if (threadIdx.x == 0) out[128 * 64] = (float)failed;
}

int main(int argc, char** argv) {
  const uint32_t lbo = argc > 1 ? atoi(argv[1]) : 4096, sbo = argc > 2 ? atoi(argv[2]) : 1024;
  const int kmajor = argc > 3 ? atoi(argv[3]) : 0;
  const int layout = argc > 4 ? atoi(argv[4]) : 2;                 // 1: 32-byte-atom 128B swizzle (TMA + UMMA)
  cudaMemcpyToSymbol(g_layout, &layout, sizeof(int));
  const int K = 32, MA = 128, NB = 64;
  std::vector<float> a(K * MA), b(K * NB);
  for (int k = 0; k < K; ++k) for (int m = 0; m < MA; ++m) a[k * MA + m] = (float)(((k * 7 + m * 3) % 11) - 5);
  for (int k = 0; k < K; ++k) for (int n = 0; n < NB; ++n) b[k * NB + n] = (float)(((k * 5 + n * 2) % 7) - 3);
  float *da, *db, *dout;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dout, (MA * NB + 1) * 4);
  if (kmajor) {                        // transposed copies [channels][K]
    std::vector<float> at(a.size()), bt(b.size());
    for (int k = 0; k < K; ++k) for (int m = 0; m < MA; ++m) at[m * K + k] = a[k * MA + m];
    for (int k = 0; k < K; ++k) for (int n = 0; n < NB; ++n) bt[n * K + k] = b[k * NB + n];
    cudaMemcpy(da, at.data(), a.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bt.data(), b.size() * 4, cudaMemcpyHostToDevice);
  } else {
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  }
  auto mk = [&](CUtensorMap* map, float* base, int ch) {
    cuuint64_t dims[2] = {(cuuint64_t)(kmajor ? K : ch), (cuuint64_t)(kmajor ? ch : K)};
    cuuint64_t strides[1] = {(cuuint64_t)(kmajor ? K : ch) * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)(kmajor ? ch : 32)}, estr[2] = {1, 1};
    return encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       layout == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUtensorMap ma, mb;
  if (mk(&ma, da, MA) || mk(&mb, db, NB)) { printf("encode failed\n"); return 1; }
  probe<<<1, 128, 6 * 4096 + 1024>>>(ma, mb, dout, lbo, sbo, kmajor);
  cudaError_t e = cudaDeviceSynchronize();
  printf("lbo %u sbo %u: %s\n", lbo, sbo, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> o(MA * NB + 1);
  cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0, first = -1;
  for (int m = 0; m < MA; ++m) for (int n = 0; n < NB; ++n) {
    float ref = 0;
    for (int k = 0; k < K; ++k) ref += a[k * MA + m] * b[k * NB + n];
    if (o[m * NB + n] != ref) { if (first < 0) first = m * NB + n; ++bad; }
  }
  printf("failed=%g mismatches %d of %d (first at m=%d n=%d)\n", o[MA * NB], bad, MA * NB, first / NB, first % NB);
  if (bad) { printf("row0 got:"); for (int n = 0; n < 8; ++n) printf(" %g", o[n]); printf("\n"); }
  return 0;
}
