// Shared device/host helpers for libflowk (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "flowk.h"

namespace flowk {

constexpr int kMaxParts = 64;        // max CTAs that cooperate on one sample's log-det sum
constexpr int kThreads = 128;        // CTA size of the coupling kernels

#define FLOWK_CUDA_OK(expr)                                          \
  do {                                                               \
    cudaError_t e_ = (expr);                                         \
    if (e_ != cudaSuccess) return FLOWK_ERR_CUDA_BASE + (int)e_;     \
  } while (0)

inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? FLOWK_OK : FLOWK_ERR_CUDA_BASE + (int)e;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

// Workspace layout: one slot of (1 + kMaxParts) words per sample: { unsigned ticket; float partial[kMaxParts] }.
// The slot of sample b sits at the same address whatever B is, so calls with different batch sizes can
// share a workspace: partials never land on another call's ticket word.
constexpr int kSlotWords = kMaxParts + 1;
struct LdjWs {
  unsigned* base;
  __host__ __device__ unsigned* ticket(int b) const { return base + (size_t)b * kSlotWords; }
  __host__ __device__ float* partial(int b) const { return reinterpret_cast<float*>(base + (size_t)b * kSlotWords + 1); }
};
inline LdjWs carve_ws(void* ws, int /*B*/) {
  LdjWs w;
  w.base = reinterpret_cast<unsigned*>(ws);
  return w;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Streaming (read-once) loads/stores: keep the 126 MB L2 for tensors that are re-read.
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ float4 ld_stream4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

// fast transcendental building blocks (MUFU.EX2 / MUFU.RCP / MUFU.LG2)
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x = hi + lo with both parts fp16 (11 significant bits each, the same as TF32, so the split is as accurate as the
// TF32 one while |x| < 65504 and lo stays above fp16's subnormal spacing 2^-24; conversions saturate instead of
// overflowing to inf).  Returned as raw 16-bit patterns.
__device__ __forceinline__ void split_f16(float x, unsigned short& hi, unsigned short& lo) {
  unsigned short h, l;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  float hf;
  asm("cvt.f32.f16 %0, %1;" : "=f"(hf) : "h"(h));
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l) : "f"(x - hf));
  hi = h;
  lo = l;
}
// two adjacent output columns -> the (hi, lo) operand arrays: TF32 pairs in fp32 containers, or fp16 pairs (out_f16)
__device__ __forceinline__ void store_pair(float* out_hi, float* out_lo, size_t off, float v0, float v1, int out_f16) {
  if (out_f16) {
    unsigned short h0, l0, h1, l1;
    split_f16(v0, h0, l0);
    split_f16(v1, h1, l1);
    *reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned short*>(out_hi) + off) = (uint32_t)h0 | ((uint32_t)h1 << 16);
    *reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned short*>(out_lo) + off) = (uint32_t)l0 | ((uint32_t)l1 << 16);
  } else {
    const float a0 = __uint_as_float((__float_as_uint(v0) + 0x1000u) & 0xffffe000u);
    const float a1 = __uint_as_float((__float_as_uint(v1) + 0x1000u) & 0xffffe000u);
    const float b0 = __uint_as_float((__float_as_uint(v0 - a0) + 0x1000u) & 0xffffe000u);
    const float b1 = __uint_as_float((__float_as_uint(v1 - a1) + 0x1000u) & 0xffffe000u);
    *reinterpret_cast<float2*>(out_hi + off) = make_float2(a0, a1);
    *reinterpret_cast<float2*>(out_lo + off) = make_float2(b0, b1);
  }
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// Per-sample log-det reduction, deterministic.
//   grid = (parts <= kMaxParts, B); every thread passes its local sum.
//   CTA sum -> ws.partial[b][part]; the CTA that takes the last ticket adds the partials in
//   index order and writes ldj_out[b] = (ldj_in ? ldj_in[b] : 0) + sign * total, then re-arms
//   the ticket counter so the workspace is clean for the next launch on this stream.
template <int THREADS>
__device__ __forceinline__ void finish_sample_ldj(float local, const float* __restrict__ ldj_in,
                                                  float* __restrict__ ldj_out, float sign, LdjWs ws) {
  __shared__ float warp_part[THREADS / 32];
  __shared__ bool is_last;
  const int b = blockIdx.y, part = blockIdx.x, parts = gridDim.x;
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) s += warp_part[w];
    if (parts == 1) {
      ldj_out[b] = (ldj_in ? ldj_in[b] : 0.f) + sign * s;
      is_last = false;
    } else {
      __stcg(ws.partial(b) + part, s);
      __threadfence();
      unsigned ticket = atomicAdd(ws.ticket(b), 1u);
      is_last = (ticket == (unsigned)parts - 1u);
    }
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    float tot = 0.f;
    for (int p = 0; p < parts; ++p) tot += __ldcg(ws.partial(b) + p);
    ldj_out[b] = (ldj_in ? ldj_in[b] : 0.f) + sign * tot;
    *ws.ticket(b) = 0u;
  }
}

// Programmatic dependent launch (PDL): a kernel launched with `launch_pdl` may start while its predecessor in the
// stream is still running; it must call griddep_wait() before its first read of the predecessor's output and before
// its first global write.  griddep_launch() lets the NEXT kernel in the stream begin its own prologue early.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// The same with a thread-block cluster of `cluster_x` CTAs along x (1 = no cluster attribute).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                      int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.numAttrs = 1;
  if (cluster_x > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  cfg.attrs = attr;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int parts_for(long long work_items, int per_cta) {
  long long p = (work_items + per_cta - 1) / per_cta;
  if (p < 1) p = 1;
  if (p > kMaxParts) p = kMaxParts;
  return (int)p;
}

}  // namespace flowk
