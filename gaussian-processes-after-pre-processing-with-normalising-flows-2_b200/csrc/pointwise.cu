// Fused pointwise layers of the Flow++ conditioner for the TRAINING path (forward + backward each in one pass):
//   concat_elu(x) = elu(cat(x, -x))                 flow_modules/mixlogcdf_nn.py:8-10
//   glu(x)        = x[:, :C] * sigmoid(x[:, C:])    flow_modules/mixlogcdf_nn.py:149-151,257-258
// Tensors are viewed as [outer, channels, inner]: inner = H*W for the NCHW convolutional branch (split along dim 1),
// inner = 1 for the NHWC attention branch (split along the last dim).  HBM-bound: 12 B / 12 B per input element.
#include "common.cuh"

namespace flowk {

__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float elu_grad(float x) { return x > 0.f ? 1.f : expf(x); }

__global__ void concat_elu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long outer, int C,
                                      long long inner) {
  const long long per = (long long)C * inner, total = outer * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long o = i / per, r = i - o * per;
    const float v = x[i];
    y[o * 2 * per + r] = elu_f(v);
    y[o * 2 * per + per + r] = elu_f(-v);
  }
}
__global__ void concat_elu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                                      long long outer, int C, long long inner) {
  const long long per = (long long)C * inner, total = outer * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long o = i / per, r = i - o * per;
    const float v = x[i];
    gx[i] = gy[o * 2 * per + r] * elu_grad(v) - gy[o * 2 * per + per + r] * elu_grad(-v);
  }
}
__global__ void glu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long outer, int C, long long inner) {
  const long long per = (long long)C * inner, total = outer * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long o = i / per, r = i - o * per;
    const float a = x[o * 2 * per + r], b = x[o * 2 * per + per + r];
    y[i] = a / (1.f + expf(-b));
  }
}
__global__ void glu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                               long long outer, int C, long long inner) {
  const long long per = (long long)C * inner, total = outer * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long o = i / per, r = i - o * per;
    const float a = x[o * 2 * per + r], b = x[o * 2 * per + per + r], g = gy[i];
    const float s = 1.f / (1.f + expf(-b));
    gx[o * 2 * per + r] = g * s;
    gx[o * 2 * per + per + r] = g * a * s * (1.f - s);
  }
}

static int grid_for(long long total) {
  long long b = (total + 255) / 256;
  return (int)(b < 148 * 16 ? b : 148 * 16);
}

}  // namespace flowk

using namespace flowk;

#define FLOWK_POINTWISE_ENTRY(NAME, KERNEL, ...)                                                            \
  if (outer < 0 || C < 1 || inner < 1) return FLOWK_ERR_SHAPE;                                              \
  if (outer == 0) return FLOWK_OK;                                                                          \
  const long long total = outer * C * inner;                                                                \
  KERNEL<<<grid_for(total), 256, 0, stream>>>(__VA_ARGS__, outer, C, inner);                                \
  return launch_status();

extern "C" int flowk_concat_elu_fwd(const float* x, float* y, long long outer, int C, long long inner, flowk_stream_t stream) {
  if (outer > 0 && (!x || !y)) return FLOWK_ERR_ARG;
  FLOWK_POINTWISE_ENTRY(concat_elu_fwd, concat_elu_fwd_kernel, x, y)
}
extern "C" int flowk_concat_elu_bwd(const float* x, const float* gy, float* gx, long long outer, int C, long long inner,
                                    flowk_stream_t stream) {
  if (outer > 0 && (!x || !gy || !gx)) return FLOWK_ERR_ARG;
  FLOWK_POINTWISE_ENTRY(concat_elu_bwd, concat_elu_bwd_kernel, x, gy, gx)
}
extern "C" int flowk_glu_fwd(const float* x, float* y, long long outer, int C, long long inner, flowk_stream_t stream) {
  if (outer > 0 && (!x || !y)) return FLOWK_ERR_ARG;
  FLOWK_POINTWISE_ENTRY(glu_fwd, glu_fwd_kernel, x, y)
}
extern "C" int flowk_glu_bwd(const float* x, const float* gy, float* gx, long long outer, int C, long long inner,
                             flowk_stream_t stream) {
  if (outer > 0 && (!x || !gy || !gx)) return FLOWK_ERR_ARG;
  FLOWK_POINTWISE_ENTRY(glu_bwd, glu_bwd_kernel, x, gy, gx)
}

// ---------------------------------------------------------------------------------------------------------------
// Weight normalisation (old-style weight_g / weight_v, mixlogcdf_nn.py:19-21 `weight_norm(nn.Conv2d)`) fused with the
// construction of the tcgen05 GEMM weight operands.  v: [N, cin, taps] (torch conv / linear weight layout).
//   norm[n] = ||v[n]||,  s[n] = g[n] / norm[n],  w = v * s
//   forward operand  [N, taps, cin_pad]  (K-major rows of W),            hi/lo TF32 split
//   dgrad operand    [cin, taps, n_pad]  = w[n, ci, taps-1-t] (taps flipped, transposed), hi/lo TF32 split
// Two launches per layer instead of ~10 torch kernels; the backward is one kernel:
//   dot = sum(gw * v);  gg = dot / norm;  gv = gw * s - v * (dot * g / norm^3)
namespace flowk {

__device__ __forceinline__ void split_tf32_rna(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xffffe000u);
}

__global__ void wn_norm_kernel(const float* __restrict__ v, float* __restrict__ norm, int rows, int cols) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* p = v + (size_t)row * cols;
  float s = 0.f;
  for (int i = threadIdx.x & 31; i < cols; i += 32) s = fmaf(p[i], p[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) norm[row] = sqrtf(s);
}

__global__ void wn_operands_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ norm,
                                   int N, int cin, int taps, int cin_pad, int n_pad, float* __restrict__ w,
                                   float* __restrict__ fwd_hi, float* __restrict__ fwd_lo, float* __restrict__ dg_hi,
                                   float* __restrict__ dg_lo) {
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (w) {
    const long long total = (long long)N * cin * taps, per = (long long)cin * taps;
    for (long long i = t0; i < total; i += stride) {
      const int n = (int)(i / per);
      w[i] = v[i] * (g[n] / norm[n]);
    }
  }
  if (fwd_hi) {
    const long long total = (long long)N * taps * cin_pad;
    for (long long i = t0; i < total; i += stride) {
      const int c = (int)(i % cin_pad);
      const int t = (int)((i / cin_pad) % taps);
      const int n = (int)(i / ((long long)cin_pad * taps));
      float val = 0.f;
      if (c < cin) val = v[((size_t)n * cin + c) * taps + t] * (g[n] / norm[n]);
      float hi, lo;
      split_tf32_rna(val, hi, lo);
      fwd_hi[i] = hi;
      fwd_lo[i] = lo;
    }
  }
  if (dg_hi) {
    const long long total = (long long)cin * taps * n_pad;
    for (long long i = t0; i < total; i += stride) {
      const int n = (int)(i % n_pad);
      const int t = (int)((i / n_pad) % taps);
      const int c = (int)(i / ((long long)n_pad * taps));
      float val = 0.f;
      if (n < N) val = v[((size_t)n * cin + c) * taps + (taps - 1 - t)] * (g[n] / norm[n]);
      float hi, lo;
      split_tf32_rna(val, hi, lo);
      dg_hi[i] = hi;
      dg_lo[i] = lo;
    }
  }
}

__global__ void wn_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ norm,
                              const float* __restrict__ gw, float* __restrict__ gv, float* __restrict__ gg, int cols) {
  __shared__ float part[4];
  const int row = blockIdx.x;
  const float* pv = v + (size_t)row * cols;
  const float* pg = gw + (size_t)row * cols;
  float d = 0.f;
  for (int i = threadIdx.x; i < cols; i += 128) d = fmaf(pg[i], pv[i], d);
  d = warp_sum(d);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = d;
  __syncthreads();
  const float dot = (part[0] + part[1]) + (part[2] + part[3]);
  const float nr = norm[row], gr = g[row];
  const float s = gr / nr, k = dot * gr / (nr * nr * nr);
  if (threadIdx.x == 0) gg[row] = dot / nr;
  for (int i = threadIdx.x; i < cols; i += 128) gv[(size_t)row * cols + i] = pg[i] * s - pv[i] * k;
}

}  // namespace flowk

extern "C" int flowk_weight_norm_operands(const float* v, const float* g, int N, int cin, int taps, int cin_pad, int n_pad,
                                          float* norm, float* w, float* fwd_hi, float* fwd_lo, float* dg_hi, float* dg_lo,
                                          flowk_stream_t stream) {
  if (N < 1 || cin < 1 || taps < 1) return FLOWK_ERR_SHAPE;
  if (!v || !g || !norm || (fwd_hi && !fwd_lo) || (dg_hi && !dg_lo)) return FLOWK_ERR_ARG;
  if ((fwd_hi && cin_pad < cin) || (dg_hi && n_pad < N)) return FLOWK_ERR_SHAPE;
  const int cols = cin * taps;
  wn_norm_kernel<<<(N + 3) / 4, 128, 0, stream>>>(v, norm, N, cols);
  long long most = w ? (long long)N * cols : 0;
  if (fwd_hi && (long long)N * taps * cin_pad > most) most = (long long)N * taps * cin_pad;
  if (dg_hi && (long long)cin * taps * n_pad > most) most = (long long)cin * taps * n_pad;
  if (most > 0)
    wn_operands_kernel<<<grid_for(most), 256, 0, stream>>>(v, g, norm, N, cin, taps, cin_pad, n_pad, w, fwd_hi, fwd_lo,
                                                            dg_hi, dg_lo);
  return launch_status();
}

extern "C" int flowk_weight_norm_bwd(const float* v, const float* g, const float* norm, const float* gw, float* gv, float* gg,
                                     int N, int cols, flowk_stream_t stream) {
  if (N < 1 || cols < 1) return FLOWK_ERR_SHAPE;
  if (!v || !g || !norm || !gw || !gv || !gg) return FLOWK_ERR_ARG;
  wn_bwd_kernel<<<N, 128, 0, stream>>>(v, g, norm, gw, gv, gg, cols);
  return launch_status();
}
