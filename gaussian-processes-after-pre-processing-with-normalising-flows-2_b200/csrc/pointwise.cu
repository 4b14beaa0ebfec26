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
