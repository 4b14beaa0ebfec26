// Once-per-weight-version preparation of the inference operands, as flowk kernels instead of chains of small library ops.
//
// flowk_pack_weight_f16: conv / linear weight [N, cin, taps] (torch layout) -> the K-major fp16 (hi, lo) operand pair
//   [N, taps, cin_pad] of FLOWK_OPERAND_F16, with the per-output-channel gain fused in:
//     FLOWK_PACK_PLAIN        w
//     FLOWK_PACK_WEIGHT_NORM  w = v g / ||v||      (mixlogcdf_nn.py:19-21, old-style weight_g / weight_v)
//     FLOWK_PACK_EXP_GAIN     w e^{factor * gain}  (ActNorm folded into Conv2d: affine_coupling.py:31-39; Conv2dZeros:
//                             affine_coupling.py:57-63 with factor = logscale_factor), bias_out = bias_in e^{factor * gain}
//   and the power-of-two pre-scaling 2^e that puts max|w 2^e| into [2^14, 2^15) (so the lo part keeps its 11 bits clear
//   of fp16's subnormal range); ws[N + 1] receives 2^-e, the `acc_scale` of flowk_conv_gemm.
//
// flowk_fold_actnorm_invconv: ActNorm followed by the LU-parametrised invertible 1x1 convolution (common_modules.py:57-127,
//   :130-187; one FlowStep's first two layers, marscf_main.py:64-68 / :95-97) folded into ONE per-pixel affine map
//   (matrix, bias, log-det term) for flowk_channel_mix_*, assembled in fp64 by a single CTA:
//     forward   M = P (L U) diag(e^{logs}),   b' = M b,    ldj = +(sum(logs) H W + sum(log_s) W^2)
//     reverse   M = diag(e^{-logs}) U^-1 L^-1 P^T,  b' = -b,  ldj = -(...)
//   with L = tril(l, -1) + I and U = triu(u, 1) + diag(sign_s e^{log_s}); the inverse is two triangular solves per column.
#include "common.cuh"

namespace flowk {

// ------------------------------------------------------------------------------------------------ weight packing
__global__ void __launch_bounds__(128) pack_rows_kernel(const float* __restrict__ w, const float* __restrict__ gain, int mode,
                                                        float factor, const float* __restrict__ bias_in,
                                                        float* __restrict__ bias_out, int N, int cols, float* __restrict__ ws) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* p = w + (size_t)row * cols;
  float ss = 0.f, mx = 0.f;
  for (int i = lane; i < cols; i += 32) {
    const float v = p[i];
    ss = fmaf(v, v, ss);
    mx = fmaxf(mx, fabsf(v));
  }
  ss = warp_sum(ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) {
    float s = 1.f;
    if (mode == FLOWK_PACK_WEIGHT_NORM) s = gain[row] / sqrtf(ss);
    else if (mode == FLOWK_PACK_EXP_GAIN) s = expf(gain[row] * factor);
    ws[row] = s;
    const float top = fabsf(s) * mx;             // NaN / inf never win the max below: the scale then stays 1
    if (top > 0.f && top <= 3.0e38f) atomicMax(reinterpret_cast<unsigned*>(ws + N), __float_as_uint(top));
    if (bias_out) bias_out[row] = bias_in[row] * s;
  }
}

__global__ void __launch_bounds__(256) pack_split_kernel(const float* __restrict__ w, float* __restrict__ ws, int N, int cin,
                                                         int taps, int cin_pad, unsigned short* __restrict__ hi,
                                                         unsigned short* __restrict__ lo) {
  const unsigned bits = reinterpret_cast<const unsigned*>(ws)[N];
  int e = 0;
  if (bits != 0u) {
    const int field = (int)((bits >> 23) & 0xffu);
    e = field == 0 ? 24 : 14 - (field - 127);
    e = e < -14 ? -14 : (e > 24 ? 24 : e);
  }
  const float scale = __int_as_float((127 + e) << 23);
  if (blockIdx.x == 0 && threadIdx.x == 0) ws[N + 1] = __int_as_float((127 - e) << 23);
  const unsigned row = (unsigned)taps * cin_pad, total = (unsigned)N * row;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned n = i / row, r = i - n * row;
    const unsigned t = r / cin_pad, c = r - t * cin_pad;
    float val = 0.f;
    if (c < (unsigned)cin) val = (w[((size_t)n * cin + c) * taps + t] * ws[n]) * scale;
    unsigned short h, l;
    split_f16(val, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// ------------------------------------------------------------------------------------------------ ActNorm + 1x1 conv fold
constexpr int FOLD_THREADS = 256;

__global__ void __launch_bounds__(FOLD_THREADS) fold_actnorm_invconv_kernel(
    const float* __restrict__ l, const float* __restrict__ u, const float* __restrict__ log_s, const float* __restrict__ p,
    const float* __restrict__ sign_s, const float* __restrict__ logs, const float* __restrict__ bias, int C, int HW, int WW,
    int reverse, float* __restrict__ mat, float* __restrict__ bias_out, float* __restrict__ ldj) {
  extern __shared__ __align__(16) double fold_sm[];
  double* T = fold_sm;                 // [C][C]
  double* diag = T + C * C;            // U's diagonal
  double* gain = diag + C;             // e^{+-logs}
  const int tid = threadIdx.x;
  auto L = [&](int i, int k) -> double { return i == k ? 1.0 : (k < i ? (double)l[i * C + k] : 0.0); };
  auto U = [&](int i, int k) -> double { return i == k ? diag[i] : (k > i ? (double)u[i * C + k] : 0.0); };
  for (int i = tid; i < C; i += FOLD_THREADS) {
    diag[i] = (double)(sign_s[i] * expf(log_s[i]));                 // rounded to fp32 like the module's `up`
    gain[i] = (double)(float)exp(reverse ? -(double)logs[i] : (double)logs[i]);
  }
  __syncthreads();
  if (!reverse) {
    // T = L U (row i needs k <= i of L, column j needs k <= j of U)
    for (int e = tid; e < C * C; e += FOLD_THREADS) {
      const int i = e / C, j = e - i * C;
      const int kmax = i < j ? i : j;
      double acc = 0.0;
      for (int k = 0; k <= kmax; ++k) acc = fma(L(i, k), U(k, j), acc);
      T[e] = (double)(float)acc;
    }
    __syncthreads();
    // M = (P T) diag(gain), written as fp32; then b' = M b
    for (int e = tid; e < C * C; e += FOLD_THREADS) {
      const int i = e / C, j = e - i * C;
      double acc = 0.0;
      for (int k = 0; k < C; ++k) acc = fma((double)p[i * C + k], T[k * C + j], acc);
      mat[e] = (float)((double)(float)acc * gain[j]);
    }
    __syncthreads();
    for (int i = tid; i < C; i += FOLD_THREADS) {
      double acc = 0.0;
      for (int j = 0; j < C; ++j) acc = fma((double)mat[i * C + j], (double)bias[j], acc);
      bias_out[i] = (float)acc;
    }
  } else {
    // column j of W^-1 = U^-1 L^-1 P^T: forward substitution L y = P^T[:, j], then back substitution U x = y in place
    for (int j = tid; j < C; j += FOLD_THREADS) {
      for (int i = 0; i < C; ++i) {
        double acc = (double)p[j * C + i];                           // P^T[i, j]
        for (int k = 0; k < i; ++k) acc = fma(-(double)l[i * C + k], T[k * C + j], acc);
        T[i * C + j] = acc;
      }
      for (int i = C - 1; i >= 0; --i) {
        double acc = T[i * C + j];
        for (int k = i + 1; k < C; ++k) acc = fma(-(double)u[i * C + k], T[k * C + j], acc);
        T[i * C + j] = acc / diag[i];
      }
    }
    __syncthreads();
    for (int e = tid; e < C * C; e += FOLD_THREADS) {
      const int i = e / C;
      mat[e] = (float)(gain[i] * (double)(float)T[e]);
    }
    for (int i = tid; i < C; i += FOLD_THREADS) bias_out[i] = -bias[i];
  }
  if (tid == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < C; ++i) {
      a += (double)logs[i];
      b += (double)log_s[i];
    }
    const float d = (float)a * (float)HW + (float)b * (float)WW;
    ldj[0] = reverse ? -d : d;
  }
}

}  // namespace flowk

using namespace flowk;

extern "C" int flowk_pack_weight_f16(const float* w, const float* gain, int mode, float factor, const float* bias_in,
                                     float* bias_out, int N, int cin, int taps, int cin_pad, void* hi, void* lo, float* ws,
                                     flowk_stream_t stream) {
  if (N < 1 || cin < 1 || taps < 1 || cin_pad < cin || (cin_pad & 7)) return FLOWK_ERR_SHAPE;
  if ((long long)N * taps * cin_pad >= (1ll << 31)) return FLOWK_ERR_SHAPE;
  if (mode != FLOWK_PACK_PLAIN && mode != FLOWK_PACK_WEIGHT_NORM && mode != FLOWK_PACK_EXP_GAIN) return FLOWK_ERR_ARG;
  if (!w || !hi || !lo || !ws || (mode != FLOWK_PACK_PLAIN && !gain) || (bias_out && !bias_in)) return FLOWK_ERR_ARG;
  FLOWK_CUDA_OK(cudaMemsetAsync(ws + N, 0, 2 * sizeof(float), stream));
  pack_rows_kernel<<<(N + 3) / 4, 128, 0, stream>>>(w, gain, mode, factor, bias_in, bias_out, N, cin * taps, ws);
  const long long total = (long long)N * taps * cin_pad;
  const int blocks = (int)((total + 255) / 256 < 148 * 4 ? (total + 255) / 256 : 148 * 4);
  pack_split_kernel<<<blocks, 256, 0, stream>>>(w, ws, N, cin, taps, cin_pad, reinterpret_cast<unsigned short*>(hi),
                                                reinterpret_cast<unsigned short*>(lo));
  return launch_status();
}

extern "C" int flowk_fold_actnorm_invconv(const float* l, const float* u, const float* log_s, const float* p,
                                          const float* sign_s, const float* logs, const float* bias, int C, int H, int W,
                                          int reverse, float* mat, float* bias_out, float* ldj, flowk_stream_t stream) {
  if (C < 1 || H < 1 || W < 1) return FLOWK_ERR_SHAPE;
  if (!l || !u || !log_s || !p || !sign_s || !logs || !bias || !mat || !bias_out || !ldj) return FLOWK_ERR_ARG;
  const size_t smem = ((size_t)C * C + 2 * (size_t)C) * sizeof(double);
  if (smem > 200 * 1024) return FLOWK_ERR_SHAPE;
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    FLOWK_CUDA_OK(cudaFuncSetAttribute(fold_actnorm_invconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  fold_actnorm_invconv_kernel<<<1, FOLD_THREADS, smem, stream>>>(l, u, log_s, p, sign_s, logs, bias, C, H * W, W * W, reverse,
                                                                  mat, bias_out, ldj);
  return launch_status();
}
