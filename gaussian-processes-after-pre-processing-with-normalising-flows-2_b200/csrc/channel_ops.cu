// squeeze2d / unsqueeze2d, ActNorm (init + per-channel affine) and the per-pixel channel mixing
// that implements InvertibleConv1x1 (optionally with ActNorm and Squeeze folded in).
// Reference semantics: flow_modules/common_modules.py:12-186.  All HBM-bound: 8 B per element.
#include "common.cuh"

namespace flowk {

// ---------------------------------------------------------------------------------------------
// squeeze / unsqueeze, factor 2 fast path: each thread moves one 2-wide input pair.
// squeeze:  thread reads float2 x[b,c,hi,2w..2w+1] -> y[b,4c+2(hi&1)+{0,1}, hi>>1, w]
// ---------------------------------------------------------------------------------------------
__global__ void squeeze2_kernel(const float* __restrict__ x, float* __restrict__ y,
                                int C, int H, int W, long long total_pairs) {
  const int Wo = W >> 1, Ho = H >> 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_pairs;
       i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % Wo);
    long long r = i / Wo;
    int hi = (int)(r % H);
    r /= H;
    int c = (int)(r % C);
    long long b = r / C;
    float2 v = __ldcs(reinterpret_cast<const float2*>(x) + i);
    size_t plane = (size_t)Ho * Wo;
    size_t o = ((size_t)(b * C + c) * 4 + 2 * (hi & 1)) * plane + (size_t)(hi >> 1) * Wo + w;
    y[o] = v.x;
    y[o + plane] = v.y;
  }
}

__global__ void unsqueeze2_kernel(const float* __restrict__ x, float* __restrict__ y,
                                  int Co, int Hi, int Wi, long long total_pairs) {
  // x: [B, 4Co, Hi, Wi] -> y: [B, Co, 2Hi, 2Wi]; thread writes float2 y[b,c,ho,2w..2w+1]
  const int Ho = Hi * 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_pairs;
       i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % Wi);
    long long r = i / Wi;
    int ho = (int)(r % Ho);
    r /= Ho;
    int c = (int)(r % Co);
    long long b = r / Co;
    size_t plane = (size_t)Hi * Wi;
    size_t s = ((size_t)(b * Co + c) * 4 + 2 * (ho & 1)) * plane + (size_t)(ho >> 1) * Wi + w;
    float2 v;
    v.x = __ldcs(x + s);
    v.y = __ldcs(x + s + plane);
    __stcs(reinterpret_cast<float2*>(y) + i, v);
  }
}

// generic factor (rare: the reference only ever uses 2, marscf_main.py:130)
__global__ void squeeze_generic_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int H, int W,
                                       int f, long long total, bool inverse) {
  // forward: x [B,C,H,W] -> y [B,C f f,H/f,W/f];  inverse: x [B,C,H,W] -> y [B,C/(f f),H f,W f]
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    if (!inverse) {
      int w = (int)(i % W);
      long long r = i / W;
      int h = (int)(r % H);
      r /= H;
      int c = (int)(r % C);
      long long b = r / C;
      int Ho = H / f, Wo = W / f;
      size_t o = (((size_t)(b * C + c) * f * f + (h % f) * f + (w % f)) * Ho + h / f) * Wo + w / f;
      y[o] = x[i];
    } else {
      int Co = C / (f * f), Ho = H * f, Wo = W * f;
      int w = (int)(i % Wo);
      long long r = i / Wo;
      int h = (int)(r % Ho);
      r /= Ho;
      int c = (int)(r % Co);
      long long b = r / Co;
      size_t s = (((size_t)(b * Co + c) * f * f + (h % f) * f + (w % f)) * H + h / f) * W + w / f;
      y[i] = x[s];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ActNorm data-dependent init: one CTA per channel, two passes (mean, then centred second moment)
// exactly as the reference does it (common_modules.py:145-147); second pass hits L2.
// ---------------------------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ double block_sum_d(double v) {
  __shared__ double sh[THREADS / 32];
  __shared__ double total;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < THREADS / 32; ++w) s += sh[w];
    total = s;
  }
  __syncthreads();
  double r = total;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(256) actnorm_init_kernel(const float* __restrict__ x, float* __restrict__ bias,
                                                           float* __restrict__ logs, int B, int C, int HW,
                                                           float scale, float eps) {
  const int c = blockIdx.x;
  const long long n = (long long)B * HW;
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    long long b = i / HW;
    int p = (int)(i - b * HW);
    acc += (double)x[((size_t)b * C + c) * HW + p];
  }
  const float mean = (float)(block_sum_d<256>(acc) / (double)n);
  const float nb = -mean;
  acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    long long b = i / HW;
    int p = (int)(i - b * HW);
    float d = x[((size_t)b * C + c) * HW + p] + nb;
    acc += (double)(d * d);
  }
  const float var = (float)(block_sum_d<256>(acc) / (double)n);
  if (threadIdx.x == 0) {
    bias[c] = nb;
    logs[c] = logf(scale / (sqrtf(var) + eps));
  }
}

// ---------------------------------------------------------------------------------------------
// per-channel affine: y = (x + pre[c]) * mul[c] + post[c]
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void channel_scale_kernel(const float* __restrict__ x, const float* __restrict__ pre,
                                     const float* __restrict__ mul, const float* __restrict__ post,
                                     float* __restrict__ y, const float* __restrict__ ldj_in,
                                     const float* __restrict__ ldj_add, float* __restrict__ ldj_out,
                                     int B, int C, int HW) {
  const int bc = blockIdx.y;           // b*C + c
  const int c = bc % C;
  const float a = pre ? pre[c] : 0.f, m = mul[c], o = post ? post[c] : 0.f;
  const float* xs = x + (size_t)bc * HW;
  float* ys = y + (size_t)bc * HW;
  if (VEC == 4) {
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < HW; i += gridDim.x * blockDim.x * 4) {
      float4 v = ld_stream4(xs + i);
      v.x = (v.x + a) * m + o; v.y = (v.y + a) * m + o; v.z = (v.z + a) * m + o; v.w = (v.w + a) * m + o;
      st_stream4(ys + i, v);
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x)
      ys[i] = (xs[i] + a) * m + o;
  }
  if (ldj_out && blockIdx.x == 0 && threadIdx.x == 0 && c == 0) {
    int b = bc / C;
    ldj_out[b] = (ldj_in ? ldj_in[b] : 0.f) + (ldj_add ? ldj_add[0] : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// channel mixing (1x1 conv): each thread owns PIX pixels, keeps all C inputs in registers and
// streams the C x C matrix through shared memory (broadcast reads).
// SQ: 0 plain, 1 squeeze-on-load, 2 unsqueeze-on-store.
// ---------------------------------------------------------------------------------------------
template <int C, int PIX, int SQ>
__global__ void __launch_bounds__(128) channel_mix_kernel(const float* __restrict__ x, const float* __restrict__ Wm,
                                                          const float* __restrict__ bias, float* __restrict__ y,
                                                          const float* __restrict__ ldj_in,
                                                          const float* __restrict__ ldj_add,
                                                          float* __restrict__ ldj_out, int B, int H, int W) {
  __shared__ __align__(16) float w_s[C * C];
  __shared__ float b_s[C];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) w_s[i] = Wm[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) b_s[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int HW = H * W;
  // Pixel groups are numbered across the WHOLE batch (groups per image = ceil(HW / PIX)): small feature maps (4x4, 8x8)
  // still fill every lane of the CTA, and the C x C matrix is staged once per 128 pixel groups instead of once per image.
  const int gpi = (HW + PIX - 1) / PIX;
  const long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= (long long)B * gpi) return;
  const int b = (int)(gidx / gpi);
  const int p0 = (int)(gidx - (long long)b * gpi) * PIX;
  if (ldj_out && p0 == 0) ldj_out[b] = (ldj_in ? ldj_in[b] : 0.f) + (ldj_add ? ldj_add[0] : 0.f);
  const float* xb = x + (size_t)b * C * HW;
  float* yb = y + (size_t)b * C * HW;

  // address of (channel ch, pixel p) under the optional squeeze maps
  auto in_off = [&](int ch, int p) -> size_t {
    if (SQ == 1) {                 // x is [C/4, 2H, 2W]
      int h = p / W, w = p - h * W;
      int c0 = ch >> 2, fh = (ch >> 1) & 1, fw = ch & 1;
      return ((size_t)c0 * (2 * H) + (2 * h + fh)) * (2 * W) + 2 * w + fw;
    }
    return (size_t)ch * HW + p;
  };
  auto out_off = [&](int ch, int p) -> size_t {
    if (SQ == 2) {                 // y is [C/4, 2H, 2W]
      int h = p / W, w = p - h * W;
      int c0 = ch >> 2, fh = (ch >> 1) & 1, fw = ch & 1;
      return ((size_t)c0 * (2 * H) + (2 * h + fh)) * (2 * W) + 2 * w + fw;
    }
    return (size_t)ch * HW + p;
  };

  float v[C][PIX];
  const bool full = (p0 + PIX <= HW);
  if (PIX == 4 && SQ != 1 && full) {
#pragma unroll
    for (int i = 0; i < C; ++i) {
      float4 t = ld_stream4(xb + (size_t)i * HW + p0);
      v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
    }
  } else if (PIX == 2 && SQ != 1 && full) {
#pragma unroll
    for (int i = 0; i < C; ++i) {
      float2 t = __ldcs(reinterpret_cast<const float2*>(xb + (size_t)i * HW + p0));
      v[i][0] = t.x; v[i][PIX - 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
      for (int q = 0; q < PIX; ++q) v[i][q] = (p0 + q < HW) ? __ldcs(xb + in_off(i, p0 + q)) : 0.f;
  }

#pragma unroll 2
  for (int o = 0; o < C; ++o) {
    float acc[PIX];
#pragma unroll
    for (int q = 0; q < PIX; ++q) acc[q] = b_s[o];
    const float* wr = w_s + o * C;
#pragma unroll
    for (int i = 0; i < C; i += 4) {
      float4 w4 = *reinterpret_cast<const float4*>(wr + i);
#pragma unroll
      for (int q = 0; q < PIX; ++q) {
        acc[q] = fmaf(w4.x, v[i][q], acc[q]);
        acc[q] = fmaf(w4.y, v[i + 1][q], acc[q]);
        acc[q] = fmaf(w4.z, v[i + 2][q], acc[q]);
        acc[q] = fmaf(w4.w, v[i + 3][q], acc[q]);
      }
    }
    if (PIX == 4 && SQ != 2 && full) {
      st_stream4(yb + (size_t)o * HW + p0, make_float4(acc[0], acc[1], acc[2], acc[3]));
    } else if (PIX == 2 && SQ != 2 && full) {
      __stcs(reinterpret_cast<float2*>(yb + (size_t)o * HW + p0), make_float2(acc[0], acc[PIX - 1]));
    } else {
#pragma unroll
      for (int q = 0; q < PIX; ++q)
        if (p0 + q < HW) yb[out_off(o, p0 + q)] = acc[q];
    }
  }
}

// any C (<= 1024): matrix row in smem, inputs re-read through L1.  Correctness path for odd widths.
template <int SQ>
__global__ void channel_mix_generic_kernel(const float* __restrict__ x, const float* __restrict__ Wm,
                                           const float* __restrict__ bias, float* __restrict__ y,
                                           const float* __restrict__ ldj_in, const float* __restrict__ ldj_add,
                                           float* __restrict__ ldj_out, int C, int H, int W) {
  const int HW = H * W, b = blockIdx.y;
  if (ldj_out && blockIdx.x == 0 && threadIdx.x == 0)
    ldj_out[b] = (ldj_in ? ldj_in[b] : 0.f) + (ldj_add ? ldj_add[0] : 0.f);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int h = p / W, w = p - h * W;
  auto off = [&](int ch, bool mapped) -> size_t {
    if (mapped) {
      int c0 = ch >> 2, fh = (ch >> 1) & 1, fw = ch & 1;
      return ((size_t)c0 * (2 * H) + (2 * h + fh)) * (2 * W) + 2 * w + fw;
    }
    return (size_t)ch * HW + p;
  };
  const float* xb = x + (size_t)b * C * HW;
  float* yb = y + (size_t)b * C * HW;
  for (int o = 0; o < C; ++o) {
    float acc = bias ? bias[o] : 0.f;
    for (int i = 0; i < C; ++i) acc = fmaf(__ldg(Wm + (size_t)o * C + i), __ldg(xb + off(i, SQ == 1)), acc);
    yb[off(o, SQ == 2)] = acc;
  }
}

template <int C, int SQ>
static int launch_mix(const float* x, const float* Wm, const float* bias, float* y, const float* ldj_in,
                      const float* ldj_add, float* ldj_out, int B, int H, int W, cudaStream_t st) {
  const int HW = H * W;
  // 4 pixels/thread (128-bit accesses) once there are enough pixels to fill the machine
  constexpr bool kCanVec4 = (C <= 24);
  const bool vec4 = kCanVec4 && (HW % 4 == 0) && ((long long)B * HW >= 4LL * 148 * 128 * 4) && aligned16(x) &&
                    aligned16(y) && SQ == 0;
  // wider matrices: two pixels per thread halve the shared-memory (matrix) reads per FMA, the bound of this kernel for C >= 32
  constexpr bool kCanVec2 = (C > 24 && C <= 48);
  const bool vec2 = kCanVec2 && (HW % 2 == 0) && ((long long)B * HW >= 2LL * 148 * 128 * 4) && SQ == 0 &&
                    (reinterpret_cast<uintptr_t>(x) & 7u) == 0 && (reinterpret_cast<uintptr_t>(y) & 7u) == 0;
  if (vec2) {
    const long long groups = (long long)B * (HW / 2);
    channel_mix_kernel<C, kCanVec2 ? 2 : 1, SQ><<<(unsigned)((groups + 127) / 128), 128, 0, st>>>(x, Wm, bias, y, ldj_in, ldj_add,
                                                                                                 ldj_out, B, H, W);
    return launch_status();
  }
  if (vec4) {
    const long long groups = (long long)B * (HW / 4);
    channel_mix_kernel<C, kCanVec4 ? 4 : 1, SQ><<<(unsigned)((groups + 127) / 128), 128, 0, st>>>(x, Wm, bias, y, ldj_in, ldj_add,
                                                                                                 ldj_out, B, H, W);
  } else {
    const long long groups = (long long)B * HW;
    channel_mix_kernel<C, 1, SQ><<<(unsigned)((groups + 127) / 128), 128, 0, st>>>(x, Wm, bias, y, ldj_in, ldj_add, ldj_out, B, H, W);
  }
  return launch_status();
}

template <int SQ>
static int dispatch_mix(const float* x, const float* Wm, const float* bias, float* y, const float* ldj_in,
                        const float* ldj_add, float* ldj_out, int B, int C, int H, int W, cudaStream_t st) {
  switch (C) {
#define FLOWK_MIX_CASE(CC) \
  case CC: return launch_mix<CC, SQ>(x, Wm, bias, y, ldj_in, ldj_add, ldj_out, B, H, W, st);
    FLOWK_MIX_CASE(4) FLOWK_MIX_CASE(8) FLOWK_MIX_CASE(12) FLOWK_MIX_CASE(16) FLOWK_MIX_CASE(24)
    FLOWK_MIX_CASE(32) FLOWK_MIX_CASE(48) FLOWK_MIX_CASE(64) FLOWK_MIX_CASE(96)
#undef FLOWK_MIX_CASE
    default: {
      dim3 grid((H * W + 127) / 128, B);
      channel_mix_generic_kernel<SQ><<<grid, 128, 0, st>>>(x, Wm, bias, y, ldj_in, ldj_add, ldj_out, C, H, W);
      return launch_status();
    }
  }
}

}  // namespace flowk

using namespace flowk;

extern "C" int flowk_squeeze2d(const float* x, float* y, int B, int C, int H, int W, int factor,
                               flowk_stream_t stream) {
  if (B < 0 || C < 1 || H < 1 || W < 1 || factor < 1) return FLOWK_ERR_SHAPE;
  if (H % factor || W % factor) return FLOWK_ERR_SHAPE;
  long long total = (long long)B * C * H * W;
  if (total == 0) return FLOWK_OK;
  if (!x || !y) return FLOWK_ERR_ARG;
  if (factor == 1) {
    FLOWK_CUDA_OK(cudaMemcpyAsync(y, x, total * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    return FLOWK_OK;
  }
  if (factor == 2 && (reinterpret_cast<uintptr_t>(x) & 7u) == 0) {
    long long pairs = total / 2;
    int blocks = (int)((pairs + 255) / 256 < 148 * 16 ? (pairs + 255) / 256 : 148 * 16);
    squeeze2_kernel<<<blocks, 256, 0, stream>>>(x, y, C, H, W, pairs);
  } else {
    int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    squeeze_generic_kernel<<<blocks, 256, 0, stream>>>(x, y, C, H, W, factor, total, false);
  }
  return launch_status();
}

extern "C" int flowk_unsqueeze2d(const float* x, float* y, int B, int C, int H, int W, int factor,
                                 flowk_stream_t stream) {
  if (B < 0 || C < 1 || H < 1 || W < 1 || factor < 1) return FLOWK_ERR_SHAPE;
  if (C % (factor * factor)) return FLOWK_ERR_SHAPE;
  long long total = (long long)B * C * H * W;
  if (total == 0) return FLOWK_OK;
  if (!x || !y) return FLOWK_ERR_ARG;
  if (factor == 1) {
    FLOWK_CUDA_OK(cudaMemcpyAsync(y, x, total * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    return FLOWK_OK;
  }
  if (factor == 2 && (reinterpret_cast<uintptr_t>(y) & 7u) == 0) {
    long long pairs = total / 2;
    int blocks = (int)((pairs + 255) / 256 < 148 * 16 ? (pairs + 255) / 256 : 148 * 16);
    unsqueeze2_kernel<<<blocks, 256, 0, stream>>>(x, y, C / 4, H, W, pairs);
  } else {
    int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    squeeze_generic_kernel<<<blocks, 256, 0, stream>>>(x, y, C, H, W, factor, total, true);
  }
  return launch_status();
}

extern "C" int flowk_actnorm_init(const float* x, float* bias, float* logs, int B, int C, int HW, float scale,
                                  float eps, flowk_stream_t stream) {
  if (!x || !bias || !logs) return FLOWK_ERR_ARG;
  if (B < 1 || C < 1 || HW < 1) return FLOWK_ERR_SHAPE;
  actnorm_init_kernel<<<C, 256, 0, stream>>>(x, bias, logs, B, C, HW, scale, eps);
  return launch_status();
}

extern "C" int flowk_channel_scale(const float* x, const float* pre, const float* mul, const float* post, float* y,
                                   const float* ldj_in, const float* ldj_add, float* ldj_out, int B, int C, int HW,
                                   flowk_stream_t stream) {
  if (B < 0 || C < 1 || HW < 1) return FLOWK_ERR_SHAPE;
  if (B == 0) return FLOWK_OK;
  if (!x || !y || !mul) return FLOWK_ERR_ARG;
  if ((long long)B * C > 65535LL * 1) {
    // gridDim.y limit: fold into several launches
    int done = 0;
    while (done < B) {
      int nb = (65535 / C) < (B - done) ? (65535 / C) : (B - done);
      if (nb < 1) return FLOWK_ERR_SHAPE;
      int st = flowk_channel_scale(x + (size_t)done * C * HW, pre, mul, post, y + (size_t)done * C * HW,
                                   ldj_in ? ldj_in + done : nullptr, ldj_add, ldj_out ? ldj_out + done : nullptr,
                                   nb, C, HW, stream);
      if (st) return st;
      done += nb;
    }
    return FLOWK_OK;
  }
  const bool vec4 = (HW % 4 == 0) && aligned16(x) && aligned16(y);
  if (vec4) {
    dim3 grid((HW / 4 + 127) / 128 > 64 ? 64 : (HW / 4 + 127) / 128, B * C);
    channel_scale_kernel<4><<<grid, 128, 0, stream>>>(x, pre, mul, post, y, ldj_in, ldj_add, ldj_out, B, C, HW);
  } else {
    dim3 grid((HW + 127) / 128 > 64 ? 64 : (HW + 127) / 128, B * C);
    channel_scale_kernel<1><<<grid, 128, 0, stream>>>(x, pre, mul, post, y, ldj_in, ldj_add, ldj_out, B, C, HW);
  }
  return launch_status();
}

extern "C" int flowk_channel_mix(const float* x, const float* Wm, const float* bias, float* y, const float* ldj_in,
                                 const float* ldj_add, float* ldj_out, int B, int C, int H, int W, int in_squeeze,
                                 int out_unsqueeze, flowk_stream_t stream) {
  if (B < 0 || C < 1 || H < 1 || W < 1 || C > 1024) return FLOWK_ERR_SHAPE;
  if ((in_squeeze || out_unsqueeze) && (C % 4)) return FLOWK_ERR_SHAPE;
  if (in_squeeze && out_unsqueeze) return FLOWK_ERR_ARG;
  if (B == 0) return FLOWK_OK;
  if (!x || !y || !Wm) return FLOWK_ERR_ARG;
  if (B > 65535) {
    for (int done = 0; done < B; done += 65535) {
      int nb = B - done < 65535 ? B - done : 65535;
      size_t off = (size_t)done * C * H * W;
      int st = flowk_channel_mix(x + off, Wm, bias, y + off, ldj_in ? ldj_in + done : nullptr, ldj_add,
                                 ldj_out ? ldj_out + done : nullptr, nb, C, H, W, in_squeeze, out_unsqueeze, stream);
      if (st) return st;
    }
    return FLOWK_OK;
  }
  if (in_squeeze) return dispatch_mix<1>(x, Wm, bias, y, ldj_in, ldj_add, ldj_out, B, C, H, W, stream);
  if (out_unsqueeze) return dispatch_mix<2>(x, Wm, bias, y, ldj_in, ldj_add, ldj_out, B, C, H, W, stream);
  return dispatch_mix<0>(x, Wm, bias, y, ldj_in, ldj_add, ldj_out, B, C, H, W, stream);
}
