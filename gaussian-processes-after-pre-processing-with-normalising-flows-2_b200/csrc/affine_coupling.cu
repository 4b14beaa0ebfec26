// AffineCoupling arithmetic (flow_modules/affine_coupling.py:100-124), one fused pass:
// split ("split" on x, "cross" on h), sigmoid(raw+2), scale/shift, log-scale per-sample reduction,
// pass-through copy of the untouched half.  16 B per transformed element + 8 B pass-through.
#include "common.cuh"

namespace flowk {

__device__ __forceinline__ float sigmoid_p2(float raw) { return 1.0f / (1.0f + __expf(-(raw + 2.0f))); }

// MODE 0: forward, 1: inverse
template <int VEC, int MODE>
__global__ void __launch_bounds__(kThreads) affine_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                          float* __restrict__ y, const float* __restrict__ ldj_in,
                                                          float* __restrict__ ldj_out, LdjWs ws, int C, int HW) {
  const int c = C >> 1;
  const int E = c * HW;
  const int b = blockIdx.y;
  const float* x1 = x + (size_t)b * C * HW;
  const float* x2 = x1 + E;
  const float* hb = h + (size_t)b * C * HW;
  float* y1 = y + (size_t)b * C * HW;
  float* y2 = y1 + E;
  float local = 0.f;
  const int step = gridDim.x * kThreads * VEC;
  for (int e = (blockIdx.x * kThreads + threadIdx.x) * VEC; e < E; e += step) {
    const int j = e / HW, p = e - j * HW;            // VEC | HW, so a vector never straddles channels
    const float* hs = hb + (size_t)(2 * j) * HW + p;
    if (VEC == 4) {
      float4 xv = ld_stream4(x2 + e), sh = ld_stream4(hs), rw = ld_stream4(hs + HW), pass = ld_stream4(x1 + e);
      float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ss[4] = {sh.x, sh.y, sh.z, sh.w}, rr[4] = {rw.x, rw.y, rw.z, rw.w};
      float o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float sc = sigmoid_p2(rr[q]);
        local += logf(sc);
        o[q] = MODE == 0 ? fmaf(xs[q], sc, ss[q]) : (xs[q] - ss[q]) / sc;
      }
      st_stream4(y2 + e, make_float4(o[0], o[1], o[2], o[3]));
      st_stream4(y1 + e, pass);
    } else {
      float xv = x2[e], sh = hs[0], rw = hs[HW];
      float sc = sigmoid_p2(rw);
      local += logf(sc);
      y2[e] = MODE == 0 ? fmaf(xv, sc, sh) : (xv - sh) / sc;
      y1[e] = x1[e];
    }
  }
  if (ldj_out) finish_sample_ldj<kThreads>(local, ldj_in, ldj_out, MODE == 0 ? 1.f : -1.f, ws);
}

// backward of the forward op:  y2 = x2*s + t, ldj = sum log s, s = sigmoid(r+2)
//   gx2 = gy2*s ; gt = gy2 ; gr = (gy2*x2*s + gldj[b]) * (1-s) ; gx1 = gy1 (conditioner grad is added outside)
template <int VEC>
__global__ void __launch_bounds__(kThreads) affine_bwd_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                              const float* __restrict__ gy,
                                                              const float* __restrict__ gldj, float* __restrict__ gx,
                                                              float* __restrict__ gh, int C, int HW) {
  const int c = C >> 1, E = c * HW, b = blockIdx.y;
  const size_t base = (size_t)b * C * HW;
  const float gl = gldj ? gldj[b] : 0.f;
  const int step = gridDim.x * kThreads * VEC;
  for (int e = (blockIdx.x * kThreads + threadIdx.x) * VEC; e < E; e += step) {
    const int j = e / HW, p = e - j * HW;
    const size_t ho = base + (size_t)(2 * j) * HW + p;
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      float x2 = x[base + E + e + q], r = h[ho + HW + q], g2 = gy[base + E + e + q];
      float s = sigmoid_p2(r);
      gx[base + E + e + q] = g2 * s;
      gx[base + e + q] = gy[base + e + q];
      gh[ho + q] = g2;
      gh[ho + HW + q] = fmaf(g2 * x2, s, gl) * (1.f - s);
    }
  }
}

template <int MODE>
static int launch_affine(const float* x, const float* h, float* y, const float* ldj_in, float* ldj_out, void* ws,
                         int B, int C, int HW, cudaStream_t st) {
  if (B < 0 || C < 2 || (C & 1) || HW < 1 || B > 65535) return FLOWK_ERR_SHAPE;
  if (B == 0) return FLOWK_OK;
  if (!x || !h || !y) return FLOWK_ERR_ARG;
  if (ldj_out && !ws) return FLOWK_ERR_ARG;
  const long long E = (long long)(C / 2) * HW;
  const bool vec4 = (HW % 4 == 0) && aligned16(x) && aligned16(h) && aligned16(y);
  LdjWs w = carve_ws(ws, B);
  if (vec4) {
    // 128-bit accesses whenever the rows allow it; a CTA takes up to 4 vectors per thread, so small samples (E = 1536 at the
    // first CIFAR level) are ONE CTA each and need no cross-CTA reduction
    dim3 grid(parts_for(E, kThreads * 4 * (B >= 2 * 148 ? 4 : 1)), B);
    affine_kernel<4, MODE><<<grid, kThreads, 0, st>>>(x, h, y, ldj_in, ldj_out, w, C, HW);
  } else {
    dim3 grid(parts_for(E, kThreads), B);
    affine_kernel<1, MODE><<<grid, kThreads, 0, st>>>(x, h, y, ldj_in, ldj_out, w, C, HW);
  }
  return launch_status();
}

}  // namespace flowk

using namespace flowk;

extern "C" int flowk_affine_coupling_fwd(const float* x, const float* h, float* y, const float* ldj_in,
                                         float* ldj_out, void* ws, int B, int C, int HW, flowk_stream_t stream) {
  return launch_affine<0>(x, h, y, ldj_in, ldj_out, ws, B, C, HW, stream);
}

extern "C" int flowk_affine_coupling_inv(const float* x, const float* h, float* y, const float* ldj_in,
                                         float* ldj_out, void* ws, int B, int C, int HW, flowk_stream_t stream) {
  return launch_affine<1>(x, h, y, ldj_in, ldj_out, ws, B, C, HW, stream);
}

extern "C" int flowk_affine_coupling_bwd(const float* x, const float* h, const float* gy, const float* gldj,
                                         float* gx, float* gh, int B, int C, int HW, flowk_stream_t stream) {
  if (B < 0 || C < 2 || (C & 1) || HW < 1 || B > 65535) return FLOWK_ERR_SHAPE;
  if (B == 0) return FLOWK_OK;
  if (!x || !h || !gy || !gx || !gh) return FLOWK_ERR_ARG;
  const long long E = (long long)(C / 2) * HW;
  dim3 grid(parts_for(E, kThreads), B);
  affine_bwd_kernel<1><<<grid, kThreads, 0, stream>>>(x, h, gy, gldj, gx, gh, C, HW);
  return launch_status();
}
