// `Transformer_attn` (the fork's invertible patch attention, reference flow_modules/transformer.py:123-326, used twice per
// FlowStep at marscf_main.py:50-51,69-70) as ONE pass over the activations, inference / sampling.
//
// The image is a 2x2 grid of patches of side p = W/2; patch n, flattened index l = (c, i, j).  Entries with (n + l) even
// (odd when `permute`) condition the attention and pass through; the others are mixed between the patches of equal
// parity by two 2x2 matrices.  With G = sum_i Wq_i^T Wk_i (the six 1x1 convs and three Q K^T products collapsed, built on
// the host side once per weight version):
//
//   score[n, m] = sum_{(i,j)} u_n(i,j)^T G u_m(i,j),   u_n(i,j) = masked channel vector of patch n at in-patch pixel (i,j)
//   attn        = sigmoid(score / scale + offset2) + offset3,   M1 = attn[{0,2},{0,2}] + offset I,  M2 = attn[{1,3},{1,3}] + offset I
//   forward     free entries of patches (0, 2) <- M1 (x0, x2),  of patches (1, 3) <- M2 (x1, x3);  ldj += (log|det M1| + log|det M2|) p (p/2) C
//   reverse     the same with the closed-form inverses; ldj -= ...
//
// One CTA per sample: the sample ([C, H, W] fp32, <= 48 KB) and G sit in shared memory, the eight scores are a fixed-order
// block reduction (bit-reproducible), the mixing pass re-reads shared memory: HBM traffic is one read + one write of the
// activations (8 B per element), the same as the fused ActNorm / 1x1-conv kernel.
#include "common.cuh"

namespace flowk {

constexpr int PA_THREADS = 256;

__global__ void __launch_bounds__(PA_THREADS) patch_attention_kernel(const float* __restrict__ x, const float* __restrict__ G,
                                                                      const float* __restrict__ prm, float* __restrict__ y,
                                                                      const float* __restrict__ ldj_in, float* __restrict__ ldj_out,
                                                                      int C, int H, int W, int permute, int reverse) {
  extern __shared__ __align__(16) float sm[];
  const int HW = H * W, n_el = C * HW, p = W >> 1, pp = p * p;
  float* xs = sm;                         // [C][H][W]
  float* xm = sm + ((n_el + 3) & ~3);     // the same with the free entries zeroed (the conditioning part the scores see)
  float* gs = xm + ((n_el + 3) & ~3);     // [C][C]
  __shared__ float red[PA_THREADS / 32][8];
  __shared__ float mat[8];                // a1 b1 c1 d1 a2 b2 c2 d2 (already inverted when reverse)
  const int b = blockIdx.x;
  const float* xb = x + (size_t)b * n_el;
  float* yb = y + (size_t)b * n_el;
  // conditioning mask of entry l = c p^2 + i p + j of patch n: (n + l) even, inverted by `permute`
  const int perm = permute ? 1 : 0;
  auto is_cond = [&](int n, int c, int i, int j) -> bool { return (((n + c * pp + i * p + j) & 1) ^ perm) == 0; };
  for (int e = threadIdx.x; e < n_el; e += PA_THREADS) {
    const float v = __ldcs(xb + e);
    const int xx = e % W, t = e / W, yy = t % H, c = t / H;
    const int lower = yy >= p, right = xx >= p;
    xs[e] = v;
    xm[e] = is_cond((lower << 1) | right, c, yy - (lower ? p : 0), xx - (right ? p : 0)) ? v : 0.f;
  }
  for (int i = threadIdx.x; i < C * C; i += PA_THREADS) gs[i] = __ldg(G + i);
  __syncthreads();

  // ---- scores: item = (patch m, in-patch pixel, output channel c): v = sum_c' G[c, c'] u_m[c'], then u_n[c] v for both n
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // (0,0) (0,2) (2,0) (2,2) (1,1) (1,3) (3,1) (3,3)
  for (int it = threadIdx.x; it < n_el; it += PA_THREADS) {
    const int c = it % C, r = it / C, m = r & 3, pix = r >> 2;
    const int i = pix / p, j = pix - i * p;
    const int ym = i + ((m >> 1) ? p : 0), xm_col = j + ((m & 1) ? p : 0);
    float v = 0.f;
    const float* grow = gs + c * C;
    const float* xcol = xm + ym * W + xm_col;
#pragma unroll 4
    for (int cc = 0; cc < C; ++cc) v = fmaf(grow[cc], xcol[cc * HW], v);
    // rows n of equal parity: n = m & 1 (upper patch) and n = (m & 1) + 2 (lower patch)
    const int n0 = m & 1, n1 = n0 + 2;
    const int y0 = i, x0 = j + (n0 ? p : 0), y1 = i + p, x1 = x0;
    const float u0 = xm[(c * H + y0) * W + x0];
    const float u1 = xm[(c * H + y1) * W + x1];
    const int base = (m & 1) ? 4 : 0, col = m >> 1;               // score[n, m]: column index of m within its parity class
    acc[base + col] += u0 * v;                                     // n = n0 (row 0 of the 2x2 block)
    acc[base + 2 + col] += u1 * v;                                 // n = n1 (row 1)
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = warp_sum(acc[k]);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[threadIdx.x >> 5][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float t = 0.f;
      for (int w = 0; w < PA_THREADS / 32; ++w) t += red[w][k];
      s[k] = t;
    }
    const float off = prm[0], off2 = prm[1], off3 = prm[2], scale = prm[3];
    float mm[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) mm[k] = 1.0f / (1.0f + expf(-(s[k] / scale + off2))) + off3;
    mm[0] += off; mm[3] += off; mm[4] += off; mm[7] += off;        // diagonals
    const float det1 = mm[0] * mm[3] - mm[1] * mm[2], det2 = mm[4] * mm[7] - mm[5] * mm[6];
    const float ld = (logf(fabsf(det1)) + logf(fabsf(det2))) * (float)(p * (p / 2) * C);
    if (reverse) {
      const float a1 = mm[3] / det1, b1 = -mm[1] / det1, c1 = -mm[2] / det1, d1 = mm[0] / det1;
      const float a2 = mm[7] / det2, b2 = -mm[5] / det2, c2 = -mm[6] / det2, d2 = mm[4] / det2;
      mm[0] = a1; mm[1] = b1; mm[2] = c1; mm[3] = d1; mm[4] = a2; mm[5] = b2; mm[6] = c2; mm[7] = d2;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) mat[k] = mm[k];
    if (ldj_out) ldj_out[b] = (ldj_in ? ldj_in[b] : 0.f) + (reverse ? -ld : ld);
  }
  __syncthreads();

  // ---- mixing: conditioning entries pass through, free entries are mixed with the same entry of the partner patch
  for (int e = threadIdx.x; e < n_el; e += PA_THREADS) {
    const int xx = e % W, t = e / W, yy = t % H, c = t / H;
    float v = xs[e];
    const int lower = yy >= p, right = xx >= p;
    if (!is_cond((lower << 1) | right, c, yy - (lower ? p : 0), xx - (right ? p : 0))) {
      const float partner = xs[(c * H + (lower ? yy - p : yy + p)) * W + xx];
      const float* mtx = mat + (right ? 4 : 0);
      // upper patch (row 0): a x_n + b x_partner;  lower patch (row 1): c x_partner + d x_n
      v = lower ? fmaf(mtx[2], partner, mtx[3] * v) : fmaf(mtx[0], v, mtx[1] * partner);
    }
    __stcs(yb + e, v);
  }
}

}  // namespace flowk

using namespace flowk;

// x, y [B, C, H, W] fp32 (H == W, even); G [C, C] = sum_i Wq_i^T Wk_i; prm = device {offset, offset2, offset3, scale};
// ldj_in / ldj_out [B] (either may be NULL).  reference: flow_modules/transformer.py:123-326.
extern "C" int flowk_patch_attention(const float* x, const float* G, const float* prm, float* y, const float* ldj_in,
                                     float* ldj_out, int B, int C, int H, int W, int permute, int reverse,
                                     flowk_stream_t stream) {
  if (B < 0 || C < 1 || H < 2 || H != W || (W & 1)) return FLOWK_ERR_SHAPE;
  if (B == 0) return FLOWK_OK;
  if (!x || !G || !prm || !y) return FLOWK_ERR_ARG;
  const size_t smem = ((size_t)2 * ((C * H * W + 3) & ~3) + (size_t)C * C) * sizeof(float);
  if (smem > 200 * 1024) return FLOWK_ERR_SHAPE;
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    FLOWK_CUDA_OK(cudaFuncSetAttribute(patch_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  patch_attention_kernel<<<B, PA_THREADS, smem, stream>>>(x, G, prm, y, ldj_in, ldj_out, C, H, W, permute, reverse);
  return launch_status();
}
