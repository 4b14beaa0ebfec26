// Multi-head self-attention core of GatedAttn (flow_modules/mixlogcdf_nn.py:134-147,154-173), inference:
//     out[b, i, h, :] = softmax_j( (q_i . k_j) / sqrt(d) ) v_j        per image b and head h, over the H*W positions.
// Input is the in_proj GEMM's output rows [M, 3C] in the reference's column order (k | v | q) (:136-139); output is
// written as the (hi, lo) TF32 operand pair of the gate GEMM, rows [M, C] with head h at columns h*d..h*d+d-1
// (the permute/view sequence of :146-147 is the identity on [b, seq, c]).
//
// seq <= 1024 and d <= 64 here (seq = 256/64/16, d = 24 at the BASELINE shapes), so the problem per (image, head) is
// tiny: K and V of a pair live in shared memory, one thread owns one query and streams the keys with an online
// softmax in registers (fp32 throughout; every lane reads the same key at the same time -> broadcast LDS.128).
#include "common.cuh"

namespace flowk {

template <int D>
__global__ void __launch_bounds__(256, 2) attention_kernel(const float* __restrict__ qkv, float* __restrict__ out_hi,
                                                        float* __restrict__ out_lo, int HW, int C, int heads,
                                                        int pairs_total, int pairs_per_block, int q_per_block,
                                                        float scale) {
  extern __shared__ __align__(16) float sm[];                  // [pairs_per_block][2][HW][D]
  griddep_launch();
  griddep_wait();                                              // PDL: qkv is the previous kernel's output
  const int row_stride = 3 * C;
  const int pair0 = blockIdx.x * pairs_per_block;
  // cooperative load of K and V of every pair of this block (rows of D contiguous floats, 16-byte vectors)
  constexpr int V4 = D / 4;
  const int vec_per_pair = 2 * HW * V4;
  for (int i = threadIdx.x; i < pairs_per_block * vec_per_pair; i += blockDim.x) {
    const int pl = i / vec_per_pair, r = i - pl * vec_per_pair;
    const int which = r / (HW * V4), rr = r - which * (HW * V4);
    const int j = rr / V4, v4 = rr - j * V4;
    const int pair = pair0 + pl;
    if (pair < pairs_total) {
      const int b = pair / heads, h = pair - b * heads;
      const float4 val = __ldg(reinterpret_cast<const float4*>(qkv + (size_t)(b * HW + j) * row_stride + which * C + h * D) + v4);
      reinterpret_cast<float4*>(sm + ((size_t)(pl * 2 + which) * HW + j) * D)[v4] = val;
    }
  }
  __syncthreads();
  // thread -> (pair, query)
  const int pl = threadIdx.x / q_per_block;
  const int qi = blockIdx.y * q_per_block + (threadIdx.x - pl * q_per_block);
  const int pair = pair0 + pl;
  if (pl >= pairs_per_block || pair >= pairs_total || qi >= HW) return;
  const int b = pair / heads, h = pair - b * heads;
  const size_t m = (size_t)b * HW + qi;
  float q[D], acc[D];
  {
    const float4* qp = reinterpret_cast<const float4*>(qkv + m * row_stride + 2 * C + h * D);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const float4 t = __ldg(qp + i);
      q[4 * i] = t.x * scale; q[4 * i + 1] = t.y * scale; q[4 * i + 2] = t.z * scale; q[4 * i + 3] = t.w * scale;
    }
  }
#pragma unroll
  for (int i = 0; i < D; ++i) acc[i] = 0.f;
  const float* Ks = sm + (size_t)(pl * 2) * HW * D;
  const float* Vs = Ks + (size_t)HW * D;
  float mx = -INFINITY, l = 0.f;
  // keys in groups of 4: four independent dot products (ILP), one running-max update per group
  for (int j = 0; j < HW; j += 4) {                            // HW % 4 == 0 (checked on the host)
    float s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4* kp = reinterpret_cast<const float4*>(Ks + (size_t)(j + u) * D);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int i = 0; i < V4; ++i) {
        const float4 t = kp[i];
        s0 = fmaf(q[4 * i], t.x, s0); s1 = fmaf(q[4 * i + 1], t.y, s1);
        s0 = fmaf(q[4 * i + 2], t.z, s0); s1 = fmaf(q[4 * i + 3], t.w, s1);
      }
      s[u] = s0 + s1;
    }
    const float gmax = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
    if (gmax > mx) {                                           // new running maximum: rescale what has been summed
      const float c = __expf(mx - gmax);
      l *= c;
#pragma unroll
      for (int i = 0; i < D; ++i) acc[i] *= c;
      mx = gmax;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float pw = __expf(s[u] - mx);
      l += pw;
      const float4* vp = reinterpret_cast<const float4*>(Vs + (size_t)(j + u) * D);
#pragma unroll
      for (int i = 0; i < V4; ++i) {
        const float4 t = vp[i];
        acc[4 * i] = fmaf(pw, t.x, acc[4 * i]); acc[4 * i + 1] = fmaf(pw, t.y, acc[4 * i + 1]);
        acc[4 * i + 2] = fmaf(pw, t.z, acc[4 * i + 2]); acc[4 * i + 3] = fmaf(pw, t.w, acc[4 * i + 3]);
      }
    }
  }
  const float inv = 1.0f / l;
  float4* oh = reinterpret_cast<float4*>(out_hi + m * C + h * D);
  float4* ol = reinterpret_cast<float4*>(out_lo + m * C + h * D);
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    float hi[4], lo[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float y = acc[4 * i + u] * inv;
      hi[u] = __uint_as_float(__float_as_uint(y) & 0xffffe000u);
      lo[u] = y - hi[u];
    }
    oh[i] = make_float4(hi[0], hi[1], hi[2], hi[3]);
    ol[i] = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

template <int D>
static int launch_attention(const float* qkv, float* out_hi, float* out_lo, int B, int HW, int C, int heads,
                            cudaStream_t st) {
  const int pairs = B * heads;
  int q_per_block = HW < 256 ? HW : 256;
  int pairs_per_block = 256 / q_per_block;                     // several small pairs share a block
  if (pairs_per_block < 1) pairs_per_block = 1;
  const size_t smem = (size_t)pairs_per_block * 2 * HW * D * sizeof(float);
  if (smem > 220 * 1024) return FLOWK_ERR_SHAPE;
  if (smem > 48 * 1024)
    FLOWK_CUDA_OK(cudaFuncSetAttribute(attention_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((pairs + pairs_per_block - 1) / pairs_per_block, (HW + q_per_block - 1) / q_per_block);
  FLOWK_CUDA_OK(launch_pdl(attention_kernel<D>, grid, dim3(q_per_block * pairs_per_block), smem, st, qkv, out_hi, out_lo,
                           HW, C, heads, pairs, pairs_per_block, q_per_block, 1.0f / sqrtf((float)D)));
  return launch_status();
}

}  // namespace flowk

using namespace flowk;

extern "C" int flowk_attention(const float* qkv, float* out_hi, float* out_lo, int B, int HW, int C, int heads,
                               flowk_stream_t stream) {
  if (B < 0 || HW < 1 || C < 1 || heads < 1 || C % heads) return FLOWK_ERR_SHAPE;
  if (B == 0) return FLOWK_OK;
  if (!qkv || !out_hi || !out_lo) return FLOWK_ERR_ARG;
  if ((HW > 256 && HW % 256) || HW % 4) return FLOWK_ERR_SHAPE;
  switch (C / heads) {
    case 8: return launch_attention<8>(qkv, out_hi, out_lo, B, HW, C, heads, stream);
    case 16: return launch_attention<16>(qkv, out_hi, out_lo, B, HW, C, heads, stream);
    case 24: return launch_attention<24>(qkv, out_hi, out_lo, B, HW, C, heads, stream);
    case 32: return launch_attention<32>(qkv, out_hi, out_lo, B, HW, C, heads, stream);
    case 40: return launch_attention<40>(qkv, out_hi, out_lo, B, HW, C, heads, stream);
    case 64: return launch_attention<64>(qkv, out_hi, out_lo, B, HW, C, heads, stream);
    default: return FLOWK_ERR_SHAPE;
  }
}
