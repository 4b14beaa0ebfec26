// Multi-head self-attention core of GatedAttn (flow_modules/mixlogcdf_nn.py:134-147,154-173), inference:
//     out[b, i, h, :] = softmax_j( (q_i . k_j) / sqrt(d) ) v_j        per image b and head h, over the H*W positions.
// Input is the in_proj GEMM's output rows [M, 3C] in the reference's column order (k | v | q) (:136-139); output is
// written as the (hi, lo) TF32 operand pair of the gate GEMM, rows [M, C] with head h at columns h*d..h*d+d-1
// (the permute/view sequence of :146-147 is the identity on [b, seq, c]).
//
// seq <= 1024 and d <= 64 here (seq = 256/64/16, d = 24 at the BASELINE shapes), so the problem per (image, head) is
// tiny: K and V of a pair live in shared memory, one thread owns one query and streams the keys with an online
// softmax in registers (fp32 throughout; every lane reads the same key at the same time -> broadcast LDS.128).
#include <stdlib.h>
#include "common.cuh"

namespace flowk {

template <int D>
__global__ void __launch_bounds__(256, 2) attention_kernel(const float* __restrict__ qkv, float* __restrict__ out_hi,
                                                        float* __restrict__ out_lo, int out_f16, int HW, int C, int heads,
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
  const size_t o0 = (size_t)m * C + h * D;
#pragma unroll
  for (int i = 0; i < D; i += 2) store_pair(out_hi, out_lo, o0 + i, acc[i] * inv, acc[i + 1] * inv, out_f16);
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-level tensor-core version (mma.sync m16n8k8 TF32, 3xTF32 split for fp32 accuracy).  The problem per
// (image, head) - seq <= 1024, d <= 64 - is far too small for a 128-row tcgen05 tile pipeline, so this is a
// flash-attention style kernel: a warp owns 16 queries, keys stream through in blocks of 8, online softmax in the
// accumulator fragments, P is re-used as the A operand of P*V by permuting the key order inside a block
// (accumulator columns (2t, 2t+1) == operand columns (t, t+4) of the permuted keys).  ~4x fewer issue slots per
// query-key pair than the scalar kernel above.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_bits(float x, uint32_t& hi, uint32_t& lo) {   // round-to-nearest TF32 split
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
}

template <int D, int NKB>      // NKB key blocks (of 8 keys) per softmax update
__global__ void __launch_bounds__(256, 2) attention_mma_kernel(const float* __restrict__ qkv, float* __restrict__ out_hi,
                                                            float* __restrict__ out_lo, int out_f16, int HW, int C, int heads,
                                                            int pairs_total, int pairs_per_block, int warps_per_pair,
                                                            int key_tile, float scale) {
  extern __shared__ __align__(16) float sm[];                  // [pairs_per_block][Khi,Klo,Vhi,Vlo][key_tile][D + 4]
  constexpr int P = D + 4;                                     // row pitch: conflict-free fragment loads
  constexpr int KS = D / 8;                                    // k-steps of QK^T == n-blocks of PV
  griddep_launch();
  griddep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int pl = warp / warps_per_pair;                        // pair handled by this warp
  const int pair = blockIdx.x * pairs_per_block + pl;
  const bool pair_ok = pl < pairs_per_block && pair < pairs_total;
  const int b = pair_ok ? pair / heads : 0, h = pair_ok ? pair - b * heads : 0;
  const int q0 = (blockIdx.y * warps_per_pair + (warp - pl * warps_per_pair)) * 16;    // first query of this warp
  const int row_stride = 3 * C;
  const bool q_lo_ok = pair_ok && q0 + g < HW, q_hi_ok = pair_ok && q0 + g + 8 < HW;

  // Q fragments (scaled), split once
  uint32_t qh[KS][4], ql[KS][4];
  {
    const float* q_lo_p = qkv + (size_t)(b * HW + (q_lo_ok ? q0 + g : 0)) * row_stride + 2 * C + h * D;
    const float* q_hi_p = qkv + (size_t)(b * HW + (q_hi_ok ? q0 + g + 8 : 0)) * row_stride + 2 * C + h * D;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const float v0 = q_lo_ok ? __ldg(q_lo_p + ks * 8 + t) * scale : 0.f, v1 = q_hi_ok ? __ldg(q_hi_p + ks * 8 + t) * scale : 0.f;
      const float v2 = q_lo_ok ? __ldg(q_lo_p + ks * 8 + t + 4) * scale : 0.f, v3 = q_hi_ok ? __ldg(q_hi_p + ks * 8 + t + 4) * scale : 0.f;
      split_bits(v0, qh[ks][0], ql[ks][0]); split_bits(v1, qh[ks][1], ql[ks][1]);
      split_bits(v2, qh[ks][2], ql[ks][2]); split_bits(v3, qh[ks][3], ql[ks][3]);
    }
  }
  float o[KS][4];
#pragma unroll
  for (int nb = 0; nb < KS; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
  float mx_lo = -INFINITY, mx_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

  // K and V are split into their (hi, lo) TF32 parts ONCE, when the tile is staged: every element is consumed by
  // all 16 warps of the pair, so splitting at the point of use would repeat the work 16 times
  const uint32_t* Kh = reinterpret_cast<const uint32_t*>(sm) + (size_t)(pl < pairs_per_block ? pl : 0) * 4 * key_tile * P;
  const uint32_t* Kl = Kh + (size_t)key_tile * P;
  const uint32_t* Vh = Kl + (size_t)key_tile * P;
  const uint32_t* Vl = Vh + (size_t)key_tile * P;
  constexpr int V4 = D / 4;
  for (int kt0 = 0; kt0 < HW; kt0 += key_tile) {
    __syncthreads();                                           // previous tile fully consumed
    {                                                          // cooperative K/V tile load, all pairs of the block
      const int vec_per_pair = 2 * key_tile * V4;
      const int total = pairs_per_block * vec_per_pair;
      constexpr int U = 4;                                     // loads in flight per thread before the first use
      for (int i0 = threadIdx.x; i0 < total; i0 += U * blockDim.x) {
        float4 val[U];
        uint32_t* dst[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * blockDim.x;
          dst[u] = nullptr;
          val[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < total) {
            const int pp = i / vec_per_pair, r = i - pp * vec_per_pair;
            const int which = r / (key_tile * V4), rr = r - which * (key_tile * V4);      // 0: K, 1: V
            const int j = rr / V4, v4 = rr - j * V4;
            const int pr = blockIdx.x * pairs_per_block + pp;
            if (pr < pairs_total && kt0 + j < HW) {
              const int bb = pr / heads, hh = pr - bb * heads;
              val[u] = __ldg(reinterpret_cast<const float4*>(qkv + (size_t)(bb * HW + kt0 + j) * row_stride + which * C + hh * D) + v4);
              dst[u] = reinterpret_cast<uint32_t*>(sm) + ((size_t)(pp * 4 + which * 2) * key_tile + j) * P + v4 * 4;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (dst[u]) {
            uint4 hi, lo;
            split_bits(val[u].x, hi.x, lo.x); split_bits(val[u].y, hi.y, lo.y);
            split_bits(val[u].z, hi.z, lo.z); split_bits(val[u].w, hi.w, lo.w);
            *reinterpret_cast<uint4*>(dst[u]) = hi;
            *reinterpret_cast<uint4*>(dst[u] + (size_t)key_tile * P) = lo;
          }
        }
      }
    }
    __syncthreads();
    const int keys_here = HW - kt0 < key_tile ? HW - kt0 : key_tile;
    for (int k0 = 0; k0 < keys_here; k0 += 8 * NKB) {
      // ---- S = Q K^T for NKB blocks of 8 keys
      float sc[NKB][4];
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        sc[kb][0] = sc[kb][1] = sc[kb][2] = sc[kb][3] = 0.f;
        const size_t ko = (size_t)(k0 + kb * 8 + g) * P + t;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint32_t bh0 = Kh[ko + ks * 8], bh1 = Kh[ko + ks * 8 + 4];
          const uint32_t bl0 = Kl[ko + ks * 8], bl1 = Kl[ko + ks * 8 + 4];
          mma_tf32(sc[kb], qh[ks], bh0, bh1);
          mma_tf32(sc[kb], ql[ks], bh0, bh1);
          mma_tf32(sc[kb], qh[ks], bl0, bl1);
        }
      }
      // ---- online softmax: rows g (c0,c1) and g+8 (c2,c3); a row lives in the 4 lanes of a quad
      float m_lo = sc[0][0], m_hi = sc[0][2];
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        m_lo = fmaxf(m_lo, fmaxf(sc[kb][0], sc[kb][1]));
        m_hi = fmaxf(m_hi, fmaxf(sc[kb][2], sc[kb][3]));
      }
      m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1)); m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
      m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1)); m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
      const float nm_lo = fmaxf(mx_lo, m_lo), nm_hi = fmaxf(mx_hi, m_hi);
      const float c_lo = __expf(mx_lo - nm_lo), c_hi = __expf(mx_hi - nm_hi);
      mx_lo = nm_lo; mx_hi = nm_hi;
      l_lo *= c_lo; l_hi *= c_hi;
#pragma unroll
      for (int nb = 0; nb < KS; ++nb) { o[nb][0] *= c_lo; o[nb][1] *= c_lo; o[nb][2] *= c_hi; o[nb][3] *= c_hi; }
      // ---- O += P V, P re-used straight from the accumulator fragment (keys permuted inside the block)
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        const float p0 = __expf(sc[kb][0] - mx_lo), p1 = __expf(sc[kb][1] - mx_lo);
        const float p2 = __expf(sc[kb][2] - mx_hi), p3 = __expf(sc[kb][3] - mx_hi);
        l_lo += p0 + p1;
        l_hi += p2 + p3;
        uint32_t ph[4], plo[4];
        split_bits(p0, ph[0], plo[0]);      // a0 = (row g,   k = t)   <- key 2t
        split_bits(p2, ph[1], plo[1]);      // a1 = (row g+8, k = t)   <- key 2t
        split_bits(p1, ph[2], plo[2]);      // a2 = (row g,   k = t+4) <- key 2t+1
        split_bits(p3, ph[3], plo[3]);      // a3 = (row g+8, k = t+4) <- key 2t+1
        const size_t vo = (size_t)(k0 + kb * 8 + 2 * t) * P + g;
#pragma unroll
        for (int nb = 0; nb < KS; ++nb) {
          const uint32_t bh0 = Vh[vo + nb * 8], bl0 = Vl[vo + nb * 8];            // b0 = (k = t,   n = g) <- V[key 2t  ][dim nb*8+g]
          const uint32_t bh1 = Vh[vo + P + nb * 8], bl1 = Vl[vo + P + nb * 8];    // b1 = (k = t+4, n = g) <- V[key 2t+1][dim nb*8+g]
          mma_tf32(o[nb], ph, bh0, bh1);
          mma_tf32(o[nb], plo, bh0, bh1);
          mma_tf32(o[nb], ph, bl0, bl1);
        }
      }
    }
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float i_lo = 1.0f / l_lo, i_hi = 1.0f / l_hi;
#pragma unroll
  for (int nb = 0; nb < KS; ++nb) {
    const int col = h * D + nb * 8 + 2 * t;
    if (q_lo_ok) store_pair(out_hi, out_lo, (size_t)(b * HW + q0 + g) * C + col, o[nb][0] * i_lo, o[nb][1] * i_lo, out_f16);
    if (q_hi_ok) store_pair(out_hi, out_lo, (size_t)(b * HW + q0 + g + 8) * C + col, o[nb][2] * i_hi, o[nb][3] * i_hi, out_f16);
  }
}

template <int D>
static int launch_attention_mma(const float* qkv, float* out_hi, float* out_lo, int out_f16, int B, int HW, int C, int heads,
                                cudaStream_t st) {
  const int pairs = B * heads;
  // 8 warps (128 queries) per CTA and 128-key tiles: ~21 K registers and <= 57 KB of shared memory per CTA, so three
  // CTAs share an SM and their dependent MMA chains overlap (one 16-warp CTA per SM left the tensor pipe ~45 % idle)
  constexpr int kWarps = 8;
  int warps_per_pair = (HW + 15) / 16;
  if (warps_per_pair > kWarps) warps_per_pair = kWarps;
  int pairs_per_block = kWarps / warps_per_pair;
  if (pairs_per_block < 1) pairs_per_block = 1;
  const int key_tile = HW < 128 ? HW : 128;
  const size_t smem = (size_t)pairs_per_block * 4 * key_tile * (D + 4) * sizeof(float);      // K, V as (hi, lo)
  if (smem > 220 * 1024) return FLOWK_ERR_SHAPE;
  dim3 grid((pairs + pairs_per_block - 1) / pairs_per_block, (HW + warps_per_pair * 16 - 1) / (warps_per_pair * 16));
  const int threads = 32 * warps_per_pair * pairs_per_block;
  const float scale = 1.0f / sqrtf((float)D);
  if (HW % 32 == 0) {
    if (smem > 48 * 1024)
      FLOWK_CUDA_OK(cudaFuncSetAttribute(attention_mma_kernel<D, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLOWK_CUDA_OK(launch_pdl(attention_mma_kernel<D, 4>, grid, dim3(threads), smem, st, qkv, out_hi, out_lo, out_f16, HW, C, heads,
                             pairs, pairs_per_block, warps_per_pair, key_tile, scale));
  } else {
    if (smem > 48 * 1024)
      FLOWK_CUDA_OK(cudaFuncSetAttribute(attention_mma_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLOWK_CUDA_OK(launch_pdl(attention_mma_kernel<D, 1>, grid, dim3(threads), smem, st, qkv, out_hi, out_lo, out_f16, HW, C, heads,
                             pairs, pairs_per_block, warps_per_pair, key_tile, scale));
  }
  return launch_status();
}

template <int D>
static int launch_attention(const float* qkv, float* out_hi, float* out_lo, int out_f16, int B, int HW, int C, int heads,
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
                           out_f16, HW, C, heads, pairs, pairs_per_block, q_per_block, 1.0f / sqrtf((float)D)));
  return launch_status();
}

}  // namespace flowk

using namespace flowk;

static int attention_impl(const float* qkv, float* out_hi, float* out_lo, int out_f16, int B, int HW, int C, int heads,
                          flowk_stream_t stream) {
  if (B < 0 || HW < 1 || C < 1 || heads < 1 || C % heads) return FLOWK_ERR_SHAPE;
  if (B == 0) return FLOWK_OK;
  if (!qkv || !out_hi || !out_lo) return FLOWK_ERR_ARG;
  if ((HW > 256 && HW % 256) || HW % 4) return FLOWK_ERR_SHAPE;
  static int use_mma = -1;
  if (use_mma < 0) { const char* e = getenv("FLOWK_ATTENTION_MMA"); use_mma = (e && e[0] == '0') ? 0 : 1; }
  if (use_mma && HW % 8 == 0) {
    switch (C / heads) {
      case 8: return launch_attention_mma<8>(qkv, out_hi, out_lo, out_f16, B, HW, C, heads, stream);
      case 16: return launch_attention_mma<16>(qkv, out_hi, out_lo, out_f16, B, HW, C, heads, stream);
      case 24: return launch_attention_mma<24>(qkv, out_hi, out_lo, out_f16, B, HW, C, heads, stream);
      case 32: return launch_attention_mma<32>(qkv, out_hi, out_lo, out_f16, B, HW, C, heads, stream);
      case 40: return launch_attention_mma<40>(qkv, out_hi, out_lo, out_f16, B, HW, C, heads, stream);
      case 64: return launch_attention_mma<64>(qkv, out_hi, out_lo, out_f16, B, HW, C, heads, stream);
      default: return FLOWK_ERR_SHAPE;
    }
  }
  switch (C / heads) {
    case 8: return launch_attention<8>(qkv, out_hi, out_lo, out_f16, B, HW, C, heads, stream);
    case 16: return launch_attention<16>(qkv, out_hi, out_lo, out_f16, B, HW, C, heads, stream);
    case 24: return launch_attention<24>(qkv, out_hi, out_lo, out_f16, B, HW, C, heads, stream);
    case 32: return launch_attention<32>(qkv, out_hi, out_lo, out_f16, B, HW, C, heads, stream);
    case 40: return launch_attention<40>(qkv, out_hi, out_lo, out_f16, B, HW, C, heads, stream);
    case 64: return launch_attention<64>(qkv, out_hi, out_lo, out_f16, B, HW, C, heads, stream);
    default: return FLOWK_ERR_SHAPE;
  }
}

extern "C" int flowk_attention(const float* qkv, float* out_hi, float* out_lo, int B, int HW, int C, int heads,
                               flowk_stream_t stream) {
  return attention_impl(qkv, out_hi, out_lo, 0, B, HW, C, heads, stream);
}

extern "C" int flowk_attention_f16(const float* qkv, void* out_hi, void* out_lo, int B, int HW, int C, int heads,
                                   flowk_stream_t stream) {
  if (C % 2) return FLOWK_ERR_SHAPE;
  return attention_impl(qkv, reinterpret_cast<float*>(out_hi), reinterpret_cast<float*>(out_lo), 1, B, HW, C, heads, stream);
}
