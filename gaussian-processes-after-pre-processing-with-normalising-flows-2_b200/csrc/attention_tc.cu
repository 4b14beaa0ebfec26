// Self-attention core of GatedAttn (flow_modules/mixlogcdf_nn.py:134-147) on the 5th-gen tensor cores, inference.
//
//   att[q, :] = softmax_k(q . k / sqrt(d)) v        per (image, head); seq = H*W in {128, 256}, d = C / heads <= 64
//
// One CTA per (image, head); its 128-query tiles run one after the other with K / V staged once:
//   1. all 16 warps stage Q (pre-scaled by log2(e) / sqrt(d)), K and V^T of the pair from the fp32 in_proj rows
//      (k | v | q column order) into shared memory as fp16 (hi, lo) operand tiles, written straight in the K-major
//      128-byte-swizzled layout tcgen05 reads (what a SWIZZLE_128B TMA load would have produced);
//   2. S = Q K^T: tcgen05.mma kind::f16, M = 128 queries, N = seq keys, accumulators in TMEM (seq columns); fp32
//      accuracy from the two-term split S = Qh Kh + Ql Kh + Qh Kl (same scheme as the conv GEMMs, csrc/tc_gemm.cu);
//   3. softmax: a thread owns half of one score row (TMEM lane): row maximum, p = 2^(s - max), row sum; P goes back
//      to shared memory as the (hi, lo) A operand of the second GEMM, over the dead Q / K tiles;
//   4. O = P V: M = 128, N = d (padded to 16), K = seq; O lands in the TMEM columns S no longer needs;
//   5. epilogue: O / rowsum -> the (hi, lo) operand pair of the gate GEMM, fp16 or TF32 containers.
// Nothing of size seq x seq touches HBM.  Replaces the mma.sync kernel (csrc/attention.cu) for seq in {128, 256}.
#include <cuda.h>
#include "common.cuh"
#include "umma.cuh"

namespace flowk {
namespace tc {

constexpr int ATT_THREADS = 512;                // 16 warps: 4 threads share a score row (TMEM lane), a quarter of the keys each
constexpr int ROW_B = 128;                      // bytes per swizzle row: 64 fp16

// byte offset of the 16-byte chunk `chunk` (0..7) of row `r` inside a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw_off(int r, int chunk) { return (uint32_t)(r * ROW_B + ((chunk ^ (r & 7)) << 4)); }

// 8 floats -> 8 fp16 hi + 8 fp16 lo (x = hi + lo), packed; |x| must be < 65504 (q, k, v and the softmax weights are)
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));      // low half = a
  return r;
}
__device__ __forceinline__ void unpack_h2(uint32_t v, float& a, float& b) {
  unsigned short lo16 = (unsigned short)(v & 0xffffu), hi16 = (unsigned short)(v >> 16);
  asm("cvt.f32.f16 %0, %1;" : "=f"(a) : "h"(lo16));
  asm("cvt.f32.f16 %0, %1;" : "=f"(b) : "h"(hi16));
}
__device__ __forceinline__ void cvt8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_h2(v[2 * i], v[2 * i + 1]);
    float a, b;
    unpack_h2(h[i], a, b);
    l[i] = pack_h2(v[2 * i] - a, v[2 * i + 1] - b);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

struct AttParams {
  const float* qkv;         // [B*HW, 3C], columns (k | v | q)
  float* out_hi;            // [B*HW, C] operand pair (fp16 when out_f16, else TF32 values in fp32 containers)
  float* out_lo;
  int out_f16;
  int HW, C, heads, D;      // D = C / heads (multiple of 8, <= 64)
  int dk;                   // D rounded up to 16: contraction length of S and MMA N of O
  float qscale;             // log2(e) / sqrt(D)
  int q_off, k_off, v_off, p_off, o_off;   // byte offsets of the tile groups in dynamic shared memory
  int* status;
  long long* trace;         // optional device [8]: clock64 stamps of CTA (0,0) (profiling), NULL in production
};

template <int KCH>           // KCH = dk / 8: 16-byte chunks per staged row the MMAs read (compile-time: index math without divisions)
__global__ void __launch_bounds__(ATT_THREADS, 1) attention_tc_kernel(const AttParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_s[2], bar_o[2];
  __shared__ uint32_t tmem_slot;
  __shared__ int failed_flag;
  __shared__ float red_max[4][128], red_sum[2][4][128];
  volatile int* failed = &failed_flag;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x, b = pair / p.heads, h = pair - b * p.heads;
  const int HW = p.HW, D = p.D, C = p.C, row_stride = 3 * C;
  const int tiles = HW >> 7;                    // 128-query tiles of this (image, head): all handled by this CTA, K / V staged once
  const int chunks = D >> 3;                    // 16-byte chunks of real data per row
  constexpr int kchunks = KCH;                  // chunks the MMAs read (a zero chunk pads D % 16 == 8)
  const int hw_shift = HW == 256 ? 8 : 7;
  const int tmem_cols = HW <= 128 ? 128 : 512;  // one S accumulator (HW columns) per query tile

  if (threadIdx.x == 0) {
    failed_flag = 0;
    mbar_init(&bar_s[0], 1);
    mbar_init(&bar_s[1], 1);
    mbar_init(&bar_o[0], 1);
    mbar_init(&bar_o[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const bool tracing = p.trace && blockIdx.x == 0 && threadIdx.x == 0;
  griddep_launch();
  griddep_wait();                               // the in_proj GEMM's rows are complete from here on
  if (tracing) p.trace[0] = clock64();

  // ---- 1. stage Q (all tiles), K (K-major rows) and V^T as fp16 (hi, lo) swizzled operand tiles ----------------------
  uint8_t* q_hi = smem + p.q_off;               // [HW rows][128 B]: query tile t = rows [128 t, 128 t + 128)
  uint8_t* q_lo = q_hi + HW * ROW_B;
  uint8_t* k_hi = smem + p.k_off;               // [HW rows][128 B]
  uint8_t* k_lo = k_hi + HW * ROW_B;
  const int vt_tile = 2 * p.dk * ROW_B;         // one 64-key block of V^T: [dk rows of hi | dk rows of lo][128 B]
  uint8_t* v_t = smem + p.v_off;                // [HW / 64 blocks][2 dk rows][128 B]
  {
    // One item = 8 consecutive floats (32 bytes) of one row of Q, K or V.  A thread issues the loads of six items before it
    // converts and scatters them (two L2 round trips for the whole staging phase).
    const float* base = p.qkv + (size_t)b * HW * row_stride + h * D;
    const int nk = HW * kchunks, total = 3 * nk;
    constexpr int MAXI = 6;
    for (int i0 = threadIdx.x; i0 < total; i0 += MAXI * ATT_THREADS) {
      float4 a[MAXI], c[MAXI];
#pragma unroll
      for (int u = 0; u < MAXI; ++u) {
        const int i = i0 + u * ATT_THREADS;
        a[u] = c[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total) {
          int r, ch, col;
          if (i < nk) { r = i / kchunks; ch = i % kchunks; col = 2 * C; }
          else if (i < 2 * nk) { r = (i - nk) / kchunks; ch = (i - nk) % kchunks; col = 0; }
          else { r = (i - 2 * nk) & (HW - 1); ch = (i - 2 * nk) >> hw_shift; col = C; }   // V: consecutive threads -> consecutive keys
          if (ch < chunks) {
            const float4* g = reinterpret_cast<const float4*>(base + (size_t)r * row_stride + col + ch * 8);
            a[u] = __ldg(g);
            c[u] = __ldg(g + 1);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < MAXI; ++u) {
        const int i = i0 + u * ATT_THREADS;
        if (i >= total) continue;
        if (i < 2 * nk) {                       // Q (scaled) and K: K-major rows as they lie
          const bool isq = i < nk;
          const int j = isq ? i : i - nk;
          const int r = j / kchunks, ch = j % kchunks;
          const float sc = isq ? p.qscale : 1.f;
          const float v[8] = {a[u].x * sc, a[u].y * sc, a[u].z * sc, a[u].w * sc, c[u].x * sc, c[u].y * sc, c[u].z * sc, c[u].w * sc};
          uint4 hi, lo;
          cvt8(v, hi, lo);
          *reinterpret_cast<uint4*>((isq ? q_hi : k_hi) + sw_off(r, ch)) = hi;
          *reinterpret_cast<uint4*>((isq ? q_lo : k_lo) + sw_off(r, ch)) = lo;
        } else {
          // V^T: rows = head dims, contraction (keys) along the 128-byte rows, 64 keys per block; the 8 dims of this key are
          // scattered to 8 rows of the transposed tile.  The lo rows of a block follow its hi rows, so [V_hi ; V_lo] is ONE
          // B operand of N = 2 dk rows.
          const int j = i - 2 * nk, key = j & (HW - 1), ch = j >> hw_shift;
          const float v[8] = {a[u].x, a[u].y, a[u].z, a[u].w, c[u].x, c[u].y, c[u].z, c[u].w};
          const int kb = key >> 6, kk = key & 63;
          uint8_t* tile = v_t + (size_t)kb * vt_tile;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int dim = ch * 8 + jj;
            unsigned short hh, ll;
            split_f16(v[jj], hh, ll);
            *reinterpret_cast<unsigned short*>(tile + sw_off(dim, kk >> 3) + ((kk & 7) << 1)) = hh;
            *reinterpret_cast<unsigned short*>(tile + sw_off(p.dk + dim, kk >> 3) + ((kk & 7) << 1)) = ll;
          }
        }
      }
    }
  }
  fence_proxy_async();                          // generic-proxy smem writes -> visible to the tensor core
  __syncthreads();
  if (tracing) p.trace[1] = clock64();

  // ---- 2. S_t = Q_t K^T for every query tile, issued back to back (tile t into TMEM columns [t HW, t HW + HW)) ---------
  // (MMAs are issued by warp 0 as a whole with uniform operands, the instruction predicated on the elected lane: umma.cuh)
  const uint32_t leader = warp == 0 ? elect_one() : 0u;
  if (warp == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(HW);
    const uint64_t dk_hi = make_smem_desc(smem_u32(k_hi)), dk_lo = make_smem_desc(smem_u32(k_lo));
    const int ksteps = p.dk >> 4;
    for (int t = 0; t < tiles; ++t) {
      const uint64_t dq_hi = make_smem_desc(smem_u32(q_hi + t * 128 * ROW_B)), dq_lo = make_smem_desc(smem_u32(q_lo + t * 128 * ROW_B));
      const uint32_t d = tmem_base + (uint32_t)(t * HW);
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t ko = (uint64_t)(ks * 2); // 32 bytes per K = 16 step, in 16-byte units
        umma_elect<true>(d, dq_hi + ko, dk_hi + ko, idesc, ks == 0 ? 0u : 1u, leader);
        umma_elect<true>(d, dq_lo + ko, dk_hi + ko, idesc, 1u, leader);
        umma_elect<true>(d, dq_hi + ko, dk_lo + ko, idesc, 1u, leader);
      }
      umma_commit_elect(&bar_s[t], leader);
    }
  }

  const int lane_grp = warp & 3, quarter = warp >> 2;
  const int row = lane_grp * 32 + lane;
  const int ncols = HW >> 2, c0 = quarter * ncols;
  uint8_t* p_hi = smem + p.p_off;               // [HW / 64 blocks][128 rows][128 B]: over the dead Q / K tiles
  uint8_t* p_lo = p_hi + (HW >> 6) * (128 * ROW_B);
  uint8_t* stage = smem + p.o_off;              // epilogue staging: [2 (hi, lo)][128 rows][chunks] 16-byte items

  // epilogue of tile t: O / rowsum -> operand pair.  fp16 output goes through shared memory so that global stores are 16-byte
  // chunks of whole head rows; TF32 output: direct pair stores.
  auto epilogue = [&](int t) {
    mbar_wait(&bar_o[t], 0, failed);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(t * HW);
    const float inv = 1.0f / ((red_sum[t][0][row] + red_sum[t][1][row]) + (red_sum[t][2][row] + red_sum[t][3][row]));
    const int q0 = t * 128;
    const size_t o0 = (size_t)(b * HW + q0 + row) * C + h * D;
    for (int ch = quarter; ch < chunks; ch += 4) {
      float a[8], c[8];
      tmem_ld8(trow + ch * 8, a);
      tmem_ld8(trow + p.dk + ch * 8, c);
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i] = (a[i] + c[i]) * inv;
      if (p.out_f16) {
        uint4 hi, lo;
        cvt8(c, hi, lo);
        *reinterpret_cast<uint4*>(stage + ((size_t)row * chunks + ch) * 16) = hi;
        *reinterpret_cast<uint4*>(stage + ((size_t)(128 + row) * chunks + ch) * 16) = lo;
      } else {
#pragma unroll
        for (int i = 0; i < 8; i += 2) store_pair(p.out_hi, p.out_lo, o0 + ch * 8 + i, c[i], c[i + 1], 0);
      }
    }
    if (p.out_f16) {
      __syncthreads();
      const int items = 2 * 128 * chunks;
      for (int i = threadIdx.x; i < items; i += ATT_THREADS) {
        const int mat = i / (128 * chunks), rr = (i / chunks) % 128, ch = i % chunks;
        const uint4 v = *reinterpret_cast<const uint4*>(stage + (size_t)i * 16);
        unsigned short* dst = reinterpret_cast<unsigned short*>(mat ? p.out_lo : p.out_hi);
        *reinterpret_cast<uint4*>(dst + (size_t)(b * HW + q0 + rr) * C + h * D + ch * 8) = v;
      }
      __syncthreads();                          // the staging area is free again
    }
  };

  for (int t = 0; t < tiles; ++t) {
    const uint32_t trow = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(t * HW);
    // ---- 3a. row maximum of S_t (thread = (row, quarter of the keys)); overlaps the previous tile's P V MMAs
    mbar_wait(&bar_s[t], 0, failed);
    tc_fence_after();
    if (tracing && t == 0) p.trace[2] = clock64();
    float mx = -INFINITY;
    for (int j = 0; j < ncols; j += 16) {
      float s[16];
      tmem_ld16(trow + c0 + j, s);
#pragma unroll
      for (int i = 0; i < 16; ++i) mx = fmaxf(mx, s[i]);
    }
    red_max[quarter][row] = mx;
    // P overwrites Q / K (tile 0: every S_t must have been computed) resp. the previous tile's P (its P V must be done)
    if (t == 0) mbar_wait(&bar_s[tiles - 1], 0, failed);
    else mbar_wait(&bar_o[t - 1], 0, failed);
    __syncthreads();
    if (tracing && t == 0) p.trace[3] = clock64();
    mx = fmaxf(fmaxf(red_max[0][row], red_max[1][row]), fmaxf(red_max[2][row], red_max[3][row]));
    // ---- 3b. p = 2^(s - max), row sums, P as the (hi, lo) A operand of the second GEMM
    float sum = 0.f;
    for (int j = 0; j < ncols; j += 16) {
      float s[16];
      tmem_ld16(trow + c0 + j, s);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        s[i] = ex2_fast(s[i] - mx);
        sum += s[i];
      }
      const int key = c0 + j, kb = key >> 6, ch = (key & 63) >> 3;       // two 8-key chunks
      uint4 hi, lo;
      cvt8(s, hi, lo);
      *reinterpret_cast<uint4*>(p_hi + kb * (128 * ROW_B) + sw_off(row, ch)) = hi;
      *reinterpret_cast<uint4*>(p_lo + kb * (128 * ROW_B) + sw_off(row, ch)) = lo;
      cvt8(s + 8, hi, lo);
      *reinterpret_cast<uint4*>(p_hi + kb * (128 * ROW_B) + sw_off(row, ch + 1)) = hi;
      *reinterpret_cast<uint4*>(p_lo + kb * (128 * ROW_B) + sw_off(row, ch + 1)) = lo;
    }
    red_sum[t][quarter][row] = sum;
    tc_fence_before();                          // our tcgen05.ld of S_t are complete before the MMAs overwrite those columns
    fence_proxy_async();
    __syncthreads();
    if (tracing && t == 0) p.trace[4] = clock64();

    // ---- 4. O_t = P V.  B operand = [V_hi ; V_lo] (N = 2 dk): P_hi [V_hi ; V_lo] fills columns [0, dk) and [dk, 2 dk) of S_t's
    //         (dead) leading columns in one MMA, P_lo V_hi adds to [0, dk); the epilogue sums the two column groups.
    if (warp == 0) {
      tc_fence_after();
      const uint32_t idesc2 = make_idesc_f16(2 * p.dk), idesc1 = make_idesc_f16(p.dk);
      const uint32_t d = tmem_base + (uint32_t)(t * HW);
      const int kblocks = HW >> 6;
      for (int kb = 0; kb < kblocks; ++kb) {
        const uint64_t dp_hi = make_smem_desc(smem_u32(p_hi + kb * (128 * ROW_B)));
        const uint64_t dp_lo = make_smem_desc(smem_u32(p_lo + kb * (128 * ROW_B)));
        const uint64_t dv = make_smem_desc(smem_u32(v_t + (size_t)kb * vt_tile));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ko = (uint64_t)(k * 2);
          umma_elect<true>(d, dp_hi + ko, dv + ko, idesc2, (kb | k) == 0 ? 0u : 1u, leader);
          umma_elect<true>(d, dp_lo + ko, dv + ko, idesc1, 1u, leader);
        }
      }
      umma_commit_elect(&bar_o[t], leader);
    }
    // ---- 5. the PREVIOUS tile's epilogue runs while this tile's P V MMAs execute
    if (t > 0) epilogue(t - 1);
  }
  if (tracing) p.trace[5] = clock64();
  epilogue(tiles - 1);
  tc_fence_before();
  __syncthreads();
  if (tracing) p.trace[6] = clock64();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  if (threadIdx.x == 0 && failed_flag && p.status) *p.status = 1;
}

}  // namespace tc
}  // namespace flowk

using namespace flowk;
using namespace flowk::tc;

// Same contract as flowk_attention / flowk_attention_f16 (include/flowk.h); returns FLOWK_ERR_SHAPE for shapes this
// kernel does not take (the caller then uses the mma.sync kernel): seq must be 128 or 256, C / heads a multiple of 8, <= 64.
extern "C" int flowk_attention_tc(const float* qkv, void* out_hi, void* out_lo, int out_f16, int B, int HW, int C, int heads,
                                  int* status, long long* trace, flowk_stream_t stream) {
  if (B < 0 || HW < 1 || C < 1 || heads < 1 || C % heads) return FLOWK_ERR_SHAPE;
  const int D = C / heads;
  if ((HW != 128 && HW != 256) || D % 8 || D > 64 || C % 4) return FLOWK_ERR_SHAPE;
  if (B == 0) return FLOWK_OK;
  if (!qkv || !out_hi || !out_lo) return FLOWK_ERR_ARG;
  if ((long long)B * heads > 0x7fffffffLL) return FLOWK_ERR_SHAPE;
  AttParams p{};
  p.qkv = qkv;
  p.out_hi = reinterpret_cast<float*>(out_hi);
  p.out_lo = reinterpret_cast<float*>(out_lo);
  p.out_f16 = out_f16;
  p.HW = HW;
  p.C = C;
  p.heads = heads;
  p.D = D;
  p.dk = (D + 15) / 16 * 16;
  p.qscale = 1.4426950408889634f / sqrtf((float)D);
  p.status = status;
  p.trace = trace;
  // shared memory: [P tiles of one query tile | over them: Q (all tiles), K], then V^T, then the epilogue staging area
  const int p_bytes = 2 * (HW / 64) * 128 * ROW_B;           // hi + lo
  const int qk_bytes = 2 * HW * ROW_B + 2 * HW * ROW_B;
  const int first = p_bytes > qk_bytes ? p_bytes : qk_bytes;
  p.p_off = 0;
  p.q_off = 0;
  p.k_off = 2 * HW * ROW_B;
  p.v_off = first;
  p.o_off = first + 2 * (HW / 64) * p.dk * ROW_B;
  const size_t smem = (size_t)p.o_off + (size_t)2 * 128 * (D / 8) * 16 + 1024;
  if (smem + 8 * 1024 > 227 * 1024) return FLOWK_ERR_SHAPE;   // (+ the kernel's static shared memory)
#define FLOWK_LAUNCH_ATT(KCH_)                                                                                          \
  do {                                                                                                                 \
    static size_t smem_set = 0;                                                                                        \
    if (smem > smem_set) {                                                                                             \
      FLOWK_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<KCH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      smem_set = smem;                                                                                                 \
    }                                                                                                                  \
    FLOWK_CUDA_OK(launch_pdl(attention_tc_kernel<KCH_>, dim3(B * heads), dim3(ATT_THREADS), smem, stream, p));          \
  } while (0)
  switch (p.dk >> 3) {
    case 2: FLOWK_LAUNCH_ATT(2); break;
    case 4: FLOWK_LAUNCH_ATT(4); break;
    case 6: FLOWK_LAUNCH_ATT(6); break;
    default: FLOWK_LAUNCH_ATT(8); break;
  }
#undef FLOWK_LAUNCH_ATT
  return launch_status();
}
