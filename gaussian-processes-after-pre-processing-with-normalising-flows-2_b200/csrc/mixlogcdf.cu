// MixLogCDF coupling arithmetic (flow_modules/mixlogcdf_coupling.py:37-57, flow_modules/log_dist.py),
// one fused pass per direction.  Per transformed element the forward reads x + 98 parameter planes
// straight out of the conditioner's raw output layout and writes one value: 400 algorithmic bytes.
//
// Numerics: the 32-component sums are evaluated in the LINEAR domain with a single shared stabiliser
// (max_k pi_k) - 4 MUFU ops per component instead of the ~6 of three separate log-sum-exps:
//     W = sum_k e^{pi_k - m},  cdf = sum_k e^{pi_k - m} sigma(z_k),  pdf = sum_k e^{pi_k - m} e^{-s_k} sigma(z_k) sigma(-z_k)
//     u = cdf / W,  log f = log pdf - log W.
// When a linear sum underflows (every component > ~70 scale-widths away) the element is redone in the
// log domain exactly as the reference does (three max-shifted log-sum-exps).  Per-element transcendentals
// (tanh, exp(a), the logit and its log-derivative) use the precise libdevice functions.
#include "common.cuh"
#include "tma.cuh"
#include <stdlib.h>

namespace flowk {

constexpr int K = 32;                 // marscf_main.py:41 (num_components=32)
constexpr int PLANES = 2 + 3 * K;     // a_raw, b, pi[K], mu[K], s[K]   (mixlogcdf_nn.py:72-73)
constexpr float kSFloor = -7.0f;      // mixlogcdf_nn.py:76
constexpr float kLogFloor = 1e-22f;   // log_dist.py:5-6
constexpr float kUnderflow = 1e-30f;

struct MixView {          // where one element's mixture parameters live: value k at ptr[k*stride]
  const float* pi;
  const float* mu;
  const float* s;
  size_t stride;
};

__device__ __forceinline__ float softplus_precise(float z) { return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z))); }

// Log-domain evaluation, the reference's own formulation (log_dist.py:9-40). Rare path.
template <bool CLAMP_S>
__device__ __noinline__ void mixture_logdomain(float x, MixView v, float* log_cdf, float* log_pdf) {
  float m = -INFINITY;
  for (int k = 0; k < K; ++k) m = fmaxf(m, v.pi[k * v.stride]);
  float W = 0.f;
  for (int k = 0; k < K; ++k) W += expf(v.pi[k * v.stride] - m);
  const float lse_pi = m + logf(W);
  float mc = -INFINITY, mp = -INFINITY;
  for (int k = 0; k < K; ++k) {
    float s = v.s[k * v.stride];
    if (CLAMP_S) s = fmaxf(s, kSFloor);
    float z = (x - v.mu[k * v.stride]) * expf(-s);
    float lp = v.pi[k * v.stride] - lse_pi;
    mc = fmaxf(mc, lp + (fminf(z, 0.f) - log1pf(expf(-fabsf(z)))));
    mp = fmaxf(mp, lp + (z - s - 2.f * softplus_precise(z)));
  }
  float sc = 0.f, sp = 0.f;
  for (int k = 0; k < K; ++k) {
    float s = v.s[k * v.stride];
    if (CLAMP_S) s = fmaxf(s, kSFloor);
    float z = (x - v.mu[k * v.stride]) * expf(-s);
    float lp = v.pi[k * v.stride] - lse_pi;
    sc += expf(lp + (fminf(z, 0.f) - log1pf(expf(-fabsf(z)))) - mc);
    sp += expf(lp + (z - s - 2.f * softplus_precise(z)) - mp);
  }
  *log_cdf = mc + logf(sc);
  *log_pdf = mp + logf(sp);
}

// Linear-domain evaluation (hot path).  Returns u = F(x) and log f(x).
template <bool CLAMP_S>
__device__ __forceinline__ void mixture_eval(float x, MixView v, float& u, float& log_pdf) {
  float pi[K];
#pragma unroll
  for (int k = 0; k < K; ++k) pi[k] = ld_stream(v.pi + k * v.stride);
  float m = pi[0];
#pragma unroll
  for (int k = 1; k < K; ++k) m = fmaxf(m, pi[k]);
  const float m2 = m * kLog2e;
  float W = 0.f, cdf = 0.f, pdf = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float mu = ld_stream(v.mu + k * v.stride);
    float s = ld_stream(v.s + k * v.stride);
    if (CLAMP_S) s = fmaxf(s, kSFloor);
    float w = ex2_fast(fmaf(pi[k], kLog2e, -m2));
    float inv = ex2_fast(-s * kLog2e);
    float z = (x - mu) * inv;
    float e = ex2_fast(-fabsf(z) * kLog2e);
    float r = rcp_fast(1.f + e);
    float er = e * r;
    W += w;
    cdf = fmaf(w, z >= 0.f ? r : er, cdf);
    pdf = fmaf(w * inv, er * r, pdf);
  }
  if (pdf > kUnderflow && cdf > kUnderflow) {
    u = cdf / W;
    log_pdf = logf(pdf / W);
  } else if (pdf != pdf || cdf != cdf) {
    u = cdf + pdf;            // NaN in, NaN out
    log_pdf = u;
  } else {
    float lc, lp;
    mixture_logdomain<CLAMP_S>(x, v, &lc, &lp);
    u = expf(lc);
    log_pdf = lp;
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 3)
mixlogcdf_fwd_kernel(const float* __restrict__ x, const float* __restrict__ raw, const float* __restrict__ rescale,
                     float* __restrict__ y, const float* __restrict__ ldj_in, float* __restrict__ ldj_out, LdjWs ws,
                     int C, int HW, int flip) {
  const int c = C >> 1, E = c * HW, b = blockIdx.y;
  const float* xc = x + (size_t)b * C * HW;        // transformed half (first), mixlogcdf_coupling.py:38
  const float* xid = xc + E;
  float* yb = y + (size_t)b * C * HW;
  float* y_out = flip ? yb + E : yb;               // TupleFlip fused into the store
  float* y_id = flip ? yb : yb + E;
  const float* rb = raw + (size_t)b * PLANES * E;
  float local = 0.f;
  for (int e = blockIdx.x * kThreads + threadIdx.x; e < E; e += gridDim.x * kThreads) {
    const float xv = ld_stream(xc + e);
    const float a_raw = ld_stream(rb + e);
    const float bb = ld_stream(rb + (size_t)E + e);
    y_id[e] = ld_stream(xid + e);
    MixView v{rb + (size_t)2 * E + e, rb + (size_t)(2 + K) * E + e, rb + (size_t)(2 + 2 * K) * E + e, (size_t)E};
    float u, log_pdf;
    mixture_eval<true>(xv, v, u, log_pdf);
    const float a = rescale[e / HW] * tanhf(a_raw);                       // mixlogcdf_nn.py:74
    const float logit = -logf(fmaxf(1.0f / u - 1.0f, kLogFloor));         // log_dist.py:81
    const float scale_ldj = -logf(fmaxf(u, kLogFloor)) - logf(fmaxf(1.0f - u, kLogFloor));   // log_dist.py:82
    y_out[e] = (logit + bb) * expf(a);                                    // mixlogcdf_coupling.py:51
    local += log_pdf + scale_ldj + a;                                     // mixlogcdf_coupling.py:53
  }
  if (ldj_out) finish_sample_ldj<kThreads>(local, ldj_in, ldj_out, 1.f, ws);
}

// ---------------------------------------------------------------------------------------------
// forward, TMA-staged: one 2-D bulk tensor load brings a [98 planes x 128 elements] tile (50 KB) of the raw
// parameter tensor into shared memory; threads then read their column conflict-free.  No registers are tied up
// by loads in flight, and 4 resident CTAs per SM keep ~200 KB per SM on the wire - that is what saturates HBM.
// Same arithmetic as mixlogcdf_fwd_kernel (shared mixture_eval body via the smem view).
// ---------------------------------------------------------------------------------------------
constexpr int kTmaTile = 128;

__device__ __forceinline__ void mixture_eval_smem(float x, const float* sm, int t, float& u, float& log_pdf) {
  const float* spi = sm + 2 * kTmaTile + t;
  const float* smu = sm + (2 + K) * kTmaTile + t;
  const float* ss = sm + (2 + 2 * K) * kTmaTile + t;
  float m = spi[0];
#pragma unroll
  for (int k = 1; k < K; ++k) m = fmaxf(m, spi[k * kTmaTile]);
  const float m2 = m * kLog2e;
  float W = 0.f, cdf = 0.f, pdf = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float mu = smu[k * kTmaTile];
    const float s = fmaxf(ss[k * kTmaTile], kSFloor);
    const float w = ex2_fast(fmaf(spi[k * kTmaTile], kLog2e, -m2));
    const float inv = ex2_fast(-s * kLog2e);
    const float z = (x - mu) * inv;
    const float e = ex2_fast(-fabsf(z) * kLog2e);
    const float r = rcp_fast(1.f + e);
    const float er = e * r;
    W += w;
    cdf = fmaf(w, z >= 0.f ? r : er, cdf);
    pdf = fmaf(w * inv, er * r, pdf);
  }
  if (pdf > kUnderflow && cdf > kUnderflow) {
    u = cdf / W;
    log_pdf = logf(pdf / W);
  } else if (pdf != pdf || cdf != cdf) {
    u = cdf + pdf;
    log_pdf = u;
  } else {
    float lc, lp;
    MixView v{spi, smu, ss, (size_t)kTmaTile};
    mixture_logdomain<true>(x, v, &lc, &lp);
    u = expf(lc);
    log_pdf = lp;
  }
}

__global__ void __launch_bounds__(kTmaTile, 4)
mixlogcdf_fwd_tma_kernel(const __grid_constant__ CUtensorMap raw_map, const float* __restrict__ x,
                         const float* __restrict__ rescale, float* __restrict__ y, const float* __restrict__ ldj_in,
                         float* __restrict__ ldj_out, LdjWs ws, int C, int HW, int flip, int* __restrict__ status) {
  extern __shared__ __align__(128) float tile[];                // [PLANES][kTmaTile]
  __shared__ __align__(8) uint64_t bar;
  const int c = C >> 1, E = c * HW, b = blockIdx.y;
  const int e0 = blockIdx.x * kTmaTile, t = threadIdx.x, e = e0 + t;
  if (t == 0) {
    tma::mbar_init(&bar, 1);
    tma::mbar_expect_tx(&bar, PLANES * kTmaTile * 4);
    tma::load_2d(tile, &raw_map, &bar, e0, b * PLANES);          // coords: (column = element, row = b*98 + plane)
  }
  const float* xc = x + (size_t)b * C * HW;
  float* yb = y + (size_t)b * C * HW;
  const float xv = ld_stream(xc + e);
  (flip ? yb : yb + E)[e] = ld_stream(xc + E + e);               // pass-through half (TupleFlip fused)
  const float rs = rescale[e / HW];
  __syncthreads();                                              // barrier init visible to the waiters
  const bool ok = tma::mbar_wait(&bar, 0);
  if (!ok && t == 0 && status) *status = 1;
  float u, log_pdf;
  mixture_eval_smem(xv, tile, t, u, log_pdf);
  const float a = rs * tanhf(tile[t]);
  const float logit = -logf(fmaxf(1.0f / u - 1.0f, kLogFloor));
  const float scale_ldj = -logf(fmaxf(u, kLogFloor)) - logf(fmaxf(1.0f - u, kLogFloor));
  (flip ? yb + E : yb)[e] = (logit + tile[kTmaTile + t]) * expf(a);
  if (ldj_out) finish_sample_ldj<kTmaTile>(log_pdf + scale_ldj + a, ldj_in, ldj_out, 1.f, ws);
}

// ---------------------------------------------------------------------------------------------
// inverse: per-element register-resident bisection (log_dist.py:43-72)
// ---------------------------------------------------------------------------------------------
struct MixRegs {           // one element's mixture, pre-digested for repeated CDF evaluation
  float w[K];              // softmax weights
  float mu[K];
  float nc[K];             // -log2(e) * e^{-s_k}
  float lb, ub;
};

template <bool CLAMP_S>
__device__ __forceinline__ void load_mixture(MixView v, MixRegs& r) {
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    r.w[k] = ld_stream(v.pi + k * v.stride);
    m = fmaxf(m, r.w[k]);
  }
  float W = 0.f, spread = 0.f, mu_lo = INFINITY, mu_hi = -INFINITY;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    r.w[k] = ex2_fast((r.w[k] - m) * kLog2e);
    W += r.w[k];
    float s = ld_stream(v.s + k * v.stride);
    if (CLAMP_S) s = fmaxf(s, kSFloor);
    spread += expf(s);                               // log_dist.py:60
    r.nc[k] = -kLog2e * ex2_fast(-s * kLog2e);
    r.mu[k] = ld_stream(v.mu + k * v.stride);
    mu_lo = fminf(mu_lo, r.mu[k]);
    mu_hi = fmaxf(mu_hi, r.mu[k]);
  }
  const float iw = 1.0f / W;
#pragma unroll
  for (int k = 0; k < K; ++k) r.w[k] *= iw;
  r.lb = mu_lo - 20.f * spread;                      // log_dist.py:61-62
  r.ub = mu_hi + 20.f * spread;
}

__device__ __forceinline__ float mixture_cdf_regs(float x, const MixRegs& r) {
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
  for (int k = 0; k < K; k += 2) {
    float e0 = ex2_fast((x - r.mu[k]) * r.nc[k]);
    float e1 = ex2_fast((x - r.mu[k + 1]) * r.nc[k + 1]);
    acc0 = fmaf(r.w[k], rcp_fast(1.f + e0), acc0);
    acc1 = fmaf(r.w[k + 1], rcp_fast(1.f + e1), acc1);
  }
  return acc0 + acc1;
}

// Bisection with the reference's update rule; every lane of the warp must call this.
// Termination is per element (|dx| <= eps or max_iters) - the reference stops on the global max |dx|,
// i.e. when the slowest element gets there; an element that is already at its fixed point stays on it,
// so the results agree to <= 2*eps.
__device__ __forceinline__ float bisect_inverse(float target, const MixRegs& r) {
  float x = 0.f, lo = r.lb, hi = r.ub;
  for (int it = 0; it < 100; ++it) {                                 // log_dist.py:44 max_iters
    const bool gt = mixture_cdf_regs(x, r) > target;
    const float nx = gt ? (x + lo) * 0.5f : (x + hi) * 0.5f;
    lo = gt ? lo : x;
    hi = gt ? x : hi;
    const bool done = !(fabsf(nx - x) > 1e-10f);                       // log_dist.py:43 eps
    x = nx;
    if (__all_sync(0xffffffffu, done)) break;
  }
  return x;
}

__global__ void __launch_bounds__(kThreads, 2)
mixlogcdf_inv_kernel(const float* __restrict__ x, const float* __restrict__ raw, const float* __restrict__ rescale,
                     float* __restrict__ y, const float* __restrict__ ldj_in, float* __restrict__ ldj_out, LdjWs ws,
                     int C, int HW, int flip) {
  const int c = C >> 1, E = c * HW, b = blockIdx.y;
  const float* xb = x + (size_t)b * C * HW;
  const float* vin = flip ? xb + E : xb;           // the transformed half of the (possibly flipped) input
  const float* xid = flip ? xb : xb + E;
  float* yb = y + (size_t)b * C * HW;
  const float* rb = raw + (size_t)b * PLANES * E;
  float local = 0.f;
  const int warp_base0 = blockIdx.x * kThreads + (threadIdx.x & ~31);
  for (int wb = warp_base0; wb < E; wb += gridDim.x * kThreads) {      // warp-uniform trip count
    const int e_raw = wb + (threadIdx.x & 31);
    const bool active = e_raw < E;
    const int e = active ? e_raw : E - 1;
    const float a = rescale[e / HW] * tanhf(ld_stream(rb + e));
    const float t = ld_stream(vin + e) * expf(-a) - ld_stream(rb + (size_t)E + e);     // mixlogcdf_coupling.py:42
    float u = 1.0f / (1.0f + expf(-t));                                                 // log_dist.py:78
    const float scale_ldj = fabsf(t) + 2.f * log1pf(expf(-fabsf(t)));                   // softplus(t)+softplus(-t)
    u = fminf(fmaxf(u, 1e-5f), (float)(1.0 - 1e-5));                                    // mixlogcdf_coupling.py:44
    MixView v{rb + (size_t)2 * E + e, rb + (size_t)(2 + K) * E + e, rb + (size_t)(2 + 2 * K) * E + e, (size_t)E};
    MixRegs regs;
    load_mixture<true>(v, regs);
    const float xs = bisect_inverse(u, regs);
    float u_chk, log_pdf;
    mixture_eval<true>(xs, v, u_chk, log_pdf);                                          // mixlogcdf_coupling.py:46
    if (active) {
      yb[e] = xs;
      yb[E + e] = ld_stream(xid + e);
      local += a + scale_ldj + log_pdf;
    }
  }
  if (ldj_out) finish_sample_ldj<kThreads>(local, ldj_in, ldj_out, -1.f, ws);
}

// ---------------------------------------------------------------------------------------------
// backward of the forward op (training).  Two passes over the components: sums, then gradients.
//   u = sum P_k sig_k, p = sum P_k i_k d_k, d_k = sig_k (1 - sig_k), i_k = e^{-s_k}
//   out = (log u - log(1-u) + b) e^a,   l = log p - log u - log(1-u) + a
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2)
mixlogcdf_bwd_kernel(const float* __restrict__ x, const float* __restrict__ raw, const float* __restrict__ rescale,
                     const float* __restrict__ gy, const float* __restrict__ gldj, float* __restrict__ gx,
                     float* __restrict__ graw, float* __restrict__ ga_tanh, int C, int HW, int flip) {
  const int c = C >> 1, E = c * HW, b = blockIdx.y;
  const size_t xo = (size_t)b * C * HW;
  const float* rb = raw + (size_t)b * PLANES * E;
  float* gb = graw + (size_t)b * PLANES * E;
  const float* gy_out = gy + xo + (flip ? E : 0);
  const float* gy_id = gy + xo + (flip ? 0 : E);
  const float gl = gldj ? gldj[b] : 0.f;
  for (int e = blockIdx.x * kThreads + threadIdx.x; e < E; e += gridDim.x * kThreads) {
    const float xv = x[xo + e];
    const float a_raw = rb[e], bb = rb[(size_t)E + e];
    const float* pi_p = rb + (size_t)2 * E + e;
    const float* mu_p = rb + (size_t)(2 + K) * E + e;
    const float* s_p = rb + (size_t)(2 + 2 * K) * E + e;
    float pi[K];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      pi[k] = pi_p[(size_t)k * E];
      m = fmaxf(m, pi[k]);
    }
    float W = 0.f, cdf = 0.f, pdf = 0.f, dpdx = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      pi[k] = __expf(pi[k] - m);                      // now holds the unnormalised weight
      float s = fmaxf(s_p[(size_t)k * E], kSFloor);
      float inv = __expf(-s);
      float z = (xv - mu_p[(size_t)k * E]) * inv;
      float sg = 1.0f / (1.0f + __expf(-z));
      float d = sg * (1.f - sg);
      W += pi[k];
      cdf = fmaf(pi[k], sg, cdf);
      pdf = fmaf(pi[k] * inv, d, pdf);
      dpdx = fmaf(pi[k] * inv * inv, d * (1.f - 2.f * sg), dpdx);
    }
    const float iW = 1.0f / W;
    const float u = fminf(fmaxf(cdf * iW, 1e-30f), 1.f - 6e-8f);
    const float p = fmaxf(pdf * iW, 1e-30f);
    dpdx *= iW;
    const float th = tanhf(a_raw), r = rescale[e / HW];
    const float a = r * th, ea = expf(a);
    const float logit = logf(u) - log1pf(-u);
    const float out = (logit + bb) * ea;
    const float go = gy_out[e];
    const float g_a = fmaf(go, out, gl);
    const float g_u = go * ea / (u * (1.f - u)) + gl * (1.f / (1.f - u) - 1.f / u);
    const float g_p = gl / p;
    gb[e] = g_a * r * (1.f - th * th);
    gb[(size_t)E + e] = go * ea;
    ga_tanh[(size_t)b * E + e] = g_a * th;
    gx[xo + e] = fmaf(g_u, p, g_p * dpdx);
    gx[xo + E + e] = gy_id[e];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float s_raw = s_p[(size_t)k * E];
      float s = fmaxf(s_raw, kSFloor);
      float inv = __expf(-s);
      float z = (xv - mu_p[(size_t)k * E]) * inv;
      float sg = 1.0f / (1.0f + __expf(-z));
      float d = sg * (1.f - sg);
      float P = pi[k] * iW;
      float f = inv * d;
      float t12 = 1.f - 2.f * sg;
      gb[(size_t)(2 + k) * E + e] = P * fmaf(g_u, sg - u, g_p * (f - p));
      gb[(size_t)(2 + K + k) * E + e] = -P * f * fmaf(g_p * inv, t12, g_u);
      float gs = -P * d * fmaf(g_u, z, g_p * inv * fmaf(t12, z, 1.f));
      gb[(size_t)(2 + 2 * K + k) * E + e] = (s_raw >= kSFloor) ? gs : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// log_dist.py functions on explicit [B,K,N] parameter tensors
// ---------------------------------------------------------------------------------------------
template <int WHAT>   // 0: log cdf, 1: log pdf
__global__ void __launch_bounds__(kThreads, 4)
mixture_log_kernel(const float* __restrict__ x, const float* __restrict__ pi, const float* __restrict__ mu,
                   const float* __restrict__ s, float* __restrict__ out, int N) {
  const int b = blockIdx.y;
  for (int e = blockIdx.x * kThreads + threadIdx.x; e < N; e += gridDim.x * kThreads) {
    const size_t po = (size_t)b * K * N + e;
    MixView v{pi + po, mu + po, s + po, (size_t)N};
    float lc, lp;
    // the standalone log functions must stay accurate deep in the tails, where a linear-domain
    // u would already have rounded: use the reference's log-domain formulation throughout.
    mixture_logdomain<false>(x[(size_t)b * N + e], v, &lc, &lp);
    out[(size_t)b * N + e] = WHAT == 0 ? lc : lp;
  }
}

__global__ void __launch_bounds__(kThreads, 3)
mixture_inv_cdf_kernel(const float* __restrict__ yv, const float* __restrict__ pi, const float* __restrict__ mu,
                       const float* __restrict__ s, float* __restrict__ out, int N) {
  const int b = blockIdx.y;
  const int warp_base0 = blockIdx.x * kThreads + (threadIdx.x & ~31);
  for (int wb = warp_base0; wb < N; wb += gridDim.x * kThreads) {
    const int e_raw = wb + (threadIdx.x & 31);
    const bool active = e_raw < N;
    const int e = active ? e_raw : N - 1;
    const size_t po = (size_t)b * K * N + e;
    MixView v{pi + po, mu + po, s + po, (size_t)N};
    MixRegs regs;
    load_mixture<false>(v, regs);
    const float xs = bisect_inverse(yv[(size_t)b * N + e], regs);
    if (active) out[(size_t)b * N + e] = xs;
  }
}

static int check_coupling_args(const void* x, const void* raw, const void* rescale, const void* y, int B, int C,
                               int HW, int Kc) {
  if (Kc != K) return FLOWK_ERR_ARG;
  if (B < 0 || C < 2 || (C & 1) || HW < 1 || B > 65535) return FLOWK_ERR_SHAPE;
  if ((long long)(C / 2) * HW * PLANES > 0x7fffffffLL) return FLOWK_ERR_SHAPE;
  if (B > 0 && (!x || !raw || !rescale || !y)) return FLOWK_ERR_ARG;
  return FLOWK_OK;
}

}  // namespace flowk

using namespace flowk;

extern "C" int flowk_mixlogcdf_fwd(const float* x, const float* raw, const float* rescale, float* y,
                                   const float* ldj_in, float* ldj_out, void* ws, int B, int C, int HW, int Kc,
                                   int flip, flowk_stream_t stream) {
  int st = check_coupling_args(x, raw, rescale, y, B, C, HW, Kc);
  if (st) return st;
  if (B == 0) return FLOWK_OK;
  if (ldj_out && !ws) return FLOWK_ERR_ARG;
  {
    // TMA-staged path: whole 128-element tiles, rows of the [B*98, E] view 16-byte aligned, <= 64 tiles per sample
    const long long E = (long long)(C / 2) * HW;
    static int use_tma = -1;
    if (use_tma < 0) { const char* env = getenv("FLOWK_MIXLOGCDF_TMA"); use_tma = (env && env[0] == '0') ? 0 : 1; }
    if (use_tma && E % kTmaTile == 0 && E / kTmaTile <= kMaxParts && aligned16(raw) && tma::encode_fn()) {
      alignas(64) CUtensorMap map;
      if (tma::make_map_2d(&map, raw, (long long)B * PLANES, E, PLANES, kTmaTile)) {
        const int smem = PLANES * kTmaTile * 4;
        static bool attr_set = false;
        if (!attr_set) {
          FLOWK_CUDA_OK(cudaFuncSetAttribute(mixlogcdf_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
          attr_set = true;
        }
        dim3 grid((unsigned)(E / kTmaTile), B);
        mixlogcdf_fwd_tma_kernel<<<grid, kTmaTile, smem, stream>>>(map, x, rescale, y, ldj_in, ldj_out, carve_ws(ws, B),
                                                                  C, HW, flip, nullptr);
        return launch_status();
      }
    }
  }
  dim3 grid(parts_for((long long)(C / 2) * HW, kThreads), B);
  mixlogcdf_fwd_kernel<<<grid, kThreads, 0, stream>>>(x, raw, rescale, y, ldj_in, ldj_out, carve_ws(ws, B), C, HW,
                                                      flip);
  return launch_status();
}

extern "C" int flowk_mixlogcdf_inv(const float* x, const float* raw, const float* rescale, float* y,
                                   const float* ldj_in, float* ldj_out, void* ws, int B, int C, int HW, int Kc,
                                   int flip, flowk_stream_t stream) {
  int st = check_coupling_args(x, raw, rescale, y, B, C, HW, Kc);
  if (st) return st;
  if (B == 0) return FLOWK_OK;
  if (ldj_out && !ws) return FLOWK_ERR_ARG;
  dim3 grid(parts_for((long long)(C / 2) * HW, kThreads), B);
  mixlogcdf_inv_kernel<<<grid, kThreads, 0, stream>>>(x, raw, rescale, y, ldj_in, ldj_out, carve_ws(ws, B), C, HW,
                                                      flip);
  return launch_status();
}

extern "C" int flowk_mixlogcdf_bwd(const float* x, const float* raw, const float* rescale, const float* gy,
                                   const float* gldj, float* gx, float* graw, float* ga_tanh, int B, int C, int HW,
                                   int Kc, int flip, flowk_stream_t stream) {
  int st = check_coupling_args(x, raw, rescale, gx, B, C, HW, Kc);
  if (st) return st;
  if (B == 0) return FLOWK_OK;
  if (!gy || !graw || !ga_tanh) return FLOWK_ERR_ARG;
  dim3 grid(parts_for((long long)(C / 2) * HW, kThreads), B);
  mixlogcdf_bwd_kernel<<<grid, kThreads, 0, stream>>>(x, raw, rescale, gy, gldj, gx, graw, ga_tanh, C, HW, flip);
  return launch_status();
}

static int check_mixture_args(const void* a, const void* pi, const void* mu, const void* s, const void* out, int B,
                              int Kc, int N) {
  if (Kc != K) return FLOWK_ERR_ARG;
  if (B < 0 || N < 1 || B > 65535 || (long long)N * K > 0x7fffffffLL) return FLOWK_ERR_SHAPE;
  if (B > 0 && (!a || !pi || !mu || !s || !out)) return FLOWK_ERR_ARG;
  return FLOWK_OK;
}

extern "C" int flowk_mixture_log_cdf(const float* x, const float* pi, const float* mu, const float* s, float* out,
                                     int B, int Kc, int N, flowk_stream_t stream) {
  int st = check_mixture_args(x, pi, mu, s, out, B, Kc, N);
  if (st || B == 0) return st;
  dim3 grid((N + kThreads - 1) / kThreads > 1024 ? 1024 : (N + kThreads - 1) / kThreads, B);
  mixture_log_kernel<0><<<grid, kThreads, 0, stream>>>(x, pi, mu, s, out, N);
  return launch_status();
}

extern "C" int flowk_mixture_log_pdf(const float* x, const float* pi, const float* mu, const float* s, float* out,
                                     int B, int Kc, int N, flowk_stream_t stream) {
  int st = check_mixture_args(x, pi, mu, s, out, B, Kc, N);
  if (st || B == 0) return st;
  dim3 grid((N + kThreads - 1) / kThreads > 1024 ? 1024 : (N + kThreads - 1) / kThreads, B);
  mixture_log_kernel<1><<<grid, kThreads, 0, stream>>>(x, pi, mu, s, out, N);
  return launch_status();
}

extern "C" int flowk_mixture_inv_cdf(const float* y, const float* pi, const float* mu, const float* s, float* out,
                                     int B, int Kc, int N, flowk_stream_t stream) {
  int st = check_mixture_args(y, pi, mu, s, out, B, Kc, N);
  if (st || B == 0) return st;
  dim3 grid((N + kThreads - 1) / kThreads > 1024 ? 1024 : (N + kThreads - 1) / kThreads, B);
  mixture_inv_cdf_kernel<<<grid, kThreads, 0, stream>>>(y, pi, mu, s, out, N);
  return launch_status();
}
