// Adamax update (torch.optim.Adamax, the reference's optimizer: marscf_main.py:302) for ALL parameter tensors of a model in
// one launch and one pass over HBM: 4 reads + 3 writes per element (28 B) instead of ~10 multi-tensor passes.
//   m <- m + (1 - beta1) (g - m)          (torch's lerp form)
//   u <- max(beta2 u, |g| + eps)
//   p <- p - clr * (m / u),   clr = lr / (1 - beta1^t), read from device memory so that a captured graph replays with
//                             the current learning rate / step (the host refreshes the scalar before each replay)
// `chunks` is a DEVICE array: every tensor is cut into pieces of at most FLOWK_ADAMAX_CHUNK elements, one CTA each.
#include "common.cuh"

namespace flowk {

__global__ void __launch_bounds__(256) adamax_kernel(const flowk_adamax_chunk* __restrict__ chunks,
                                                     const float* __restrict__ clr_ptr, float beta1, float beta2, float eps) {
  const flowk_adamax_chunk c = chunks[blockIdx.x];
  const float clr = *clr_ptr, w = 1.f - beta1;
  auto update = [&](float& p, float g, float& m, float& u) {
    m = m + w * (g - m);
    u = fmaxf(beta2 * u, fabsf(g) + eps);
    p = p + (-clr) * (m / u);
  };
  const bool vec = (((uintptr_t)c.p | (uintptr_t)c.g | (uintptr_t)c.m | (uintptr_t)c.u) & 15u) == 0;
  long long done = 0;
  if (vec) {
    const long long n4 = c.n >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
      float4 p = reinterpret_cast<float4*>(c.p)[i], m = reinterpret_cast<float4*>(c.m)[i], u = reinterpret_cast<float4*>(c.u)[i];
      const float4 g = reinterpret_cast<const float4*>(c.g)[i];
      update(p.x, g.x, m.x, u.x);
      update(p.y, g.y, m.y, u.y);
      update(p.z, g.z, m.z, u.z);
      update(p.w, g.w, m.w, u.w);
      reinterpret_cast<float4*>(c.p)[i] = p;
      reinterpret_cast<float4*>(c.m)[i] = m;
      reinterpret_cast<float4*>(c.u)[i] = u;
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < c.n; i += 256) {
    float p = c.p[i], m = c.m[i], u = c.u[i];
    update(p, c.g[i], m, u);
    c.p[i] = p;
    c.m[i] = m;
    c.u[i] = u;
  }
}

}  // namespace flowk

extern "C" int flowk_adamax_step(const flowk_adamax_chunk* chunks_device, int nchunks, const float* clr_device, float beta1,
                                 float beta2, float eps, flowk_stream_t stream) {
  if (nchunks < 0) return FLOWK_ERR_SHAPE;
  if (nchunks == 0) return FLOWK_OK;
  if (!chunks_device || !clr_device) return FLOWK_ERR_ARG;
  flowk::adamax_kernel<<<nchunks, 256, 0, stream>>>(chunks_device, clr_device, beta1, beta2, eps);
  return flowk::launch_status();
}
