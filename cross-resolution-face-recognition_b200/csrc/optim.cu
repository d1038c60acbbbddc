// Fused RMSprop step over a flat fp32 parameter arena (HBM-bound: 3 reads + 2 writes per element, float4).
// ref: torch.optim.RMSprop(lr, alpha=0.99, eps=1e-8, weight_decay=1e-5, momentum=0, centered=False) as built at
//      FSR_main.py:185 and distill_main.py:222-225:  g += wd*p; sq = alpha*sq + (1-alpha)*g*g; p -= lr*g/(sqrt(sq)+eps)
#include "common.cuh"
#include "crfr.h"

namespace {
__global__ void rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ sq,
                               long long n, float lr, float alpha, float eps, float wd, float gscale) {
  long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n4 = n >> 2;
  for (long long v = i; v < n4; v += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[v];
    float4 gv = reinterpret_cast<const float4*>(g)[v];
    float4 sv = reinterpret_cast<float4*>(sq)[v];
    float* pp = &pv.x; float* gg = &gv.x; float* ss = &sv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gr = fmaf(wd, pp[j], gg[j] * gscale);
      ss[j] = alpha * ss[j] + (1.f - alpha) * gr * gr;
      pp[j] -= lr * gr / (sqrtf(ss[j]) + eps);
    }
    reinterpret_cast<float4*>(p)[v] = pv;
    reinterpret_cast<float4*>(sq)[v] = sv;
  }
  for (long long e = (n4 << 2) + i; e < n; e += stride) {
    float gr = fmaf(wd, p[e], g[e] * gscale);
    float s = alpha * sq[e] + (1.f - alpha) * gr * gr;
    sq[e] = s;
    p[e] -= lr * gr / (sqrtf(s) + eps);
  }
}
}  // namespace

extern "C" int crfr_rmsprop_step(float* p, const float* g, float* sq, long long n, float lr, float alpha, float eps,
                                 float weight_decay, float gscale, void* stream) {
  CRFR_CHECK_ARG(p && g && sq && n > 0, "rmsprop_step: bad argument");
  CRFR_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)sq) & 15) == 0, "rmsprop_step: pointers must be 16B aligned");
  long long want = (n / 4 + 255) / 256;
  int blocks = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  rmsprop_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, sq, n, lr, alpha, eps, weight_decay, gscale);
  CRFR_COUNT_LAUNCH();
  CRFR_LAUNCH_CHECK();
  return CRFR_OK;
}
