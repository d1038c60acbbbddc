// Landmark heat-map target of the FSRNet prior loss, generated on the device (input pipeline).
//
// ref: helen_loader.py:118-143 - hm = zeros(float32); for every landmark: hm += exp(-((x-x0)^2 + (y-y0)^2) / (2 s^2))
// with the Gaussian evaluated in float64 (numpy arange(float)) and the running sum rounded to float32 after every
// landmark (in-place += on a float32 array).  One thread per pixel reproduces exactly that order of operations.
#include "common.cuh"
#include "crfr.h"

namespace {

__global__ void landmark_heatmap_kernel(const float* __restrict__ lm, int k, double inv2s2, int h, int w,
                                        float* __restrict__ hm, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % w);
  const long long q = i / w;
  const int y = (int)(q % h);
  const long long n = q / h;
  const float* p = lm + n * k * 2;
  float acc = 0.f;
  for (int j = 0; j < k; ++j) {
    const double dx = (double)x - (double)p[2 * j], dy = (double)y - (double)p[2 * j + 1];
    acc = (float)((double)acc + exp(-(dx * dx + dy * dy) * inv2s2));
  }
  hm[i] = acc;
}

}  // namespace

extern "C" int crfr_landmark_heatmap(const float* landmarks, int n, int k, float sigma, int h, int w, float* hm,
                                     void* stream) {
  CRFR_CHECK_ARG(landmarks && hm && n > 0 && k > 0 && h > 0 && w > 0 && sigma > 0.f, "landmark_heatmap: bad argument");
  const long long total = (long long)n * h * w;
  landmark_heatmap_kernel<<<crfr_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
      landmarks, k, 1.0 / (2.0 * (double)sigma * (double)sigma), h, w, hm, total);
  CRFR_COUNT_LAUNCH();
  CRFR_LAUNCH_CHECK();
  return CRFR_OK;
}
