// Training-time augmentation of the uint8 face / parsing-map batches on the device, bit-exact with the Pillow calls of
// the reference's loader (helen_loader.py:75-104): Image.rotate(angle) - NEAREST, no expand, zero fill, i.e. Pillow's
// 16.16 fixed-point affine gather - followed by ImageEnhance.Contrast(...).enhance(f) once per factor (the reference
// builds Contrast for its "contrast", "brightness" and "sharpness" draws alike, :84-91): the image is blended with its
// rounded mean luma in float32, truncated (0 <= f <= 1) or clipped (extrapolation) exactly as ImagingBlend does.
//
// One CTA per image: the gather writes the rotated image, then every enhancement is a CTA-wide exact integer sum of the
// luma followed by an in-place blend (each thread rewrites the bytes it read).  HBM-bound byte work: h*w*c bytes in,
// the same out, the enhancement passes run out of L2.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "crfr.h"

namespace {

constexpr int kT = 1024;

__device__ __forceinline__ long long block_sum(long long v, long long* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();   // previous use of sm is over
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  long long t = 0;
  for (int i = 0; i < kT / 32; ++i) t += sm[i];   // same order in every thread: one exact integer
  return t;
}

__global__ void __launch_bounds__(kT)
augment_kernel(const uint8_t* __restrict__ src, int h, int w, int c, const int* __restrict__ coef,
               const float* __restrict__ factors, int nfac, uint8_t* __restrict__ dst) {
  __shared__ long long sm[kT / 32];
  const long long n = blockIdx.x;
  const int npix = h * w;
  const uint8_t* s = src + n * npix * c;
  uint8_t* d = dst + n * npix * c;
  // Pillow affine_fixed: xx = a2 + a1 * y + a0 * x, source = (xx >> 16, yy >> 16), outside -> 0
  const int a0 = coef[n * 6 + 0], a1 = coef[n * 6 + 1], a2 = coef[n * 6 + 2];
  const int a3 = coef[n * 6 + 3], a4 = coef[n * 6 + 4], a5 = coef[n * 6 + 5];
  for (int p = threadIdx.x; p < npix; p += kT) {
    const int y = p / w, x = p - y * w;
    const int xin = (a2 + a1 * y + a0 * x) >> 16, yin = (a5 + a4 * y + a3 * x) >> 16;
    const bool ok = xin >= 0 && xin < w && yin >= 0 && yin < h;
    for (int k = 0; k < c; ++k) d[p * c + k] = ok ? s[(yin * w + xin) * c + k] : (uint8_t)0;
  }
  for (int fi = 0; fi < nfac; ++fi) {
    const float f = factors[n * nfac + fi];
    __syncthreads();   // the bytes written above / by the previous enhancement are visible to the whole CTA
    long long part = 0;
    for (int p = threadIdx.x; p < npix; p += kT) {
      if (c == 3) {
        const int r = d[p * 3], g = d[p * 3 + 1], b = d[p * 3 + 2];
        part += (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16;    // Pillow's RGB -> L
      } else {
        part += d[p];
      }
    }
    const long long total = block_sum(part, sm);
    const int mean = (int)((double)total / (double)npix + 0.5);       // int(ImageStat.mean + 0.5)
    if (f == 1.0f) continue;
    const float in1 = (float)mean;
    const bool interp = f >= 0.0f && f <= 1.0f;
    for (int p = threadIdx.x; p < npix; p += kT)
      for (int k = 0; k < c; ++k) {
        uint8_t o;
        if (f == 0.0f) {
          o = (uint8_t)mean;
        } else {
          // float32 multiply then add, each rounded (no FMA): (int)in1 + alpha * ((int)in2 - (int)in1)
          const float t = __fadd_rn(in1, __fmul_rn(f, (float)((int)d[p * c + k] - mean)));
          if (interp) o = (uint8_t)(int)t;
          else o = t <= 0.0f ? (uint8_t)0 : (t >= 255.0f ? (uint8_t)255 : (uint8_t)(int)t);
        }
        d[p * c + k] = o;
      }
  }
}

// per-image window copy: dst[n][y][x][:] = src[n][off_y[n] + y][off_x[n] + x][:]; one thread per output byte
__global__ void crop_kernel(const uint8_t* __restrict__ src, int h, int w, int c, const int* __restrict__ off, int oh, int ow,
                            uint8_t* __restrict__ dst, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int row = ow * c;
  const int xb = (int)(i % row);
  const long long q = i / row;
  const int y = (int)(q % oh);
  const long long n = q / oh;
  dst[i] = src[((n * h + off[2 * n] + y) * w + off[2 * n + 1]) * c + xb];
}

// Python's round(x, 15): correctly rounded 15-decimal string, back to the nearest double
double round15(double x) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%.15f", x);
  return strtod(buf, nullptr);
}
int fix16(double v) {
  const double t = v * 65536.0 + 0.5;
  return t >= 0.0 ? (int)t : (int)floor(t);
}

}  // namespace

extern "C" int crfr_rotate_coeffs(int h, int w, double angle_deg, int32_t* host_coef6) {
  CRFR_CHECK_ARG(h > 0 && w > 0 && host_coef6, "rotate_coeffs: bad argument");
  CRFR_CHECK_ARG(h < 16384 && w < 16384, "rotate_coeffs: image too large for the 16.16 fixed-point path");
  double angle = fmod(angle_deg, 360.0);
  if (angle < 0.0) angle += 360.0;                       // Python's % on floats
  const double a = -(angle * (M_PI / 180.0));            // -math.radians(angle)
  double m[6] = {round15(cos(a)), round15(sin(a)), 0.0, round15(-sin(a)), round15(cos(a)), 0.0};
  const double cx = w / 2.0, cy = h / 2.0;
  m[2] = m[0] * (-cx) + m[1] * (-cy) + m[2];
  m[5] = m[3] * (-cx) + m[4] * (-cy) + m[5];
  m[2] += cx;
  m[5] += cy;
  host_coef6[0] = fix16(m[0]);
  host_coef6[1] = fix16(m[1]);
  host_coef6[2] = fix16(m[2] + m[0] * 0.5 + m[1] * 0.5);
  host_coef6[3] = fix16(m[3]);
  host_coef6[4] = fix16(m[4]);
  host_coef6[5] = fix16(m[5] + m[3] * 0.5 + m[4] * 0.5);
  return CRFR_OK;
}

extern "C" int crfr_augment_u8(const uint8_t* src, int n, int h, int w, int c, const int32_t* coef, const float* factors,
                               int nfac, uint8_t* dst, void* stream) {
  CRFR_CHECK_ARG(src && dst && coef && n > 0 && h > 0 && w > 0, "augment_u8: bad argument");
  CRFR_CHECK_ARG(c == 1 || c == 3, "augment_u8: %d channels (1 or 3 supported)", c);
  CRFR_CHECK_ARG(nfac == 0 || factors, "augment_u8: %d factors but no factor array", nfac);
  CRFR_CHECK_ARG(src != dst, "augment_u8: the rotation is a gather, it cannot run in place");
  CRFR_CHECK_ARG((long long)h * w * 3 < (1ll << 31) / 2, "augment_u8: image too large");
  augment_kernel<<<n, kT, 0, (cudaStream_t)stream>>>(src, h, w, c, coef, factors, nfac, dst);
  CRFR_COUNT_LAUNCH();
  CRFR_LAUNCH_CHECK();
  return CRFR_OK;
}

extern "C" int crfr_crop_u8(const uint8_t* src, int n, int h, int w, int c, const int32_t* offsets_yx, int oh, int ow,
                            uint8_t* dst, void* stream) {
  CRFR_CHECK_ARG(src && dst && offsets_yx && n > 0 && c > 0 && oh > 0 && ow > 0 && oh <= h && ow <= w,
                 "crop_u8: bad argument (window %dx%d of %dx%d)", oh, ow, h, w);
  const long long total = (long long)n * oh * ow * c;
  crop_kernel<<<crfr_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(src, h, w, c, offsets_yx, oh, ow, dst, total);
  CRFR_COUNT_LAUNCH();
  CRFR_LAUNCH_CHECK();
  return CRFR_OK;
}
