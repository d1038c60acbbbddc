// PIL-exact bicubic resampling of uint8 faces on the GPU (integer fixed-point, bit-exact with Pillow's 8-bit path).
// ref: Image.resize((W,H), Image.BICUBIC) at bicubic_interpolation.py:188 and SUPER_RESOLUTION/FHN_loader.py:66; the
// arithmetic is Pillow's ImagingResample (third party): horizontal pass then vertical pass with a uint8 intermediate,
// Keys cubic a=-0.5, window clipped at the borders and renormalised, weights rounded to 2^22 fixed point.
// HBM-bound and tiny (768 B in, 48 KiB out per 16->128 face): one thread per output byte, coalesced along (x, c).
#include <math.h>

#include "common.cuh"
#include "crfr.h"

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

double keys(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

int table_ksize(int in_size, int out_size) {
  double scale = (double)in_size / out_size;
  double filterscale = scale < 1.0 ? 1.0 : scale;
  double support = 2.0 * filterscale;
  return (int)ceil(support) * 2 + 1;
}

__device__ __forceinline__ uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// resample the `axis_len` axis: src [outer][axis_len][inner] -> dst [outer][out_len][inner]
__global__ void resample_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int axis_len, int out_len,
                                int inner, const int32_t* __restrict__ tab, int tab_stride, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int in = (int)(i % inner);
  long long q = i / inner;
  int o = (int)(q % out_len);
  long long outer = q / out_len;
  const int32_t* t = tab + (long long)o * tab_stride;
  const int xmin = t[0], cnt = t[1];
  int acc = 1 << (kPrecisionBits - 1);
  const uint8_t* s = src + (outer * axis_len + xmin) * inner + in;
  for (int k = 0; k < cnt; ++k) acc += (int)s[(long long)k * inner] * t[2 + k];
  dst[i] = clip8(acc >> kPrecisionBits);
}

// u8 NHWC -> fp32 NCHW, (v/255 - 0.5)/0.5
__global__ void normalise_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int hw, int c,
                                 long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over n*c*hw (output order)
  if (i >= total) return;
  int p = (int)(i % hw);
  long long q = i / hw;
  int ch = (int)(q % c);
  long long n = q / c;
  float v = (float)src[(n * hw + p) * c + ch] / 255.0f;
  dst[i] = (v - 0.5f) / 0.5f;
}

}  // namespace

extern "C" int crfr_bicubic_table_size(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return 0;
  return out_size * (2 + table_ksize(in_size, out_size));
}

extern "C" int crfr_bicubic_tables(int in_size, int out_size, int32_t* host_tab) {
  CRFR_CHECK_ARG(in_size > 0 && out_size > 0 && host_tab, "bicubic_tables: bad argument");
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const int ksize = table_ksize(in_size, out_size);
  const double ss = 1.0 / filterscale;
  double k[64];
  CRFR_CHECK_ARG(ksize <= 64, "bicubic_tables: kernel too wide");
  for (int xx = 0; xx < out_size; ++xx) {
    int32_t* t = host_tab + (long long)xx * (2 + ksize);
    double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    int n = xmax - xmin;
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
      k[x] = keys((x + xmin - center + 0.5) * ss);
      ww += k[x];
    }
    for (int x = 0; x < n; ++x)
      if (ww != 0.0) k[x] /= ww;
    t[0] = xmin;
    t[1] = n;
    for (int x = 0; x < ksize; ++x) {
      if (x < n) {
        double v = k[x] * (double)(1 << kPrecisionBits);
        t[2 + x] = k[x] < 0 ? (int)(-0.5 + v) : (int)(0.5 + v);
      } else {
        t[2 + x] = 0;
      }
    }
  }
  return CRFR_OK;
}

extern "C" int crfr_bicubic_u8(const uint8_t* src, int n, int ih, int iw, int c, const int32_t* tab_h,
                               const int32_t* tab_w, int oh, int ow, uint8_t* tmp, uint8_t* dst, float* dst_f32,
                               void* stream) {
  CRFR_CHECK_ARG(src && tab_h && tab_w && tmp && dst && n > 0 && ih > 0 && iw > 0 && c > 0 && oh > 0 && ow > 0,
                 "bicubic_u8: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  // horizontal: [n*ih][iw][c] -> [n*ih][ow][c]
  long long t1 = (long long)n * ih * ow * c;
  resample_kernel<<<crfr_cdiv(t1, 256), 256, 0, st>>>(src, tmp, iw, ow, c, tab_w, 2 + table_ksize(iw, ow), t1);
  CRFR_COUNT_LAUNCH();
  // vertical: [n][ih][ow*c] -> [n][oh][ow*c]
  long long t2 = (long long)n * oh * ow * c;
  resample_kernel<<<crfr_cdiv(t2, 256), 256, 0, st>>>(tmp, dst, ih, oh, ow * c, tab_h, 2 + table_ksize(ih, oh), t2);
  CRFR_COUNT_LAUNCH();
  if (dst_f32) {
    normalise_kernel<<<crfr_cdiv(t2, 256), 256, 0, st>>>(dst, dst_f32, oh * ow, c, t2);
    CRFR_COUNT_LAUNCH();
  }
  CRFR_LAUNCH_CHECK();
  return CRFR_OK;
}
