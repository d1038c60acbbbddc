// Hourglass resampling (HBM-bound, 16-byte vectors over NHWC channel groups).
// ref: F.max_pool2d(x, 2, stride=2) model/FSRnet.py:202; F.interpolate(scale_factor=2) (nearest) + add :210-211.
#include "common.cuh"
#include "crfr.h"

namespace {

__global__ void maxpool2_fwd_kernel(const bf16* __restrict__ x, int x_ld, bf16* __restrict__ out, int out_ld, int h,
                                    int w, int groups, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int cg = (int)(i % groups);
  long long q = i / groups;  // output pixel (n, oy, ox)
  int ow = w >> 1, oh = h >> 1;
  int ox = (int)(q % ow);
  long long r = q / ow;
  int oy = (int)(r % oh);
  long long n = r / oh;
  const bf16* s = x + ((n * h + 2 * oy) * w + 2 * ox) * x_ld + cg * 8;
  float a[8], b[8], c[8], d[8];
  unpack8(*reinterpret_cast<const bf16x8*>(s), a);
  unpack8(*reinterpret_cast<const bf16x8*>(s + x_ld), b);
  unpack8(*reinterpret_cast<const bf16x8*>(s + (long long)w * x_ld), c);
  unpack8(*reinterpret_cast<const bf16x8*>(s + (long long)(w + 1) * x_ld), d);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = fmaxf(fmaxf(a[j], b[j]), fmaxf(c[j], d[j]));
  *reinterpret_cast<bf16x8*>(out + q * out_ld + cg * 8) = pack8(a);
}

// dx[window] = dout at the first maximum of the window (scan order), 0 elsewhere
__global__ void maxpool2_bwd_kernel(const bf16* __restrict__ x, int x_ld, const bf16* __restrict__ dout, int dout_ld,
                                    bf16* __restrict__ dx, int dx_ld, int h, int w, int groups, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int cg = (int)(i % groups);
  long long q = i / groups;
  int ow = w >> 1, oh = h >> 1;
  int ox = (int)(q % ow);
  long long r = q / ow;
  int oy = (int)(r % oh);
  long long n = r / oh;
  long long pix = (n * h + 2 * oy) * w + 2 * ox;
  const bf16* s = x + pix * x_ld + cg * 8;
  float v[4][8], g[8], o[4][8];
  unpack8(*reinterpret_cast<const bf16x8*>(s), v[0]);
  unpack8(*reinterpret_cast<const bf16x8*>(s + x_ld), v[1]);
  unpack8(*reinterpret_cast<const bf16x8*>(s + (long long)w * x_ld), v[2]);
  unpack8(*reinterpret_cast<const bf16x8*>(s + (long long)(w + 1) * x_ld), v[3]);
  unpack8(*reinterpret_cast<const bf16x8*>(dout + q * dout_ld + cg * 8), g);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int best = 0;
    float m = v[0][j];
#pragma unroll
    for (int k = 1; k < 4; ++k)
      if (v[k][j] > m) {
        m = v[k][j];
        best = k;
      }
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k][j] = (k == best) ? g[j] : 0.f;
  }
  bf16* d = dx + pix * dx_ld + cg * 8;
  *reinterpret_cast<bf16x8*>(d) = pack8(o[0]);
  *reinterpret_cast<bf16x8*>(d + dx_ld) = pack8(o[1]);
  *reinterpret_cast<bf16x8*>(d + (long long)w * dx_ld) = pack8(o[2]);
  *reinterpret_cast<bf16x8*>(d + (long long)(w + 1) * dx_ld) = pack8(o[3]);
}

// out[2h x 2w] = up + nearest2x(low)
__global__ void upadd_fwd_kernel(const bf16* __restrict__ up, int up_ld, const bf16* __restrict__ low, int low_ld,
                                 bf16* __restrict__ out, int out_ld, int h, int w, int groups, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int cg = (int)(i % groups);
  long long q = i / groups;  // output pixel
  int W = 2 * w, H = 2 * h;
  int X = (int)(q % W);
  long long r = q / W;
  int Y = (int)(r % H);
  long long n = r / H;
  float a[8], b[8];
  unpack8(*reinterpret_cast<const bf16x8*>(up + q * up_ld + cg * 8), a);
  unpack8(*reinterpret_cast<const bf16x8*>(low + ((n * h + (Y >> 1)) * w + (X >> 1)) * low_ld + cg * 8), b);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] += b[j];
  *reinterpret_cast<bf16x8*>(out + q * out_ld + cg * 8) = pack8(a);
}

// dlow[h x w] = 2x2 sum of dout[2h x 2w]
__global__ void up_bwd_kernel(const bf16* __restrict__ dout, int dout_ld, bf16* __restrict__ dlow, int dlow_ld, int h,
                              int w, int groups, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int cg = (int)(i % groups);
  long long q = i / groups;  // low pixel
  int x = (int)(q % w);
  long long r = q / w;
  int y = (int)(r % h);
  long long n = r / h;
  int W = 2 * w;
  const bf16* s = dout + ((n * 2 * h + 2 * y) * W + 2 * x) * dout_ld + cg * 8;
  float a[8], b[8], c[8], d[8];
  unpack8(*reinterpret_cast<const bf16x8*>(s), a);
  unpack8(*reinterpret_cast<const bf16x8*>(s + dout_ld), b);
  unpack8(*reinterpret_cast<const bf16x8*>(s + (long long)W * dout_ld), c);
  unpack8(*reinterpret_cast<const bf16x8*>(s + (long long)(W + 1) * dout_ld), d);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = (a[j] + b[j]) + (c[j] + d[j]);
  *reinterpret_cast<bf16x8*>(dlow + q * dlow_ld + cg * 8) = pack8(a);
}

__global__ void add_n_kernel(const bf16* __restrict__ a, int a_ld, const bf16* __restrict__ b, int b_ld,
                             const bf16* __restrict__ c3, int c_ld, bf16* __restrict__ out, int out_ld, int groups,
                             long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int cg = (int)(i % groups);
  long long q = i / groups;
  float x[8], y[8];
  unpack8(*reinterpret_cast<const bf16x8*>(a + q * a_ld + cg * 8), x);
  unpack8(*reinterpret_cast<const bf16x8*>(b + q * b_ld + cg * 8), y);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] += y[j];
  if (c3) {
    unpack8(*reinterpret_cast<const bf16x8*>(c3 + q * c_ld + cg * 8), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
  }
  *reinterpret_cast<bf16x8*>(out + q * out_ld + cg * 8) = pack8(x);
}

// 4-channel (8-byte) variant for the 3-channel image gradients
__global__ void add_n4_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, const bf16* __restrict__ c3,
                              bf16* __restrict__ out, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  bf16x4 x = *reinterpret_cast<const bf16x4*>(a + i * 4), y = *reinterpret_cast<const bf16x4*>(b + i * 4);
  float2 x0 = __bfloat1622float2(x.v[0]), x1 = __bfloat1622float2(x.v[1]);
  float2 y0 = __bfloat1622float2(y.v[0]), y1 = __bfloat1622float2(y.v[1]);
  x0.x += y0.x; x0.y += y0.y; x1.x += y1.x; x1.y += y1.y;
  if (c3) {
    bf16x4 z = *reinterpret_cast<const bf16x4*>(c3 + i * 4);
    float2 z0 = __bfloat1622float2(z.v[0]), z1 = __bfloat1622float2(z.v[1]);
    x0.x += z0.x; x0.y += z0.y; x1.x += z1.x; x1.y += z1.y;
  }
  bf16x4 o;
  o.v[0] = __floats2bfloat162_rn(x0.x, x0.y);
  o.v[1] = __floats2bfloat162_rn(x1.x, x1.y);
  *reinterpret_cast<bf16x4*>(out + i * 4) = o;
}

inline bool vec_ok(int c, int ld) { return c > 0 && (c & 7) == 0 && ld >= c && (ld & 7) == 0; }

}  // namespace

#define LAUNCH_1D(kernel, total, st, ...)                                             \
  do {                                                                                \
    kernel<<<crfr_cdiv((total), 256), 256, 0, (st)>>>(__VA_ARGS__);                   \
    CRFR_COUNT_LAUNCH();                                                              \
    CRFR_LAUNCH_CHECK();                                                              \
  } while (0)

extern "C" int crfr_maxpool2_fwd(const void* x, int x_ld, void* out, int out_ld, int n, int h, int w, int c,
                                 void* stream) {
  CRFR_CHECK_ARG(x && out && n > 0 && h > 0 && w > 0 && !(h & 1) && !(w & 1), "maxpool2_fwd: bad argument");
  CRFR_CHECK_ARG(vec_ok(c, x_ld) && vec_ok(c, out_ld), "maxpool2_fwd: channels %d", c);
  long long total = (long long)n * (h / 2) * (w / 2) * (c / 8);
  LAUNCH_1D(maxpool2_fwd_kernel, total, (cudaStream_t)stream, (const bf16*)x, x_ld, (bf16*)out, out_ld, h, w, c / 8,
            total);
  return CRFR_OK;
}

extern "C" int crfr_maxpool2_bwd(const void* x, int x_ld, const void* dout, int dout_ld, void* dx, int dx_ld, int n,
                                 int h, int w, int c, void* stream) {
  CRFR_CHECK_ARG(x && dout && dx && n > 0 && !(h & 1) && !(w & 1), "maxpool2_bwd: bad argument");
  CRFR_CHECK_ARG(vec_ok(c, x_ld) && vec_ok(c, dout_ld) && vec_ok(c, dx_ld), "maxpool2_bwd: channels %d", c);
  long long total = (long long)n * (h / 2) * (w / 2) * (c / 8);
  LAUNCH_1D(maxpool2_bwd_kernel, total, (cudaStream_t)stream, (const bf16*)x, x_ld, (const bf16*)dout, dout_ld,
            (bf16*)dx, dx_ld, h, w, c / 8, total);
  return CRFR_OK;
}

extern "C" int crfr_upnearest2_add_fwd(const void* up, int up_ld, const void* low, int low_ld, void* out, int out_ld,
                                       int n, int h, int w, int c, void* stream) {
  CRFR_CHECK_ARG(up && low && out && n > 0 && h > 0 && w > 0, "upnearest2_add_fwd: bad argument");
  CRFR_CHECK_ARG(vec_ok(c, up_ld) && vec_ok(c, low_ld) && vec_ok(c, out_ld), "upnearest2_add_fwd: channels %d", c);
  long long total = (long long)n * (2 * h) * (2 * w) * (c / 8);
  LAUNCH_1D(upadd_fwd_kernel, total, (cudaStream_t)stream, (const bf16*)up, up_ld, (const bf16*)low, low_ld,
            (bf16*)out, out_ld, h, w, c / 8, total);
  return CRFR_OK;
}

extern "C" int crfr_upnearest2_bwd(const void* dout, int dout_ld, void* dlow, int dlow_ld, int n, int h, int w, int c,
                                   void* stream) {
  CRFR_CHECK_ARG(dout && dlow && n > 0 && h > 0 && w > 0, "upnearest2_bwd: bad argument");
  CRFR_CHECK_ARG(vec_ok(c, dout_ld) && vec_ok(c, dlow_ld), "upnearest2_bwd: channels %d", c);
  long long total = (long long)n * h * w * (c / 8);
  LAUNCH_1D(up_bwd_kernel, total, (cudaStream_t)stream, (const bf16*)dout, dout_ld, (bf16*)dlow, dlow_ld, h, w, c / 8,
            total);
  return CRFR_OK;
}

extern "C" int crfr_add_n(const void* a, int a_ld, const void* b, int b_ld, const void* c3, int c_ld, void* out,
                          int out_ld, long long pixels, int c, void* stream) {
  CRFR_CHECK_ARG(a && b && out && pixels > 0, "add_n: bad argument");
  if (c == 4 && a_ld == 4 && b_ld == 4 && out_ld == 4 && (!c3 || c_ld == 4)) {
    LAUNCH_1D(add_n4_kernel, pixels, (cudaStream_t)stream, (const bf16*)a, (const bf16*)b, (const bf16*)c3,
              (bf16*)out, pixels);
    return CRFR_OK;
  }
  CRFR_CHECK_ARG(vec_ok(c, a_ld) && vec_ok(c, b_ld) && vec_ok(c, out_ld) && (!c3 || vec_ok(c, c_ld)),
                 "add_n: channels %d", c);
  long long total = pixels * (c / 8);
  LAUNCH_1D(add_n_kernel, total, (cudaStream_t)stream, (const bf16*)a, a_ld, (const bf16*)b, b_ld, (const bf16*)c3,
            c_ld, (bf16*)out, out_ld, c / 8, total);
  return CRFR_OK;
}
