// Reflection padding and Tanh for the SUPER_RESOLUTION FSRNet variant (SUPER_RESOLUTION/model/FSRnet.py:251-416:
// nn.ReflectionPad2d(p) in front of every convolution, nn.Tanh() behind the two image heads).  HBM-bound gathers:
// NHWC bf16, one 16-byte channel vector (or one 8-byte 3(+1)-channel pixel) per thread, coalesced along channels.
#include "common.cuh"
#include "crfr.h"

namespace {

__device__ __forceinline__ int reflect(int i, int n) {   // nn.ReflectionPad2d: -1 -> 1, n -> n - 2 (pad < n)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

template <int V>   // V bf16 elements per thread (8: 16-byte vector, 4: one 3(+1)-channel pixel)
__global__ void __launch_bounds__(256)
reflect_pad_fwd_kernel(const bf16* __restrict__ x, int x_ld, bf16* __restrict__ out, int out_ld, int h, int w, int groups,
                       int pad, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int g = (int)(i % groups);
  long long p = i / groups;
  const int ow = w + 2 * pad, oh = h + 2 * pad;
  const int ox = (int)(p % ow);
  p /= ow;
  const int oy = (int)(p % oh);
  const long long n = p / oh;
  const int sy = reflect(oy - pad, h), sx = reflect(ox - pad, w);
  const bf16* src = x + ((n * h + sy) * (long long)w + sx) * x_ld + g * V;
  bf16* dst = out + ((n * oh + oy) * (long long)ow + ox) * out_ld + g * V;
  if (V == 8) *reinterpret_cast<bf16x8*>(dst) = *reinterpret_cast<const bf16x8*>(src);
  else *reinterpret_cast<bf16x4*>(dst) = *reinterpret_cast<const bf16x4*>(src);
}

// dx[y][x] = sum of dout over the (up to 2 x 2) padded positions that read input pixel (y, x); fp32 sum, one rounding
template <int V>
__global__ void __launch_bounds__(256)
reflect_pad_bwd_kernel(const bf16* __restrict__ dout, int dout_ld, bf16* __restrict__ dx, int dx_ld, int h, int w,
                       int groups, int pad, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int g = (int)(i % groups);
  long long p = i / groups;
  const int xx = (int)(p % w);
  p /= w;
  const int yy = (int)(p % h);
  const long long n = p / h;
  const int ow = w + 2 * pad, oh = h + 2 * pad;
  int ys[3], xs[3], ny = 0, nx = 0;
  ys[ny++] = yy + pad;
  if (yy >= 1 && yy <= pad) ys[ny++] = pad - yy;
  if (yy <= h - 2 && yy >= h - 1 - pad) ys[ny++] = 2 * (h - 1) + pad - yy;
  xs[nx++] = xx + pad;
  if (xx >= 1 && xx <= pad) xs[nx++] = pad - xx;
  if (xx <= w - 2 && xx >= w - 1 - pad) xs[nx++] = 2 * (w - 1) + pad - xx;
  float acc[V];
#pragma unroll
  for (int e = 0; e < V; ++e) acc[e] = 0.f;
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) {
      const bf16* src = dout + ((n * oh + ys[a]) * (long long)ow + xs[b]) * dout_ld + g * V;
#pragma unroll
      for (int e = 0; e < V; e += 2) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src + e));
        acc[e] += f.x;
        acc[e + 1] += f.y;
      }
    }
  bf16* dst = dx + ((n * h + yy) * (long long)w + xx) * dx_ld + g * V;
#pragma unroll
  for (int e = 0; e < V; e += 2) *reinterpret_cast<__nv_bfloat162*>(dst + e) = __floats2bfloat162_rn(acc[e], acc[e + 1]);
}

__global__ void __launch_bounds__(256) tanh_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = tanhf(x[i]);
}
__global__ void __launch_bounds__(256)
tanh_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = dy[i] * (1.f - y[i] * y[i]);
}

int check_pad(const char* who, const void* a, const void* b, int n, int h, int w, int c, int ld_in, int ld_out, int pad) {
  CRFR_CHECK_ARG(a && b && n > 0 && h > 0 && w > 0 && c > 0, "%s: bad argument", who);
  CRFR_CHECK_ARG(pad >= 0 && pad < h && pad < w, "%s: padding %d must be smaller than the image (%d x %d)", who, pad, h, w);
  CRFR_CHECK_ARG((c % 8 == 0 && ld_in >= c && ld_out >= c && ld_in % 8 == 0 && ld_out % 8 == 0) ||
                     (c <= 4 && ld_in == 4 && ld_out == 4),
                 "%s: channels %d (ld %d / %d) must be a multiple of 8, or <= 4 with ld 4", who, c, ld_in, ld_out);
  return CRFR_OK;
}

}  // namespace

extern "C" int crfr_reflect_pad_fwd(const void* x, int x_ld, void* out, int out_ld, int n, int h, int w, int c, int pad,
                                    void* stream) {
  CRFR_TRY(check_pad("reflect_pad_fwd", x, out, n, h, w, c, x_ld, out_ld, pad));
  cudaStream_t st = (cudaStream_t)stream;
  const int groups = c % 8 == 0 ? c / 8 : 1;
  const long long total = (long long)n * (h + 2 * pad) * (w + 2 * pad) * groups;
  if (c % 8 == 0)
    reflect_pad_fwd_kernel<8><<<crfr_cdiv(total, 256), 256, 0, st>>>((const bf16*)x, x_ld, (bf16*)out, out_ld, h, w, groups, pad, total);
  else
    reflect_pad_fwd_kernel<4><<<crfr_cdiv(total, 256), 256, 0, st>>>((const bf16*)x, x_ld, (bf16*)out, out_ld, h, w, groups, pad, total);
  CRFR_COUNT_LAUNCH();
  CRFR_LAUNCH_CHECK();
  return CRFR_OK;
}

extern "C" int crfr_reflect_pad_bwd(const void* dout, int dout_ld, void* dx, int dx_ld, int n, int h, int w, int c, int pad,
                                    void* stream) {
  CRFR_TRY(check_pad("reflect_pad_bwd", dout, dx, n, h, w, c, dx_ld, dout_ld, pad));
  cudaStream_t st = (cudaStream_t)stream;
  const int groups = c % 8 == 0 ? c / 8 : 1;
  const long long total = (long long)n * h * w * groups;
  if (c % 8 == 0)
    reflect_pad_bwd_kernel<8><<<crfr_cdiv(total, 256), 256, 0, st>>>((const bf16*)dout, dout_ld, (bf16*)dx, dx_ld, h, w, groups, pad, total);
  else
    reflect_pad_bwd_kernel<4><<<crfr_cdiv(total, 256), 256, 0, st>>>((const bf16*)dout, dout_ld, (bf16*)dx, dx_ld, h, w, groups, pad, total);
  CRFR_COUNT_LAUNCH();
  CRFR_LAUNCH_CHECK();
  return CRFR_OK;
}

extern "C" int crfr_tanh_fwd(const float* x, float* y, long long numel, void* stream) {
  CRFR_CHECK_ARG(x && y && numel > 0, "tanh_fwd: bad argument");
  tanh_fwd_kernel<<<crfr_cdiv(numel, 256), 256, 0, (cudaStream_t)stream>>>(x, y, numel);
  CRFR_COUNT_LAUNCH();
  CRFR_LAUNCH_CHECK();
  return CRFR_OK;
}

extern "C" int crfr_tanh_bwd(const float* y, const float* dy, float* dx, long long numel, void* stream) {
  CRFR_CHECK_ARG(y && dy && dx && numel > 0, "tanh_bwd: bad argument");
  tanh_bwd_kernel<<<crfr_cdiv(numel, 256), 256, 0, (cudaStream_t)stream>>>(y, dy, dx, numel);
  CRFR_COUNT_LAUNCH();
  CRFR_LAUNCH_CHECK();
  return CRFR_OK;
}
