// Persistent row-streaming weight gradient for the dominant FSRNet layer shape (3x3, 64 -> 64 channels, width 128).
//
//   dW[ky][kx][ci][co] = sum_{n,y,x} X[n][y+ky-1][x+kx-1][ci] * dY[n][y][x][co]
//
// Pixels are the GEMM K dimension, so both operands are MN-major views of NHWC rows (channels contiguous).  As in
// rowconv.cu the natural N = 64 MMA is avoided (it is capped at ~50 % of the tensor pipe, tools/micro/mma_bench.cu):
// for one input row q of X the three dY rows q-1, q, q+1 (= ky 2, 1, 0) are stacked along N through the descriptor's
// leading-byte-offset (consecutive slots of the dY ring), and two kx taps are stacked along M the same way - the kx
// shift of the X row (loaded once, with halo) is a 128-byte start offset, so the second M atom is simply LBO = 128 B
// away:
//     D1[(kx 0|1) x 64 ci][(ky 2|1|0) x 64 co] += X_q(shift 0|1)^T * [dY_{q-1} | dY_q | dY_{q+1}]      (M=128, N=192)
//     D2[(kx 2|-) x 64 ci][...]                 += X_q(shift 2|3)^T * [...]                               (upper half unused)
// 16 MMAs (K = 16 pixels each) per row instead of 48 N = 64 MMAs.  Rows outside the image are TMA zero fill.
// Each CTA streams a contiguous range of rows, keeps its partial dW in TMEM for its whole lifetime and writes one
// fp32 slab; a second kernel sums the slabs in fixed order (deterministic, no float atomics) and accumulates into
// the reference-layout gradient dW[co][ci][ky][kx].
//
// ref: the weight gradients of the nn.Conv2d sites model/FSRnet.py:79,85 (coarse / decoder residual stacks).
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "crfr.h"
#include "internal.h"
#include "sm100.cuh"

using namespace sm100;

namespace {

constexpr int kW = 128;
constexpr int kC = 64;
constexpr int kXRowBytes = 132 * 128;      // X row with halo: x = -1 .. 130 (shift 3 of the unused atom stays inside)
constexpr int kXSlotBytes = 17 * 1024;
constexpr int kXSlots = 4;
constexpr int kDyRowBytes = 128 * 128;
constexpr int kDySlots = 7;
constexpr int kThreads = 192;
constexpr int kSlabFloats = 9 * kC * kC;
constexpr int kSmemBytes = kXSlots * kXSlotBytes + kDySlots * kDyRowBytes + 1024 + 512;

struct WgParams {
  int n, h, total_rows;
  float* slabs;   // [gridDim.x][9][64][64]
};

__global__ void __launch_bounds__(kThreads, 1)
rowwgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = base;
  uint8_t* sDY = base + kXSlots * kXSlotBytes;
  uint64_t* full_x = (uint64_t*)(sDY + kDySlots * kDyRowBytes);
  uint64_t* empty_x = full_x + kXSlots;
  uint64_t* full_dy = empty_x + kXSlots;
  uint64_t* empty_dy = full_dy + kDySlots;
  uint64_t* done = empty_dy + kDySlots;
  uint32_t* tmem_slot = (uint32_t*)(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kXSlots; ++s) {
      mbar_init(&full_x[s], 1);
      mbar_init(&empty_x[s], 1);
    }
    for (int s = 0; s < kDySlots; ++s) {
      mbar_init(&full_dy[s], 1);
      mbar_init(&empty_dy[s], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmDY);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const long long r_begin = (long long)p.total_rows * blockIdx.x / gridDim.x;
  const long long r_end = (long long)p.total_rows * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    const bool leader = elect_one();
    int gx = 0, gd = 0;
    long long r = r_begin;
    while (r < r_end) {
      const int n = (int)(r / p.h), y0 = (int)(r % p.h);
      const int seg = (int)min((long long)(p.h - y0), r_end - r);
      // dY rows y0-1 .. y0+seg (out-of-image rows are zero filled), X rows y0 .. y0+seg-1; interleaved in use order
      for (int i = -1; i <= seg; ++i) {
        {
          const int s = gd % kDySlots;
          mbar_wait(&empty_dy[s], ((gd / kDySlots) & 1) ^ 1);
          if (leader) {
            mbar_expect_tx(&full_dy[s], kDyRowBytes);
            tma_load_4d(sDY + s * kDyRowBytes, &tmDY, &full_dy[s], 0, 0, y0 + i, n);
          }
          ++gd;
        }
        if (i >= 1) {   // X row y0+i-1 is first needed once dY row y0+i has been requested
          const int s = gx % kXSlots;
          mbar_wait(&empty_x[s], ((gx / kXSlots) & 1) ^ 1);
          if (leader) {
            mbar_expect_tx(&full_x[s], kXRowBytes);
            tma_load_4d(sX + s * kXSlotBytes, &tmX, &full_x[s], 0, -1, y0 + i - 1, n);
          }
          ++gx;
        }
      }
      r += seg;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    // A: two 64-channel atoms 128 B apart (the kx and kx+1 views of the same row); B: N atoms one dY slot apart
    const uint64_t xdesc0 = make_smem_desc_sw128(smem_u32(sX), 128, 1024);
    const uint64_t dydesc0 = make_smem_desc_sw128(smem_u32(sDY), kDyRowBytes, 1024);
    int gx = 0, gd = 0;   // gd: ring index of dY row (q-1) of the current X row
    bool first = true;
    long long r = r_begin;
    while (r < r_end) {
      const int y0 = (int)(r % p.h);
      const int seg = (int)min((long long)(p.h - y0), r_end - r);
      for (int i = 0; i < seg; ++i, ++gx, ++gd) {
        const int sx = gx % kXSlots;
        if (i == 0) {
          mbar_wait(&full_dy[gd % kDySlots], (gd / kDySlots) & 1);
          mbar_wait(&full_dy[(gd + 1) % kDySlots], ((gd + 1) / kDySlots) & 1);
        }
        mbar_wait(&full_dy[(gd + 2) % kDySlots], ((gd + 2) / kDySlots) & 1);
        mbar_wait(&full_x[sx], (gx / kXSlots) & 1);
        tc_fence_after();
        const uint64_t xd = xdesc0 + (uint64_t)((sx * kXSlotBytes) >> 4);
        const int s0 = gd % kDySlots;
        const int cnt1 = min(3, kDySlots - s0);     // dY rows before the ring wraps
        const uint64_t b0 = dydesc0 + (uint64_t)((s0 * kDyRowBytes) >> 4);
        const uint32_t id1 = make_idesc_bf16(128, kC * cnt1, 1, 1);
        const uint32_t id2 = make_idesc_bf16(128, kC * (cnt1 < 3 ? 3 - cnt1 : 1), 1, 1);
        const uint32_t d2 = tmem + kC * cnt1;
        const uint32_t acc0 = first ? 0u : 1u;
        if (leader) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {             // 8 x 16 pixels
            const uint32_t acc = k == 0 ? acc0 : 1u;
            const uint64_t a01 = xd + (uint64_t)(k * (2048 >> 4));
            const uint64_t a2x = a01 + (uint64_t)(256 >> 4);
            umma_bf16(tmem, a01, b0 + (uint64_t)(k * (2048 >> 4)), id1, acc);
            umma_bf16(tmem + 256, a2x, b0 + (uint64_t)(k * (2048 >> 4)), id1, acc);
            if (cnt1 < 3) {
              umma_bf16(d2, a01, dydesc0 + (uint64_t)(k * (2048 >> 4)), id2, acc);
              umma_bf16(d2 + 256, a2x, dydesc0 + (uint64_t)(k * (2048 >> 4)), id2, acc);
            }
          }
        }
        first = false;
        if (leader) {
          umma_commit(&empty_x[sx]);
          umma_commit(&empty_dy[s0]);               // dY row q-1 is not needed by later X rows
        }
        __syncwarp();
      }
      if (leader) {                                  // the last two dY rows of the segment
        umma_commit(&empty_dy[gd % kDySlots]);
        umma_commit(&empty_dy[(gd + 1) % kDySlots]);
      }
      __syncwarp();
      gd += 2;
      r += seg;
    }
    if (leader) umma_commit(done);
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> this CTA's slab
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ci = row & 63;
    mbar_wait(done, 0);
    tc_fence_after();
    float* slab = p.slabs + (size_t)blockIdx.x * kSlabFloats;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int kx = half == 0 ? (row >> 6) : 2;
      const bool live = half == 0 || row < 64;      // warp-uniform
#pragma unroll
      for (int c0 = 0; c0 < 192; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + half * 256 + c0, v);
        tmem_ld_wait();
        if (live) {
          const int ky = 2 - c0 / 64;
          float* dst = slab + ((size_t)(ky * 3 + kx) * kC + ci) * kC + (c0 & 63);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                              __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// dW[co][ci][t] += sum_slabs slab[s][t][ci][co]   (fixed order)
__global__ void __launch_bounds__(256)
slab_reduce_kernel(const float* __restrict__ slabs, int nslabs, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over [t][ci][co]
  if (i >= kSlabFloats) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int s = 0;
  for (; s + 4 <= nslabs; s += 4) {
    s0 += slabs[(size_t)s * kSlabFloats + i];
    s1 += slabs[(size_t)(s + 1) * kSlabFloats + i];
    s2 += slabs[(size_t)(s + 2) * kSlabFloats + i];
    s3 += slabs[(size_t)(s + 3) * kSlabFloats + i];
  }
  for (; s < nslabs; ++s) s0 += slabs[(size_t)s * kSlabFloats + i];
  const int co = i & 63, ci = (i >> 6) & 63, t = i >> 12;
  dw[((size_t)co * kC + ci) * 9 + t] += (s0 + s1) + (s2 + s3);
}

int encode(CUtensorMap* m, const void* ptr, int n, int h, int ld, int box_w, const char* what) {
  const unsigned long long dims[4] = {(unsigned long long)kC, (unsigned long long)kW, (unsigned long long)h,
                                      (unsigned long long)n};
  const unsigned long long strides[3] = {(unsigned long long)ld * 2, (unsigned long long)kW * ld * 2,
                                         (unsigned long long)h * kW * ld * 2};
  const unsigned int box[4] = {64, (unsigned int)box_w, 1, 1};
  return crfr_tmap_encode_bf16(m, ptr, 4, dims, strides, box, what);
}

int grid_for(int total_rows) {
  const int sms = crfr_sm_count();
  return sms < total_rows ? sms : total_rows;
}

}  // namespace

int crfr_rowwgrad_supported(int h, int w, int cin, int cout, int k, int stride, int pad) {
  return w == kW && cin == kC && cout == kC && k == 3 && stride == 1 && pad == 1 && h >= 1;
}

size_t crfr_rowwgrad_ws_bytes(int n, int h) { return sizeof(float) * (size_t)grid_for(n * h) * kSlabFloats + 256; }

// x, dy: NHWC bf16 [n][h][128][64] views; dw: fp32 [64][64][3][3], accumulated.
int crfr_rowwgrad(const void* x, int x_ld, const void* dy, int dy_ld, int n, int h, float* dw, void* ws,
                  size_t ws_bytes, cudaStream_t st) {
  CRFR_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && (x_ld & 7) == 0 && (dy_ld & 7) == 0,
                 "rowwgrad: pointers must be 16B aligned and ld a multiple of 8");
  const size_t need = crfr_rowwgrad_ws_bytes(n, h);
  if (!ws || ws_bytes < need) {
    crfr_set_error("rowwgrad: workspace %zu < %zu", ws_bytes, need);
    return CRFR_EWORKSPACE;
  }
  CUtensorMap tmX, tmDY;
  CRFR_TRY(encode(&tmX, x, n, h, x_ld, 132, "x"));
  CRFR_TRY(encode(&tmDY, dy, n, h, dy_ld, 128, "dy"));
  static std::atomic<unsigned long long> attr_done{0};
  CRFR_CUDA((cudaError_t)crfr_smem_attr(rowwgrad_kernel, kSmemBytes, attr_done));
  WgParams p;
  p.n = n; p.h = h; p.total_rows = n * h;
  p.slabs = (float*)ws;
  const int grid = grid_for(p.total_rows);
  rowwgrad_kernel<<<grid, kThreads, kSmemBytes, st>>>(tmX, tmDY, p);
  CRFR_COUNT_LAUNCH();
  CRFR_LAUNCH_CHECK();
  slab_reduce_kernel<<<crfr_cdiv(kSlabFloats, 256), 256, 0, st>>>((const float*)ws, grid, dw);
  CRFR_COUNT_LAUNCH();
  CRFR_LAUNCH_CHECK();
  return CRFR_OK;
}
