// Persistent row-streaming 3x3 convolution for the dominant FSRNet layer shape (64 -> 64 channels, width 128:
// 36 of the 98 convolutions and 87 % of the MACs, model/FSRnet.py:79,85 inside the coarse and decoder stacks).
//
// Each CTA owns a contiguous range of output rows of the flattened (image, row) space (perfect balance over the 148
// SMs) and streams through it:
//   * the 9 x [64 x 64] weight taps (72 KB) are loaded once per CTA and stay in shared memory;
//   * every INPUT row is fetched exactly once per CTA (TMA box 64ch x 130px: the row plus its left/right halo, zero
//     filled outside the image) into a ring of row buffers; the 9 taps of an output row are just 9 shared-memory
//     descriptors into three ring slots - the +-1 pixel shift is a +-128-byte start offset inside the
//     SWIZZLE_128B tile (measured on B200: the UMMA swizzle is a function of the absolute smem address, so shifted
//     views stay consistent with what TMA wrote; the descriptor's base_offset field must stay 0);
//     L2 -> smem traffic is therefore ~1.15x the input instead of 9x for the per-tap tiling of tc_conv.cu;
//   * one elected thread issues 36 tcgen05.mma (M=128 pixels, N=64, K=16) per output row into one of two TMEM
//     accumulators, so the epilogue warps (tcgen05.ld -> +bias -> bf16 -> NHWC row store) of row y overlap the MMAs
//     of row y+1.
// dgrad is the same kernel with the tap table flipped.
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "crfr.h"
#include "internal.h"
#include "sm100.cuh"

using namespace sm100;

namespace {

constexpr int kW = 128;                    // image width handled by this kernel
constexpr int kC = 64;                     // channels in = out
constexpr int kRowBytes = 130 * 128;       // one input row with halo, 128 B per pixel
constexpr int kSlotBytes = 17 * 1024;      // ring slot stride (1024-aligned)
constexpr int kSlots = 6;
constexpr int kWeightBytes = 9 * kC * 128; // 9 taps x 64 rows x 128 B
constexpr int kThreads = 192;
constexpr int kStageOutBytes = 128 * 128;  // one output row staged for the TMA store (x2: double buffered)
constexpr int kSmemBytes = kWeightBytes + kSlots * kSlotBytes + 2 * kStageOutBytes + 1024 + 256;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

struct RowParams {
  int n, h;              // images, rows per image (width is kW)
  int total_rows;        // n * h
  int flip;              // 1: dgrad (tap table flipped)
  const float* bias;
};

__global__ void __launch_bounds__(kThreads, 1)
rowconv_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmY, RowParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = base;
  uint8_t* sRing = base + kWeightBytes;
  uint8_t* sOut = sRing + kSlots * kSlotBytes;
  uint64_t* full = (uint64_t*)(sOut + 2 * kStageOutBytes);
  uint64_t* empty = full + kSlots;
  uint64_t* w_full = empty + kSlots;
  uint64_t* acc_full = w_full + 1;     // [2]
  uint64_t* acc_empty = acc_full + 2;  // [2]
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);
    }
    fence_barrier_init();
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmY);
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // contiguous range of flattened output rows for this CTA
  const long long r_begin = (long long)p.total_rows * blockIdx.x / gridDim.x;
  const long long r_end = (long long)p.total_rows * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(w_full, kWeightBytes);
      for (int j = 0; j < 3; ++j) tma_load_2d(sW + j * 192 * 128, &tmW, w_full, 0, j * 192);
    }
    int g = 0;  // running index of loaded input rows
    long long r = r_begin;
    while (r < r_end) {
      const int n = (int)(r / p.h), y0 = (int)(r % p.h);
      const int seg = (int)min((long long)(p.h - y0), r_end - r);
      for (int iy = y0 - 1; iy <= y0 + seg; ++iy, ++g) {
        const int s = g % kSlots;
        mbar_wait(&empty[s], ((g / kSlots) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&full[s], kRowBytes);
          tma_load_4d(sRing + s * kSlotBytes, &tmX, &full[s], 0, -1, iy, n);
        }
      }
      r += seg;
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, kC, 0, 0);
    const uint64_t wdesc0 = make_smem_desc_sw128(smem_u32(sW), 16, 1024);
    const uint64_t rdesc0 = make_smem_desc_sw128(smem_u32(sRing), 16, 1024);
    mbar_wait(w_full, 0);
    int g = 0, orow = 0;  // g: ring index of the segment's first input row; orow: output rows issued so far
    long long r = r_begin;
    while (r < r_end) {
      const int y0 = (int)(r % p.h);
      const int seg = (int)min((long long)(p.h - y0), r_end - r);
      for (int j = 0; j < seg; ++j, ++orow) {
        const int buf = orow & 1;
        mbar_wait(&acc_empty[buf], ((orow >> 1) & 1) ^ 1);
        {  // only the newest of the three input rows can still be in flight
          const int gi = g + j + 2;
          if (j == 0) {
            mbar_wait(&full[(gi - 2) % kSlots], ((gi - 2) / kSlots) & 1);
            mbar_wait(&full[(gi - 1) % kSlots], ((gi - 1) / kSlots) & 1);
          }
          mbar_wait(&full[gi % kSlots], (gi / kSlots) & 1);
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem + buf * kC;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int gi = g + j + ky;
          const uint64_t rowd = rdesc0 + (uint64_t)(((gi % kSlots) * kSlotBytes) >> 4);
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int tap = p.flip ? (8 - (ky * 3 + kx)) : (ky * 3 + kx);
            const uint64_t ad = rowd + (uint64_t)((kx * 128) >> 4);
            const uint64_t bd = wdesc0 + (uint64_t)((tap * (kC * 128)) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (leader) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (uint32_t)((ky | kx | k) != 0));
          }
        }
        if (leader) {
          umma_commit(&acc_full[buf]);
          umma_commit(&empty[(g + j) % kSlots]);        // input row j of the segment is no longer needed
        }
        __syncwarp();
      }
      if (leader) {  // the last two input rows of the segment
        umma_commit(&empty[(g + seg) % kSlots]);
        umma_commit(&empty[(g + seg + 1) % kSlots]);
      }
      __syncwarp();
      g += seg + 2;
      r += seg;
    }
  } else {
    // epilogue: TMEM -> registers -> (+bias, bf16) -> swizzled smem row -> one TMA store per output row
    const int q = warp & 3;
    const int x = q * 32 + lane;  // pixel within the row
    const bool issuer = (warp == 2) && (lane == 0);
    float bias[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) bias[c] = p.bias ? p.bias[c] : 0.f;
    int orow = 0;
    for (long long r = r_begin; r < r_end; ++r, ++orow) {
      const int buf = orow & 1;
      mbar_wait(&acc_full[buf], (orow >> 1) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * kC, v0);
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * kC + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      // the TMA store issued two rows ago must have finished reading this staging buffer
      if (issuer) tma_store_wait_read<1>();
      named_bar_sync(1, 128);
      uint8_t* srow = sOut + buf * kStageOutBytes + x * 128;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v0[j + e]) + bias[j + e];
        *reinterpret_cast<bf16x8*>(srow + (((j >> 3) ^ (x & 7)) << 4)) = pack8(f);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v1[j + e]) + bias[32 + j + e];
        *reinterpret_cast<bf16x8*>(srow + ((((32 + j) >> 3) ^ (x & 7)) << 4)) = pack8(f);
      }
      fence_proxy_async();
      named_bar_sync(1, 128);
      if (issuer) {
        tma_store_4d(&tmY, sOut + buf * kStageOutBytes, 0, 0, (int)(r % p.h), (int)(r / p.h));
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<128>(tmem);
  }
}

}  // namespace

int crfr_rowconv_supported(int h, int w, int cin, int cout, int k, int stride, int pad) {
  return w == kW && cin == kC && cout == kC && k == 3 && stride == 1 && pad == 1 && h >= 1;
}

// src/dst: NHWC bf16 [n][h][128][64] views; w_packed: [9][64][64] bf16 ([tap][n][k]); flip = 1 for dgrad
int crfr_rowconv(const void* src, int src_ld, int n, int h, const void* w_packed, int flip, const float* bias, void* dst,
                 int dst_ld, cudaStream_t st) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    crfr_set_error("cuTensorMapEncodeTiled entry point not available");
    return CRFR_ECUDA;
  }
  CRFR_CHECK_ARG(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 &&
                     (src_ld & 7) == 0 && (dst_ld & 7) == 0,
                 "rowconv: pointers must be 16B aligned and ld a multiple of 8");
  CUtensorMap tmX, tmW, tmY;
  {
    cuuint64_t dims[4] = {(cuuint64_t)kC, (cuuint64_t)kW, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)dst_ld * 2, (cuuint64_t)kW * dst_ld * 2, (cuuint64_t)h * kW * dst_ld * 2};
    cuuint32_t box[4] = {64, 128, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dst, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      crfr_set_error("rowconv: cuTensorMapEncodeTiled(output) failed: %d", (int)r);
      return CRFR_ECUDA;
    }
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)kC, (cuuint64_t)kW, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)src_ld * 2, (cuuint64_t)kW * src_ld * 2, (cuuint64_t)h * kW * src_ld * 2};
    cuuint32_t box[4] = {64, 130, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(src), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      crfr_set_error("rowconv: cuTensorMapEncodeTiled(activation) failed: %d", (int)r);
      return CRFR_ECUDA;
    }
  }
  {
    cuuint64_t dims[2] = {64, 9 * 64};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 192};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      crfr_set_error("rowconv: cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
      return CRFR_ECUDA;
    }
  }
  static bool attr_done = false;
  if (!attr_done) {
    CRFR_CUDA(cudaFuncSetAttribute(rowconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_done = true;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    CRFR_CUDA(cudaGetDevice(&dev));
    CRFR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  RowParams p;
  p.n = n; p.h = h; p.total_rows = n * h; p.flip = flip;
  p.bias = bias;
  int grid = sms < p.total_rows ? sms : p.total_rows;
  rowconv_kernel<<<grid, kThreads, kSmemBytes, st>>>(tmX, tmW, tmY, p);
  CRFR_COUNT_LAUNCH();
  CRFR_LAUNCH_CHECK();
  return CRFR_OK;
}
