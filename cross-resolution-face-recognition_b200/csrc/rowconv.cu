// Persistent row-streaming 3x3 convolution for the dominant FSRNet layer shape (64 -> 64 channels, width 128:
// 36 of the 98 convolutions and 87 % of the MACs, model/FSRnet.py:79,85 inside the coarse and decoder stacks).
//
// Measured on B200 (tools/micro/mma_bench.cu): a tcgen05.mma with both operands in shared memory costs
// ~89 / 103 / 166 cycles at N = 64 / 128 / 256 (M = 128, K = 16), i.e. an N = 64 tile cannot exceed ~50 % of the
// tensor-pipe peak.  The kernel therefore never issues the natural [128 pixels x 64 cout] MMA per tap.  Instead the
// three ky taps of one kx are stacked along N:
//
//     D[128 pixels of input row r][ (row r-1 | row r | row r+1) x 64 cout ]  +=  X_r(shifted by kx) * [W(ky=2,kx) | W(ky=1,kx) | W(ky=0,kx)]
//
// One N = 192 MMA scatters an input row into the accumulators of the three output rows it contributes to.  The
// accumulators of consecutive output rows are consecutive 64-column slots of a ring over the whole TMEM (8 slots x
// 64 columns), so "three output rows" is one contiguous 192-column destination (split in two MMAs where the ring
// wraps).  12 MMAs per row instead of 36, each ~1.5x the cost: 2x the throughput of the per-tap formulation.
//
// Structure (one CTA per SM, each owning a contiguous range of the flattened (image, row) space):
//   * the 9 x [64 x 64] weight taps (72 KB) are loaded once per CTA, ordered [kx][ky descending] so that every
//     (kx) is one 192-row K-major B operand;
//   * every INPUT row is fetched once per CTA (TMA box 64 ch x 130 px: the row plus left/right halo, zero filled)
//     into a ring of row buffers; the +-1 pixel shift of kx is a +-128-byte start offset of the A descriptor inside
//     the SWIZZLE_128B tile (the UMMA swizzle is a function of the absolute smem address);
//   * one elected thread issues the MMAs; an output row is complete after the input row below it has been issued,
//     then 4 epilogue warps drain its TMEM slot (tcgen05.ld -> +bias -> bf16), stage it in swizzled smem, accumulate
//     the per-channel InstanceNorm sum / sum-of-squares of the stored (rounded) values from that tile, and hand it
//     to one TMA store.  Epilogue of row y overlaps the MMAs of rows y+2...
//   * InstanceNorm partials are written per (image, CTA) in fixed slots and finalised in fixed order
//     (crfr_norm_finalize): deterministic, no float atomics.
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
constexpr int kSlots = 5;
constexpr int kTapBytes = kC * 128;        // one [64 x 64] weight tap
constexpr int kWeightBytes = 9 * kTapBytes;
constexpr int kThreads = 192;
constexpr int kStageOutBytes = 128 * 128;  // one output row staged for the TMA store (x2: double buffered)
constexpr int kAccSlots = 8;               // TMEM ring: 8 x 64 columns = all 512 columns
constexpr int kSmemBytes = kWeightBytes + kSlots * kSlotBytes + 2 * kStageOutBytes + 4 * 128 * 4 /*stats*/ +
                           kC * 4 /*bias*/ + 1024 /*align*/ + 512 /*barriers*/;

struct RowParams {
  int n, h;              // images, rows per image (width is kW)
  int total_rows;        // n * h
  int flip;              // 1: dgrad (tap table flipped)
  const float* bias;
  float* partial;        // InstanceNorm partials [n][parts][2][64] or nullptr
  int parts;             // partial slots per image
};

// first CTA whose row range [R*b/G, R*(b+1)/G) contains row x
__device__ __forceinline__ int first_cta_of_row(long long x, int R, int G) {
  return (int)(((x + 1) * G + R - 1) / R) - 1;
}

__global__ void __launch_bounds__(kThreads, 1)
rowconv_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmY, RowParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = base;
  uint8_t* sRing = base + kWeightBytes;
  uint8_t* sOut = sRing + kSlots * kSlotBytes;
  float* sStat = (float*)(sOut + 2 * kStageOutBytes);   // [4 pixel groups][2][64]
  float* sBias = sStat + 4 * 128;
  uint64_t* full = (uint64_t*)(sBias + kC);
  uint64_t* empty = full + kSlots;
  uint64_t* w_full = empty + kSlots;
  uint64_t* acc_full = w_full + 1;             // [kAccSlots]
  uint64_t* acc_empty = acc_full + kAccSlots;  // [kAccSlots]
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + kAccSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(w_full, 1);
    for (int b = 0; b < kAccSlots; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);
    }
    fence_barrier_init();
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmY);
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kC) sBias[threadIdx.x - 64] = p.bias ? p.bias[threadIdx.x - 64] : 0.f;
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp >= 2) {   // all accumulators start at zero: every MMA of this kernel accumulates
    for (int c = 0; c < 512; c += 32) tmem_st32_zero(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // contiguous range of flattened output rows for this CTA
  const long long r_begin = (long long)p.total_rows * blockIdx.x / gridDim.x;
  const long long r_end = (long long)p.total_rows * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(w_full, kWeightBytes);
      for (int kx = 0; kx < 3; ++kx)
        for (int j = 0; j < 3; ++j) {      // B rows [64 j, 64 j + 64) of kx: the tap that maps input row r to output row r-1+j
          const int tap = (2 - j) * 3 + kx;
          tma_load_2d(sW + (kx * 3 + j) * kTapBytes, &tmW, w_full, 0, (p.flip ? 8 - tap : tap) * kC);
        }
    }
    int g = 0;  // running index of loaded input rows
    long long r = r_begin;
    while (r < r_end) {
      const int n = (int)(r / p.h), y0 = (int)(r % p.h);
      const int seg = (int)min((long long)(p.h - y0), r_end - r);
      const int iy0 = max(y0 - 1, 0), iy1 = min(y0 + seg, p.h - 1);
      for (int iy = iy0; iy <= iy1; ++iy, ++g) {
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
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    const uint64_t wdesc0 = make_smem_desc_sw128(smem_u32(sW), 16, 1024);
    const uint64_t rdesc0 = make_smem_desc_sw128(smem_u32(sRing), 16, 1024);
    mbar_wait(w_full, 0);
    int g = 0;      // running input-row index (ring position)
    int obase = 0;  // output rows of the previous segments
    long long r = r_begin;
    while (r < r_end) {
      const int y0 = (int)(r % p.h);
      const int seg = (int)min((long long)(p.h - y0), r_end - r);
      const int iy0 = max(y0 - 1, 0), iy1 = min(y0 + seg, p.h - 1);
      for (int iy = iy0; iy <= iy1; ++iy, ++g) {
        // output rows fed by this input row, clipped to the segment; rows >= t_new are touched for the first time
        const int t_lo = max(iy - 1, y0), t_hi = min(iy + 1, y0 + seg - 1);
        const int t_new = (iy == iy0) ? t_lo : iy + 1;
        for (int t = max(t_new, t_lo); t <= t_hi; ++t) {   // wait until the epilogue has drained (and zeroed) the slot
          const int o = obase + (t - y0);
          mbar_wait(&acc_empty[o % kAccSlots], ((o / kAccSlots) & 1) ^ 1);
        }
        // Destination = rows [t_lo, t_hi] = 1..3 consecutive TMEM slots; part A up to the end of the ring, part B the
        // wrapped remainder.  Every MMA accumulates (drained slots are re-zeroed by the epilogue), so all 12 steps
        // of the row are identical; everything below is scalar and warp-uniform to keep the single issuing warp on
        // the uniform datapath (its instruction count per MMA is what bounds this kernel).
        const int slot = (obase + (t_lo - y0)) % kAccSlots;
        const int cnt = t_hi - t_lo + 1;
        const int cnt_a = min(cnt, kAccSlots - slot), cnt_b = cnt - cnt_a;
        const int j0 = t_lo - (iy - 1);
        const uint32_t d_a = tmem + slot * kC;
        const uint32_t id_a = make_idesc_bf16(128, kC * cnt_a, 0, 0);
        const uint32_t id_b = make_idesc_bf16(128, kC * max(cnt_b, 1), 0, 0);
        const uint64_t b_a = wdesc0 + (uint64_t)((j0 * kTapBytes) >> 4);
        const uint64_t b_b = wdesc0 + (uint64_t)(((j0 + cnt_a) * kTapBytes) >> 4);
        const int s = g % kSlots;
        mbar_wait(&full[s], (g / kSlots) & 1);
        tc_fence_after();
        const uint64_t rowd = rdesc0 + (uint64_t)((s * kSlotBytes) >> 4);
        if (leader) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = rowd + (uint64_t)(((kx * 128) >> 4) + 2 * k);
              const uint64_t bo = (uint64_t)(((kx * 3 * kTapBytes) >> 4) + 2 * k);
              umma_bf16(d_a, ad, b_a + bo, id_a, 1u);
              if (cnt_b) umma_bf16(tmem, ad, b_b + bo, id_b, 1u);
            }
          }
        }
        if (leader) {
          umma_commit(&empty[s]);                       // the input row is no longer needed
          // completed output rows: the one above this input row, plus this row itself at the bottom of the image
          if (iy - 1 >= y0) umma_commit(&acc_full[(obase + (iy - 1 - y0)) % kAccSlots]);
          if (iy == iy1 && iy1 == y0 + seg - 1) umma_commit(&acc_full[(obase + (iy - y0)) % kAccSlots]);
        }
        __syncwarp();
      }
      obase += seg;
      r += seg;
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    // TMEM -> registers -> (+bias, bf16) -> swizzled smem row -> statistics from the staged tile -> TMA store
    const int q = warp & 3;
    const int x = q * 32 + lane;              // pixel within the row (TMEM lane)
    const int et = (warp - 2) * 32 + lane;    // 0..127
    const int cp = et & 31, pg = et >> 5;     // statistics role: channel pair cp, pixel group pg
    const bool issuer = (warp == 2) && (lane == 0);
    const bool has_bias = p.bias != nullptr;
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    int orow = 0;
    for (long long r = r_begin; r < r_end; ++r, ++orow) {
      const int buf = orow & 1;
      const int slot = orow % kAccSlots;
      const int y = (int)(r % p.h), n = (int)(r / p.h);
      mbar_wait(&acc_full[slot], (orow / kAccSlots) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + slot * kC, v0);
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + slot * kC + 32, v1);
      tmem_ld_wait();
      tmem_st32_zero(tmem + ((uint32_t)(q * 32) << 16) + slot * kC);        // hand the slot back zeroed
      tmem_st32_zero(tmem + ((uint32_t)(q * 32) << 16) + slot * kC + 32);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[slot]);
      // the TMA store issued two rows ago must have finished reading this staging buffer
      if (issuer) tma_store_wait_read<1>();
      named_bar_sync(1, 128);
      uint8_t* stile = sOut + buf * kStageOutBytes;
      uint8_t* srow = stile + x * 128;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v0[j + e]) + (has_bias ? sBias[j + e] : 0.f);
        *reinterpret_cast<bf16x8*>(srow + (((j >> 3) ^ (x & 7)) << 4)) = pack8(f);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v1[j + e]) + (has_bias ? sBias[32 + j + e] : 0.f);
        *reinterpret_cast<bf16x8*>(srow + ((((32 + j) >> 3) ^ (x & 7)) << 4)) = pack8(f);
      }
      fence_proxy_async();
      named_bar_sync(1, 128);
      if (issuer) {
        tma_store_4d(&tmY, stile, 0, 0, y, n);
        tma_store_commit();
      }
      if (p.partial) {
        // per-channel sums over this row from the staged (rounded) values: thread = (channel pair, 32-pixel group);
        // a warp reads one 128-byte pixel row per step -> conflict free
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
          const int px = pg * 32 + i;
          const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(
              stile + px * 128 + ((((cp >> 2) ^ (px & 7)) << 4) | ((cp & 3) << 2)));
          const float2 f = __bfloat1622float2(v);
          s0 += f.x; q0 = fmaf(f.x, f.x, q0);
          s1 += f.y; q1 = fmaf(f.y, f.y, q1);
        }
        const bool seg_end = (y == p.h - 1) || (r + 1 == r_end);
        if (seg_end) {
          // fold the 4 pixel groups in fixed order and publish this CTA's partial of image n
          named_bar_sync(2, 128);   // previous use of sStat is over
          sStat[(pg * 2 + 0) * kC + 2 * cp] = s0;
          sStat[(pg * 2 + 0) * kC + 2 * cp + 1] = s1;
          sStat[(pg * 2 + 1) * kC + 2 * cp] = q0;
          sStat[(pg * 2 + 1) * kC + 2 * cp + 1] = q1;
          named_bar_sync(2, 128);
          const int b0 = first_cta_of_row((long long)n * p.h, p.total_rows, gridDim.x);
          const int part = blockIdx.x - b0;
          float* dst = p.partial + ((long long)n * p.parts + part) * 2 * kC;
          dst[et] = (sStat[et] + sStat[128 + et]) + (sStat[256 + et] + sStat[384 + et]);
          if (y == p.h - 1)   // last CTA of this image: the unused slots must read as zero
            for (int z = part + 1; z < p.parts; ++z) p.partial[((long long)n * p.parts + z) * 2 * kC + et] = 0.f;
          s0 = s1 = q0 = q1 = 0.f;
        }
      }
    }
    if (issuer) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

int encode(CUtensorMap* m, const void* ptr, int rank, const unsigned long long* dims, const unsigned long long* strides,
           const unsigned int* box, const char* what) {
  return crfr_tmap_encode_bf16(m, ptr, rank, dims, strides, box, what);
}

int sm_count() { return crfr_sm_count(); }

int grid_for(int total_rows) {
  const int sms = sm_count();
  return sms < total_rows ? sms : total_rows;
}

// partial slots per image: an image of h rows is covered by at most this many CTAs
int parts_for(int n, int h) {
  const int total = n * h, grid = grid_for(total);
  const int min_rows = total / grid;   // every CTA owns at least floor(total / grid) >= 1 rows
  return (h + min_rows - 1) / min_rows + 1;
}

}  // namespace

int crfr_rowconv_supported(int h, int w, int cin, int cout, int k, int stride, int pad) {
  return w == kW && cin == kC && cout == kC && k == 3 && stride == 1 && pad == 1 && h >= 1;
}

size_t crfr_rowconv_ws_bytes(int n, int h) {
  const size_t one = sizeof(float) * (size_t)n * parts_for(n, h) * 2 * kC + 256, two = crfr_rowconv_pair_ws_bytes(n, h);
  return one > two ? one : two;
}

// src/dst: NHWC bf16 [n][h][128][64] views; w_packed: [9][64][64] bf16 ([tap][n][k]); flip = 1 for dgrad.
// stats (optional, forward only): InstanceNorm (mean, rstd) [n][64][2] of the stored output; needs ws.
int crfr_rowconv(const void* src, int src_ld, int n, int h, const void* w_packed, int flip, const float* bias, void* dst,
                 int dst_ld, float* stats, float eps, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (crfr_opt(CRFR_OPT_ROWCONV_PAIR) && crfr_rowconv_pair_supported(n, h))   // CTA-pair kernel (rowconv2.cu)
    return crfr_rowconv_pair(src, src_ld, n, h, w_packed, flip, bias, dst, dst_ld, stats, eps, ws, ws_bytes, st);
  CRFR_CHECK_ARG(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 &&
                     (src_ld & 7) == 0 && (dst_ld & 7) == 0,
                 "rowconv: pointers must be 16B aligned and ld a multiple of 8");
  CUtensorMap tmX, tmW, tmY;
  {
    unsigned long long dims[4] = {(unsigned long long)kC, (unsigned long long)kW, (unsigned long long)h, (unsigned long long)n};
    unsigned long long strides[3] = {(unsigned long long)dst_ld * 2, (unsigned long long)kW * dst_ld * 2, (unsigned long long)h * kW * dst_ld * 2};
    unsigned int box[4] = {64, 128, 1, 1};
    CRFR_TRY(encode(&tmY, dst, 4, dims, strides, box, "output"));
  }
  {
    unsigned long long dims[4] = {(unsigned long long)kC, (unsigned long long)kW, (unsigned long long)h, (unsigned long long)n};
    unsigned long long strides[3] = {(unsigned long long)src_ld * 2, (unsigned long long)kW * src_ld * 2, (unsigned long long)h * kW * src_ld * 2};
    unsigned int box[4] = {64, 130, 1, 1};
    CRFR_TRY(encode(&tmX, src, 4, dims, strides, box, "activation"));
  }
  {
    unsigned long long dims[2] = {64, 9 * 64};
    unsigned long long strides[1] = {128};
    unsigned int box[2] = {64, 64};
    CRFR_TRY(encode(&tmW, w_packed, 2, dims, strides, box, "weights"));
  }
  static std::atomic<unsigned long long> attr_done{0};
  CRFR_CUDA((cudaError_t)crfr_smem_attr(rowconv_kernel, kSmemBytes, attr_done));
  RowParams p;
  p.n = n; p.h = h; p.total_rows = n * h; p.flip = flip;
  p.bias = bias;
  p.partial = nullptr;
  p.parts = 0;
  if (stats) {
    const size_t need = sizeof(float) * (size_t)n * parts_for(n, h) * 2 * kC + 256;
    if (!ws || ws_bytes < need) {
      crfr_set_error("rowconv: workspace %zu < %zu", ws_bytes, need);
      return CRFR_EWORKSPACE;
    }
    p.partial = (float*)ws;
    p.parts = parts_for(n, h);
  }
  rowconv_kernel<<<grid_for(p.total_rows), kThreads, kSmemBytes, st>>>(tmX, tmW, tmY, p);
  CRFR_COUNT_LAUNCH();
  CRFR_LAUNCH_CHECK();
  if (stats) CRFR_TRY(crfr_norm_finalize(p.partial, n, p.parts, h * kW, kC, eps, stats, st));
  return CRFR_OK;
}
