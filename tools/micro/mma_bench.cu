// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, bf16 in, fp32 accumulate) on sm_100a for the tile shapes the
// convolution kernels use.  One CTA per SM; one elected thread issues `iters` MMAs back to back (operands are
// zero-filled SWIZZLE_128B K-major tiles in shared memory, or A in TMEM), then commits and waits.
// Prints cycles per MMA (clock64 on SM 0's CTA) and the implied dense TFLOP/s at the measured wall time.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cross-resolution-face-recognition_b200/csrc
//        tools/micro/mma_bench.cu -o build_tmp/mma_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "sm100.cuh"
using namespace sm100;

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct Res { long long cycles; };

// mode 0: SS, one accumulator; mode 1: SS, two accumulators alternating; mode 2: TS (A in TMEM), one accumulator
// a_tiles: number of distinct 16 KB A tiles cycled through (1 = always the same tile)
template <int N>
__global__ void __launch_bounds__(128, 1) mma_kernel(int iters, int mode, int a_tiles, Res* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  // 8 A tiles of 16 KB then B tile(s) of N*128 bytes
  const int a_bytes = 8 * 16384;
  for (int i = threadIdx.x; i < (a_bytes + N * 128) / 16; i += blockDim.x) ((uint4*)base)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(base), 16, 1024);
    const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(base + a_bytes), 16, 1024);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int tile = it % a_tiles;
      const uint64_t ad = adesc0 + (uint64_t)((tile * 16384) >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t d = tmem + ((mode == 1) ? ((it & 1) * N) : 0);
        if (leader) {
          if (mode == 2) umma_bf16_ts(d, tmem + 256 + k * 8, bdesc0 + 2 * k, idesc, 1u);
          else umma_bf16(d, ad + 2 * k, bdesc0 + 2 * k, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (leader) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (leader && blockIdx.x == 0) out->cycles = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int N>
void run(int iters, int mode, int a_tiles, Res* d_res) {
  const int smem = 8 * 16384 + N * 128 + 2048;
  cudaFuncSetAttribute(mma_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  mma_kernel<N><<<148, 128, smem>>>(iters, mode, a_tiles, d_res);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  mma_kernel<N><<<148, 128, smem>>>(iters, mode, a_tiles, d_res);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  Res r;
  cudaMemcpy(&r, d_res, sizeof(r), cudaMemcpyDeviceToHost);
  const double flops = 2.0 * 128 * N * 16 * 4.0 * iters * 148;
  printf("N=%3d mode=%d a_tiles=%d: %7.1f cycles/MMA  %8.1f TFLOP/s (wall %.3f ms) %s\n", N, mode, a_tiles,
         (double)r.cycles / (4.0 * iters), flops / (ms * 1e-3) / 1e12, ms, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

// Replica of the rowconv issue pattern: per "row" 12 MMAs of N=192 (3 kx shifts x 4 k-steps) from a 5-slot A ring and
// three 24 KB B blocks into a rotating 192-column window of TMEM.  variant bit0: apply the +128 B kx shift to A;
// bit1: four extra warps hammer shared memory with 16-byte stores + loads (epilogue / TMA traffic stand-in);
// bit2: the four extra warps read TMEM (tcgen05.ld) in a loop instead.
__global__ void __launch_bounds__(256, 1) rowpattern_kernel(int rows, int variant, Res* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int ring_bytes = 5 * 17408, w_bytes = 3 * 24576, extra = 32768;
  for (int i = threadIdx.x; i < (ring_bytes + w_bytes + extra) / 16; i += blockDim.x) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (variant & 8) {   // pseudo-random bf16 values in (-2, 2): realistic switching activity (power)
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      uint32_t w[4];
      for (int j = 0; j < 4; ++j) {
        h = h * 1664525u + 1013904223u;
        const uint32_t lo = 0x3f00u | ((h >> 9) & 0x80ffu), hi = 0x3f00u | ((h >> 17) & 0x80ffu);
        w[j] = lo | (hi << 16);
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    ((uint4*)base)[i] = v;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, 192, 0, 0);
    const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(base), 16, 1024);
    const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(base + ring_bytes), 16, 1024);
    long long t0 = clock64();
    for (int it = 0; it < rows; ++it) {
      const uint64_t rowd = adesc0 + (uint64_t)(((it % 5) * 17408) >> 4);
      const uint32_t d = tmem + (it % 6) * 64;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = rowd + (uint64_t)((((variant & 1) ? kx * 128 : 0) >> 4) + 2 * k);
          const uint64_t bd = bdesc0 + (uint64_t)(((kx * 24576) >> 4) + 2 * k);
          if (leader) umma_bf16(d, ad, bd, idesc, 1u);
        }
      __syncwarp();
    }
    if (leader) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (leader && blockIdx.x == 0) out->cycles = t1 - t0;
    stop = 1;
  } else if (warp >= 4) {
    uint8_t* scratch = base + ring_bytes + w_bytes;
    uint4 v = make_uint4(lane, 1, 2, 3);
    uint32_t sink = 0;
    if (variant & 2) {
      while (!stop) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          *reinterpret_cast<uint4*>(scratch + ((warp - 4) * 8192) + ((j * 32 + lane) * 16)) = v;
          v.x += reinterpret_cast<uint4*>(scratch + ((warp - 4) * 8192) + (((7 - j) * 32 + lane) * 16))->y;
        }
      }
    } else if (variant & 4) {
      while (!stop) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 448, r);
        tmem_ld_wait();
        sink += r[lane & 31];
      }
    }
    if (sink == 0x12345678 || v.x == 0x9abcdef) out->cycles = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

void run_pattern(int rows, int variant, Res* d_res) {
  const int smem = 5 * 17408 + 3 * 24576 + 32768 + 2048;
  cudaFuncSetAttribute(rowpattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rowpattern_kernel<<<148, 256, smem>>>(rows, variant, d_res);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  rowpattern_kernel<<<148, 256, smem>>>(rows, variant, d_res);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  Res r;
  cudaMemcpy(&r, d_res, sizeof(r), cudaMemcpyDeviceToHost);
  printf("rowpattern variant=%d: %7.1f cycles/MMA  %7.1f cycles/row (wall %.3f ms) %s\n", variant,
         (double)r.cycles / (12.0 * rows), (double)r.cycles / rows, ms, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

// ---- cta_group::2: a CTA pair issues M = 256 (128 rows of A per CTA), N columns, with each CTA holding N/2 rows of B ----
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// a_tiles distinct 16 KB A tiles per CTA; B half of N/2 rows per CTA.  Only rank 0 issues; the commit is multicast to the
// barrier of both CTAs.  Cycles per MMA are per PAIR-MMA (2 x 128 x N x 16 MACs x 2).
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma2_kernel(int iters, int a_tiles, Res* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int a_bytes = 8 * 16384;
  for (int i = threadIdx.x; i < (a_bytes + (N / 2) * 128) / 16; i += blockDim.x) ((uint4*)base)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync();                       // both CTAs' operands and barriers are ready before the first pair MMA
  const uint32_t tmem = tmem_slot;
  const uint32_t rank = cluster_rank();
  if (warp == 0) {
    const bool leader = elect_one();
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(256, N, 0, 0);
      const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(base), 16, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(base + a_bytes), 16, 1024);
      long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint64_t ad = adesc0 + (uint64_t)(((it % a_tiles) * 16384) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (leader) umma2_bf16(tmem, ad + 2 * k, bdesc0 + 2 * k, idesc, 1u);
        __syncwarp();
      }
      if (leader)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
      __syncwarp();
      mbar_wait(&bar, 0);
      long long t1 = clock64();
      if (leader && blockIdx.x == 0) out->cycles = t1 - t0;
    } else {
      mbar_wait(&bar, 0);               // the multicast commit arrives here too
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                       // nobody frees TMEM while the peer may still be inside the MMAs
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

template <int N>
void run2(int iters, int a_tiles, Res* d_res) {
  const int smem = 8 * 16384 + (N / 2) * 128 + 2048;
  cudaFuncSetAttribute(mma2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  mma2_kernel<N><<<148, 128, smem>>>(iters, a_tiles, d_res);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("cta_group::2 N=%d: %s\n", N, cudaGetErrorString(err)); return; }
  cudaEventRecord(e0);
  mma2_kernel<N><<<148, 128, smem>>>(iters, a_tiles, d_res);
  cudaEventRecord(e1);
  err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  Res r;
  cudaMemcpy(&r, d_res, sizeof(r), cudaMemcpyDeviceToHost);
  const double flops = 2.0 * 256 * N * 16 * 4.0 * iters * 74;
  printf("cta_group::2 M=256 N=%3d a_tiles=%d: %7.1f cycles/pair-MMA  %8.1f TFLOP/s (wall %.3f ms) %s\n", N, a_tiles,
         (double)r.cycles / (4.0 * iters), flops / (ms * 1e-3) / 1e12, ms, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

// ---- replica of the rowconv PAIR issue pattern (rowconv2.cu): per row 12 pair-MMAs of N = 192 (3 kx x 4 k-steps), A from
// a 4-slot ring of 17 KB rows, B = the per-rank 96-row half of three 40 KB kx blocks, D rotating over 8 x 64 columns.
// variant bit0: +128 B kx shift on A; bit1: four extra warps hammer shared memory; bit2: B fixed (kx block 0 only);
// bit3: random operand data; bit4: D fixed
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) rowpattern2_kernel(int rows, int variant, Res* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int ring_bytes = 4 * 17408, w_bytes = 3 * 40960, extra = 32768;
  for (int i = threadIdx.x; i < (ring_bytes + w_bytes + extra) / 16; i += blockDim.x) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (variant & 8) {
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      uint32_t w[4];
      for (int j = 0; j < 4; ++j) {
        h = h * 1664525u + 1013904223u;
        const uint32_t lo = 0x3f00u | ((h >> 9) & 0x80ffu), hi = 0x3f00u | ((h >> 17) & 0x80ffu);
        w[j] = lo | (hi << 16);
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    ((uint4*)base)[i] = v;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t rank = cluster_rank();
  if (warp == 0) {
    const bool leader = elect_one();
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(256, 192, 0, 0);
      const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(base), 16, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(base + ring_bytes), 16, 1024);
      long long t0 = clock64();
      for (int it = 0; it < rows; ++it) {
        const uint64_t rowd = adesc0 + (uint64_t)(((it % 4) * 17408) >> 4);
        const uint32_t d = tmem + ((variant & 16) ? 0 : (it % 6) * 64);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = rowd + (uint64_t)((((variant & 1) ? kx * 128 : 0) >> 4) + 2 * k);
            const uint64_t bd = bdesc0 + (uint64_t)(((((variant & 4) ? 0 : kx) * 40960) >> 4) + 2 * k);
            if (leader) umma2_bf16(d, ad, bd, idesc, 1u);
          }
        __syncwarp();
      }
      if (leader)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
      __syncwarp();
      mbar_wait(&bar, 0);
      long long t1 = clock64();
      if (leader && blockIdx.x == 0) out->cycles = t1 - t0;
    } else {
      mbar_wait(&bar, 0);
    }
    stop = 1;
  } else if (warp >= 4) {
    uint8_t* scratch = base + ring_bytes + w_bytes;
    uint4 v = make_uint4(lane, 1, 2, 3);
    if (variant & 2) {
      while (!stop) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          *reinterpret_cast<uint4*>(scratch + ((warp - 4) * 8192) + ((j * 32 + lane) * 16)) = v;
          v.x += reinterpret_cast<uint4*>(scratch + ((warp - 4) * 8192) + (((7 - j) * 32 + lane) * 16))->y;
        }
      }
    }
    if (v.x == 0x9abcdef) out->cycles = 0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

void run_pattern2(int rows, int variant, Res* d_res) {
  const int smem = 4 * 17408 + 3 * 40960 + 32768 + 2048;
  cudaFuncSetAttribute(rowpattern2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rowpattern2_kernel<<<148, 256, smem>>>(rows, variant, d_res);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("rowpattern2 variant=%d: %s\n", variant, cudaGetErrorString(err)); return; }
  cudaEventRecord(e0);
  rowpattern2_kernel<<<148, 256, smem>>>(rows, variant, d_res);
  cudaEventRecord(e1);
  err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  Res r;
  cudaMemcpy(&r, d_res, sizeof(r), cudaMemcpyDeviceToHost);
  printf("rowpattern2 (pair) variant=%2d: %7.1f cycles/pair-MMA  %7.1f cycles/row  %7.1f TFLOP/s (wall %.3f ms) %s\n", variant,
         (double)r.cycles / (12.0 * rows), (double)r.cycles / rows, 2.0 * 256 * 192 * 16 * 12.0 * rows * 74 / (ms * 1e-3) / 1e12,
         ms, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main(int argc, char** argv) {
  Res* d_res;
  cudaMalloc(&d_res, sizeof(Res));
  const int iters = 4000;
  if (argc > 1 && atoi(argv[1]) == 3) {   // pair row pattern: what costs what
    for (int v : {0, 1, 4, 5, 16, 17, 8, 9, 2, 3, 11}) run_pattern2(4000, v, d_res);
    for (int v : {0, 1, 8, 9}) run_pattern(4000, v, d_res);
    return 0;
  }
  if (argc > 1 && atoi(argv[1]) == 2) {   // cta_group::2 issue rate next to the single-CTA SS figures
    run<128>(iters, 0, 8, d_res);
    run<192>(iters, 0, 8, d_res);
    run<256>(iters, 0, 8, d_res);
    run2<128>(iters, 8, d_res);
    run2<192>(iters, 8, d_res);
    run2<256>(iters, 8, d_res);
    return 0;
  }
  for (int mode = 0; mode < 0; ++mode) {
    for (int a_tiles : {8}) {
      run<64>(iters, mode, a_tiles, d_res);
      run<96>(iters, mode, a_tiles, d_res);
      run<128>(iters, mode, a_tiles, d_res);
      run<160>(iters, mode, a_tiles, d_res);
      run<192>(iters, mode, a_tiles, d_res);
      run<256>(iters, mode, a_tiles, d_res);
    }
  }
  for (int v : {0, 1, 8, 9}) run_pattern(2000, v, d_res);
  for (int v : {8, 9}) run_pattern(20000, v, d_res);
  return 0;
}
