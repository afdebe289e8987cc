// micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM on sm_100a (how fast can softmax / epilogue warps
// move fp32 accumulators between TMEM and registers?)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../mixermdm_b200/csrc/common.cuh"
using namespace mmdm;

// MODE 0: ld x32, wait after each;  1: 5 x ld x32 then one wait;  2: st x16 + wait;  3: ld x32 + st x16 interleaved
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  const uint32_t trow = base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t half = (warp >> 2) * 160;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(trow + half + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
      }
    } else if (MODE == 1) {
      uint32_t r[5][32];
#pragma unroll
      for (int c = 0; c < 5; ++c) tmem_ld_32x32b_x32(trow + half + c * 32, r[c]);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 5; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[c][i]);
    } else if (MODE == 2) {
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(acc + i);
        tmem_st_32x32b_x16(trow + half + c * 16, r);
      }
      tmem_st_wait();
      acc += 1.f;
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(base);
}

template <int MODE>
void run(const char* name, int warps) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<MODE><<<148, warps * 32>>>(out, 10, cyc);
  k<MODE><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double bytes = (MODE == 2 ? 5.0 * 16 : 5.0 * 32) * 32 * 4 * warps * (double)iters;
  printf("%-34s warps %2d : %8.1f cycles/iter  %7.1f B/clk/SM  (%s)\n", name, warps, (double)h / iters, bytes / h,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 8, 16}) {
    run<0>("tcgen05.ld x32, wait each", w);
    run<1>("tcgen05.ld 5 x x32, one wait", w);
    run<2>("tcgen05.st 5 x x16, one wait", w);
  }
  return 0;
}
