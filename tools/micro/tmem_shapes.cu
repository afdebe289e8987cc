// micro-benchmark: tcgen05.ld throughput per warp for different shapes (how fast can a softmax warp pull a
// 32-lane x 64-column fp32 block out of TMEM?), and whether loads overlap with FMA work issued behind them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/tmem_shapes tools/micro/tmem_shapes.cu mixermdm_b200/csrc/host_util.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../mixermdm_b200/csrc/common.cuh"
using namespace mmdm;

#define R4(r, o) "=r"(r[o]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3])
__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t* r) {   // 16 lanes x 64 columns -> 32 regs
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : R4(r, 0), R4(r, 4), R4(r, 8), R4(r, 12), R4(r, 16), R4(r, 20), R4(r, 24), R4(r, 28)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld_16x128b_x16(uint32_t taddr, uint32_t* r) {  // 16 lanes x 64 columns -> 32 regs
  asm volatile(
      "tcgen05.ld.sync.aligned.16x128b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : R4(r, 0), R4(r, 4), R4(r, 8), R4(r, 12), R4(r, 16), R4(r, 20), R4(r, 24), R4(r, 28)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld_32x32b_x64(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : R4(r, 0), R4(r, 4), R4(r, 8), R4(r, 12), R4(r, 16), R4(r, 20), R4(r, 24), R4(r, 28), R4(r, 32), R4(r, 36),
        R4(r, 40), R4(r, 44), R4(r, 48), R4(r, 52), R4(r, 56), R4(r, 60)
      : "r"(taddr)
      : "memory");
}

// MODE 0: 2 x 32x32b.x32   1: 32x32b.x64   2: 2 x 16x256b.x8 (lanes 0-15, 16-31)   3: 2 x 16x128b.x16
// MODE 4: 2 x 32x32b.x32 followed by 256 dependent-free FFMAs, then the wait (does the load overlap with math?)
// MODE 5: the 256 FFMAs alone
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  const uint32_t trow = base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  float acc = 0.f;
  float f[8];
  for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-3f + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[64];
    if (MODE == 0 || MODE == 4) {
      tmem_ld_32x32b_x32(trow, r);
      tmem_ld_32x32b_x32(trow + 32, r + 32);
    } else if (MODE == 1) {
      ld_32x32b_x64(trow, r);
    } else if (MODE == 2) {
      ld_16x256b_x8(trow, r);
      ld_16x256b_x8(trow + (16u << 16), r + 32);
    } else if (MODE == 3) {
      ld_16x128b_x16(trow, r);
      ld_16x128b_x16(trow + (16u << 16), r + 32);
    }
    if (MODE == 4 || MODE == 5) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], 1.0001f, 0.5f);
    }
    if (MODE != 5) {
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 64; ++i) acc += __uint_as_float(r[i]);
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  for (int i = 0; i < 8; ++i) acc += f[i];
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
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s warps %2d : %8.1f cycles per 32x64 block per warp  (%s)\n", name, warps, (double)h / iters, cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 8}) {
    run<0>("2 x tcgen05.ld.32x32b.x32", w);
    run<1>("tcgen05.ld.32x32b.x64", w);
    run<2>("2 x tcgen05.ld.16x256b.x8", w);
    run<3>("2 x tcgen05.ld.16x128b.x16", w);
    run<4>("2 x 32x32b.x32 + 256 FFMA before the wait", w);
    run<5>("256 FFMA alone", w);
  }
  return 0;
}
