// micro-benchmark: the per-tile softmax chain of the attention kernel in isolation, with 8 or 16 softmax warps.
// The attention kernel spends ~7700 cycles per 128-row query tile although its MMAs need ~2400 and its exponentials
// ~2400 (MUFU floor): the tile time is the latency of ONE thread's serial work on its share of the 304 logits
// (TMEM load -> row max -> exchange -> exp2 -> sum -> bf16 pack -> TMEM store). This measures that chain alone
// - no MMA, no TMA, logits are whatever the TMEM holds after a zero fill - for
//   NW = 8 : thread = (query row, key half),    152 logits per thread   (what attention.cu does today)
//   NW = 16: thread = (query row, key quarter),  76 logits per thread   (plan (1))
// to tell whether halving the per-thread chain halves the tile time before the kernel is rewritten.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/softmax_chain tools/micro/softmax_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../mixermdm_b200/csrc/common.cuh"
using namespace mmdm;

constexpr int NKEYS = 304;   // keys of a tile (T = 300 + zero-attention key, rounded to 16)

// VAR bits (which part of the chain is left out, to see what the ~3500 cycles are made of):
//   1: no row-max pass / exchange   2: no TMEM store of P   4: no bf16 pack   8: every fourth exp2 as an FMA polynomial
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(0.05517132f, f, 0.24261054f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992812f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int NW, int VAR>
__global__ void __launch_bounds__(NW * 32 + 32, 1) chain_kernel(long long* cycles, float* sink, int tiles, float scale_log2) {
  constexpr int PARTS = NW / 4;            // warps per TMEM lane quarter = key parts
  constexpr int KPT = NKEYS / PARTS;       // logits per thread: 152 or 76
  constexpr int CH = (KPT + 31) / 32;      // 32-column chunks per thread: 5 or 3 (the last one partly beyond the part)
  __shared__ uint32_t tmem_slot;
  __shared__ float xmax[PARTS][128], xsum[PARTS][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == NW) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp < NW) {
    const int q = warp & 3, part = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t scol = trow + (uint32_t)(part * KPT);            // this thread's logits
    const uint32_t pcol = trow + 320u + (uint32_t)(part * (KPT / 2));   // its packed probabilities (columns 320..471)
    // zero fill of the logits so that exp2 sees sane inputs
    {
      uint32_t z[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = 0u;
      for (int c = 0; c < KPT; c += 16) tmem_st_32x32b_x16(scol + (uint32_t)c, z);
      tmem_st_wait();
    }
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
    tc_fence_after();
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < tiles; ++it) {
      uint32_t sv[CH][32];
#pragma unroll
      for (int c = 0; c < CH; ++c) tmem_ld_32x32b_x32(scol + (uint32_t)(c * 32), sv[c]);
      tmem_ld_wait();
      // row max over my part (independent chains), exchanged between the parts of the row through shared memory
      float mx = 0.f;
      if (!(VAR & 1)) {
        float m4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c * 32 + 2 * i < KPT)
              m4[i & 3] = fmax3(m4[i & 3], __uint_as_float(sv[c][2 * i]), __uint_as_float(sv[c][2 * i + 1]));
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        xmax[part][row] = mx;
        asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(PARTS * 32) : "memory");
#pragma unroll
        for (int p = 0; p < PARTS; ++p) mx = fmaxf(mx, xmax[p][row]);
      }
      const float moff = mx * scale_log2;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (c * 32 + 2 * i < KPT) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(sv[c][2 * i]), scale_log2, -moff));
            const float x1 = fmaf(__uint_as_float(sv[c][2 * i + 1]), scale_log2, -moff);
            const float e1 = ((VAR & 8) && (i & 1)) ? ex2_poly(x1) : ex2_approx(x1);
            s0 += e0;
            s1 += e1;
            pk[i] = (VAR & 4) ? __float_as_uint(e0 + e1) : pack_bf16x2(e0, e1);
          } else {
            pk[i] = 0u;
          }
        }
        if (VAR & 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) acc += __uint_as_float(pk[i]) * 1e-30f;
        } else if (c * 32 + 32 <= KPT) {
          tmem_st_32x32b_x16(pcol + (uint32_t)(c * 16), pk);
        } else {
          tmem_st_32x32b_x8(pcol + (uint32_t)(c * 16), pk);   // (the tail chunk: 24 / 12 keys -> 8 columns cover it)
        }
      }
      if (!(VAR & 2)) tmem_st_wait();
      xsum[part][row] = s0 + s1;
      asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(PARTS * 32) : "memory");
      float tot = 0.f;
#pragma unroll
      for (int p = 0; p < PARTS; ++p) tot += xsum[p][row];
      acc += 1.0f / tot;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 12345.f) sink[threadIdx.x] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NW) tmem_dealloc<512>(tmem_base);
}

template <int NW, int VAR>
static void run(const char* name) {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaMalloc(&sink, 4096);
  const int tiles = 2000;
  chain_kernel<NW, VAR><<<148, NW * 32 + 32>>>(cyc, sink, tiles, 0.1275f);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: %s\n", name, cudaGetErrorString(e));
    return;
  }
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < 148; ++i) mean += (double)h[i] / tiles;
  printf("%-44s %8.0f cycles per tile\n", name, mean / 148);
}

int main() {
  printf("softmax chain of one 128 x 304 tile, no MMA (MUFU floor 2432 cycles; the attention kernel needs ~7700 per tile)\n");
  run<8, 0>("8 softmax warps (152 keys per thread)");
  run<16, 0>("16 softmax warps (76 keys per thread)");
  run<8, 1>("8 warps, no max pass / exchange");
  run<8, 2>("8 warps, no TMEM store of P");
  run<8, 6>("8 warps, no pack, no TMEM store");
  run<8, 7>("8 warps, TMEM load + exp2 + sum only");
  run<8, 8>("8 warps, every 4th exp2 as FMA polynomial");
  run<8, 9>("8 warps, polynomial, no max pass");
  return 0;
}
