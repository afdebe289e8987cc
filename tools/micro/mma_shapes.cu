// micro-benchmark: how long do the attention kernel's tcgen05.mma streams take on the tensor pipe of one SM?
//   S  = Q K^T : 8 x SS-MMA  M=128 N=64  K=16 (A: Q tile in smem, B: 64-key block in smem), one accumulator
//   PV = P V   : 4 x TS-MMA  M=128 N=128 K=16 (A: packed bf16 P in TMEM, B: V^T block in smem), one accumulator
// issued by one elected thread per "warpgroup stream" exactly as attention.cu does, with a commit per group of MMAs.
// Operand contents are irrelevant (whatever is in shared memory / TMEM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/mma_shapes tools/micro/mma_shapes.cu mixermdm_b200/csrc/host_util.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../mixermdm_b200/csrc/common.cuh"
using namespace mmdm;

constexpr int HD = 128;
constexpr uint32_t Q_BYTES = 2 * 128 * 128, KB_BYTES = 2 * 64 * 128, VB_BYTES = HD * 128;

// MODE 0: S only   1: PV only   2: S then PV (one stream)   3: two streams (two issuing warps), S then PV each
// MODE 4: S with N=128 (a 128-key block), 8 MMAs   5: PV as SS (A = P from smem, 16 KB tile) 4 MMAs
// MODE 6: S as TS (A = Q from TMEM, 64 packed columns)   7: two streams, S as TS + PV
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;                    // 2 x Q tile
  uint8_t* Ks = Qs + 2 * Q_BYTES;        // 2 x 128-key block (MODE 4 reads two adjacent 64-key blocks as one)
  uint8_t* Vs = Ks + 4 * KB_BYTES;       // 2 x V^T block
  uint8_t* Ps = Vs + 2 * VB_BYTES;       // 2 x P tile in smem (MODE 5)
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = slot;
  constexpr bool two = MODE == 3 || MODE == 7;
  if (warp < (two ? 2 : 1)) {
    const int g = warp;
    if (elect_one()) {
      const uint32_t ids = umma_idesc_bf16(128, MODE == 4 ? 128 : 64);
      const uint32_t ido = umma_idesc_bf16(128, HD);
      const uint64_t dq0 = umma_desc_k_sw128(smem_u32(Qs) + g * Q_BYTES);
      const uint64_t dk0 = umma_desc_k_sw128(smem_u32(Ks) + g * 2 * KB_BYTES);
      const uint64_t dv0 = umma_desc_k_sw128(smem_u32(Vs) + g * VB_BYTES);
      const uint64_t dp0 = umma_desc_k_sw128(smem_u32(Ps) + g * 128 * 128);
      const uint32_t scol = tmem_base + (uint32_t)g * 128u;
      const uint32_t ocol = tmem_base + 256u + (uint32_t)g * HD;
      const uint32_t qcol = tmem_base + (uint32_t)g * 128u + 64u;   // MODE 6/7: Q in TMEM (packed), S in [0,64)
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 4) {
          uint64_t da = dq0, db = dk0;
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            umma_bf16_ss(scol + (uint32_t)(it & 1) * (MODE == 4 ? 0u : 64u), da, db, ids, kk > 0 ? 1u : 0u);
            const bool wrap = (kk & 3) == 3;
            da += wrap ? (uint64_t)(128 * 128 / 16 - 6) : 2u;
            db += wrap ? (uint64_t)((MODE == 4 ? 128 : 64) * 128 / 16 - 6) : 2u;
          }
        }
        if (MODE == 6 || MODE == 7) {
          uint64_t db = dk0;
          uint32_t qa = qcol;
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            umma_bf16_ts(scol, qa, db, ids, kk > 0 ? 1u : 0u);
            qa += 8u;
            db += ((kk & 3) == 3) ? (uint64_t)(64 * 128 / 16 - 6) : 2u;
          }
        }
        if (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 7) {
          uint32_t pa = scol + (uint32_t)(it & 1) * 64u;
          if (MODE == 7) pa = qcol;   // (read-only columns: no dependency on the S accumulator)
          uint64_t db = dv0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            umma_bf16_ts(ocol, pa, db, ido, (it | i) != 0 ? 1u : 0u);
            pa += 8u;
            db += 2u;
          }
        }
        if (MODE == 5) {
          uint64_t da = dp0, db = dv0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            umma_bf16_ss(ocol, da, db, ido, (it | i) != 0 ? 1u : 0u);
            da += 2u;
            db += 2u;
          }
        }
        if ((it & 7) == 7) {   // bound the queue depth like the kernel's barriers do: wait for every 8th group
          umma_commit(&bar[g]);
          mbar_wait(&bar[g], (it >> 3) & 1);
        }
      }
      const long long t1 = clock64();
      cyc[g] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

template <int MODE>
void run(const char* name, int mmas_per_iter, double flop_per_iter) {
  const int iters = 512;   // multiple of 8
  long long* cyc;
  cudaMalloc(&cyc, 16);
  cudaMemset(cyc, 0, 16);
  const size_t smem = 2 * Q_BYTES + 4 * KB_BYTES + 2 * VB_BYTES + 2 * 128 * 128 + 1024;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<MODE><<<1, 128, smem>>>(iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
  const double c = (double)(h[0] > h[1] ? h[0] : h[1]) / iters;
  printf("%-58s %8.1f cycles/iter  %6.1f cycles/MMA  %6.0f FLOP/clk (%s)\n", name, c, c / mmas_per_iter,
         flop_per_iter / c, cudaGetErrorString(e));
  cudaFree(cyc);
}

int main() {
  const double fs = 2.0 * 128 * 64 * 128, fpv = 2.0 * 128 * 128 * 64;
  run<0>("S: 8 x SS M128 N64 K16", 8, fs);
  run<1>("PV: 4 x TS M128 N128 K16", 4, fpv);
  run<2>("S then PV, one stream", 12, fs + fpv);
  run<3>("S then PV, two streams (per-stream iteration)", 24, 2 * (fs + fpv));
  run<4>("S over a 128-key block: 8 x SS M128 N128 K16", 8, 2 * fs);
  run<5>("PV with P from shared memory: 4 x SS M128 N128 K16", 4, fpv);
  run<6>("S with Q from TMEM: 8 x TS M128 N64 K16", 8, fs);
  run<7>("TS-S then PV, two streams", 24, 2 * (fs + fpv));
  return 0;
}
