// micro-benchmark: how fast can one SM stream rows from HBM into a shared-memory ring with cp.async.bulk (1-D,
// no tensor map)? Sweeps copy size / alignment, ring depth and how the producer warp issues the copies. The fused
// mix kernel reads 2096-byte rows (16-byte aligned only); this tells what the ring can deliver per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/bulk_ring tools/micro/bulk_ring.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../mixermdm_b200/csrc/common.cuh"
using namespace mmdm;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void st_vol(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_vol(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}

struct Args {
  const uint8_t* src;
  float* sink;
  int items_per_cta;   // items streamed by every CTA
  int copies;          // copies per item
  int copy_bytes;      // bytes per copy
  int pitch;           // distance between consecutive copies in global memory
  int nslots;
  int group;           // items per producer group (lane-parallel issue); 0 = one thread issues everything
  int consumers;       // consumer warps
  int touch;           // consumers read the slot (one float4 per lane per copy)
  int streams;         // 1: the copies of an item are consecutive in memory; 0: copy k comes from its own region
                       //    (k * region_bytes apart), like the out1a/out2a/x1/x2/g rows of the mix kernel
  size_t region_bytes;
  float* wdst;         // != null: consumers also write `wrows` rows of copy_bytes per item (coalesced float4)
  int wrows;
};

__global__ void __launch_bounds__(512, 1) ring_kernel(const Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 64;
  int* issued = reinterpret_cast<int*>(empty + 64);
  uint8_t* ring = smem + 2048;
  const int slot_bytes = (a.copies * a.copy_bytes + 127) / 128 * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nslots; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    *issued = 0;
    fence_barrier_init();
  }
  __syncthreads();
  const uint8_t* base = a.src + (size_t)blockIdx.x * a.items_per_cta * (a.streams ? a.copies : 1) * a.pitch;
  auto src_of = [&](int c, int k) {
    return a.streams ? base + ((size_t)c * a.copies + k) * a.pitch : base + (size_t)k * a.region_bytes + (size_t)c * a.pitch;
  };
  const int NS = a.nslots;
  if (warp == a.consumers) {
    if (a.group == 0) {
      if (lane == 0) {
        for (int c = 0; c < a.items_per_cta; ++c) {
          const int lap = c / NS, slot = c - lap * NS;
          if (lap > 0) mbar_wait(&empty[slot], (uint32_t)(lap - 1) & 1u);
          mbar_arrive_expect_tx(&full[slot], (uint32_t)(a.copies * a.copy_bytes));
          for (int k = 0; k < a.copies; ++k)
            bulk_g2s(smem_u32(ring + (size_t)slot * slot_bytes + k * a.copy_bytes), src_of(c, k),
                     (uint32_t)a.copy_bytes, smem_u32(&full[slot]));
          st_vol(issued, c + 1);
        }
      }
    } else {
      const int G = a.group;
      for (int g = 0; g * G < a.items_per_cta; ++g) {
        const int li = lane / a.copies, k = lane - li * a.copies;
        const int c = g * G + li;
        const bool valid = li < G && c < a.items_per_cta;
        int slot = 0;
        if (valid) {
          const int lap = c / NS;
          slot = c - lap * NS;
          if (k == 0) {
            if (lap > 0) mbar_wait(&empty[slot], (uint32_t)(lap - 1) & 1u);
            mbar_arrive_expect_tx(&full[slot], (uint32_t)(a.copies * a.copy_bytes));
          }
        }
        __syncwarp();
        if (valid)
          bulk_g2s(smem_u32(ring + (size_t)slot * slot_bytes + k * a.copy_bytes), src_of(c, k),
                   (uint32_t)a.copy_bytes, smem_u32(&full[slot]));
        __syncwarp();
        if (lane == 0) st_vol(issued, min((g + 1) * G, a.items_per_cta));
      }
    }
  } else if (warp < a.consumers) {
    float acc = 0.f;
    for (int c = warp; c < a.items_per_cta; c += a.consumers) {
      const int lap = c / NS, slot = c - lap * NS;
      if (lane == 0)
        while (ld_vol(issued) <= c) __nanosleep(32);
      __syncwarp();
      mbar_wait(&full[slot], (uint32_t)lap & 1u);
      if (a.touch) {
        const float4* p = reinterpret_cast<const float4*>(ring + (size_t)slot * slot_bytes);
        const int n4 = a.copies * a.copy_bytes / 16;
        for (int i = lane; i < n4; i += 32) {
          const float4 v = p[i];
          acc += v.x + v.y + v.z + v.w;
        }
      }
      if (a.wdst != nullptr) {
        const int n4 = a.copy_bytes / 16;
        for (int r = 0; r < a.wrows; ++r) {
          float4* w = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(a.wdst) + (size_t)r * a.region_bytes +
                                                ((size_t)blockIdx.x * a.items_per_cta + c) * a.pitch);
          for (int i = lane; i < n4; i += 32) w[i] = make_float4(acc, 1.f, 2.f, 3.f);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
    if (acc == 123.456f) a.sink[threadIdx.x] = acc;
  }
}

static void run(const char* name, Args a, int grid, uint8_t* flush, size_t flush_bytes) {
  const int slot_bytes = (a.copies * a.copy_bytes + 127) / 128 * 128;
  const size_t smem = 2048 + (size_t)a.nslots * slot_bytes;
  cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 4; ++it) {
    cudaMemsetAsync(flush, it, flush_bytes);
    cudaEventRecord(e0);
    ring_kernel<<<grid, (a.consumers + 1) * 32, smem>>>(a);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", name, cudaGetErrorString(e));
      exit(1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it > 0 && ms < best) best = ms;
  }
  const double bytes = (double)grid * a.items_per_cta * (a.copies + (a.wdst ? a.wrows : 0)) * a.copy_bytes;
  printf("%-58s %8.1f us  %7.1f GB/s  (%5.1f GB/s per SM, %d B in flight per SM)\n", name, best * 1e3, bytes / best / 1e6,
         bytes / best / 1e6 / grid, a.nslots * a.copies * a.copy_bytes);
}

int main() {
  const size_t total = (size_t)1 << 30;
  uint8_t *src, *flush;
  float* sink;
  cudaMalloc(&src, total);
  cudaMemset(src, 1, total);
  cudaMalloc(&flush, (size_t)256 << 20);
  cudaMalloc(&sink, 4096);
  const int grid = 148;
  float* wbuf;
  cudaMalloc(&wbuf, total);
  const size_t region = total / 8;   // 128 MB per stream
  auto mk = [&](int copies, int bytes, int pitch, int nslots, int group, int consumers, int touch, int streams = 1,
                int wrows = 0) {
    Args a;
    a.src = src; a.sink = sink; a.copies = copies; a.copy_bytes = bytes; a.pitch = pitch; a.nslots = nslots;
    a.group = group; a.consumers = consumers; a.touch = touch;
    a.streams = streams; a.region_bytes = region; a.wdst = wrows ? wbuf : nullptr; a.wrows = wrows;
    a.items_per_cta = streams ? (int)((size_t)4 << 20) / (copies * pitch) : (int)(region / 148 / pitch);
    if (!streams && a.items_per_cta > 400) a.items_per_cta = 400;
    return a;
  };
  char name[128];
  struct Cfg { int copies, bytes, pitch, nslots, group, consumers, touch; };
  const Cfg cfgs[] = {
      {2, 2096, 2096, 14, 0, 14, 0},  {2, 2096, 2096, 14, 8, 14, 0},  {2, 2096, 2096, 14, 14, 14, 0},
      {2, 2096, 2096, 14, 8, 14, 1},  {2, 2096, 2096, 28, 8, 14, 0},  {2, 2096, 2096, 48, 16, 14, 0},
      {2, 2048, 2048, 14, 8, 14, 0},  {2, 2048, 2048, 28, 8, 14, 0},  {1, 4192, 4192, 14, 8, 14, 0},
      {1, 4192, 4192, 28, 14, 14, 0}, {1, 4096, 4096, 14, 8, 14, 0},  {1, 4096, 4096, 28, 14, 14, 0},
      {1, 8192, 8192, 14, 8, 14, 0},  {1, 16384, 16384, 12, 8, 14, 0}, {4, 2096, 2096, 7, 7, 14, 0},
      {2, 2096, 2096, 14, 8, 4, 0},   {1, 2096, 2096, 28, 16, 14, 0}, {1, 2096, 2096, 56, 28, 14, 0},
  };
  for (const Cfg& c : cfgs) {
    snprintf(name, sizeof(name), "copies=%d bytes=%5d pitch=%5d slots=%2d group=%2d cons=%2d touch=%d", c.copies, c.bytes,
             c.pitch, c.nslots, c.group, c.consumers, c.touch);
    run(name, mk(c.copies, c.bytes, c.pitch, c.nslots, c.group, c.consumers, c.touch), grid, flush, (size_t)256 << 20);
  }
  // the mix kernel's access pattern: every item = 4 (or 7) rows of 2096 B from separate arrays, optionally with 3
  // rows written back per item; 132 or 148 CTAs
  printf("--- separate-array streams (mix-kernel pattern)\n");
  struct Cfg2 { int copies, nslots, group, wrows, grid; };
  const Cfg2 c2[] = {{4, 25, 8, 0, 148}, {4, 25, 8, 0, 132}, {4, 25, 8, 2, 132}, {7, 14, 4, 3, 132}, {4, 12, 8, 0, 132}};
  for (const Cfg2& c : c2) {
    snprintf(name, sizeof(name), "arrays=%d bytes=2096 slots=%2d group=%d written rows=%d grid=%d", c.copies, c.nslots, c.group,
             c.wrows, c.grid);
    run(name, mk(c.copies, 2096, 2096, c.nslots, c.group, 14, 1, 0, c.wrows), c.grid, flush, (size_t)256 << 20);
  }
  return 0;
}
