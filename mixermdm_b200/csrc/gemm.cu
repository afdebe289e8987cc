// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//   warp 0 : TMA producer (one elected lane)      A/B tiles -> 128B-swizzled smem ring
//   warp 1 : MMA issuer   (one elected lane)      tcgen05.mma 128 x BN x 16, fp32 accumulators in TMEM
//   warp 2 : TMEM allocator / deallocator
//   warps 4-7 : epilogue  (tcgen05.ld -> registers -> fused epilogue -> global), TMEM double buffered so
//               the epilogue of tile i overlaps the MMAs of tile i+1.
// Replaces every nn.Linear on the hot path of the reference (models/utils/layers.py:11,33,74,95-96,112;
// models/in2in.py:389; models/mixermdm.py:652) with the listed fused epilogues.
#include "gemm.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace mmdm {

static int g_debug_simt = -1;
void gemm_set_debug_simt(int on) { g_debug_simt = on; }
static bool use_simt() {
  if (g_debug_simt < 0) {
    const char* e = getenv("MMDM_GEMM_SIMT");
    g_debug_simt = (e && e[0] == '1') ? 1 : 0;
  }
  return g_debug_simt == 1;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ------------------------------------------------------------------------------------------------
// fused epilogue: one thread owns row `row`, 32 consecutive columns starting at `col0`
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_store(const GemmEpi& ep, int row, int col0, const float* v, int M,
                                               int N) {
  if (row >= M || col0 >= N) return;
  const int nvalid = min(32, N - col0);
  float b[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) b[j] = (ep.bias != nullptr && j < nvalid) ? __ldg(ep.bias + col0 + j) : 0.f;

  switch (ep.mode) {
    case EPI_BF16:
    case EPI_GELU_BF16: {
      float o[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = v[j] + b[j];
        o[j] = (ep.mode == EPI_GELU_BF16) ? gelu_erf(x) : x;
      }
      if (ep.mode == EPI_BF16 && ep.vt_col0 >= 0 && col0 >= ep.vt_col0) {
        const int s = row / ep.T, t = row - s * ep.T;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < nvalid) {
            const int n = col0 + j - ep.vt_col0;
            const int h = n / ep.hd, d = n - h * ep.hd;
            ep.vt[((size_t)(s * ep.H + h) * ep.hd + d) * ep.Tp + t] = __float2bfloat16_rn(o[j]);
          }
        }
      } else {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ldo + col0;
        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 pk;
            pk.x = pack_bf16x2(o[j], o[j + 1]);
            pk.y = pack_bf16x2(o[j + 2], o[j + 3]);
            pk.z = pack_bf16x2(o[j + 4], o[j + 5]);
            pk.w = pack_bf16x2(o[j + 6], o[j + 7]);
            *reinterpret_cast<uint4*>(dst + j) = pk;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) dst[j] = __float2bfloat16_rn(o[j]);
        }
      }
      break;
    }
    case EPI_RESID: {
      const float* rs = ep.resid + (size_t)row * ep.ldo + col0;
      float* dst = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo + col0;
      float o[32];
      if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) &&
          ((reinterpret_cast<uintptr_t>(rs) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 r4 = *reinterpret_cast<const float4*>(rs + j);
          o[j] = r4.x + v[j] + b[j];
          o[j + 1] = r4.y + v[j + 1] + b[j + 1];
          o[j + 2] = r4.z + v[j + 2] + b[j + 2];
          o[j + 3] = r4.w + v[j + 3] + b[j + 3];
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) {
            o[j] = rs[j] + v[j] + b[j];
            dst[j] = o[j];
          }
      }
      if (ep.out_bf16 != nullptr) {
        __nv_bfloat16* d2 = ep.out_bf16 + (size_t)row * ep.ldo + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) d2[j] = __float2bfloat16_rn(o[j]);
      }
      break;
    }
    case EPI_EMBED_PE: {
      const int t = row % ep.T;
      const float* pe = ep.pe + (size_t)t * ep.pe_ld + col0;
      float* dst = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo + col0;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) {
          const float o = v[j] + b[j] + __ldg(pe + j);
          dst[j] = o;
          if (ep.dup_row_offset != 0) dst[ep.dup_row_offset * ep.ldo + j] = o;
        }
      break;
    }
    case EPI_FINAL: {
      const int s = row / ep.T, t = row - s * ep.T;
      float* dst;
      if (ep.persons_per_row == 2)
        dst = reinterpret_cast<float*>(ep.out) + ((size_t)(s >> 1) * ep.T + t) * (2 * ep.nfeats) +
              (s & 1) * ep.nfeats + col0;
      else
        dst = reinterpret_cast<float*>(ep.out) + ((size_t)s * ep.T + t) * ep.nfeats + col0;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) {
          const float y = v[j] + b[j];
          dst[j] = ep.std != nullptr ? y * __ldg(ep.std + col0 + j) + __ldg(ep.mean + col0 + j) : y;
        }
      break;
    }
    case EPI_F32:
    default: {
      float* dst = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo + col0;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) dst[j] = v[j] + b[j];
      break;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------------
template <int BN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BN;  // double-buffered fp32 accumulator (512 or 256)
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmEpi ep, int M, int N, int K) {
  using C = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + C::BM - 1) / C::BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + C::BK - 1) / C::BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * C::BM;
        const int n0 = (tile % num_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * C::BK, m0);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * C::BK, n0);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(C::BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * C::STAGE_BYTES);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < C::BK / 16; ++k)
            umma_bf16_ss(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * C::BM;
      const int n0 = (tile % num_n) * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + (uint32_t)(acc * BN + c * 32) + ((uint32_t)(q * 32) << 16);
        tmem_ld_32x32b_x32(taddr, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        epilogue_store(ep, row, n0 + c * 32, v, M, N);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// SIMT bring-up kernel (same epilogue); used only when MMDM_GEMM_SIMT=1 to bisect tcgen05 problems.
// ------------------------------------------------------------------------------------------------
__global__ void gemm_simt_debug_kernel(const __nv_bfloat16* __restrict__ A, int lda,
                                       const __nv_bfloat16* __restrict__ W, int ldw, const GemmEpi ep, int M,
                                       int N, int K) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int col0 = blockIdx.y * 32;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  if (row < M) {
    for (int k = 0; k < K; ++k) {
      const float a = __bfloat162float(A[(size_t)row * lda + k]);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < N) acc[j] += a * __bfloat162float(W[(size_t)(col0 + j) * ldw + k]);
    }
  }
  epilogue_store(ep, row, col0, acc, M, N);
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <int BN>
static int launch_tc(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
                     const GemmEpi& ep, cudaStream_t stream) {
  using C = GemmCfg<BN>;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {(uint32_t)C::BK, (uint32_t)C::BM};
    MMDM_TRY(make_tensor_map(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {(uint32_t)C::BK, (uint32_t)BN};
    MMDM_TRY(make_tensor_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B));
  }
  const int tiles = ((M + C::BM - 1) / C::BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_bf16_tn_kernel<BN><<<grid, 256, C::SMEM_BYTES, stream>>>(tmA, tmB, ep, M, N, K);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int gemm_init() {
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)GemmCfg<256>::SMEM_BYTES));
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)GemmCfg<128>::SMEM_BYTES));
  return 0;
}

int launch_gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
                const GemmEpi& ep, cudaStream_t stream) {
  MMDM_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  static int inited = gemm_init();
  MMDM_REQUIRE(inited == 0, "gemm: kernel attribute setup failed");
  MMDM_REQUIRE((lda % 8) == 0 && (ldw % 8) == 0, "gemm: lda/ldw must be multiples of 8 (got %d, %d)", lda, ldw);
  MMDM_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
               "gemm: operands must be 16-byte aligned");
  if (use_simt()) {
    dim3 grid((M + 127) / 128, (N + 31) / 32);
    gemm_simt_debug_kernel<<<grid, 128, 0, stream>>>(A, lda, W, ldw, ep, M, N, K);
    MMDM_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  // 128x256 tiles when N fills them; 128x128 for narrow / ragged N (e.g. the 262-wide output layer)
  if (N % 256 == 0 || N >= 1024) return launch_tc<256>(A, lda, W, ldw, M, N, K, ep, stream);
  return launch_tc<128>(A, lda, W, ldw, M, N, K, ep, stream);
}

}  // namespace mmdm
