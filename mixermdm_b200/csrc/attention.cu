// Fused attention for short motion sequences (T <= 319) on sm_100a, with the reference's zero-attention
// sink: nn.MultiheadAttention(add_zero_attn=True) appends one key with logit 0 and value 0
// (reference: models/utils/layers.py:33-34,74-75; torch F.multi_head_attention_forward).
//
// The whole key range fits one tile, so there is no online-softmax rescaling:
//   S[128 x NK] = Q K^T        tcgen05.mma, fp32 in TMEM (NK = round_up(T+1,16), two N<=256 instructions)
//   P = exp2((S - max) * c)    each of the 128 threads owns one query row (tcgen05.ld 32x32b)
//   O[128 x hd] = P V          tcgen05.mma with the A operand (P, packed bf16) read from TMEM where it overwrote S,
//                              V consumed K-major from the transposed V^T buffer the QKV GEMM writes.
// Key row T is zero-filled by TMA out-of-bounds handling, which *is* the zero-attention key; key columns
// > T are masked. Rows/keys beyond the sequence never touch HBM.
#include "attention.cuh"

namespace mmdm {

struct AttnParams {
  int S, T, H, NK, n1, n2, kv_xor, ldo;
  float scale_log2;
  __nv_bfloat16* out;
};

template <int HD>
struct AttnCfg {
  static constexpr int KA = HD / 64;                       // 64-wide K atoms along the head dim
  static constexpr uint32_t Q_BYTES = KA * 128 * 128;      // one Q tile  [KA][128 rows][128 B]   (x2 buffers)
  static constexpr uint32_t K_BYTES = KA * 320 * 128;      // keys        [KA][320 rows][128 B]
  static constexpr uint32_t V_BYTES = 5 * HD * 128;        // V^T         [5][HD rows][128 B]
  static constexpr size_t SMEM_BYTES = 2 * Q_BYTES + K_BYTES + V_BYTES + 1024 + 256;
  static constexpr uint32_t O_COL = 320;                   // TMEM: S/P in [0,320), O in [320, 320+HD)
  static constexpr int THREADS = 160;                      // 4 softmax warps + 1 TMA/MMA warp
};

// Persistent kernel. Work item = (sequence, head); its K / V^T tiles stay in shared memory while the
// ceil(T/128) query tiles stream through a double-buffered Q slot. Per query tile:
//   MMA1  S = Q K^T               (SS, fp32 in TMEM)                      warp 4, one lane
//   softmax, one thread per row   (two TMEM passes; P written back to TMEM as packed bf16, in place of S)
//   MMA2  O = P V                 (TS: A operand read from TMEM)          warp 4, one lane
//   O / rowsum -> bf16 -> global                                          softmax warps
template <int HD>
__global__ void __launch_bounds__(AttnCfg<HD>::THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = AttnCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;                       // 2 buffers
  uint8_t* Ks = Qs + 2 * C::Q_BYTES;
  uint8_t* Vs = Ks + C::K_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + C::V_BYTES);
  uint64_t* qfull = bars;        // [2]
  uint64_t* kfull = bars + 2;
  uint64_t* vfull = bars + 3;
  uint64_t* sfull = bars + 4;
  uint64_t* pfull = bars + 5;
  uint64_t* ofull = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int QT = (p.T + 127) / 128;
  const int total_items = p.S * p.H;
  const int nitems = (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int njobs = nitems * QT;
  const int kgroups = (p.NK + 63) / 64;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&qfull[0], 1);
    mbar_init(&qfull[1], 1);
    mbar_init(kfull, 1);
    mbar_init(vfull, 1);
    mbar_init(sfull, 1);
    mbar_init(pfull, 4);
    mbar_init(ofull, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32(smem_u32(tmem_slot));

  if (warp == 4) {
    if (lane == 0) {
      auto item_sh = [&](int k, int& s, int& h) {
        const int item = (int)blockIdx.x + k * (int)gridDim.x;
        s = item / p.H;
        h = item - s * p.H;
      };
      auto issue_q = [&](int j) {
        if (j >= njobs) return;
        int s, h;
        item_sh(j / QT, s, h);
        const int qt = j % QT, b = j & 1;
        mbar_arrive_expect_tx(&qfull[b], C::Q_BYTES);
        for (int a = 0; a < C::KA; ++a)
          tma_load_3d(Qs + b * C::Q_BYTES + a * (128 * 128), &tmQ, &qfull[b], h * HD + a * 64, qt * 128, s);
      };
      auto issue_k = [&](int k) {
        int s, h;
        item_sh(k, s, h);
        const int ks = s ^ p.kv_xor;
        mbar_arrive_expect_tx(kfull, (uint32_t)(C::KA * kgroups) * (64 * 128));
        for (int a = 0; a < C::KA; ++a)
          for (int r = 0; r < kgroups; ++r)
            tma_load_3d(Ks + a * (320 * 128) + r * (64 * 128), &tmK, kfull, h * HD + a * 64, r * 64, ks);
      };
      auto issue_v = [&](int k) {
        int s, h;
        item_sh(k, s, h);
        const int ks = s ^ p.kv_xor;
        mbar_arrive_expect_tx(vfull, (uint32_t)kgroups * (HD * 128));
        for (int kt = 0; kt < kgroups; ++kt) tma_load_3d(Vs + kt * (HD * 128), &tmV, vfull, kt * 64, 0, ks * p.H + h);
      };
      const uint32_t id1 = umma_idesc_bf16(128, p.n1);
      const uint32_t id2 = umma_idesc_bf16(128, p.n2);
      const uint32_t ido = umma_idesc_bf16(128, HD);
      const uint32_t k0 = smem_u32(Ks), v0 = smem_u32(Vs);
      const int ksteps = p.NK / 16;
      if (nitems > 0) {
        issue_k(0);
        issue_q(0);
        issue_q(1);
        issue_v(0);
      }
      for (int j = 0; j < njobs; ++j) {
        const int k = j / QT, qt = j - k * QT;
        if (qt == 0) mbar_wait(kfull, k & 1);
        mbar_wait(&qfull[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t q0 = smem_u32(Qs + (j & 1) * C::Q_BYTES);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t aoff = (uint32_t)(kk >> 2) * (128 * 128) + (uint32_t)(kk & 3) * 32;
          const uint32_t boff = (uint32_t)(kk >> 2) * (320 * 128) + (uint32_t)(kk & 3) * 32;
          const uint64_t da = umma_desc_k_sw128(q0 + aoff);
          umma_bf16_ss(tmem_base, da, umma_desc_k_sw128(k0 + boff), id1, kk > 0 ? 1u : 0u);
          if (p.n2 > 0)
            umma_bf16_ss(tmem_base + (uint32_t)p.n1, da, umma_desc_k_sw128(k0 + boff + (uint32_t)p.n1 * 128), id2,
                         kk > 0 ? 1u : 0u);
        }
        umma_commit(sfull);
        // P(j) ready in TMEM => MMA1(j) has completed: its Q slot (and, after the last tile, K) can be refilled
        mbar_wait(pfull, j & 1);
        tc_fence_after();
        issue_q(j + 2);
        if (qt == QT - 1 && k + 1 < nitems) issue_k(k + 1);
        if (qt == 0) mbar_wait(vfull, k & 1);
        for (int i = 0; i < ksteps; ++i) {
          const uint32_t boff = (uint32_t)(i >> 2) * (HD * 128) + (uint32_t)(i & 3) * 32;
          umma_bf16_ts(tmem_base + C::O_COL, tmem_base + (uint32_t)(i * 8), umma_desc_k_sw128(v0 + boff), ido,
                       i > 0 ? 1u : 0u);
        }
        umma_commit(ofull);
        if (qt == QT - 1 && k + 1 < nitems) {
          mbar_wait(ofull, j & 1);   // MMA2 of the last tile done: V^T can be replaced
          issue_v(k + 1);
        }
      }
    }
  } else {
    // ---- softmax / output warps: thread <-> query row ------------------------------------------------
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int nchunks = (p.NK + 31) / 32;
    const int row = tid;
    for (int j = 0; j < njobs; ++j) {
      const int k = j / QT, qt = j - k * QT;
      const int item = (int)blockIdx.x + k * (int)gridDim.x;
      const int s = item / p.H, h = item - s * p.H;
      mbar_wait(sfull, j & 1);
      tc_fence_after();
      float mx = 0.f;   // the zero-attention key contributes logit 0
      for (int c = 0; c < nchunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(trow + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (c * 32 + 32 <= p.T) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < p.T) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      const float moff = mx * p.scale_log2;
      float lsum = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(trow + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        float pv[32];
        if (c * 32 + 32 <= p.T) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            pv[i] = ex2_approx(fmaf(__uint_as_float(r[i]), p.scale_log2, -moff));
            lsum += pv[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = ex2_approx(fmaf(__uint_as_float(r[i]), p.scale_log2, -moff));
            pv[i] = (c * 32 + i <= p.T) ? e : 0.f;
            lsum += pv[i];
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(pv[2 * i], pv[2 * i + 1]);
        tmem_st_32x32b_x16(trow + (uint32_t)(c * 16), pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pfull);

      mbar_wait(ofull, j & 1);
      tc_fence_after();
      const float inv = 1.0f / lsum;
      const int t = qt * 128 + row;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(trow + C::O_COL + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (t < p.T) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)s * p.T + t) * p.ldo + h * HD + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(r[8 * i]) * inv, __uint_as_float(r[8 * i + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv);
            dst[i] = o;
          }
        }
      }
      tc_fence_before();   // O reads retire before the next P arrive lets MMA2(j+1) overwrite O
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

static int num_sms_attn() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int HD>
static int launch_attention_t(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk,
                              const __nv_bfloat16* vt, int Tp, __nv_bfloat16* out, int ldo, int S, int T, int H,
                              int kv_xor, cudaStream_t stream) {
  using C = AttnCfg<HD>;
  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)T, (uint64_t)S};
    uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)T * ldq * 2};
    uint32_t box[3] = {64, 128, 1};
    MMDM_TRY(make_tensor_map(&tmQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)T, (uint64_t)S};
    uint64_t str[2] = {(uint64_t)ldk * 2, (uint64_t)T * ldk * 2};
    uint32_t box[3] = {64, 64, 1};
    MMDM_TRY(make_tensor_map(&tmK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    uint64_t dims[3] = {(uint64_t)T, (uint64_t)HD, (uint64_t)S * H};
    uint64_t str[2] = {(uint64_t)Tp * 2, (uint64_t)HD * Tp * 2};
    uint32_t box[3] = {64, (uint32_t)HD, 1};
    MMDM_TRY(make_tensor_map(&tmV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, vt, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  AttnParams p;
  p.S = S; p.T = T; p.H = H;
  p.NK = ((T + 1 + 15) / 16) * 16;
  if (p.NK <= 256) { p.n1 = p.NK; p.n2 = 0; }
  else { p.n1 = ((p.NK / 2 + 15) / 16) * 16; p.n2 = p.NK - p.n1; }
  p.kv_xor = kv_xor;
  p.ldo = ldo;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  p.out = out;
  const int items = S * H;
  const int grid = items < num_sms_attn() ? items : num_sms_attn();
  attention_kernel<HD><<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int attention_init() {
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)AttnCfg<128>::SMEM_BYTES));
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)AttnCfg<64>::SMEM_BYTES));
  return 0;
}

int launch_attention(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* vt,
                     int Tp, __nv_bfloat16* out, int ldo, int S, int T, int H, int hd, int kv_xor,
                     cudaStream_t stream) {
  static int inited = attention_init();
  MMDM_REQUIRE(inited == 0, "attention: kernel attribute setup failed");
  MMDM_REQUIRE(T >= 1 && T <= 319, "attention: T=%d outside the single-tile range [1,319]", T);
  MMDM_REQUIRE(hd == 64 || hd == 128, "attention: head dim %d not in {64,128}", hd);
  MMDM_REQUIRE((ldq % 8) == 0 && (ldk % 8) == 0 && (Tp % 8) == 0 && (ldo % 8) == 0 && Tp >= T,
               "attention: pitches must be multiples of 8 elements");
  if (hd == 128) return launch_attention_t<128>(q, ldq, k, ldk, vt, Tp, out, ldo, S, T, H, kv_xor, stream);
  return launch_attention_t<64>(q, ldq, k, ldk, vt, Tp, out, ldo, S, T, H, kv_xor, stream);
}

}  // namespace mmdm
