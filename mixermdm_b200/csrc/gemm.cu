// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//   warp 0 : TMA producer (one elected lane)      A/B tiles -> 128B-swizzled smem ring
//   warp 1 : MMA issuer   (one elected lane)      tcgen05.mma 128 x BN x 16, fp32 accumulators in TMEM
//   warp 2 : TMEM allocator / deallocator
//   warps 4-7 : epilogue  (tcgen05.ld -> registers -> fused epilogue -> global), TMEM double buffered so
//               the epilogue of tile i overlaps the MMAs of tile i+1.
// Replaces every nn.Linear on the hot path of the reference (models/utils/layers.py:11,33,74,95-96,112;
// models/in2in.py:389; models/mixermdm.py:652) with the listed fused epilogues.
// File map: fused epilogues (one thread = one output row; bias / GELU / V^T scatter / LN-fold consumer) - residual
// epilogue with a quarter-warp transpose (single-CTA kernels, embedding modes) and its LN-fold producer part - single-CTA
// kernel - residual epilogue through TMA (CTA-pair kernel) - CTA-pair kernel (cta_group::2; warps 2 and 3 are the aux
// warps of the LN-fold consumers: one extra k step of the MMA carries the mean / shift terms) - host launchers.
#include "gemm.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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

int num_sms() { return device_sm_count(); }

// ------------------------------------------------------------------------------------------------
// fused epilogue. One thread owns one output row of the tile and walks it in chunks of 32 columns (the
// tcgen05.ld 32x32b shape). Everything that does not depend on the column (row validity, sequence / frame
// index of the row) is computed once per tile in RowCtx; the per-chunk fast path has no per-element guards.
// ------------------------------------------------------------------------------------------------
// modes whose fp32 output goes through the shared-memory transpose (full-line loads / stores)
__host__ __device__ constexpr bool transposed_mode(int mode) { return mode_resid(mode) || mode_embed(mode); }

struct RowCtx {
  int row;
  bool valid;
  int s, t;   // row = s * T + t (only for the modes that need it)
  // LN-fold consumer: y = rho * acc + nrm * G_s[n] + Bc_s[n]   (nrm = -rho * mu)
  float rho, nrm;
  const float* gp;
  const float* bp;
  uint32_t g_smem, b_smem;   // the warp's staged G_s / Bc_s values of this row's sequence (chunk 0), 0: not staged
};

template <int MODE, bool AUX = false>
__device__ __forceinline__ RowCtx make_row_ctx(const GemmEpi& ep, int row, int M) {
  RowCtx rc;
  rc.row = row;
  rc.valid = row < M;
  rc.s = 0;
  rc.t = 0;
  rc.rho = 1.f;
  rc.nrm = 0.f;
  rc.gp = rc.bp = nullptr;
  rc.g_smem = rc.b_smem = 0;
  if (mode_bf16(MODE) || mode_embed(MODE) || MODE == EPI_FINAL || mode_ln_consumer(MODE)) {
    rc.s = row / ep.T;
    rc.t = row - rc.s * ep.T;
  }
  if (mode_ln_consumer(MODE) && !AUX && rc.valid) {
    // row statistics from the producer's partial sums, always added in slot order (bitwise reproducible)
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < ep.lnc_slots; ++k) {
      const float2 p = __ldg(ep.lnc_stats + (size_t)k * ep.lnc_pitch + row);
      s1 += p.x;
      s2 += p.y;
    }
    const float mu = s1 * ep.lnc_inv_d;
    const float var = fmaxf(fmaf(-mu, mu, s2 * ep.lnc_inv_d), 0.f);
    rc.rho = rsqrtf(var + 1e-6f);
    rc.nrm = -rc.rho * mu;
    rc.gp = ep.lnc_g + (size_t)rc.s * ep.lnc_ld;
    rc.bp = ep.lnc_b + (size_t)rc.s * ep.lnc_ld;
  }
  return rc;
}

// The bias of the 32 columns of a chunk, read (as a broadcast) from the warp's shared-memory slice that
// epilogue_tile() fills once per tile: 8 LDS.128 instead of 8 dependent global loads per chunk.
__device__ __forceinline__ void load_bias32(uint32_t bias_smem, float* b) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 q = lds128(bias_smem + (uint32_t)j * 16);
    b[4 * j] = q.x; b[4 * j + 1] = q.y; b[4 * j + 2] = q.z; b[4 * j + 3] = q.w;
  }
}

// GELU of the FFN-1 epilogue: exact-erf definition (Abramowitz-Stegun, 1.5e-7). -DMMDM_GELU_TANH selects the one-MUFU
// tanh form of common.cuh (|error| < 3e-4 |x|): measured 137.4 against 143.0 us per FFN-1 launch (a plain bf16 epilogue:
// 130.0 us), i.e. 0.5 % of a DDIM step - not worth leaving the reference's definition for.
#ifdef MMDM_GELU_TANH
#define MMDM_GELU1 gelu_tanh5
#define MMDM_GELU2 gelu_tanh2_bf16
#else
#define MMDM_GELU1 gelu_fast
#define MMDM_GELU2 gelu_fast2_bf16
#endif

// generic (edge) path: per-element guards; used only for ragged N (the 262-wide output layer) and by the
// SIMT bring-up kernel
template <int MODE>
__device__ __forceinline__ void epilogue_chunk_edge(const GemmEpi& ep, const RowCtx& rc, int col0, const float* v, int N) {
  const int nvalid = min(32, N - col0);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j >= nvalid) break;
    const int n = col0 + j;
    const float x = mode_ln_consumer(MODE) ? fmaf(rc.rho, v[j], fmaf(rc.nrm, __ldg(rc.gp + n), __ldg(rc.bp + n)))
                                           : v[j] + (ep.bias ? __ldg(ep.bias + n) : 0.f);
    if (mode_bf16(MODE) || mode_gelu(MODE) || MODE == EPI_QGELU_BF16) {
      const float o = mode_gelu(MODE) ? MMDM_GELU1(x) : (MODE == EPI_QGELU_BF16) ? quick_gelu(x) : x;
      if (mode_bf16(MODE) && ep.vt_col0 >= 0 && n >= ep.vt_col0) {
        const int nn = n - ep.vt_col0, h = nn / ep.hd, d = nn - h * ep.hd;
        ep.vt[((size_t)(rc.s * ep.H + h) * ep.hd + d) * ep.Tp + rc.t] = __float2bfloat16_rn(o);
      } else {
        reinterpret_cast<__nv_bfloat16*>(ep.out)[(size_t)rc.row * ep.ldo + n] = __float2bfloat16_rn(o);
      }
    } else if (mode_resid(MODE)) {   // (LN-fold producers never come here: launch_gemm requires the vector path)
      const float o = ep.resid[(size_t)rc.row * ep.ldo + n] + x;
      reinterpret_cast<float*>(ep.out)[(size_t)rc.row * ep.ldo + n] = o;
      if (ep.out_bf16) ep.out_bf16[(size_t)rc.row * ep.ldo + n] = __float2bfloat16_rn(o);
    } else if (mode_embed(MODE)) {
      const float o = x + __ldg(ep.pe + (size_t)rc.t * ep.pe_ld + n);
      float* dst = reinterpret_cast<float*>(ep.out) + (size_t)rc.row * ep.ldo + n;
      *dst = o;
      if (ep.dup_row_offset != 0) dst[ep.dup_row_offset * ep.ldo] = o;
    } else if (MODE == EPI_FINAL) {
      float* dst = (ep.persons_per_row == 2)
                       ? reinterpret_cast<float*>(ep.out) + ((size_t)(rc.s >> 1) * ep.T + rc.t) * (2 * ep.nfeats) +
                             (rc.s & 1) * ep.nfeats + n
                       : reinterpret_cast<float*>(ep.out) + ((size_t)rc.s * ep.T + rc.t) * ep.nfeats + n;
      *dst = ep.std != nullptr ? x * __ldg(ep.std + n) + __ldg(ep.mean + n) : x;
    } else if (MODE == EPI_SIGMOID_F32) {
      reinterpret_cast<float*>(ep.out)[(size_t)rc.row * ep.ldo + n] = 1.f / (1.f + expf(-x));
    } else {
      reinterpret_cast<float*>(ep.out)[(size_t)rc.row * ep.ldo + n] = x;
    }
  }
}

template <int MODE, bool AUX = false>
__device__ __forceinline__ void epilogue_chunk(const GemmEpi& ep, const RowCtx& rc, int col0, const float* v, int N,
                                               uint32_t bias_smem) {
  if (!rc.valid || col0 >= N) return;
  if (col0 + 32 > N || MODE == EPI_FINAL || MODE == EPI_SIGMOID_F32 || transposed_mode(MODE) || !ep.vec_ok ||
      (mode_ln_consumer(MODE) && !AUX && rc.g_smem == 0)) {   // FINAL rows are only 8-byte aligned in the scattered layout
    epilogue_chunk_edge<MODE>(ep, rc, col0, v, N);
    return;
  }
  if (mode_ln_consumer(MODE)) {
    // x = rho * acc + (nrm * G_s[n] + Bc_s[n]); G_s / Bc_s come from the warp's staged rows (epilogue_tile made sure this
    // row's sequence is one of the two staged ones; bias_smem carries the chunk's byte offset inside them)
    const f32x2 rho2 = pack2(rc.rho, rc.rho), nrm2 = pack2(rc.nrm, rc.nrm);
    const uint32_t gs = rc.g_smem + bias_smem, bs = rc.b_smem + bias_smem;
    const bool to_vt = mode_bf16(MODE) && ep.vt_col0 >= 0 && col0 >= ep.vt_col0;
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)rc.row * ep.ldo + col0);
    __nv_bfloat16* dvt = nullptr;
    if (to_vt) {   // V^T: the 32 columns of a chunk lie inside one head
      const int nn = col0 - ep.vt_col0;
      const int h = nn / ep.hd, d0 = nn - h * ep.hd;
      dvt = ep.vt + ((size_t)(rc.s * ep.H + h) * ep.hd + d0) * ep.Tp + rc.t;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // 8 columns per step
      f32x2 x[4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (AUX) {   // the accumulator already holds acc - mu G_s + Bc_s / rho (an extra k step of the MMA, see the aux warp)
          x[2 * u] = mul2(rho2, pack2(v[8 * j + 4 * u], v[8 * j + 4 * u + 1]));
          x[2 * u + 1] = mul2(rho2, pack2(v[8 * j + 4 * u + 2], v[8 * j + 4 * u + 3]));
        } else {
          const float4 g4 = lds128(gs + (uint32_t)(2 * j + u) * 16), b4 = lds128(bs + (uint32_t)(2 * j + u) * 16);
          x[2 * u] = fma2(rho2, pack2(v[8 * j + 4 * u], v[8 * j + 4 * u + 1]), fma2(nrm2, pack2(g4.x, g4.y), pack2(b4.x, b4.y)));
          x[2 * u + 1] = fma2(rho2, pack2(v[8 * j + 4 * u + 2], v[8 * j + 4 * u + 3]), fma2(nrm2, pack2(g4.z, g4.w), pack2(b4.z, b4.w)));
        }
      }
      if (mode_gelu(MODE)) {
        uint4 pk;
        pk.x = MMDM_GELU2(x[0]);
        pk.y = MMDM_GELU2(x[1]);
        pk.z = MMDM_GELU2(x[2]);
        pk.w = MMDM_GELU2(x[3]);
        dst[j] = pk;
      } else {
        float o[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) unpack2(x[u], o[2 * u], o[2 * u + 1]);
        if (to_vt) {
#pragma unroll
          for (int u = 0; u < 8; ++u) dvt[(size_t)(8 * j + u) * ep.Tp] = __float2bfloat16_rn(o[u]);
        } else {
          uint4 pk;
          pk.x = pack_bf16x2(o[0], o[1]);
          pk.y = pack_bf16x2(o[2], o[3]);
          pk.z = pack_bf16x2(o[4], o[5]);
          pk.w = pack_bf16x2(o[6], o[7]);
          dst[j] = pk;
        }
      }
    }
    return;
  }
  float b[32];
  load_bias32(bias_smem, b);
  if (MODE == EPI_GELU_BF16) {
    // packed fp32x2 evaluation: the epilogue was issue-bound on the scalar form
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)rc.row * ep.ldo + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 pk;
      pk.x = MMDM_GELU2(add2(pack2(v[8 * j], v[8 * j + 1]), pack2(b[8 * j], b[8 * j + 1])));
      pk.y = MMDM_GELU2(add2(pack2(v[8 * j + 2], v[8 * j + 3]), pack2(b[8 * j + 2], b[8 * j + 3])));
      pk.z = MMDM_GELU2(add2(pack2(v[8 * j + 4], v[8 * j + 5]), pack2(b[8 * j + 4], b[8 * j + 5])));
      pk.w = MMDM_GELU2(add2(pack2(v[8 * j + 6], v[8 * j + 7]), pack2(b[8 * j + 6], b[8 * j + 7])));
      dst[j] = pk;
    }
  } else if (MODE == EPI_QGELU_BF16) {
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)rc.row * ep.ldo + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 pk;
      pk.x = pack_bf16x2(quick_gelu(v[8 * j] + b[8 * j]), quick_gelu(v[8 * j + 1] + b[8 * j + 1]));
      pk.y = pack_bf16x2(quick_gelu(v[8 * j + 2] + b[8 * j + 2]), quick_gelu(v[8 * j + 3] + b[8 * j + 3]));
      pk.z = pack_bf16x2(quick_gelu(v[8 * j + 4] + b[8 * j + 4]), quick_gelu(v[8 * j + 5] + b[8 * j + 5]));
      pk.w = pack_bf16x2(quick_gelu(v[8 * j + 6] + b[8 * j + 6]), quick_gelu(v[8 * j + 7] + b[8 * j + 7]));
      dst[j] = pk;
    }
  } else if (MODE == EPI_BF16) {
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = v[j] + b[j];
    if (MODE == EPI_BF16 && ep.vt_col0 >= 0 && col0 >= ep.vt_col0) {
      // V^T: the 32 columns of a chunk lie inside one head (hd is 64 or 128, chunks are 32-aligned)
      const int nn = col0 - ep.vt_col0;
      const int h = nn / ep.hd, d0 = nn - h * ep.hd;
      __nv_bfloat16* dst = ep.vt + ((size_t)(rc.s * ep.H + h) * ep.hd + d0) * ep.Tp + rc.t;
#pragma unroll
      for (int j = 0; j < 32; ++j) dst[(size_t)j * ep.Tp] = __float2bfloat16_rn(o[j]);
    } else {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)rc.row * ep.ldo + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        pk.x = pack_bf16x2(o[8 * j], o[8 * j + 1]);
        pk.y = pack_bf16x2(o[8 * j + 2], o[8 * j + 3]);
        pk.z = pack_bf16x2(o[8 * j + 4], o[8 * j + 5]);
        pk.w = pack_bf16x2(o[8 * j + 6], o[8 * j + 7]);
        dst[j] = pk;
      }
    }
  } else if (MODE == EPI_RESID) {
    const float4* rs = reinterpret_cast<const float4*>(ep.resid + (size_t)rc.row * ep.ldo + col0);
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (size_t)rc.row * ep.ldo + col0);
    float o[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 r4 = rs[j];
      o[4 * j] = r4.x + v[4 * j] + b[4 * j];
      o[4 * j + 1] = r4.y + v[4 * j + 1] + b[4 * j + 1];
      o[4 * j + 2] = r4.z + v[4 * j + 2] + b[4 * j + 2];
      o[4 * j + 3] = r4.w + v[4 * j + 3] + b[4 * j + 3];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    if (ep.out_bf16 != nullptr) {
      uint4* d2 = reinterpret_cast<uint4*>(ep.out_bf16 + (size_t)rc.row * ep.ldo + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        pk.x = pack_bf16x2(o[8 * j], o[8 * j + 1]);
        pk.y = pack_bf16x2(o[8 * j + 2], o[8 * j + 3]);
        pk.z = pack_bf16x2(o[8 * j + 4], o[8 * j + 5]);
        pk.w = pack_bf16x2(o[8 * j + 6], o[8 * j + 7]);
        d2[j] = pk;
      }
    }
  } else if (MODE == EPI_EMBED_PE) {
    const float4* pe = reinterpret_cast<const float4*>(ep.pe + (size_t)rc.t * ep.pe_ld + col0);
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (size_t)rc.row * ep.ldo + col0);
    float4* dup = ep.dup_row_offset != 0
                      ? reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) +
                                                  ((size_t)rc.row + ep.dup_row_offset) * ep.ldo + col0)
                      : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 q = __ldg(pe + j);
      const float4 o = make_float4(v[4 * j] + b[4 * j] + q.x, v[4 * j + 1] + b[4 * j + 1] + q.y,
                                   v[4 * j + 2] + b[4 * j + 2] + q.z, v[4 * j + 3] + b[4 * j + 3] + q.w);
      dst[j] = o;
      if (dup) dup[j] = o;
    }
  } else {  // EPI_F32
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (size_t)rc.row * ep.ldo + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      dst[j] = make_float4(v[4 * j] + b[4 * j], v[4 * j + 1] + b[4 * j + 1], v[4 * j + 2] + b[4 * j + 2],
                           v[4 * j + 3] + b[4 * j + 3]);
  }
}

// Residual epilogue with a shared-memory transpose: accumulators arrive one ROW per thread, but the fp32
// residual stream is read and written one row per QUARTER-WARP (4 rows x 128 B per instruction), so every
// global access is a full 128-byte line instead of 32 scattered 16-byte pieces.
// stg: per-warp staging of 32 x 32 floats, float4 slots XOR-swizzled -> conflict-free both ways.
// bias4: the bias of this lane's 4 columns (col0 + 4*(lane&7) ...), the same for all 8 rows the lane writes.
// The same path serves the motion-embedding epilogue (MODE == EPI_EMBED_PE): the addend is the positional-encoding
// row pe[row % T] instead of the residual row, and the result is optionally written a second time dup_row_offset
// rows further down (the unconditional CFG half shares the embedding).
// LN-fold producer state of a lane: in every chunk of a tile the lane touches the same 8 rows, m_base + 4 i + (lane >> 3)
// (4 columns of each), so their sequence indices are computed once per tile and the row sums accumulate in registers.
struct LnProd {
  float s1[8], s2[8];
  int sa, sb;   // sequences of the warp's first / last participating row (warp-uniform; a warp's 32 rows span two
                // sequences at most unless T < 32): their scale values are fetched once per chunk, ahead of use
  int bnd;      // first row of sequence sb
  int lim;      // rows >= lim take no part (beyond M or beyond lnp_rows)
  float m[8];   // row offsets (GemmEpi::lnp_moff_in) of the lane's 8 rows
};
__device__ __forceinline__ int ln_prod_seq(const GemmEpi& ep, const LnProd& lp, int row) {
  if (ep.lnp_T < 32) return row / ep.lnp_T;   // (more than two sequences per warp possible)
  return row < lp.bnd ? lp.sa : lp.sb;
}
__device__ __forceinline__ void ln_prod_begin(const GemmEpi& ep, LnProd& lp, int m_base, int lane, int M) {
  lp.lim = min(M, ep.lnp_rows);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    lp.s1[i] = lp.s2[i] = 0.f;
    const int row = m_base + 4 * i + (lane >> 3);
    lp.m[i] = (ep.lnp_moff_in != nullptr && row < lp.lim) ? __ldg(ep.lnp_moff_in + row) : 0.f;
  }
  lp.sa = min(m_base, lp.lim - 1) / ep.lnp_T;
  lp.sb = min(m_base + 31, lp.lim - 1) / ep.lnp_T;
  lp.bnd = (lp.sa + 1) * ep.lnp_T;
}
// scale values of this lane's 4 columns for the warp's two sequences (+ the duplicated rows' sequences)
struct LnScale {
  float4 a, b;
};
__device__ __forceinline__ LnScale ln_scale_fetch(const float* scale, int ld, int sa, int sb, int cg) {
  LnScale r;
  r.a = __ldg(reinterpret_cast<const float4*>(scale + (size_t)sa * ld + cg));
  r.b = sb != sa ? __ldg(reinterpret_cast<const float4*>(scale + (size_t)sb * ld + cg)) : r.a;
  return r;
}
__device__ __forceinline__ float4 ln_scale_pick(const LnScale& sc, const float* scale, int ld, int seq, int sa, int sb, int cg) {
  if (seq == sa) return sc.a;
  if (seq == sb) return sc.b;
  return __ldg(reinterpret_cast<const float4*>(scale + (size_t)seq * ld + cg));   // T < 32 only
}
// partial sums of the last 64 columns -> stats[slot][row] (8 lanes hold 4 + 4 columns each of a row)
__device__ __forceinline__ void ln_prod_flush(const GemmEpi& ep, LnProd& lp, int m_base, int lane, int slot) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a = lp.s1[i], b = lp.s2[i];
    a += __shfl_xor_sync(0xffffffffu, a, 1); b += __shfl_xor_sync(0xffffffffu, b, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2); b += __shfl_xor_sync(0xffffffffu, b, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 4); b += __shfl_xor_sync(0xffffffffu, b, 4);
    if ((lane & 7) == 0 && m_base + 4 * i + (lane >> 3) < lp.lim) {
      float2* dst = ep.lnp_stats + (size_t)slot * ep.lnp_pitch + (m_base + 4 * i + (lane >> 3));
      *dst = make_float2(a, b);
      if (ep.dup_row_offset != 0) dst[ep.dup_row_offset] = make_float2(a, b);
    }
    lp.s1[i] = lp.s2[i] = 0.f;
  }
}
__device__ __forceinline__ uint2 ln_prod_pack(const float4 o, const float4 sc) {
  uint2 pk;   // bf16(h * (1 + scale))
  pk.x = pack_bf16x2(fmaf(o.x, sc.x, o.x), fmaf(o.y, sc.y, o.y));
  pk.y = pack_bf16x2(fmaf(o.z, sc.z, o.z), fmaf(o.w, sc.w, o.w));
  return pk;
}

template <int MODE>
__device__ __forceinline__ void epilogue_chunk_resid_t(const GemmEpi& ep, int m_base, int lane, int col0,
                                                       const float* v, int M, uint32_t stg, float4 bias4, LnProd& lp) {
#pragma unroll
  for (int g = 0; g < 8; ++g)
    sts128(stg + (uint32_t)(lane * 8 + (g ^ (lane & 7))) * 16,
           make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]));
  __syncwarp();
  const int g = lane & 7;
  // LN-fold producer: the scale values go out with the residual loads (behind the stores they would wait for L2)
  LnScale sc0 = {}, sc1 = {}, sd0 = {}, sd1 = {};
  const int cg = col0 + 4 * g;
  if (mode_ln_producer(MODE)) {
    sc0 = ln_scale_fetch(ep.lnp_scale0, ep.lnp_ld0, lp.sa, lp.sb, cg);
    if (ep.lnp_out1 != nullptr) sc1 = ln_scale_fetch(ep.lnp_scale1, ep.lnp_ld1, lp.sa, lp.sb, cg);
    if (mode_embed(MODE) && ep.dup_row_offset != 0) {
      sd0 = ln_scale_fetch(ep.lnp_scale0, ep.lnp_ld0, lp.sa + ep.lnp_dup_seq, lp.sb + ep.lnp_dup_seq, cg);
      if (ep.lnp_out1 != nullptr) sd1 = ln_scale_fetch(ep.lnp_scale1, ep.lnp_ld1, lp.sa + ep.lnp_dup_seq, lp.sb + ep.lnp_dup_seq, cg);
    }
  }
  // all residual loads first (the output may alias the residual, so loads cannot be hoisted over stores)
  float4 r4[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m_base + 4 * i + (lane >> 3);
    if (mode_embed(MODE))
      r4[i] = (row < M) ? __ldg(reinterpret_cast<const float4*>(ep.pe + (size_t)(row % ep.T) * ep.pe_ld + col0 + 4 * g))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    else
      r4[i] = (row < M) ? *reinterpret_cast<const float4*>(ep.resid + (size_t)row * ep.ldo + col0 + 4 * g)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = 4 * i + (lane >> 3);
    const int row = m_base + rr;
    if (row < M) {
      const float4 a = lds128(stg + (uint32_t)(rr * 8 + (g ^ (rr & 7))) * 16);
      const size_t off = (size_t)row * ep.ldo + col0 + 4 * g;
      const float4 o = make_float4(r4[i].x + a.x + bias4.x, r4[i].y + a.y + bias4.y, r4[i].z + a.z + bias4.z,
                                   r4[i].w + a.w + bias4.w);
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + off) = o;
      if (mode_embed(MODE) && ep.dup_row_offset != 0)
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + off + (size_t)ep.dup_row_offset * ep.ldo) = o;
      if (mode_resid(MODE) && ep.out_bf16 != nullptr) {
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        *reinterpret_cast<uint2*>(ep.out_bf16 + off) = pk;
      }
      if (mode_ln_producer(MODE) && row < lp.lim) {
        const float4 os = make_float4(o.x - lp.m[i], o.y - lp.m[i], o.z - lp.m[i], o.w - lp.m[i]);   // h - m_r
        lp.s1[i] += (os.x + os.y) + (os.z + os.w);
        lp.s2[i] += fmaf(os.x, os.x, os.y * os.y) + fmaf(os.z, os.z, os.w * os.w);
        const int sq = ln_prod_seq(ep, lp, row);
        *reinterpret_cast<uint2*>(ep.lnp_out0 + off) =
            ln_prod_pack(os, ln_scale_pick(sc0, ep.lnp_scale0, ep.lnp_ld0, sq, lp.sa, lp.sb, cg));
        if (ep.lnp_out1 != nullptr)
          *reinterpret_cast<uint2*>(ep.lnp_out1 + off) =
              ln_prod_pack(os, ln_scale_pick(sc1, ep.lnp_scale1, ep.lnp_ld1, sq, lp.sa, lp.sb, cg));
        if (mode_embed(MODE) && ep.dup_row_offset != 0) {   // the other CFG half: same h, its own modulation
          const size_t off2 = off + (size_t)ep.dup_row_offset * ep.ldo;
          const int dq = ep.lnp_dup_seq;
          *reinterpret_cast<uint2*>(ep.lnp_out0 + off2) =
              ln_prod_pack(os, ln_scale_pick(sd0, ep.lnp_scale0, ep.lnp_ld0, sq + dq, lp.sa + dq, lp.sb + dq, cg));
          if (ep.lnp_out1 != nullptr)
            *reinterpret_cast<uint2*>(ep.lnp_out1 + off2) =
                ln_prod_pack(os, ln_scale_pick(sd1, ep.lnp_scale1, ep.lnp_ld1, sq + dq, lp.sa + dq, lp.sb + dq, cg));
        }
      }
    }
  }
  __syncwarp();
}

// Output-layer epilogue (MODE == EPI_FINAL), same transpose: (acc + bias) * std + mean written to the (B, T,
// persons * nfeats) layout. Rows of that layout are only 8-byte aligned (nfeats = 262), hence float2 stores; a
// quarter-warp still covers 128 contiguous bytes of one row instead of 32 lanes writing 32 different rows.
__device__ __forceinline__ void epilogue_chunk_final_t(const GemmEpi& ep, int m_base, int lane, int col0,
                                                       const float* v, int M, int N, uint32_t stg) {
#pragma unroll
  for (int g = 0; g < 8; ++g)
    sts128(stg + (uint32_t)(lane * 8 + (g ^ (lane & 7))) * 16,
           make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]));
  __syncwarp();
  const int g = lane & 7;
  const int c = col0 + 4 * g;
  float bs[4], sd[4], mn[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const bool ok = c + e < N;
    bs[e] = (ok && ep.bias) ? __ldg(ep.bias + c + e) : 0.f;
    sd[e] = (ok && ep.std) ? __ldg(ep.std + c + e) : 1.f;
    mn[e] = (ok && ep.std) ? __ldg(ep.mean + c + e) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = 4 * i + (lane >> 3);
    const int row = m_base + rr;
    if (row < M && c < N) {
      const float4 a = lds128(stg + (uint32_t)(rr * 8 + (g ^ (rr & 7))) * 16);
      const int sq = row / ep.T, t = row - sq * ep.T;
      float* dst = (ep.persons_per_row == 2)
                       ? reinterpret_cast<float*>(ep.out) + ((size_t)(sq >> 1) * ep.T + t) * (2 * ep.nfeats) +
                             (sq & 1) * ep.nfeats + c
                       : reinterpret_cast<float*>(ep.out) + ((size_t)sq * ep.T + t) * ep.nfeats + c;
      const float o0 = (a.x + bs[0]) * sd[0] + mn[0], o1 = (a.y + bs[1]) * sd[1] + mn[1];
      const float o2 = (a.z + bs[2]) * sd[2] + mn[2], o3 = (a.w + bs[3]) * sd[3] + mn[3];
      if (c + 1 < N) *reinterpret_cast<float2*>(dst) = make_float2(o0, o1); else dst[0] = o0;
      if (c + 3 < N) *reinterpret_cast<float2*>(dst + 2) = make_float2(o2, o3); else if (c + 2 < N) dst[2] = o2;
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Epilogue of one 128 x BN accumulator tile by one of the 8 epilogue warps (TMEM lane quarter q, column half
// `half`). The TMEM read of chunk c+1 is in flight while chunk c is converted and stored.
// ------------------------------------------------------------------------------------------------
// Bias of the warp's BN/2 columns, fetched from global memory ahead of use. Residual mode (transposed stores): the
// float4 of this lane's column group per chunk. Other modes (one row per thread): lane l holds column l of each
// chunk in .x; bias_publish() spreads it through the warp's shared-memory slice (which aliases the transpose
// staging that only the residual mode uses).
template <int BN>
struct EpiBias {
  float4 b[BN / 64];
};
template <int BN, int MODE>
__device__ __forceinline__ EpiBias<BN> bias_fetch(const GemmEpi& ep, int n0, int c0, int lane, int N) {
  EpiBias<BN> r;   // c0: first 32-column chunk of the warp; up to BN/64 chunks are fetched
#pragma unroll
  for (int i = 0; i < BN / 64; ++i) {
    const int col0 = n0 + (c0 + i) * 32;
    r.b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.bias != nullptr) {
      if (transposed_mode(MODE)) {
        if (col0 + 32 <= N && ep.vec_ok) r.b[i] = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + (lane & 7));
      } else if (col0 + lane < N) {
        r.b[i].x = __ldg(ep.bias + col0 + lane);
      }
    }
  }
  return r;
}
template <int BN>
__device__ __forceinline__ void bias_publish(const EpiBias<BN>& b, uint32_t bias_smem, int lane) {
  __syncwarp();   // the previous tile's readers are done
#pragma unroll
  for (int i = 0; i < BN / 64; ++i)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_smem + (uint32_t)(i * 32 + lane) * 4), "f"(b.b[i].x) : "memory");
  __syncwarp();
}

// LN-fold consumer: G_s | Bc_s of the warp's (at most two) sequences over its column range, fetched one tile ahead and
// published in the warp's shared-memory slice as [sequence a | b][G | Bc][128 floats] (2 KB; a row of a third sequence -
// possible only for T < 32 - takes the generic per-element path)
template <int BN>
struct EpiLnc {
  float4 v[4];
};
template <int BN>
__device__ __forceinline__ EpiLnc<BN> lnc_fetch(const GemmEpi& ep, int row0, int n0, int c0, int cnt, int lane, int M, int N) {
  EpiLnc<BN> r;
  static_assert(BN / 64 <= 4, "one float4 per lane and block");
  const int sa = min(row0, M - 1) / ep.T, sb = min(row0 + 31, M - 1) / ep.T;
  const int col = n0 + c0 * 32 + lane * 4;
  const bool on = lane < cnt * 8 && col + 4 <= N;
#pragma unroll
  for (int j = 0; j < 4; ++j) {   // block j: sequence (j >> 1 ? sb : sa), array (j & 1 ? Bc : G); lane = float4 inside it
    const float* base = ((j & 1) ? ep.lnc_b : ep.lnc_g) + (size_t)((j & 2) ? sb : sa) * ep.lnc_ld;
    r.v[j] = on ? __ldg(reinterpret_cast<const float4*>(base + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return r;
}
template <int BN>
__device__ __forceinline__ void lnc_publish(const EpiLnc<BN>& r, uint32_t smem, int cnt, int lane) {
  __syncwarp();   // the previous tile's readers are done
#pragma unroll
  for (int j = 0; j < 4; ++j) sts128(smem + (uint32_t)(j * 32 + lane) * 16, r.v[j]);
  __syncwarp();
}

template <int BN, int MODE, bool AUX = false>
__device__ __forceinline__ void epilogue_tile(const GemmEpi& ep, int m0, int n0, int M, int N, uint32_t tmem_acc,
                                              int q, int c0, int cnt, int lane, uint32_t stg, const EpiBias<BN>& bias,
                                              RowCtx rc) {
  const uint32_t bias_smem = stg;   // non-residual modes: the staging slice holds the published bias
  constexpr int NCH = BN / 64;   // at most this many chunks of 32 columns per warp: chunks [c0, c0 + cnt)
  if (mode_ln_consumer(MODE) && !AUX && ep.vec_ok) {
    const int row0 = m0 + q * 32;
    const int sa = min(row0, M - 1) / ep.T, sb = min(row0 + 31, M - 1) / ep.T;
    const int sel = rc.s == sa ? 0 : (rc.s == sb ? 1 : -1);
    if (sel >= 0) {
      rc.g_smem = stg + (uint32_t)sel * 1024u;
      rc.b_smem = rc.g_smem + 512u;
    }
  }
  const uint32_t taddr = tmem_acc + (uint32_t)(c0 * 32) + ((uint32_t)(q * 32) << 16);
  LnProd lp;
  if (mode_ln_producer(MODE)) ln_prod_begin(ep, lp, m0 + q * 32, lane, M);
  uint32_t r[2][32];
  tmem_ld_32x32b_x32(taddr, r[0]);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c >= cnt) break;
    // (the LN-fold producers keep one TMEM buffer: their epilogue is bound by global memory and short of registers)
    if (c + 1 < cnt && !mode_ln_producer(MODE)) tmem_ld_32x32b_x32(taddr + (uint32_t)((c + 1) * 32), r[(c + 1) & 1]);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[c & 1][j]);
    const int col0 = n0 + (c0 + c) * 32;
    if (transposed_mode(MODE) && ep.vec_ok && col0 + 32 <= N) {
      epilogue_chunk_resid_t<MODE>(ep, m0 + q * 32, lane, col0, v, M, stg, bias.b[c], lp);
      // (a warp's chunks come in pairs that cover one 64-column statistics slot: c0 is even, cnt is even)
      if (mode_ln_producer(MODE) && (c & 1)) ln_prod_flush(ep, lp, m0 + q * 32, lane, (col0 - 32) / kLnSlotCols);
    }
    else if (MODE == EPI_FINAL && ep.final_vec_ok && col0 < N)
      epilogue_chunk_final_t(ep, m0 + q * 32, lane, col0, v, M, N, stg);
    else
      epilogue_chunk<MODE, AUX>(ep, rc, col0, v, N, mode_ln_consumer(MODE) ? (uint32_t)c * 128 : bias_smem + (uint32_t)c * 128);
    if (c + 1 < cnt) {
      if (mode_ln_producer(MODE)) tmem_ld_32x32b_x32(taddr + (uint32_t)((c + 1) * 32), r[(c + 1) & 1]);
      tmem_ld_wait();
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
  static constexpr size_t STG_BYTES = 8 * 32 * 32 * 4;   // per-epilogue-warp transpose staging
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + STG_BYTES;
};

template <int BN, int MODE>
__global__ void __launch_bounds__(384, 1)
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
  const uint32_t stg_base = smem_u32(smem + (size_t)C::STAGES * C::STAGE_BYTES + 256);

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
      mbar_init(&tempty_bar[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32(smem_u32(tmem_slot));

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
    // 8 epilogue warps: warps w and w+4 share TMEM lane quarter (w & 3) and split the tile's column chunks
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t stg = stg_base + (uint32_t)(warp - 4) * 4096;
    int acc = 0;
    uint32_t acc_phase = 0;
    EpiBias<BN> bias = {}, bias_next = {};
    EpiLnc<BN> lnc_next = {};
    if (!transposed_mode(MODE) && !mode_ln_consumer(MODE) && (int)blockIdx.x < num_tiles)
      bias_next = bias_fetch<BN, MODE>(ep, ((int)blockIdx.x % num_n) * BN, half * (BN / 64), lane, N);
    if (mode_ln_consumer(MODE) && ep.vec_ok && (int)blockIdx.x < num_tiles)
      lnc_next = lnc_fetch<BN>(ep, ((int)blockIdx.x / num_n) * C::BM + q * 32, ((int)blockIdx.x % num_n) * BN,
                               half * (BN / 64), BN / 64, lane, M, N);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * C::BM;
      const int n0 = (tile % num_n) * BN;
      if (mode_ln_consumer(MODE) && ep.vec_ok) {
        lnc_publish<BN>(lnc_next, stg, BN / 64, lane);
        const int nt = tile + (int)gridDim.x;
        if (nt < num_tiles)
          lnc_next = lnc_fetch<BN>(ep, (nt / num_n) * C::BM + q * 32, (nt % num_n) * BN, half * (BN / 64), BN / 64, lane, M, N);
      }
      if (transposed_mode(MODE)) {
        bias = bias_fetch<BN, MODE>(ep, n0, half * (BN / 64), lane, N);   // in flight while waiting for the accumulator
      } else if (MODE != EPI_FINAL && !mode_ln_consumer(MODE)) {   // (the output layer reads bias / std / mean per column group itself)
        bias_publish<BN>(bias_next, stg, lane);
        if (tile + (int)gridDim.x < num_tiles)   // next tile's bias: its latency hides behind this tile
          bias_next = bias_fetch<BN, MODE>(ep, ((tile + (int)gridDim.x) % num_n) * BN, half * (BN / 64), lane, N);
      }
      if (mode_resid(MODE) && ep.vec_ok) {
        // pull this warp's share of the residual tile towards L2 while the MMAs of the tile are still running
        const int col = n0 + half * (BN / 2);
        if (col < N) {
          const int r = m0 + q * 32 + lane;
          if (r < M) {
            const float* p = ep.resid + (size_t)r * ep.ldo + col;
#pragma unroll
            for (int i = 0; i < (BN / 2) * 4 / 128; ++i)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i * 32));
          }
        }
      }
      const RowCtx rc = make_row_ctx<MODE>(ep, m0 + q * 32 + lane, M);   // (LN fold: the row statistics load while we wait)
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<BN, MODE>(ep, m0, n0, M, N, tmem_base + (uint32_t)(acc * BN), q, half * (BN / 64), BN / 64, lane,
                              stg, bias, rc);
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
// Residual epilogue through TMA (CTA-pair kernel, EPI_RESID / EPI_RESID_LN).
// The epilogue of the residual GEMMs moves 8 bytes per FLOP-free element (fp32 row in, fp32 row out) and was bound by
// the serial instruction latency of its 8 warps: per 32 x 32 chunk ~700 instructions of address arithmetic, predicated
// 16-byte loads / stores and the quarter-warp transpose, at ~15 cycles each. Here a warp's chunk of the residual
// stream arrives by one bulk tensor copy (128B-swizzled, so that "thread = row" shared-memory accesses are conflict
// free), is updated IN PLACE by the thread that owns the row and leaves by one bulk tensor store; the bf16 operand(s)
// for the next GEMM leave the same way. Row clipping at M (and at lnp_rows for the bf16 operands) is the tensor map's.
// Per warp: two 4 KB residual buffers (chunk c in buffer c & 1; the next chunk's tile is requested as soon as the
// stores of the previous chunk have left its buffer - across tiles too, so the first chunk of a tile is in shared
// memory long before its accumulator is) and one 2 KB buffer per bf16 output.
// ------------------------------------------------------------------------------------------------
// finished per-row values of the LN fold (GemmEpi::lnp_rowa): operand words of the aux k step + rho
__device__ __forceinline__ void ln_split_bf16(float x, uint32_t& hi, uint32_t& lo) {   // bf16 pair, raw 16-bit patterns
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  hi = (uint32_t)__bfloat16_as_ushort(h);
  lo = (uint32_t)__bfloat16_as_ushort(l);
}
__host__ __device__ __forceinline__ void ln_row_moments(float s1, float s2, float inv_d, float& mu, float& rho, float& sg) {
  mu = s1 * inv_d;
  const float var = fmaxf(fmaf(-mu, mu, s2 * inv_d), 0.f) + 1e-6f;
  rho = rsqrtf(var);
  sg = var * rho;   // 1 / rho
}
__device__ __forceinline__ uint4 ln_row_finalize(float s1, float s2, float inv_d) {
  float mu, rho, sg;
  ln_row_moments(s1, s2, inv_d, mu, rho, sg);
  uint32_t mh, ml, sh, sl;
  ln_split_bf16(-mu, mh, ml);
  ln_split_bf16(sg, sh, sl);
  return make_uint4(mh | (mh << 16), ml | (sh << 16), sh | (sl << 16), __float_as_uint(rho));
}
// row pass behind the producers that do not finalise in their epilogue (single-CTA kernels, the embedding modes)
__global__ void ln_rowstat_kernel(const float2* __restrict__ stats, long long pitch, int rows, int slots, float inv_d,
                                  uint4* __restrict__ rowa, const float* __restrict__ moff_in, float* __restrict__ moff_out) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  float s1 = 0.f, s2 = 0.f;
  const float2* sp = stats + row;
  for (int k0 = 0; k0 < slots; k0 += 8) {   // 8 loads in flight, added in slot order
    float2 p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j] = k0 + j < slots ? __ldg(sp + (size_t)(k0 + j) * pitch) : make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1 += p[j].x;
      s2 += p[j].y;
    }
  }
  rowa[row] = ln_row_finalize(s1, s2, inv_d);
  // the row's mean (offset the producer subtracted + mean of what was left): the next site's offset
  if (moff_out != nullptr) moff_out[row] = (moff_in != nullptr ? moff_in[row] : 0.f) + s1 * inv_d;
}

// L2 policies on the TMA epilogue's traffic (evict-first for the fp32 residual stream, evict-last for the bf16 operands the
// next GEMM reads): measured twice in one job, 33.44 / 33.56 with against 33.53 / 33.67 samples/s without - off.
#ifndef MMDM_TMA_EPI_HINTS
#define MMDM_TMA_EPI_HINTS 0
#endif
struct EpiMaps {
  CUtensorMap r, o, h0, h1;   // residual in, fp32 out: box 32 x 32, 128B swizzle; bf16 out: box 32 x 32, 64B swizzle
};
constexpr uint32_t kTmaEpiWarpBytes = 2 * 4096 + 2 * 2048;

struct TmaEpi {
  uint8_t* buf;    // the warp's slice: R0 | R1 | H0 | H1
  uint64_t* bar;   // [2]: residual tile landed in R0 / R1
  uint32_t ph;     // bit b: parity of the next completion of bar[b]
  uint64_t pol_stream, pol_keep;   // L2 policies: the fp32 residual stream is not touched again for several kernels
                                   // (evict first), the bf16 operand is the next kernel's A matrix (evict last)
};
__device__ __forceinline__ void tma_epi_policies(TmaEpi& te) {
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(te.pol_stream));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(te.pol_keep));
}
__device__ __forceinline__ void tma_epi_request(const EpiMaps& mp, const TmaEpi& te, int b, int col, int row) {   // one lane
  mbar_arrive_expect_tx(&te.bar[b], 4096);
  if (MMDM_TMA_EPI_HINTS) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(te.buf + b * 4096)), "l"(reinterpret_cast<uint64_t>(&mp.r)), "r"(smem_u32(&te.bar[b])), "r"(col), "r"(row),
        "l"(te.pol_stream)
        : "memory");
  } else {
    tma_load_2d(te.buf + b * 4096, &mp.r, &te.bar[b], col, row);
  }
}
__device__ __forceinline__ void tma_epi_store(const CUtensorMap* m, uint32_t src, int c0, int c1, uint64_t pol) {
  if (MMDM_TMA_EPI_HINTS) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(src), "r"(c0), "r"(c1), "l"(pol)
                 : "memory");
  } else {
    tma_store_2d(m, src, c0, c1);
  }
}

template <int MODE>
__device__ __forceinline__ void epilogue_tile_tma(const GemmEpi& ep, const EpiMaps& mp, TmaEpi& te, int m0, int n0, int M, int N,
                                                  uint32_t tmem_acc, int q, int c0, int cnt, int lane, int nxt_row0,
                                                  int nxt_col0, float moff) {
  const int row0 = m0 + q * 32, row = row0 + lane;
  const uint32_t taddr = tmem_acc + (uint32_t)(c0 * 32) + ((uint32_t)(q * 32) << 16);
  const uint32_t sbuf = smem_u32(te.buf);
  constexpr bool LN = mode_ln_producer(MODE);
  const bool copy16 = LN || ep.out_bf16 != nullptr;      // H0 in use
  const bool dual = LN && ep.lnp_out1 != nullptr;        // H1 in use
  bool ln_row = false;
  const float* sc0 = nullptr;
  const float* sc1 = nullptr;
  float s1 = 0.f, s2 = 0.f;   // (moff: this row's offset m_r, loaded by the caller while the accumulator was in the making)
  if (LN) {
    const int lim = min(M, ep.lnp_rows);
    ln_row = row < lim;
    const int sq = min(row, lim - 1) / ep.lnp_T;
    sc0 = ep.lnp_scale0 + (size_t)sq * ep.lnp_ld0;
    if (dual) sc1 = ep.lnp_scale1 + (size_t)sq * ep.lnp_ld1;
  }
  const int sw = lane & 7, sw2 = (lane >> 1) & 3;
  uint32_t r[2][32];
  tmem_ld_32x32b_x32(taddr, r[0]);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c >= cnt) break;
    const int b = c & 1;
    const int col0 = n0 + (c0 + c) * 32;
    if (c + 1 < cnt) tmem_ld_32x32b_x32(taddr + (uint32_t)((c + 1) * 32), r[(c + 1) & 1]);
    if (lane == 0) {
      bulk_wait_group_read0();   // the previous chunk's stores have left shared memory ...
      if (c + 1 < cnt) tma_epi_request(mp, te, b ^ 1, col0 + 32, row0);   // ... its buffer takes the next residual tile
      else if (nxt_col0 >= 0) tma_epi_request(mp, te, b ^ 1, nxt_col0, nxt_row0);
    }
    __syncwarp();
    mbar_wait(&te.bar[b], (te.ph >> b) & 1u);
    te.ph ^= 1u << b;
    const uint32_t R = sbuf + (uint32_t)b * 4096u + (uint32_t)lane * 128u;
    const uint32_t H0 = sbuf + 8192u + (uint32_t)lane * 64u, H1 = H0 + 2048u;
    const float4* bias4 = reinterpret_cast<const float4*>(ep.bias + col0);
    uint2 k0 = make_uint2(0u, 0u), k1 = make_uint2(0u, 0u);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t addr = R + ((uint32_t)(j ^ sw) << 4);
      const float4 a = lds128(addr);
      const float4 bb = ep.bias != nullptr ? __ldg(bias4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 o = make_float4(a.x + __uint_as_float(r[b][4 * j]) + bb.x, a.y + __uint_as_float(r[b][4 * j + 1]) + bb.y,
                                   a.z + __uint_as_float(r[b][4 * j + 2]) + bb.z, a.w + __uint_as_float(r[b][4 * j + 3]) + bb.w);
      sts128(addr, o);
      const float4 os = make_float4(o.x - moff, o.y - moff, o.z - moff, o.w - moff);
      if (LN) {
        s1 += (os.x + os.y) + (os.z + os.w);
        s2 += fmaf(os.x, os.x, os.y * os.y) + fmaf(os.z, os.z, os.w * os.w);
      }
      if (copy16) {
        uint2 p0, p1 = make_uint2(0u, 0u);
        if (LN) {
          p0 = ln_prod_pack(os, __ldg(reinterpret_cast<const float4*>(sc0 + col0) + j));
          if (dual) p1 = ln_prod_pack(os, __ldg(reinterpret_cast<const float4*>(sc1 + col0) + j));
        } else {
          p0.x = pack_bf16x2(o.x, o.y);
          p0.y = pack_bf16x2(o.z, o.w);
        }
        if (j & 1) {   // 8 bf16 = one 16-byte slot of the 64B-swizzled row
          const uint32_t slot = (uint32_t)((j >> 1) ^ sw2) << 4;
          sts128u(H0 + slot, make_uint4(k0.x, k0.y, p0.x, p0.y));
          if (dual) sts128u(H1 + slot, make_uint4(k1.x, k1.y, p1.x, p1.y));
        } else {
          k0 = p0;
          k1 = p1;
        }
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_epi_store(&mp.o, sbuf + (uint32_t)b * 4096u, col0, row0, te.pol_stream);
      if (copy16) tma_epi_store(&mp.h0, sbuf + 8192u, col0, row0, te.pol_keep);
      if (dual) tma_epi_store(&mp.h1, sbuf + 10240u, col0, row0, te.pol_keep);
      bulk_commit_group();
    }
    if (LN && (c & 1)) {   // two chunks = one 64-column statistics slot
      if (ln_row) ep.lnp_stats[(size_t)((col0 - 32) / kLnSlotCols) * ep.lnp_pitch + row] = make_float2(s1, s2);
      s1 = s2 = 0.f;
    }
    if (c + 1 < cnt) tmem_ld_wait();
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair kernel (cta_group::2): one 256 x 256 output tile per pair of SMs. Each CTA loads its own 128 rows of
// A and HALF of the B tile (128 of the 256 weight rows); tcgen05.mma.cta_group::2 reads both halves across the
// pair, so a CTA pulls 32 KB instead of 48 KB through L2 per 64-wide k block - the 128 x 256 single-CTA kernel is
// bound by exactly that traffic. The leader CTA (rank 0) issues the MMAs and owns the full / accumulator-empty
// barriers; both CTAs' TMA loads signal the leader's full barrier, the leader's commits are multicast to both
// CTAs' empty / accumulator-full barriers. Epilogue as in the single-CTA kernel (each CTA drains its own TMEM).
// ------------------------------------------------------------------------------------------------
template <int MODE>
struct Gemm2Cfg {
  static constexpr int BM = 128;    // rows per CTA (256 per pair)
  static constexpr int BN = 256;
  static constexpr int BK = 64;
  static constexpr bool TMA_EPI = mode_resid(MODE);   // residual tiles through bulk tensor copies (epilogue_tile_tma)
  // LN-fold consumers: the row / sequence dependent part of the normalisation, - mu_r G_s[n] + Bc_s[n] / rho_r, enters
  // the accumulator through one extra k step of the MMA whose operands the aux warp (warp 3) writes (see the kernel)
  static constexpr bool AUX = mode_ln_consumer(MODE);
  static constexpr int STAGES = TMA_EPI ? 4 : AUX ? 5 : 6;   // (96 KB of epilogue buffers / the aux operands take the room)
  // Epilogue warps: 8 (two per TMEM lane quarter, 4 chunks of 32 columns each); the GELU epilogue costs ~20
  // instructions per element and paced FFN-1 at 55 % tensor-pipe activity, so it gets 12 (3 + 3 + 2 chunks).
  static constexpr int EPI_WARPS = mode_gelu(MODE) ? 12 : 8;   // (12 for the QKV-class epilogue: measured, no change)
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = (BN / 2) * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
  // per-warp staging: a 32 x 32 fp32 transpose slab for the full-line store modes, else just the bias slice
  static constexpr uint32_t STG_WARP_BYTES =
      TMA_EPI ? kTmaEpiWarpBytes : (transposed_mode(MODE) || MODE == EPI_FINAL) ? 4096 : mode_ln_consumer(MODE) ? 2048 : 512;
  static constexpr size_t STG_BYTES = (size_t)EPI_WARPS * STG_WARP_BYTES;
  static constexpr size_t SMEM_BYTES =
      (size_t)(STAGES + (AUX ? 1 : 0)) * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + STG_BYTES;
};

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gemm2Cfg<MODE>::THREADS, 1)
gemm_bf16_tn_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ EpiMaps emaps, const GemmEpi ep, int M, int N, int K) {
  using C = Gemm2Cfg<MODE>;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // [pipeline stages][TMA epilogue: per-warp buffers, 1024-byte aligned][barriers, 256 B][other modes: per-warp staging]
  // (AUX: one more stage-sized buffer behind the ring holds the aux warp's operands)
  constexpr size_t kRing = (size_t)(C::STAGES + (C::AUX ? 1 : 0)) * C::STAGE_BYTES;
  uint8_t* const aux_ptr = smem + (size_t)C::STAGES * C::STAGE_BYTES;
  uint8_t* const stg_ptr = smem + kRing + (C::TMA_EPI ? 0 : 256);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kRing + (C::TMA_EPI ? C::STG_BYTES : 0));
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* rbar = tempty_bar + 3;   // TMA epilogue: [8 warps][2] "residual tile landed" (fits the 256-byte barrier area at 4 stages)
  // AUX: the operands of tile t live in k columns [16 p, 16 p + 16) of the aux buffer, p = t & 1, so the aux warps work
  // a whole tile ahead of the MMA; every barrier exists per p and completes once per two tiles
  uint64_t* aux_ready = tempty_bar + 3;   // [2] (leader CTA): the other three aux warps' quarters are written
  uint64_t* aux_full = tempty_bar + 5;    // [2] (leader CTA): all of it is
  uint64_t* aux_empty = tempty_bar + 7;   // [2] (both CTAs): the aux k step that read it has been executed
  const uint32_t stg_base = smem_u32(stg_ptr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)blockIdx.x >> 1, npairs = (int)gridDim.x >> 1;
  const int num_m = (M + 2 * C::BM - 1) / (2 * C::BM);
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + C::BK - 1) / C::BK;
  // Work items. The tiles that fill whole waves of the grid are 256 x 256; a short last wave (e.g. 8 of 600 tiles
  // on 74 pairs - it would cost a full tile time with 66 pairs idle) is cut into 256 x 64 quarter tiles that spread
  // over the grid. A quarter tile re-reads its A block and over-fetches B (the box stays 128 rows), so it is only
  // used where it removes idle time.
  const int full_tiles = (num_tiles / npairs) * npairs;
  const int rem_tiles = num_tiles - full_tiles;
  const bool split_tail = ep.split_tail && full_tiles > 0 && rem_tiles > 0 && 4 * rem_tiles <= npairs;
  const int num_items = split_tail ? full_tiles + 4 * rem_tiles : num_tiles;
  auto item_of = [&](int it, int& m_idx, int& n0, int& ncols) {
    int tile = it, quarter = 0;
    ncols = BN;
    if (split_tail && it >= full_tiles) {
      const int sub = it - full_tiles;
      tile = full_tiles + (sub >> 2);
      quarter = sub & 3;
      ncols = BN / 4;
    }
    if (ep.rev) tile = num_tiles - 1 - tile;   // walk the tiles from the last row block down (see GemmEpi::rev)
    m_idx = tile / num_n;
    n0 = (tile % num_n) * BN + quarter * (BN / 4);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (C::TMA_EPI) {
      tma_prefetch_desc(&emaps.r);
      tma_prefetch_desc(&emaps.o);
      tma_prefetch_desc(&emaps.h0);
      tma_prefetch_desc(&emaps.h1);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * C::EPI_WARPS);   // the epilogue warps of both CTAs arrive on the leader's barrier
    }
    if (C::TMA_EPI)
      for (int i = 0; i < 2 * C::EPI_WARPS; ++i) mbar_init(&rbar[i], 1);
    if (C::AUX) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&aux_ready[i], 3);   // the other three aux warps of the pair (the leader's warp 3 waits for them)
        mbar_init(&aux_full[i], 1);
        mbar_init(&aux_empty[i], 1);
      }
    }
    fence_barrier_init();
  }
  pdl_trigger();   // the next kernel of the stream may take this SM as soon as this CTA is gone
  if (warp == 2) tmem_alloc_2sm<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = lds32(smem_u32(tmem_slot));
  pdl_wait();      // everything above overlapped the previous kernel's tail; its results are needed from here on

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = pair; it < num_items; it += npairs) {
        int m_idx, n_item, ncols;
        item_of(it, m_idx, n_item, ncols);
        const int m0 = m_idx * (2 * C::BM) + (int)rank * C::BM;
        const int n0 = n_item + (int)rank * (ncols / 2);   // each CTA supplies half of the item's weight rows
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
          tma_load_2d_2sm(sa, &tmA, fb, kb * C::BK, m0);
          tma_load_2d_2sm(sb, &tmB, fb, kb * C::BK, n0);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ---- aux warps (LN-fold consumers; each takes half of the CTA's rows of A2 and B2) ----------------------------------------------------------------------
    // y[r, n] = rho_r (acc[r, n] - mu_r G_s[n] + sigma_r Bc_s[n]), sigma = 1 / rho, s = sequence of row r. The bracket's
    // second part is a rank-12 product over the (at most two, T >= 256) sequences a, b of the pair's 256 rows:
    //   A2[r, :] = [s_r == a] (-mu_h, -mu_h, -mu_l, sg_h, sg_h, sg_l, 0 ...) + [s_r == b] (0 x 6, the same six, 0 ...)
    //   B2[n, :] = (Ga_h, Ga_l, Ga_h, Bca_h, Bca_l, Bca_h, Gb_h, Gb_l, Gb_h, Bcb_h, Bcb_l, Bcb_h, 0 ...)
    // with bf16 pairs x_h + x_l of every factor (products of the pairs' leading three terms: 2^-16 relative). This warp
    // writes both into a stage-sized buffer of its own - 16 of 64 k columns, K-major 128B-swizzled rows like the TMA
    // tiles - and the MMA warp issues one more k step on it, so the epilogue only multiplies by rho_r. (Own buffer and
    // barriers, one phase per tile: a warp that does not take part in every block of the ring cannot wait on the ring's
    // barriers - a parity wait is only meaningful within one phase of the barrier's current one.)
    if (C::AUX) {
      float* const rho_sm = reinterpret_cast<float*>(stg_ptr);   // [8 tiles in flight][128 rows]
      int titer = 0;
      for (int it = pair; it < num_items; it += npairs, ++titer) {
        int m_idx, n_item, ncols;
        item_of(it, m_idx, n_item, ncols);
        const int m0 = m_idx * (2 * C::BM) + (int)rank * C::BM;
        const int nb0 = n_item + (int)rank * (ncols / 2);
        // (a, b): the sequences of the PAIR's 256 rows - both CTAs' rows meet both CTAs' halves of B2
        const int sa = min(m_idx * (2 * C::BM), M - 1) / ep.T, sb = min(m_idx * (2 * C::BM) + 2 * C::BM - 1, M - 1) / ep.T;
        const int ib = 2 * (warp - 2);   // this warp: rows / B2 rows lane + 32 (ib + {0, 1})
        // (this warp runs ~1 instruction per 16 cycles and has a tile's time, ~8000 cycles, per item: the rows' finished
        // values come as one 16-byte load each, GemmEpi::lnc_rowa)
        uint4 a0[2], a1[2], b0[2], b1[2];
        float rho_v[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int row = m0 + lane + 32 * (ib + i);
          const uint4 ra = __ldg(ep.lnc_rowa + min(row, M - 1));
          const bool valid = row < M, second = valid && row / ep.T != sa;
          const uint32_t w0 = valid ? ra.x : 0u, w1 = valid ? ra.y : 0u, w2 = valid ? ra.z : 0u;
          rho_v[i] = valid ? __uint_as_float(ra.w) : 0.f;
          a0[i] = second ? make_uint4(0u, 0u, 0u, w0) : make_uint4(w0, w1, w2, 0u);
          a1[i] = second ? make_uint4(w1, w2, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
          const int li = lane + 32 * (ib + i), n = nb0 + li;
          float ga = 0.f, ba = 0.f, gb = 0.f, bb = 0.f;
          if (li < ncols / 2 && n < N) {
            ga = __ldg(ep.lnc_g + (size_t)sa * ep.lnc_ld + n);
            ba = __ldg(ep.lnc_b + (size_t)sa * ep.lnc_ld + n);
            gb = __ldg(ep.lnc_g + (size_t)sb * ep.lnc_ld + n);
            bb = __ldg(ep.lnc_b + (size_t)sb * ep.lnc_ld + n);
          }
          uint32_t gh, gl, bh, bl;
          ln_split_bf16(ga, gh, gl);
          ln_split_bf16(ba, bh, bl);
          const uint32_t x0 = gh | (gl << 16), x1 = gh | (bh << 16), x2 = bl | (bh << 16);
          ln_split_bf16(gb, gh, gl);
          ln_split_bf16(bb, bh, bl);
          b0[i] = make_uint4(x0, x1, x2, gh | (gl << 16));
          b1[i] = make_uint4(gh | (bh << 16), bl | (bh << 16), 0u, 0u);
        }
        const uint32_t ap = (uint32_t)titer & 1u, aph = ((uint32_t)titer >> 1) & 1u;
        mbar_wait(&aux_empty[ap], aph ^ 1u);   // the aux k step of tile t - 2 has read these columns
        const uint32_t pa = smem_u32(aux_ptr), pb = pa + C::A_BYTES;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t r = (uint32_t)(lane + 32 * (ib + i)), sw = r & 7u;
          sts128u(pa + r * 128u + (((2u * ap) ^ sw) << 4), a0[i]);
          sts128u(pa + r * 128u + (((2u * ap + 1u) ^ sw) << 4), a1[i]);
          sts128u(pb + r * 128u + (((2u * ap) ^ sw) << 4), b0[i]);
          sts128u(pb + r * 128u + (((2u * ap + 1u) ^ sw) << 4), b1[i]);
          rho_sm[(titer & 7) * C::BM + (int)r] = rho_v[i];
        }
        fence_proxy_async_smem();   // generic-proxy writes -> the tensor core's (async proxy) reads
        __syncwarp();
        if (lane == 0) {
          if (rank != 0) {
            mbar_arrive_cluster_release(mapa_u32(smem_u32(&aux_ready[ap]), 0));
          } else if (warp == 2) {
            mbar_arrive(&aux_ready[ap]);
          } else {   // one arrival per tile, after all four quarters are written
            mbar_wait_cluster(&aux_ready[ap], aph);
            mbar_arrive(&aux_full[ap]);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc_full = umma_idesc_bf16(2 * C::BM, BN), idesc_quarter = umma_idesc_bf16(2 * C::BM, BN / 4);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // The issuing thread must stay well ahead of the tensor pipe (4 MMAs = 512 cycles per k block): descriptors
      // are advanced by a running stage offset (start address field, 16-byte units) instead of being rebuilt from
      // the stage index, which cost ~25 dependent uniform-datapath instructions per k block.
      const uint64_t da0 = umma_desc_k_sw128(smem_u32(smem));
      const uint64_t db0 = umma_desc_k_sw128(smem_u32(smem) + C::A_BYTES);
      uint32_t soff = 0, aux_t = 0;
      const uint64_t da_aux = umma_desc_k_sw128(smem_u32(aux_ptr)), db_aux = umma_desc_k_sw128(smem_u32(aux_ptr) + C::A_BYTES);
      for (int it = pair; it < num_items; it += npairs) {
        const uint32_t idesc = (split_tail && it >= full_tiles) ? idesc_quarter : idesc_full;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = da0 + soff, db = db0 + soff;
#pragma unroll
          for (int k = 0; k < C::BK / 16; ++k)
            umma_bf16_ss_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage], 3);
          soff += C::STAGE_BYTES >> 4;
          if (++stage == C::STAGES) {
            stage = 0;
            soff = 0;
            phase ^= 1;
          }
        }
        if (C::AUX) {   // one k step (16 columns) on the aux warp's operands
          const uint32_t ap = aux_t & 1u;
          mbar_wait(&aux_full[ap], (aux_t >> 1) & 1u);
          ++aux_t;
          tc_fence_after();
          umma_bf16_ss_2sm(d_tmem, da_aux + (uint64_t)(ap * 2u), db_aux + (uint64_t)(ap * 2u), idesc, 1u);
          umma_commit_2sm(&aux_empty[ap], 3);
        }
        umma_commit_2sm(&tfull_bar[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int part = (warp - 4) >> 2;                       // column part of the tile this warp drains
    constexpr int NPART = C::EPI_WARPS / 4, NCHUNK = BN / 32;
    // chunks (32 columns) of an item this warp drains: a full tile has 8, a quarter tile 2
    auto my_chunks = [&](int ncols, int& c0, int& cnt) {
      if (ncols == BN) {
        c0 = NPART == 2 ? part * 4 : part * 3;
        cnt = NPART == 2 ? 4 : (part < 2 ? 3 : NCHUNK - 6);
      } else if (mode_ln_producer(MODE) || C::TMA_EPI) {   // a quarter tile is one 64-column statistics slot (and an even
                                                           // number of chunks for the TMA epilogue): one warp per lane quarter
        c0 = 0;
        cnt = part == 0 ? 2 : 0;
      } else {
        c0 = part;
        cnt = part < 2 ? 1 : 0;
      }
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t tempty_leader = mapa_u32(smem_u32(&tempty_bar[0]), 0);
    const uint32_t stg = stg_base + (uint32_t)(warp - 4) * C::STG_WARP_BYTES;
    TmaEpi te;
    te.buf = stg_ptr + (size_t)(warp - 4) * C::STG_WARP_BYTES;
    te.bar = rbar + 2 * (warp - 4);
    te.ph = 0;
    te.pol_stream = te.pol_keep = 0;
    if (C::TMA_EPI) tma_epi_policies(te);
    if (C::TMA_EPI && pair < num_items) {   // the first chunk's residual tile
      int mi, ni, nc, c0n, cntn;
      item_of(pair, mi, ni, nc);
      my_chunks(nc, c0n, cntn);
      if (cntn > 0 && lane == 0) tma_epi_request(emaps, te, 0, ni + c0n * 32, mi * (2 * C::BM) + (int)rank * C::BM + q * 32);
    }
    EpiBias<BN> bias = {}, bias_next = {};
    EpiLnc<BN> lnc_next = {};
    if (!transposed_mode(MODE) && !mode_ln_consumer(MODE) && pair < num_items) {
      int mi, ni, nc, c0n, cntn;
      item_of(pair, mi, ni, nc);
      my_chunks(nc, c0n, cntn);
      bias_next = bias_fetch<BN, MODE>(ep, ni, c0n, lane, N);
    }
    if (mode_ln_consumer(MODE) && !C::AUX && ep.vec_ok && pair < num_items) {
      int mi, ni, nc, c0n, cntn;
      item_of(pair, mi, ni, nc);
      my_chunks(nc, c0n, cntn);
      lnc_next = lnc_fetch<BN>(ep, mi * (2 * C::BM) + (int)rank * C::BM + q * 32, ni, c0n, cntn, lane, M, N);
    }
    int titer = 0;
    for (int it = pair; it < num_items; it += npairs, ++titer) {
      int m_idx, n0, ncols, c0, cnt;
      item_of(it, m_idx, n0, ncols);
      my_chunks(ncols, c0, cnt);
      const int m0 = m_idx * (2 * C::BM) + (int)rank * C::BM;
      if (mode_ln_consumer(MODE) && !C::AUX && ep.vec_ok) {
        lnc_publish<BN>(lnc_next, stg, cnt, lane);
        if (it + npairs < num_items) {
          int mi, ni, nc, c0n, cntn;
          item_of(it + npairs, mi, ni, nc);
          my_chunks(nc, c0n, cntn);
          lnc_next = lnc_fetch<BN>(ep, mi * (2 * C::BM) + (int)rank * C::BM + q * 32, ni, c0n, cntn, lane, M, N);
        }
      }
      int nxt_row0 = 0, nxt_col0 = -1;   // TMA epilogue: the first chunk of this warp's next item
      if (C::TMA_EPI && it + npairs < num_items) {
        int mi, ni, nc, c0n, cntn;
        item_of(it + npairs, mi, ni, nc);
        my_chunks(nc, c0n, cntn);
        if (cntn > 0) {
          nxt_row0 = mi * (2 * C::BM) + (int)rank * C::BM + q * 32;
          nxt_col0 = ni + c0n * 32;
        }
      }
      if (C::TMA_EPI) {
      } else if (transposed_mode(MODE)) {
        bias = bias_fetch<BN, MODE>(ep, n0, c0, lane, N);
      } else if (MODE != EPI_FINAL && !mode_ln_consumer(MODE)) {   // (LN-fold consumers: the bias is part of Bc_s)
        bias_publish<BN>(bias_next, stg, lane);
        if (it + npairs < num_items) {
          int mi, ni, nc, c0n, cntn;
          item_of(it + npairs, mi, ni, nc);
          my_chunks(nc, c0n, cntn);
          bias_next = bias_fetch<BN, MODE>(ep, ni, c0n, lane, N);
        }
      }
      if (mode_resid(MODE) && !C::TMA_EPI && ep.vec_ok && cnt > 0) {
        const int col = n0 + c0 * 32;
        if (col < N) {
          const int r = m0 + q * 32 + lane;
          if (r < M) {
            const float* p = ep.resid + (size_t)r * ep.ldo + col;
            for (int i = 0; i < cnt; ++i)   // one 128-byte line per chunk of the row
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i * 32));
          }
        }
      }
      RowCtx rc = make_row_ctx<MODE, C::AUX>(ep, m0 + q * 32 + lane, M);
      float moff = 0.f;   // LN-fold producers: this row's operand offset (GemmEpi::lnp_moff_in)
      if (mode_ln_producer(MODE) && C::TMA_EPI && ep.lnp_moff_in != nullptr && m0 + q * 32 + lane < min(M, ep.lnp_rows))
        moff = __ldg(ep.lnp_moff_in + m0 + q * 32 + lane);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (C::AUX) rc.rho = reinterpret_cast<const float*>(stg_ptr)[(titer & 7) * C::BM + q * 32 + lane];   // the aux warp's
      if (C::TMA_EPI) {
        if (cnt > 0) epilogue_tile_tma<MODE>(ep, emaps, te, m0, n0, M, N, tmem_base + (uint32_t)(acc * BN), q, c0, cnt, lane, nxt_row0, nxt_col0, moff);
      } else if (cnt > 0) {
        epilogue_tile<BN, MODE, C::AUX>(ep, m0, n0, M, N, tmem_base + (uint32_t)(acc * BN), q, c0, cnt, lane, stg, bias, rc);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + (uint32_t)acc * 8u);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (C::TMA_EPI && lane == 0) bulk_wait_group0();   // the last stores have left shared memory and are complete
  }

  tc_fence_before();
  cluster_sync_all();   // the peer may still read this CTA's shared memory / signal its barriers until here
  if (warp == 2) tmem_dealloc_2sm<C::TMEM_COLS>(tmem_base);
}

template <int MODE>
static int launch_tc2(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
                      const GemmEpi& ep, cudaStream_t stream) {
  using C = Gemm2Cfg<MODE>;
  static PerDeviceFlag attr;   // kernel attributes are per device
  if (attr.first())
    MMDM_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_2cta_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::SMEM_BYTES));
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
    uint32_t box[2] = {(uint32_t)C::BK, (uint32_t)(C::BN / 2)};
    MMDM_TRY(make_tensor_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B));
  }
  EpiMaps em;
  memset(&em, 0, sizeof(em));
  if (C::TMA_EPI) {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t st32[1] = {(uint64_t)ep.ldo * 4}, st16[1] = {(uint64_t)ep.ldo * 2};
    uint32_t box[2] = {32, 32};
    MMDM_TRY(make_tensor_map(&em.r, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ep.resid, dims, st32, box, CU_TENSOR_MAP_SWIZZLE_128B));
    MMDM_TRY(make_tensor_map(&em.o, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ep.out, dims, st32, box, CU_TENSOR_MAP_SWIZZLE_128B));
    em.h0 = em.o;
    em.h1 = em.o;
    const void* b0 = mode_ln_producer(MODE) ? (const void*)ep.lnp_out0 : (const void*)ep.out_bf16;
    if (mode_ln_producer(MODE)) dims[1] = (uint64_t)(ep.lnp_rows < M ? ep.lnp_rows : M);   // rows beyond get no operand
    if (b0 != nullptr)
      MMDM_TRY(make_tensor_map(&em.h0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b0, dims, st16, box, CU_TENSOR_MAP_SWIZZLE_64B));
    if (mode_ln_producer(MODE) && ep.lnp_out1 != nullptr)
      MMDM_TRY(make_tensor_map(&em.h1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ep.lnp_out1, dims, st16, box, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  const int tiles = ((M + 2 * C::BM - 1) / (2 * C::BM)) * ((N + C::BN - 1) / C::BN);
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  MMDM_CHECK_CUDA(launch_maybe_pdl(gemm_bf16_tn_2cta_kernel<MODE>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, stream, tmA, tmB, em,
                                   ep, M, N, K));
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int g_gemm_2cta = -1;
static bool use_2cta() {
  if (g_gemm_2cta < 0) {
    const char* e = getenv("MMDM_GEMM_1CTA");   // debug / A-B comparison: force the single-CTA kernel
    g_gemm_2cta = (e && e[0] == '1') ? 0 : 1;
  }
  return g_gemm_2cta == 1;
}

// ------------------------------------------------------------------------------------------------
// SIMT bring-up kernel (same epilogue); used only when MMDM_GEMM_SIMT=1 to bisect tcgen05 problems.
// ------------------------------------------------------------------------------------------------
template <int MODE>
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
  __shared__ float bias_s[32];
  if (threadIdx.x < 32) bias_s[threadIdx.x] = (ep.bias != nullptr && col0 + (int)threadIdx.x < N) ? ep.bias[col0 + threadIdx.x] : 0.f;
  __syncthreads();
  const RowCtx rc = make_row_ctx<MODE>(ep, row, M);
  epilogue_chunk<MODE>(ep, rc, col0, acc, N, smem_u32(bias_s));
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <int BN, int MODE>
static int launch_tc(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
                     const GemmEpi& ep, cudaStream_t stream) {
  using C = GemmCfg<BN>;
  static PerDeviceFlag attr;   // kernel attributes are per device
  if (attr.first())
    MMDM_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::SMEM_BYTES));
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
  gemm_bf16_tn_kernel<BN, MODE><<<grid, 384, C::SMEM_BYTES, stream>>>(tmA, tmB, ep, M, N, K);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int MODE>
static int launch_mode(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
                       const GemmEpi& ep, cudaStream_t stream) {
  if (use_simt()) {
    MMDM_REQUIRE(!mode_ln_producer(MODE), "gemm: the SIMT bring-up kernel has no LN-fold producer epilogue");
    dim3 grid((M + 127) / 128, (N + 31) / 32);
    gemm_simt_debug_kernel<MODE><<<grid, 128, 0, stream>>>(A, lda, W, ldw, ep, M, N, K);
    MMDM_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  // CTA pairs on 256x256 tiles for the big projections; 128x256 single-CTA tiles when N fills them but M is
  // small; 128x128 for narrow / ragged N (e.g. the 262-wide output layer)
  // (the residual modes' TMA epilogue needs 16-byte aligned rows; launch_gemm's vec_ok says so)
  // (the LN-fold consumers' aux k step covers two sequences per 256 rows: T >= 256)
  if (N % 256 == 0 && M >= 512 && use_2cta() && (!mode_resid(MODE) || (ep.vec_ok && ep.bias_ok16)) &&
      (!mode_ln_consumer(MODE) || (ep.T >= 256 && ep.vec_ok && ep.lnc_rowa != nullptr)))
    return launch_tc2<MODE>(A, lda, W, ldw, M, N, K, ep, stream);
  if (N % 256 == 0 || N >= 1024) {
    // One row of tiles (the conditioning GEMMs: M = sequences, weight-streaming bound): take the tile width that
    // fills the last wave better, e.g. 81920 columns = 320 wide tiles (2.16 waves on 148 SMs) or 640 narrow ones (4.32)
    if (M <= 128 && N % 128 == 0) {
      const int sms = num_sms();
      const int t256 = (N + 255) / 256, t128 = N / 128;
      const double e256 = (double)t256 / (((t256 + sms - 1) / sms) * sms), e128 = (double)t128 / (((t128 + sms - 1) / sms) * sms);
      if (e128 > e256 + 0.05) return launch_tc<128, MODE>(A, lda, W, ldw, M, N, K, ep, stream);
    }
    return launch_tc<256, MODE>(A, lda, W, ldw, M, N, K, ep, stream);
  }
  return launch_tc<128, MODE>(A, lda, W, ldw, M, N, K, ep, stream);
}

// Kernel attributes are set lazily, once per instantiation and device, at the first launch. cudaFuncSetAttribute is
// not a stream operation, so it is legal while mmdm_sample captures its step graph (which it does on first use).
int gemm_init() { return 0; }

static int launch_gemm_inner(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
                             const GemmEpi& ep, cudaStream_t stream);

int launch_gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
                const GemmEpi& ep, cudaStream_t stream) {
  MMDM_TRY(launch_gemm_inner(A, lda, W, ldw, M, N, K, ep, stream));
  if (mode_ln_producer(ep.mode) && ep.lnp_rowa != nullptr) {
    // The finished per-row values come from a row pass behind the GEMM (~3 us). Finishing them inside the epilogue - the
    // warp that completes a 32-row group's slots, found with a device-wide counter - was measured: the fence in front
    // of the counter costs every warp ~1-2 us per tile, 16-23 us per launch.
    const int lim = (ep.lnp_rows > 0 && ep.lnp_rows < M) ? ep.lnp_rows : M;
    const int rows = ep.dup_row_offset != 0 ? (int)ep.dup_row_offset + lim : lim;
    ln_rowstat_kernel<<<(rows + 63) / 64, 64, 0, stream>>>(ep.lnp_stats, ep.lnp_pitch, rows, N / kLnSlotCols, 1.0f / (float)N,
                                                            ep.lnp_rowa, ep.lnp_moff_in, ep.lnp_moff_out);
    MMDM_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

static int launch_gemm_inner(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
                             const GemmEpi& ep, cudaStream_t stream) {
  MMDM_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  MMDM_REQUIRE((lda % 8) == 0 && (ldw % 8) == 0, "gemm: lda/ldw must be multiples of 8 (got %d, %d)", lda, ldw);
  MMDM_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
               "gemm: operands must be 16-byte aligned");
  MMDM_REQUIRE(ep.bias == nullptr || (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0, "gemm: bias must be 16-byte aligned");
  GemmEpi e2 = ep;
  static int split_tail = -1;
  if (split_tail < 0) {
    const char* e = getenv("MMDM_GEMM_SPLIT_TAIL");   // A/B measurements: 0 keeps whole tiles in the last wave
    split_tail = (e && e[0] == '0') ? 0 : 1;
  }
  e2.split_tail = split_tail;
  e2.vec_ok = ((reinterpret_cast<uintptr_t>(ep.out) & 15) == 0 && (ep.ldo % 8) == 0 &&
               (ep.resid == nullptr || (reinterpret_cast<uintptr_t>(ep.resid) & 15) == 0) &&
               (ep.out_bf16 == nullptr || (reinterpret_cast<uintptr_t>(ep.out_bf16) & 15) == 0) &&
               (ep.pe == nullptr || ((reinterpret_cast<uintptr_t>(ep.pe) & 15) == 0 && (ep.pe_ld % 4) == 0)) &&
               (!mode_ln_consumer(ep.mode) || ((reinterpret_cast<uintptr_t>(ep.lnc_g) & 15) == 0 &&
                                               (reinterpret_cast<uintptr_t>(ep.lnc_b) & 15) == 0 && (ep.lnc_ld % 4) == 0)))
                  ? 1 : 0;
  if (mode_ln_consumer(ep.mode)) {
    MMDM_REQUIRE(ep.lnc_stats && ep.lnc_g && ep.lnc_b && ep.lnc_slots >= 1 && ep.lnc_slots * kLnSlotCols == K && ep.T >= 1 &&
                     ep.lnc_pitch >= M,
                 "gemm: LN-fold consumer needs statistics (K / 64 slots), G and Bc rows and the sequence length");
    e2.lnc_inv_d = 1.0f / (float)K;
  }
  if (mode_ln_producer(ep.mode)) {
    if (e2.lnp_rows <= 0 || e2.lnp_rows > M) e2.lnp_rows = M;
    MMDM_REQUIRE(ep.lnp_moff_out == nullptr || ep.lnp_rowa != nullptr, "gemm: lnp_moff_out is written by the row pass that fills lnp_rowa");
    MMDM_REQUIRE(ep.lnp_stats && ep.lnp_scale0 && ep.lnp_out0 && ep.lnp_T >= 1 && (N % kLnSlotCols) == 0 && ep.ldo == N &&
                     ep.lnp_pitch >= e2.lnp_rows + ep.dup_row_offset && (ep.lnp_ld0 % 4) == 0 && (ep.lnp_ld1 % 4) == 0 &&
                     (reinterpret_cast<uintptr_t>(ep.lnp_scale0) & 15) == 0 && (reinterpret_cast<uintptr_t>(ep.lnp_scale1) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(ep.lnp_out0) & 15) == 0 && (reinterpret_cast<uintptr_t>(ep.lnp_out1) & 15) == 0 &&
                     ep.out_bf16 == nullptr &&
                     (ep.lnp_out1 == nullptr || ep.lnp_scale1 != nullptr) && e2.vec_ok,
                 "gemm: LN-fold producer needs a dense, aligned (M, N %% 64 == 0) output, scale rows and a statistics buffer");
  }
  e2.bias_ok16 = 1;   // (bias alignment is required above)
  e2.final_vec_ok = (ep.mode == EPI_FINAL && (reinterpret_cast<uintptr_t>(ep.out) & 7) == 0 && (ep.nfeats % 2) == 0) ? 1 : 0;
  switch (ep.mode) {
    case EPI_BF16:
      if (ep.vt_col0 >= 0)
        MMDM_REQUIRE((ep.vt_col0 % 32) == 0 && (ep.hd % 32) == 0 && ep.vt != nullptr, "gemm: V^T split must be 32-aligned");
      return launch_mode<EPI_BF16>(A, lda, W, ldw, M, N, K, e2, stream);
    case EPI_GELU_BF16: return launch_mode<EPI_GELU_BF16>(A, lda, W, ldw, M, N, K, e2, stream);
    case EPI_RESID:
      MMDM_REQUIRE(ep.resid != nullptr, "gemm: residual pointer missing");
      return launch_mode<EPI_RESID>(A, lda, W, ldw, M, N, K, e2, stream);
    case EPI_EMBED_PE:
      MMDM_REQUIRE(ep.pe != nullptr, "gemm: pe table missing");
      return launch_mode<EPI_EMBED_PE>(A, lda, W, ldw, M, N, K, e2, stream);
    case EPI_FINAL: return launch_mode<EPI_FINAL>(A, lda, W, ldw, M, N, K, e2, stream);
    case EPI_F32: return launch_mode<EPI_F32>(A, lda, W, ldw, M, N, K, e2, stream);
    case EPI_SIGMOID_F32: return launch_mode<EPI_SIGMOID_F32>(A, lda, W, ldw, M, N, K, e2, stream);
    case EPI_QGELU_BF16: return launch_mode<EPI_QGELU_BF16>(A, lda, W, ldw, M, N, K, e2, stream);
    case EPI_BF16_LN:
      if (ep.vt_col0 >= 0)
        MMDM_REQUIRE((ep.vt_col0 % 32) == 0 && (ep.hd % 32) == 0 && ep.vt != nullptr, "gemm: V^T split must be 32-aligned");
      return launch_mode<EPI_BF16_LN>(A, lda, W, ldw, M, N, K, e2, stream);
    case EPI_GELU_BF16_LN: return launch_mode<EPI_GELU_BF16_LN>(A, lda, W, ldw, M, N, K, e2, stream);
    case EPI_RESID_LN:
      MMDM_REQUIRE(ep.resid != nullptr, "gemm: residual pointer missing");
      return launch_mode<EPI_RESID_LN>(A, lda, W, ldw, M, N, K, e2, stream);
    case EPI_EMBED_PE_LN:
      MMDM_REQUIRE(ep.pe != nullptr, "gemm: pe table missing");
      return launch_mode<EPI_EMBED_PE_LN>(A, lda, W, ldw, M, N, K, e2, stream);
    default: break;
  }
  set_last_error("gemm: unknown epilogue mode %d", ep.mode);
  return -1;
}

}  // namespace mmdm
