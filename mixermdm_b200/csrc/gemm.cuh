// GEMM interface shared by the engine: D[M,N] = A[M,K] (bf16, row-major) * W[N,K]^T (bf16, row-major,
// i.e. the nn.Linear weight layout) with a fused epilogue. fp32 accumulation in TMEM.
#pragma once
#include "common.cuh"

namespace mmdm {

enum EpiMode : int {
  EPI_BF16 = 0,       // out_bf16 = acc + bias                         (QKV / q / kv projections)
  EPI_GELU_BF16 = 1,  // out_bf16 = gelu_erf(acc + bias)               (FFN linear1)
  EPI_RESID = 2,      // out_f32  = resid + acc + bias  [+ bf16 copy]  (attention out-proj, FFN linear2)
  EPI_EMBED_PE = 3,   // out_f32  = acc + bias + pe[row % T]  [dup]    (motion_embed + positional enc.)
  EPI_FINAL = 4,      // out_f32  = (acc + bias) * std + mean, scattered to (B,T,2*nfeats) layout
  EPI_F32 = 5,        // out_f32  = acc + bias                         (AdaLN conditioning GEMM)
  EPI_SIGMOID_F32 = 6,  // out_f32 = sigmoid(acc + bias)               (Influence head, influence.py:120-125)
  EPI_QGELU_BF16 = 7,   // out_bf16 = x * sigmoid(1.702 x), x = acc + bias (CLIP's QuickGELU: text tower MLP)
  // ---- AdaLN folded into the GEMMs on either side of it ("LN fold", see GemmEpi::lnp_* / lnc_*) ----------------
  EPI_BF16_LN = 8,       // EPI_BF16 on the operand bf16(h (1+scale)): out = rho (acc - mu G_s) + Bc_s
  EPI_GELU_BF16_LN = 9,  // EPI_GELU_BF16, same
  EPI_RESID_LN = 10,     // EPI_RESID that also emits bf16(h (1+scale_s)) for the next AdaLN site(s) + row sums
  EPI_EMBED_PE_LN = 11,  // EPI_EMBED_PE, same
};

// AdaLN (reference models/utils/layers.py:15-25): xn = (h - mu) rho (1 + scale_s) + shift_s feeds a Linear W, c:
//   y[r, n] = rho_r * sum_k [h[r,k] (1+scale_s[k])] W[n,k]  -  rho_r mu_r G_s[n]  +  Bc_s[n]
//   G_s[n] = sum_k (1+scale_s[k]) W[n,k],   Bc_s[n] = sum_k shift_s[k] W[n,k] + c[n]       (s = sequence of row r)
// so the GEMM that WRITES h can emit the bf16 operand h (1+scale_s) (no row statistics needed for that) together with
// per-row partial sums, and the GEMM that READS it applies rho, mu, G, Bc in its epilogue: the stand-alone
// LayerNorm pass over the fp32 residual stream (one read + one write of the activation per site) disappears.
__host__ __device__ constexpr bool mode_resid(int m) { return m == EPI_RESID || m == EPI_RESID_LN; }
__host__ __device__ constexpr bool mode_embed(int m) { return m == EPI_EMBED_PE || m == EPI_EMBED_PE_LN; }
__host__ __device__ constexpr bool mode_bf16(int m) { return m == EPI_BF16 || m == EPI_BF16_LN; }
__host__ __device__ constexpr bool mode_gelu(int m) { return m == EPI_GELU_BF16 || m == EPI_GELU_BF16_LN; }
__host__ __device__ constexpr bool mode_ln_producer(int m) { return m == EPI_RESID_LN || m == EPI_EMBED_PE_LN; }
__host__ __device__ constexpr bool mode_ln_consumer(int m) { return m == EPI_BF16_LN || m == EPI_GELU_BF16_LN; }
constexpr int kLnSlotCols = 64;   // columns per partial-sum slot

struct GemmEpi {
  int mode = EPI_BF16;
  const float* bias = nullptr;  // [N] or null
  void* out = nullptr;          // bf16* or float* depending on mode
  int ldo = 0;                  // row pitch of out (elements)
  // EPI_BF16: columns >= vt_col0 are written transposed per (sequence, head): vt[s*H+h][d][t]
  int vt_col0 = -1;
  __nv_bfloat16* vt = nullptr;
  int T = 1, H = 1, hd = 1, Tp = 1;
  // EPI_RESID
  const float* resid = nullptr;          // may alias out
  __nv_bfloat16* out_bf16 = nullptr;     // optional bf16 copy, same pitch
  // EPI_EMBED_PE (uses T)
  const float* pe = nullptr;             // [>=T, N]
  int pe_ld = 0;
  long long dup_row_offset = 0;          // if != 0 also write rows shifted by this amount
  // EPI_FINAL (uses T): v*std[n]+mean[n] -> out[((s>>1)*T + t)*2*nfeats + (s&1)*nfeats + n]
  const float* std = nullptr;
  const float* mean = nullptr;
  int nfeats = 0;
  int persons_per_row = 2;               // 2: sequences alternate persons (s&1); 1: plain (s*T+t)*nfeats
  // LN-fold producer (EPI_RESID_LN / EPI_EMBED_PE_LN; N = latent dim, a multiple of 64; out pitch = ldo for all):
  const float* lnp_scale0 = nullptr;     // AdaLN scale rows of the next site: scale0[s * lnp_ld0 + k], s = row / lnp_T
  int lnp_ld0 = 0;
  __nv_bfloat16* lnp_out0 = nullptr;     // bf16(h * (1 + scale0))
  const float* lnp_scale1 = nullptr;     // optional second site fed by the same rows (SA of a stream + CA-KV of its partner)
  int lnp_ld1 = 0;
  __nv_bfloat16* lnp_out1 = nullptr;
  float2* lnp_stats = nullptr;           // [N / 64][lnp_pitch]: (sum h, sum h^2) of a row over the slot's 64 columns
  long long lnp_pitch = 0;
  int lnp_T = 1;
  int lnp_rows = 0;                      // rows >= lnp_rows write h only (Influence: the m_I half of the embed GEMM)
  // optional: finished per-row values for the CTA-pair consumers' aux k step, rowa[r] = (bf16 pairs of -mu and
  // 1/rho packed as the three operand words, rho as float bits), written by a small row pass behind the GEMM
  uint4* lnp_rowa = nullptr;
  // optional per-row offsets: the operand(s) and the row sums are those of h - m_r, m_r = lnp_moff_in[r] (an estimate of
  // the row mean: the previous AdaLN site's mean of the same row). The consumers need not know - they see a row whose mean
  // is small - and the bf16 rounding of the operand no longer carries the row mean (it would be amplified by
  // sqrt(1 + (mu/sigma)^2)). The row pass stores m_r + mean(h - m_r) = the row's mean in lnp_moff_out for the next site.
  const float* lnp_moff_in = nullptr;
  float* lnp_moff_out = nullptr;
  int lnp_dup_seq = 0;                   // EPI_EMBED_PE_LN with dup_row_offset: sequence offset of the duplicated rows
  // LN-fold consumer (EPI_BF16_LN / EPI_GELU_BF16_LN; uses T): K = latent dim of the normalised rows
  const float2* lnc_stats = nullptr;     // [lnc_slots][lnc_pitch]
  long long lnc_pitch = 0;
  int lnc_slots = 0;
  float lnc_inv_d = 0.f;
  const uint4* lnc_rowa = nullptr;       // the producer's lnp_rowa (needed by the CTA-pair kernel, T >= 256)
  const float* lnc_g = nullptr;          // G_s[n] = lnc_g[s * lnc_ld + n]
  const float* lnc_b = nullptr;          // Bc_s[n] (bias included; GemmEpi::bias is ignored)
  int lnc_ld = 0;
  // Walk the output tiles from the last row block down (CTA-pair kernel only). The engine alternates the direction
  // from launch to launch so that a kernel starts on the rows its producer wrote last, which are still in L2.
  int rev = 0;
  int split_tail = 0;                    // set by launch_gemm: cut a short last wave into quarter tiles (CTA-pair kernel)
  int bias_ok16 = 0;                     // set by launch_gemm
  int vec_ok = 0;                        // set by launch_gemm: pointers/pitches allow 16-byte vector access
  int final_vec_ok = 0;                  // set by launch_gemm: EPI_FINAL rows allow 8-byte vector stores
};

// A: [M, K] bf16 with row pitch lda (elements, lda*2 % 16 == 0); W: [N, K] bf16 pitch ldw.
int launch_gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
                const GemmEpi& ep, cudaStream_t stream);

// debug: 0 = tcgen05 (product path), 1 = SIMT reference kernel (bring-up only; MMDM_GEMM_SIMT=1)
void gemm_set_debug_simt(int on);
int num_sms();
int gemm_init();   // one-time kernel attribute setup (also done lazily)

}  // namespace mmdm
