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
};

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
  // Walk the output tiles from the last row block down (CTA-pair kernel only). The engine alternates the direction
  // from launch to launch so that a kernel starts on the rows its producer wrote last, which are still in L2.
  int rev = 0;
  int split_tail = 0;                    // set by launch_gemm: cut a short last wave into quarter tiles (CTA-pair kernel)
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
