/* mixermdm_b200 — C ABI of the B200-native MixerMDM denoising loop.
 *
 * The reference (pabloruizponce/MixerMDM) has no FFI: its boundary for this path is the Python model API
 * (src/models/mixermdm.py:490-602, 660-811; src/models/in2in.py:401-462; src/models/utils/influence.py:93-126;
 * src/models/utils/cfg_sampler.py:31-56; src/models/utils/gaussian_diffusion.py:1769-2091). The Python mirror
 * in mixermdm_b200/ keeps those signatures and calls the entry points below through ctypes. Conventions:
 *   - every function returns 0 on success, a negative code on failure; mmdm_last_error() gives the message;
 *   - the caller owns every device buffer (weights, workspace, state); nothing is allocated behind its back
 *     except the handle itself; no host synchronisation; everything is enqueued on `stream`
 *     (a cudaStream_t passed as void*) and is CUDA-graph capturable;
 *   - bf16 buffers are raw uint16 storage (`void*`), fp32 buffers are `float*`.
 */
#ifndef MIXERMDM_B200_H
#define MIXERMDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMDM_ABI_VERSION 1

int mmdm_abi_version(void);
const char* mmdm_last_error(void);
/* bring-up switch: 1 routes GEMMs through a slow SIMT kernel with the same epilogue (tests only). */
void mmdm_debug_gemm_simt(int on);

/* ------------------------------------------------------------------------------------------------
 * Operator level (each replaces one PyTorch library call of the reference; used by the parity tests)
 * ---------------------------------------------------------------------------------------------- */

enum mmdm_epilogue_mode {
  MMDM_EPI_BF16 = 0,      /* out_bf16 = acc + bias                      nn.MultiheadAttention in-proj      */
  MMDM_EPI_GELU_BF16 = 1, /* out_bf16 = gelu_erf(acc + bias)            FFN.linear1 + nn.GELU              */
  MMDM_EPI_RESID = 2,     /* out_f32  = resid + acc + bias (+bf16 copy) out_proj / FFN.linear2 + residual   */
  MMDM_EPI_EMBED_PE = 3,  /* out_f32  = acc + bias + pe[row % T]        motion_embed + PositionalEncoding   */
  MMDM_EPI_FINAL = 4,     /* out_f32  = (acc + bias) * std + mean       FinalLayer + normalizer.backward    */
  MMDM_EPI_F32 = 5        /* out_f32  = acc + bias                      AdaLN.emb_layers[1]                 */
};

typedef struct mmdm_epilogue {
  int32_t mode;
  int32_t ldo;            /* row pitch of out, elements */
  const float* bias;      /* [N] or NULL */
  void* out;
  /* MMDM_EPI_BF16: columns >= vt_col0 go, transposed, to vt[(s*H+h)*hd+d][t] (pitch Tp); -1 disables */
  int32_t vt_col0;
  int32_t T, H, hd, Tp;
  void* vt;
  /* MMDM_EPI_RESID */
  const float* resid;     /* may alias out */
  void* out_bf16;         /* optional bf16 copy (same pitch) or NULL */
  /* MMDM_EPI_EMBED_PE */
  const float* pe;
  int32_t pe_ld;
  int64_t dup_row_offset; /* != 0: also write row + offset (CFG halves share the embedding) */
  /* MMDM_EPI_FINAL */
  const float* std_;
  const float* mean;
  int32_t nfeats;
  int32_t persons_per_row;
} mmdm_epilogue;

/* D[M,N] = A[M,K] * W[N,K]^T, bf16 operands (row pitches lda/ldw in elements), fp32 accumulation (tcgen05).
 * replaces nn.Linear (reference models/utils/layers.py:11,95-96,112; in2in.py:389) */
int mmdm_linear_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                     const mmdm_epilogue* ep, void* stream);

/* out0 = bf16(LayerNorm(h)*(1+mod0[s,0:D]) + mod0[s,D:2D]), s = row / T; optional second modulation (out1).
 * replaces AdaLN.forward (reference models/utils/layers.py:15-25) */
int mmdm_ln_modulate(const float* h, int M, int D, int T, const float* mod0, int ld_mod0, void* out0,
                     const float* mod1, int ld_mod1, void* out1, void* stream);

/* softmax([Q K^T / sqrt(hd), 0]) [V; 0] per (sequence, head); keys of sequence s come from s ^ kv_xor.
 * replaces nn.MultiheadAttention(add_zero_attn=True) core (reference models/utils/layers.py:33-45,74-88) */
int mmdm_attention(const void* q, int ldq, const void* k, int ldk, const void* vt, int Tp, void* out, int ldo,
                   int S, int T, int H, int hd, int kv_xor, void* stream);

/* fp32 out[r,n] = act(a[r,:].w[n,:] + b[n]) (act 0: none, 1: SiLU); text_embed / time_embed (utils.py:41-55) */
int mmdm_small_linear(const float* a, int lda, const float* w, const float* b, float* out, int ldo, int R,
                      int N, int K, int act, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MIXERMDM_B200_H */
