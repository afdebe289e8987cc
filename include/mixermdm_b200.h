/* mixermdm_b200 — C ABI of the B200-native MixerMDM denoising loop.
 *
 * The reference (pabloruizponce/MixerMDM) has no FFI: its boundary for this path is the Python model API
 * (src/models/mixermdm.py:490-602, 660-811; src/models/in2in.py:401-462; src/models/utils/influence.py:93-126;
 * src/models/utils/cfg_sampler.py:31-56; src/models/utils/gaussian_diffusion.py:1769-2091). The Python mirror
 * in mixermdm_b200/ keeps those signatures and calls the entry points below through ctypes. Conventions:
 *   - every function returns 0 on success, a negative code on failure; mmdm_last_error(h) gives the message;
 *   - a handle belongs to the device that was current in mmdm_create; its entry points make that device current
 *     for the duration of the call, so the caller may sit on another GPU of the same process;
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

#define MMDM_ABI_VERSION 4

typedef struct mmdm_handle mmdm_handle;

int mmdm_abi_version(void);
/* message of the last failed call on handle h; with h == NULL the calling thread's last failure of a call that has
 * no handle (mmdm_create, the operator-level entry points). No process-global error state. */
const char* mmdm_last_error(const mmdm_handle* h);
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
  MMDM_EPI_F32 = 5,       /* out_f32  = acc + bias                      AdaLN.emb_layers[1]                 */
  MMDM_EPI_SIGMOID_F32 = 6, /* out_f32 = sigmoid(acc + bias)             Influence.out + sigmoid            */
  MMDM_EPI_QGELU_BF16 = 7, /* out_bf16 = x sigmoid(1.702 x), x = acc + bias   CLIP QuickGELU (text tower MLP) */
  /* AdaLN (layers.py:15-25) folded into the GEMMs on either side of it. With xn = (h - mu) rho (1 + scale_s) + shift_s
   * feeding a Linear (W, c):  y[r,n] = rho_r sum_k [h (1+scale_s)][r,k] W[n,k] - rho_r mu_r G_s[n] + Bc_s[n],
   * G_s[n] = sum_k (1+scale_s[k]) W[n,k],  Bc_s[n] = sum_k shift_s[k] W[n,k] + c[n]  (s = row / T).
   * The *_LN producers write h and bf16(h (1+scale_s)) plus per-row (sum h, sum h^2) per 64-column slot; the *_LN
   * consumers take that operand and finish the normalisation in their epilogue. */
  MMDM_EPI_BF16_LN = 8,       /* MMDM_EPI_BF16 on the folded operand                                         */
  MMDM_EPI_GELU_BF16_LN = 9,  /* MMDM_EPI_GELU_BF16 on the folded operand                                    */
  MMDM_EPI_RESID_LN = 10,     /* MMDM_EPI_RESID + operand(s) and row sums for the next AdaLN site            */
  MMDM_EPI_EMBED_PE_LN = 11   /* MMDM_EPI_EMBED_PE + the same                                                */
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
  /* *_LN producers (N = latent dim, N % 64 == 0, ldo == N; the bf16 operands share that pitch) */
  const float* lnp_scale0; int32_t lnp_ld0; void* lnp_out0;   /* scale rows [s*ld + k] of the next site, its operand  */
  const float* lnp_scale1; int32_t lnp_ld1; void* lnp_out1;   /* optional second site on the same rows, or NULL       */
  float* lnp_stats;        /* float2 [N/64][lnp_pitch]: (sum h, sum h^2) of row r over columns [64 j, 64 j + 64)      */
  int64_t lnp_pitch;
  int32_t lnp_T;           /* rows per sequence                                                                       */
  int32_t lnp_rows;        /* rows >= lnp_rows only get h (0 = all rows)                                              */
  int32_t lnp_dup_seq;     /* EMBED_PE_LN with dup_row_offset: sequence offset of the duplicated rows                 */
  /* *_LN consumers (K = latent dim; uses T; `bias` is ignored, it is part of Bc) */
  const float* lnc_stats; int64_t lnc_pitch;                  /* the producer's statistics                            */
  const float* lnc_g; const float* lnc_b; int32_t lnc_ld;     /* G_s[n] = lnc_g[s*ld + n], Bc_s[n] = lnc_b[s*ld + n]  */
  /* optional, producers: 16 bytes per row with the finished statistics (-mean and 1/rstd as bf16 pairs, rstd as float),
   * written by a row pass behind the GEMM; consumers: the producer's lnp_rowa. With it the CTA-pair consumer kernel
   * applies the mean / shift terms as one extra k step of the MMA (needs T >= 256). */
  void* lnp_rowa;
  const void* lnc_rowa;
  /* optional, producers: per-row offsets. Operand(s) and row sums are then those of h - lnp_moff_in[r] (an estimate of
   * the row mean keeps the bf16 rounding of the operand free of it; any value gives the same result up to rounding); the
   * row pass stores the row's mean, lnp_moff_in[r] + mean(h - lnp_moff_in[r]), in lnp_moff_out for the next site. */
  const float* lnp_moff_in;
  float* lnp_moff_out;
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
/* the same with a key-padding mask (key_padding_mask of nn.MultiheadAttention, layers.py:36-44): words from
 * mmdm_build_key_mask, [S][ceil((T+1)/64)] uint64, indexed by the sequence the keys come from */
int mmdm_attention_masked(const void* q, int ldq, const void* k, int ldk, const void* vt, int Tp, void* out, int ldo,
                          int S, int T, int H, int hd, int kv_xor, const uint64_t* key_mask_words, void* stream);
/* mask: device (rows, T) fp32, frame attendable where > 0.5 (reference: key_padding_mask = ~(mask > 0.5),
 * in2in.py:433-435); sequence s of nseq reads mask row (s / persons) % rows; the zero-attention key stays attendable */
int mmdm_build_key_mask(const float* mask, int rows, int nseq, int T, int persons, uint64_t* words, void* stream);

/* fp32 out[r,n] = act(a[r,:].w[n,:] + b[n]) (act 0: none, 1: SiLU); text_embed / time_embed (utils.py:41-55) */
int mmdm_small_linear(const float* a, int lda, const float* w, const float* b, float* out, int ldo, int R,
                      int N, int K, int act, void* stream);


/* ih_to_smpl -> align_motions -> smpl_to_ih on the de-normalised predictions (mixermdm.py:702-712,
 * utils/alignment.py:11-158) + bf16 packing of the Mixer motion_embed operand rows ((be*2+p)*T+t, kpad). */
int mmdm_align_motions(const float* out1, const float* out2, float* out1a, float* out2a, void* emb_i, void* emb_I,
                       int Be, int T, int kpad, int do_align, void* stream);

/* The fused HBM pass: Influence head + sigmoid, weight expansion, mix, classifier-free guidance, process_xstart
 * (centre + both normalisers) and the DDIM update (mixermdm.py:739-801; cfg_sampler.py:52-55;
 * gaussian_diffusion.py:2031-2062, 1936-1965). One 8-CTA cluster per sample. */
typedef struct mmdm_mixstep_args {
  int32_t B, T, mixing_mode;
  int32_t ncfg;        /* 2: rows [0,B) cond, [B,2B) uncond; 1: plain batch (mix_only = 1 only) */
  int32_t mix_only;    /* 0: full sampler step; 1: write the mix; 2: write the guided prediction */
  int32_t center, has_force;
  float force_val, cfg_scale;
  const int32_t* step_ptr;   /* device int: DDIM index (nsteps-1 .. 0); 0 = last step, raw result to final_out */
  int32_t nsteps;
  const float* coef;         /* device (nsteps,4) */
  const float* weights;      /* fp32 Mixer weights sigmoid(head), pitch 24: row ((c*B+b)*2+p)*T+t (modes 2,4) or
                                ((c*B+b)*2+p) (pooled modes 1,3); 1 (modes 1,2) or 23 (modes 3,4) valid columns */
  const float* out1a; const float* out2a;  /* (ncfg*B,T,524) aligned predictions */
  float* x1; float* x2;      /* (B,T,524) state streams, updated in place */
  /* out1a/out2a and x1/x2 are dense (.,T,524) buffers, or interleaved row by row - [out1a row | out2a row] with
     out2a == out1a + 524, likewise x2 == x1 + 524 - which the kernel fetches with one bulk copy per pair of rows */
  float* final_out;          /* (B,T,524): the guided prediction of the step. The result at index 0; at every other
                                index the kernel's own hand-over buffer between its two phases (contents = the raw
                                guided prediction of that step) */
  float* mix_out;
  const float* mean1; const float* std1; const float* mean2; const float* std2;
  float* hist_i1; float* hist_i2;
} mmdm_mixstep_args;
int mmdm_mix_cfg_ddim(const mmdm_mixstep_args* args, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Engine level: the whole Mixer step / DDIM loop behind one handle
 * ---------------------------------------------------------------------------------------------- */

/* one transformer block (reference models/utils/blocks.py:30-63, influence.py:6-48); all matrices bf16 in the
 * nn.Linear (out_features, in_features) layout, biases fp32. CA pointers are NULL for mode "individual". */
typedef struct mmdm_block_weights {
  const void* sa_in_w;  const float* sa_in_b;   /* (3D, D) = sa_block.attention.in_proj_{weight,bias}   */
  const void* sa_out_w; const float* sa_out_b;  /* (D, D)  = sa_block.attention.out_proj                */
  const void* ca_q_w;   const float* ca_q_b;    /* (D, D)  = ca_block.attention.in_proj rows [0:D]      */
  const void* ca_kv_w;  const float* ca_kv_b;   /* (2D, D) = ca_block.attention.in_proj rows [D:3D]     */
  const void* ca_out_w; const float* ca_out_b;  /* (D, D)  = ca_block.attention.out_proj                */
  const void* ffn_w1;   const float* ffn_b1;    /* (F, D)  = ffn.linear1                                */
  const void* ffn_w2;   const float* ffn_b2;    /* (D, F)  = ffn.linear2                                */
} mmdm_block_weights;

enum mmdm_net_kind { MMDM_NET_INDIVIDUAL = 0, MMDM_NET_INTERACTION = 1, MMDM_NET_INFLUENCE = 2 };

/* One network. The AdaLN modulation linears (layers.py:11) of all blocks are concatenated along their output
 * dimension into two matrices, grouped by the conditioning vector that drives them:
 *   ada0 <- E0 (per-person embedding: emb_individual{1,2} / cond_i{1,2})
 *   ada1 <- E1 (interaction embedding: emb / cond_I)
 * Row order inside each, per block l (each entry is 2D rows = [scale; shift]):
 *   INDIVIDUAL : ada0 = [sa.norm, ffn.norm]                        ada1 empty
 *   INTERACTION: ada0 = [sa.norm, ffn.norm]                        ada1 = [ca.norm, ca.xf_norm]
 *   INFLUENCE  : ada0 = [sa.norm]                                  ada1 = [ca.norm, ca.xf_norm, ffn.norm]     */
typedef struct mmdm_net_weights {
  int32_t kind, D, F, L, H;
  const void* ada0_w; const float* ada0_b; int32_t ada0_n;
  const void* ada1_w; const float* ada1_b; int32_t ada1_n;
  const void* embed_w; const float* embed_b;    /* motion_embed (D, kpad) bf16 zero padded from (D, 262)      */
  const float* pe; int32_t pe_rows;             /* sequence_pos_encoder.pe (pe_rows, D) fp32                   */
  const void* out_w; const float* out_b;        /* out.linear (262, D) bf16 (denoisers) / NULL                 */
  const float* time_w0; const float* time_b0;   /* embed_timestep.time_embed.0 (D, D) fp32                     */
  const float* time_w2; const float* time_b2;   /* embed_timestep.time_embed.2                                 */
  const float* text_w;  const float* text_b;    /* text_embed (D, 768) fp32                                    */
  const mmdm_block_weights* blocks;             /* host array of L entries                                     */
  /* Optional: composite conditioning tables for the LN fold (MMDM_EPI_*_LN). The AdaLN (scale|shift) vectors are linear
   * in E = bf16(SiLU(emb)): scale_s = Wsc E_s + bsc, shift_s = Wsh E_s + bsh (rows of ada0 / ada1), hence so are
   *   G_s  = W (1 + scale_s) = (W Wsc) E_s + W (1 + bsc)       Bc_s = W shift_s + c = (W Wsh) E_s + W bsh + c
   * for the Linear (W, c) behind the norm. comp{0,1}_w holds, per block and per site in the ada{0,1} order, the rows
   * [W Wsc (N rows); W Wsh (N rows)] - computed from the bf16 values of W, Wsc, Wsh in >= fp32 - each row stored as a
   * bf16 pair [hi (D) | lo (D)] with hi + lo = the fp32 value, so the table is (comp_n, 2D) and the engine multiplies
   * it by [E_s | E_s]; comp{0,1}_b the fp32 constants [W (1 + bsc); W bsh + c]. Sites (N = rows of the consuming Linear):
   *   INDIVIDUAL : comp0 = [sa.norm -> in_proj (3D), ffn.norm -> linear1 (F)]
   *   INTERACTION: comp0 = the same;  comp1 = [ca.norm -> in_proj[0:D] (D), ca.xf_norm -> in_proj[D:3D] (2D)]
   *   INFLUENCE  : comp0 = [sa.norm -> in_proj (3D)];  comp1 = [ca.norm -> in_proj[0:D] (D), ffn.norm -> linear1 (F)]
   * NULL comp0_w: the engine runs the stand-alone LayerNorm kernel (mmdm_ln_modulate) in front of those GEMMs. */
  const void* comp0_w; const float* comp0_b; int32_t comp0_n;
  const void* comp1_w; const float* comp1_b; int32_t comp1_n;
} mmdm_net_weights;

typedef struct mmdm_config {
  int32_t nfeats;        /* 262 */
  int32_t kpad;          /* 264: nfeats rounded up to a 16-byte row pitch in bf16 */
  int32_t text_dim;      /* 768 */
  int32_t nsteps;        /* DDIM steps (50 for "ddim50") */
  int32_t mixing_mode;   /* 1 global, 2 temporal, 3 spatial, 4 spatio-temporal (influence.py:61-65) */
  int32_t align;         /* Mixer(align=...)          : mixermdm.py:702-712 */
  int32_t center;        /* MixerDiffusion(align=...) : gaussian_diffusion.py:2043-2049 */
  int32_t has_force;     /* Mixer.force_influence_val is not None */
  float force_val;
  float cfg_scale;       /* ClassifierFreeSampleModelX2.s */
  const int32_t* timestep_map;  /* host, nsteps entries (_WrappedModelTest, gaussian_diffusion.py:2193-2205) */
  const float* coef;     /* device, (nsteps,4): sqrt_recip_ab, sqrt_recipm1_ab, sqrt(ab_prev), sqrt(1-ab_prev) */
  const void* head_w;    /* device, influence.out.weight (1 or 23, 512) bf16 */
  const float* head_b;
  const float* mean1; const float* std1;  /* MotionNormalizerTorchHML3D (individual model) */
  const float* mean2; const float* std2;  /* MotionNormalizerTorch      (interaction model) */
} mmdm_config;

int mmdm_create(const mmdm_config* cfg, const mmdm_net_weights nets[3], mmdm_handle** out);
void mmdm_destroy(mmdm_handle* h);
/* bytes of device workspace needed for an (un-doubled) batch B of T frames */
size_t mmdm_workspace_bytes(const mmdm_handle* h, int B, int T);
int mmdm_bind_workspace(mmdm_handle* h, void* ws, size_t bytes, int B, int T);

/* text_embed of the 8 condition slices for the bound batch (cond: device (B, 6144) fp32); once per batch */
int mmdm_set_condition(mmdm_handle* h, const float* cond, void* stream);
/* frame mask for every attention of the following forwards / steps on this handle (in2INDenoiser / Influence
 * `mask` argument, in2in.py:410-435, influence.py:107-116): device (rows, T) fp32, frame kept where > 0.5; network
 * batch row bx reads mask row bx % rows (so the sampler's B rows serve both CFG halves). NULL clears it. Cleared by
 * mmdm_bind_workspace. The Mixer entry points refuse it with cfg.align (the reference's masked alignment does not run,
 * alignment.py:86-89); the denoiser entry point takes it always. */
int mmdm_set_key_mask(mmdm_handle* h, const float* mask, int rows, void* stream);
/* state streams x1 (individual space) / x2 (interaction space), device (B,T,524) fp32, and the DDIM index */
int mmdm_set_state(mmdm_handle* h, const float* x1, const float* x2, int step_index, void* stream);
int mmdm_get_state(mmdm_handle* h, float* x1, float* x2, void* stream);
/* one DDIM step: CFG-doubled Mixer.forward + process_xstart + DDIM update; decrements the device step index.
 * replaces MixerDiffusion.ddim_sample (gaussian_diffusion.py:1902-1965) */
int mmdm_step(mmdm_handle* h, float* final_out /* (B,T,524) guided prediction of the step (the sample when the index is 0), or NULL */, void* stream);
/* state += scale * noise on both streams (noise1 / noise2: device (B,T,524) fp32, may be the same buffer): the
 * sigma_t * z term of an ancestral (DDPM) update, GaussianDiffusion.p_sample (gaussian_diffusion.py:683-691). With the
 * handle's coefficient table set to the posterior-mean form (schedule.py: ddpm_coefficients) a DDPM step is
 * mmdm_step followed by this call; the reference's own MixerDiffusion cannot run p_sample (SURVEY 8a note 2), so the
 * parity of this variant is formula-level. */
int mmdm_add_state_noise(mmdm_handle* h, const float* noise1, const float* noise2, float scale, void* stream);
/* whole loop from noise; result = pred_xstart2 of the last step (gaussian_diffusion.py:1769-1820).
 * use_graph != 0 replays one captured step graph nsteps times (graphs are cached per batch / frame count, so
 * re-binding the same workspace to another motion length and back does not capture again). hist_i1/hist_i2: optional
 * (nsteps, 2B, T, 262) fp32 device buffers receiving the expanded Mixer weights, or NULL.
 * hist_out1/out2/mix: optional (nsteps, 2B, T, 524) buffers with the aligned predictions and out_influenced of
 * every step (Mixer mode "eval", mixermdm.py:805-809); all three or none; forces the eager (non-graph) loop. */
int mmdm_sample(mmdm_handle* h, const float* cond, const float* noise, float* out, float* hist_i1,
                float* hist_i2, float* hist_out1, float* hist_out2, float* hist_mix, int use_graph, void* stream);

/* Mixer.forward in mode "eval_intermediate" (mixermdm.py:660-811) on an explicit batch of Bx rows
 * (the caller did any CFG doubling): returns out_influenced (Bx,T,524); optional out1/out2 (aligned) and the
 * expanded weights infl_i1/infl_i2 (Bx,T,262). With cfg_combine != 0 the rows are [cond; uncond] halves and
 * out_influenced receives cfg_scale*cond + (1-cfg_scale)*uncond, (Bx/2,T,524)
 * (ClassifierFreeSampleModelX2.forward, cfg_sampler.py:38-56). */
int mmdm_mixer_forward(mmdm_handle* h, const float* x1, const float* x2, const int64_t* timesteps,
                       const float* cond, float* out_influenced, float* out1, float* out2, float* infl_i1,
                       float* infl_i2, int Bx, int cfg_combine, float cfg_scale, void* stream);
/* in2INDenoiser.forward (in2in.py:401-462): which = MMDM_NET_INDIVIDUAL (x: (Bx,T,262), cond (Bx,768)) or
 * MMDM_NET_INTERACTION (x: (Bx,T,524), cond (Bx,2304)); output is the raw (normalised-space) prediction. */
int mmdm_denoiser_forward(mmdm_handle* h, int which, const float* x, const int64_t* timesteps,
                          const float* cond, float* out, int Bx, void* stream);

/* Post-processing of a generated motion (reference: scripts/infer/mixermdm.py:127-132, the step after the sampler):
 * scipy.ndimage.gaussian_filter1d(motion, sigma, axis = time, mode = 'nearest') on (B,T,F) fp32 device buffers
 * (F = 524 covers the script's (T,2,262) view). weights: the 2*radius+1 normalised Gaussian taps, HOST memory
 * (radius = int(4 sigma + 0.5) = 4 for the script's sigma = 1; radius <= 8); out must not alias x. */
int mmdm_smooth_motion(const float* x, float* out, int B, int T, int F, const float* weights, int radius,
                       void* stream);

/* Output format of the evaluation generators (reference: evaluation/datasets.py:121-134, 452-465): the sampled
 * motion (B,T,2*nfeats) is optionally de-normalised per person (MotionNormalizer.backward: x*std+mean; mean/std are
 * (nfeats) DEVICE arrays, both or neither) and padded with zero frames to max_length -> out (B,max_length,2*nfeats),
 * i.e. the (B,max_length,2,nfeats) array the generators store. */
int mmdm_format_motion(const float* x, float* out, int B, int T, int max_length, int nfeats, const float* mean,
                       const float* std_, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Text conditioning, the step in front of the loop (reference: MixerMDM.text_process / generate_cond,
 * models/mixermdm.py:283-356; in2IN.text_process, models/in2in.py:109-135): frozen CLIP ViT-L/14 text tower
 * (OpenAI clip 1.0, clip/model.py) -> 2-layer post-norm nn.TransformerEncoder -> LayerNorm -> end-of-text row.
 * All matrices bf16 in the nn.Linear (out, in) layout, everything else fp32, all device pointers.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mmdm_clip_block {            /* clip/model.py ResidualAttentionBlock */
  const float* ln1_g; const float* ln1_b;   /* ln_1 */
  const void* attn_in_w; const float* attn_in_b;    /* attn.in_proj_{weight,bias} (3W, W) */
  const void* attn_out_w; const float* attn_out_b;  /* attn.out_proj (W, W) */
  const float* ln2_g; const float* ln2_b;   /* ln_2 */
  const void* fc_w; const float* fc_b;      /* mlp.c_fc (4W, W), QuickGELU behind it */
  const void* proj_w; const float* proj_b;  /* mlp.c_proj (W, 4W) */
} mmdm_clip_block;

typedef struct mmdm_text_tower {
  int32_t width, layers, heads, ctx, vocab; /* 768, 12, 12, 77, 49408 */
  const float* tok_emb;                     /* token_embedding.weight (vocab, W) */
  const float* pos_emb;                     /* positional_embedding (ctx, W) */
  const mmdm_clip_block* blocks;            /* host array of `layers` entries */
  const float* lnf_g; const float* lnf_b;   /* ln_final */
} mmdm_text_tower;

typedef struct mmdm_enc_layer {             /* nn.TransformerEncoderLayer(norm_first=False, activation="gelu") */
  const void* in_w; const float* in_b;      /* self_attn.in_proj (3W, W) */
  const void* out_w; const float* out_b;    /* self_attn.out_proj */
  const float* norm1_g; const float* norm1_b;
  const void* lin1_w; const float* lin1_b;  /* linear1 (ff, W) */
  const void* lin2_w; const float* lin2_b;  /* linear2 (W, ff) */
  const float* norm2_g; const float* norm2_b;
} mmdm_enc_layer;

typedef struct mmdm_text_encoder {
  int32_t layers, heads, ff;                /* 2, 8, 2048 */
  const mmdm_enc_layer* layers_;            /* host array */
  const float* ln_g; const float* ln_b;     /* clip_ln / clip_ln_individual / clip_ln_interaction */
} mmdm_text_encoder;

size_t mmdm_text_workspace_bytes(int nseq, int ctx, int width, int ff_max);
/* tokens: device (nseq, ctx) int32 (clip.tokenize output); eot: device (nseq) int32 = argmax of the token ids per row;
 * cond_out: host array of n_enc device pointers, each (nseq, cond_ld >= width) fp32: one tower pass, n_enc encoders. */
int mmdm_text_encode(const mmdm_text_tower* tower, const mmdm_text_encoder* encs, int n_enc, const int32_t* tokens,
                     const int32_t* eot, int nseq, float* const* cond_out, int cond_ld, void* ws, size_t ws_bytes,
                     void* stream);

/* Test hook (host only, no GPU): the index arithmetic of the fused mix kernel's frame stream (csrc/mix_stream.cuh).
 * For item `item` of the stream of S samples x nfr frames (with or without a phase 2) in a ring of ring_slots <= nfr
 * slots: out[0..9] = phase (1|2), sample, frame, ring slot, lap, item index recomputed from (phase, sample, frame),
 * successor valid, phase / sample / frame of the item ring_slots positions later (the one that takes over the slot). */
int mmdm_debug_mix_stream(int S, int nfr, int has_phase2, int ring_slots, int item, int32_t* out);

/* per-kernel-class device timing (CUDA events on the launch stream, no graph): enable, run steps, read. */
int mmdm_profile_enable(mmdm_handle* h, int on);
/* writes up to max_entries records; returns the number written */
typedef struct mmdm_profile_entry {
  char name[32];
  int64_t launches;
  double ms;       /* summed device time */
  double flops;    /* algorithmic flops (2*M*N*K, attention 4*T*Tk*hd per head) */
  double bytes;    /* algorithmic bytes where defined (mix kernel), else 0 */
} mmdm_profile_entry;
int mmdm_profile_read(mmdm_handle* h, mmdm_profile_entry* entries, int max_entries);
int64_t mmdm_launch_count(const mmdm_handle* h);
/* step graphs captured so far by mmdm_sample (one per distinct workspace / batch / frame count, cached) */
int64_t mmdm_graph_captures(const mmdm_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* MIXERMDM_B200_H */
