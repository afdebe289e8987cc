#pragma once
#include "common.cuh"

namespace mmdm {

// out0 = bf16(LN(h) * (1 + mod0[s, 0:D]) + mod0[s, D:2D]), s = row / T; optional second modulation.
// rev != 0: rows are visited from the last one down (L2 reuse of the producer's tail, see GemmEpi::rev).
int launch_ln_modulate(const float* h, int M, int D, int T, const float* mod0, int ld_mod0,
                       __nv_bfloat16* out0, const float* mod1, int ld_mod1, __nv_bfloat16* out1,
                       cudaStream_t stream, int rev = 0);

// LN fold, first block: out_m = bf16(h * (1 + sc_m[s, :])) (one or two sites) + per-row (sum, sum of squares) per 64-column
// slot, stats[slot * pitch + row] - what the residual GEMMs' EPI_RESID_LN epilogue emits for the later blocks; here of
// the exactly centred row (h - mean), mean -> moff_out[row] (GemmEpi::lnp_moff_in of the next site's producer)
int launch_ln_fold_first(const float* h, int M, int D, int T, const float* sc0, int ld0, __nv_bfloat16* out0,
                         const float* sc1, int ld1, __nv_bfloat16* out1, float2* stats, long long pitch, uint4* rowa,
                         float* moff_out, cudaStream_t stream, int rev = 0);

// E[s, 0:D] = E[s, D:2D] = bf16(SiLU(temb[trow,:] + text[text_row[s],:]))   (row pitch 2D)
// trow = step_ptr ? *step_ptr : temb_idx[s]
int launch_cond_silu(const float* temb, const int* temb_idx, const int* step_ptr, const float* text,
                     const int* text_row, int S, int D, __nv_bfloat16* E, cudaStream_t stream);
int launch_build_maps(int* map0, int* map1, int* tmap, int S, int persons, int Bt, int dup, int nslices,
                      int k0_base, cudaStream_t stream);
int launch_gather_rows(const float* table, int rows, const long long* idx64, const int* idx32, int R, int D,
                       float* out, cudaStream_t stream);

// fp32 out[r,n] = act(a[r,:] . w[n,:] + b[n]); act: 0 none, 1 SiLU
int launch_small_linear(const float* a, int lda, const float* w, const float* b, float* out, int ldo, int R,
                        int N, int K, int act, cudaStream_t stream);

// (B,T,persons*nf) fp32 [rows row_pitch floats apart; 0 = dense] -> ((b*persons+p)*T+t, kpad) bf16, zero padded
int launch_pack_state(const float* x, int B, int T, int nf, int kpad, int persons, __nv_bfloat16* out,
                      cudaStream_t stream, int row_pitch = 0);

int launch_set_int(int* p, int v, cudaStream_t stream);
int launch_add_int(int* p, int v, cudaStream_t stream);
int launch_add_state_noise(float* x1, float* x2, const float* n1, const float* n2, long long nrows, int rowf, float scale,
                           cudaStream_t stream);

}  // namespace mmdm
