#pragma once
#include "common.cuh"

namespace mmdm {

// out1/out2: (Be,T,524) fp32 de-normalised predictions (individual / interaction model).
// out1a/out2a: aligned copies; emb_i/emb_I: bf16 rows ((be*2+p)*T+t, kpad) for the Mixer's motion_embed.
// pitch_a: floats between consecutive (be,t) rows of out1a / out2a (524, or 1048 when the two are interleaved row by row)
int launch_align(const float* out1, const float* out2, float* out1a, float* out2a, __nv_bfloat16* emb_i,
                 __nv_bfloat16* emb_I, int Be, int T, int kpad, int do_align, cudaStream_t stream, int pitch_a = 524);

struct MixStepParams {
  int B, T, mode;                    // un-doubled batch, frames, mixing mode 1..4
  int ncfg;                          // 2: rows [0,B) cond + [B,2B) uncond (sampler); 1: plain batch (mix_only)
  int mix_only;                      // 1: stop after the mix, mix_out = out_influenced (ncfg*B,T,524)  [Mixer.forward]
                                     // 2: stop after guidance, mix_out = s*cond+(1-s)*uncond (B,T,524) [CFG wrapper]
  float* mix_out;
  int center;                        // MixerDiffusion(align=...) : centre stream 1 in process_xstart
  int has_force;                     // Mixer.force_influence_val is not None
  float force_val;
  float cfg_scale;
  const int* step_ptr;               // device: current DDIM index (nsteps-1 ... 0)
  int nsteps;
  const float* coef;                 // [nsteps][4] = sqrt_recip_ab, sqrt_recipm1_ab, sqrt(ab_prev), sqrt(1-ab_prev)
  const float* weights;              // sigmoid(head) rows, pitch 24: per token (modes 2,4) or per sequence (1,3)
  const float* out1a;                // (2B,T,524) rows with pitch pred_pitch
  const float* out2a;
  int pred_pitch;                    // floats between consecutive (b,t) rows of out1a / out2a: 524, or 1048 when the two
                                     // buffers are interleaved row by row (out2a == out1a + 524): one bulk copy per pair
  int state_pitch;                   // same for x1 / x2
  float* x1;                         // (B,T,524) state, individual space (updated in place)
  float* x2;                         // (B,T,524) state, interaction space
  float* final_out;                  // (B,T,524) guided prediction: the result at step index 0, otherwise the
                                     // kernel's phase-1 -> phase-2 hand-over buffer (stays in L2)
  const float* mean1; const float* std1;   // individual-model normaliser (HumanML3D stats)
  const float* mean2; const float* std2;   // interaction-model normaliser (InterHuman stats)
  float* hist_i1;                    // optional (nsteps, 2B, T, 262) expanded weights, or null
  float* hist_i2;
};

int launch_mixstep(const MixStepParams& P, cudaStream_t stream);
// temporal Gaussian blur (scripts/infer/mixermdm.py:131): x, out (B,T,F) fp32, weights[2R+1] symmetric taps (host)
int launch_smooth_time(const float* x, float* out, int B, int T, int F, const float* weights, int R,
                       cudaStream_t stream);
// evaluation output format (evaluation/datasets.py:121-134): (B,T,2*nf) -> (B,Tmax,2*nf), x*std+mean per person when
// mean/std (nf, device) are given, zero frames beyond T
int launch_format_motion(const float* x, float* out, int B, int T, int Tmax, int nf, const float* mean,
                         const float* stdv, cudaStream_t stream);
int launch_pool_mean(const __nv_bfloat16* h, int S, int T, int D, __nv_bfloat16* out, cudaStream_t stream);
int motion_init();

}  // namespace mmdm
