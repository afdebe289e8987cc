// Motion-space kernels of the Mixer step (HBM-bound, fp32):
//   align_kernel   : ih_to_smpl -> align_motions -> smpl_to_ih on the four de-normalised predictions and
//                    bf16 packing of the Mixer's motion_embed operand (models/mixermdm.py:702-726).
//   mixstep_kernel : Influence head (Linear 512->{1,23} + sigmoid, models/utils/influence.py:120-125) fused with
//                    the weight expansion + mix (models/mixermdm.py:739-801), classifier-free guidance
//                    (models/utils/cfg_sampler.py:52-55), process_xstart (centre + both normalisers,
//                    models/utils/gaussian_diffusion.py:2031-2062) and the DDIM update (:1936-1965) in one pass.
//                    One 4-CTA thread-block CLUSTER per sample: every CTA keeps its frames of the guided prediction in
//                    shared memory, the per-sample reductions centring needs (min-y over all frames, frame-0
//                    root and hips) are exchanged through distributed shared memory, so the prediction never
//                    makes a round trip through HBM.
#include "motion.cuh"

#include <cooperative_groups.h>

#include "motion_math.cuh"

namespace cg = cooperative_groups;

namespace mmdm {

// ------------------------------------------------------------------------------------------------
// align_kernel: one warp per (sequence-pair be, frame t); handles both persons of both models.
// ------------------------------------------------------------------------------------------------
constexpr int ALIGN_WARPS = 8;

__global__ void __launch_bounds__(ALIGN_WARPS * 32)
align_kernel(const float* __restrict__ out1, const float* __restrict__ out2, float* __restrict__ out1a,
             float* __restrict__ out2a, __nv_bfloat16* __restrict__ emb_i, __nv_bfloat16* __restrict__ emb_I,
             int Be, int T, int kpad, int do_align) {
  __shared__ float buf[ALIGN_WARPS][4][NFEAT + 2];  // [model*2 + person][feature]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * ALIGN_WARPS + warp;
  if (item >= (long long)Be * T) return;
  const int be = (int)(item / T), t = (int)(item % T);
  const float* r1 = out1 + ((size_t)be * T + t) * (2 * NFEAT);
  const float* r2 = out2 + ((size_t)be * T + t) * (2 * NFEAT);
  float(*b)[NFEAT + 2] = buf[warp];
  for (int f = lane; f < 2 * NFEAT; f += 32) {
    const int p = f >= NFEAT, fi = f - p * NFEAT;
    b[p][fi] = r1[f];
    b[2 + p][fi] = r2[f];
  }
  __syncwarp();
  if (do_align) {
    // rot6d round trip on 4 x 21 joints
    for (int idx = lane; idx < 4 * NROT; idx += 32) {
      const int m = idx / NROT, j = idx - m * NROT;
      float o[6];
      rot6d_roundtrip(&b[m][OFF_ROT + j * 6], o);
#pragma unroll
      for (int e = 0; e < 6; ++e) b[m][OFF_ROT + j * 6 + e] = o[e];
    }
    // trajectory alignment of the individual prediction (model 0) onto the interaction one (model 1)
    if (lane < NJ) {
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const float* i0p = out2 + ((size_t)be * T) * (2 * NFEAT) + p * NFEAT;
        const float* iLp = out2 + ((size_t)be * T + (T - 1)) * (2 * NFEAT) + p * NFEAT;
        const float* s0p = out1 + ((size_t)be * T) * (2 * NFEAT) + p * NFEAT;
        const float* sLp = out1 + ((size_t)be * T + (T - 1)) * (2 * NFEAT) + p * NFEAT;
        const AlignXf a = make_align(v3(__ldg(i0p), __ldg(i0p + 1), __ldg(i0p + 2)),
                                     v3(__ldg(iLp), __ldg(iLp + 1), __ldg(iLp + 2)),
                                     v3(__ldg(s0p), __ldg(s0p + 1), __ldg(s0p + 2)),
                                     v3(__ldg(sLp), __ldg(sLp + 1), __ldg(sLp + 2)));
        float* pp = &b[p][lane * 3];
        V3 pos = v3(pp[0], pp[1], pp[2]) + a.d0;
        pos = qrot(a.q, pos) + a.d1;
        pp[0] = pos.x; pp[1] = pos.y; pp[2] = pos.z;
        float* vp = &b[p][OFF_VEL + lane * 3];
        const V3 vel = qrot(a.q, v3(vp[0], vp[1], vp[2]));
        vp[0] = vel.x; vp[1] = vel.y; vp[2] = vel.z;
      }
    }
    if (lane < 8) b[lane >> 2][OFF_FC + (lane & 3)] = 0.f;  // align_motions drops the foot contacts (201-d)
  }
  __syncwarp();
  float* w1 = out1a + ((size_t)be * T + t) * (2 * NFEAT);
  float* w2 = out2a + ((size_t)be * T + t) * (2 * NFEAT);
  for (int f = lane; f < 2 * NFEAT; f += 32) {
    const int p = f >= NFEAT, fi = f - p * NFEAT;
    w1[f] = b[p][fi];
    w2[f] = b[2 + p][fi];
  }
  // bf16 operand rows: sequence index be*2+p
  for (int idx = lane; idx < 2 * kpad; idx += 32) {
    const int p = idx >= kpad, k = idx - p * kpad;
    const size_t row = ((size_t)(be * 2 + p) * T + t) * kpad + k;
    emb_i[row] = __float2bfloat16_rn(k < NFEAT ? b[p][k] : 0.f);
    emb_I[row] = __float2bfloat16_rn(k < NFEAT ? b[2 + p][k] : 0.f);
  }
}

int launch_align(const float* out1, const float* out2, float* out1a, float* out2a, __nv_bfloat16* emb_i,
                 __nv_bfloat16* emb_I, int Be, int T, int kpad, int do_align, cudaStream_t stream) {
  const long long items = (long long)Be * T;
  const unsigned blocks = (unsigned)((items + ALIGN_WARPS - 1) / ALIGN_WARPS);
  align_kernel<<<blocks, ALIGN_WARPS * 32, 0, stream>>>(out1, out2, out1a, out2a, emb_i, emb_I, Be, T, kpad, do_align);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// mixstep_kernel
// ------------------------------------------------------------------------------------------------
constexpr int MIX_THREADS = 512;    // 16 warps per CTA (one CTA per SM): latency hiding matters more than per-lane batch depth
constexpr int MIX_WARPS = MIX_THREADS / 32;
constexpr int MIX_CLUSTER = 4;     // 4 SMs per sample: ~36 clusters resident (8-CTA clusters only reach 15)
constexpr int HD_MIX = 512;  // Mixer latent dim the head reads

struct MixExport {        // per-CTA data other CTAs of the cluster read through DSMEM
  float ymin[2];
  float f0[2][9];         // frame-0 joints 0,1,2 (xyz) per person (valid on cluster rank 0)
};

static size_t mix_smem_bytes(int frames_per_cta) {
  size_t b = (size_t)frames_per_cta * 2 * NFEAT * 4;         // guided prediction of my frames
  b += (size_t)MIX_WARPS * 4 * 24 * 4;                       // per-warp weights [4][24]
  b += (size_t)MIX_WARPS * 2 * 4;                            // per-warp y-min
  b += sizeof(MixExport);
  b += 4 * NFEAT * 4;                                        // 1/std and mean tables of both normalisers
  return b + 64;
}

template <int NW>
__global__ void __launch_bounds__(MIX_THREADS, 1)
mixstep_kernel(const MixStepParams P) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / MIX_CLUSTER;
  const int T = P.T, B = P.B;
  const int FR = (T + MIX_CLUSTER - 1) / MIX_CLUSTER;
  const int t_begin = rank * FR;
  const int nfr = max(0, min(FR, T - t_begin));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool pooled = (P.mode == 1 || P.mode == 3);

  extern __shared__ __align__(16) uint8_t smem[];
  float* cfg = reinterpret_cast<float*>(smem);
  float* wwarp = cfg + (size_t)FR * 2 * NFEAT;               // [MIX_WARPS][4][24]
  float* ymin_w = wwarp + MIX_WARPS * 4 * 24;                // [MIX_WARPS][2]
  MixExport* ex = reinterpret_cast<MixExport*>(ymin_w + MIX_WARPS * 2);
  float* tab = reinterpret_cast<float*>(ex + 1);             // [mean1 | rstd1 | mean2 | rstd2] x 262

  const int step = P.mix_only ? 0 : *P.step_ptr;
  const bool final_step = (step == 0) && !P.mix_only;
  if (!P.mix_only) {
    for (int i = threadIdx.x; i < NFEAT; i += MIX_THREADS) {
      tab[i] = P.mean1[i];
      tab[NFEAT + i] = 1.0f / P.std1[i];
      tab[2 * NFEAT + i] = P.mean2[i];
      tab[3 * NFEAT + i] = 1.0f / P.std2[i];
    }
  }
  // row of the Mixer weights for (cfg half c, person p, frame t): sequence ((c*B + b)*2 + p), per token unless pooled
  auto wrow = [&](int c, int p, int t) {
    const size_t seq = (size_t)(c * B + b) * 2 + p;
    return P.weights + (pooled ? seq : seq * T + t) * 24;
  };
  __syncthreads();

  // ---- phase 1: head + mix + guidance for my frames ---------------------------------------------
  // One warp per frame. All global loads of a stage are issued before the first one is consumed (the kernel is
  // HBM-latency bound otherwise): 4 x 17 coalesced words per lane for the aligned predictions.
  float ym0 = INFINITY, ym1 = INFINITY;
  const float s_cfg = P.cfg_scale, s_unc = 1.f - P.cfg_scale;
  const bool two = (P.ncfg == 2);
  constexpr int NIT = (2 * NFEAT + 31) / 32;   // 17 words per lane per row
  // A lane always handles the same 17 features of a row (f = lane + 32 i): which Mixer weight multiplies them,
  // whether they are a joint height (floor search) and their per-person index are computed once, not per frame.
  int meta[NIT];   // bits 0-5: person * 24 + weight index, bit 6: y coordinate of a joint position, bits 8-16: fi
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    const int f = lane + 32 * i;
    const int p = f >= NFEAT, fi = f - p * NFEAT;
    const int wi = (NW == 1) ? 0 : weight_index(fi);
    meta[i] = (p * 24 + wi) | ((fi < OFF_VEL && (fi % 3) == 1) ? 64 : 0) | (fi << 8);
  }
  for (int fl = warp; fl < nfr; fl += MIX_WARPS) {
    const int t = t_begin + fl;
    float* wv = wwarp + warp * 4 * 24;
    if (lane < 24) {
#pragma unroll
      for (int cp = 0; cp < 4; ++cp)
        if (cp < 2 * P.ncfg) wv[cp * 24 + lane] = (lane < NW) ? __ldg(wrow(cp >> 1, cp & 1, t) + lane) : 0.f;
    }
    __syncwarp();
    if (P.has_force)
      for (int i = lane; i < 4 * 24; i += 32) wv[i] = P.force_val;
    __syncwarp();
    const size_t rc = ((size_t)b * T + t) * (2 * NFEAT);            // cond half row
    const size_t ru = ((size_t)(B + b) * T + t) * (2 * NFEAT);      // uncond half row
    float* crow = cfg + (size_t)fl * 2 * NFEAT;
    float o1c[NIT], o2c[NIT], o1u[NIT], o2u[NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int f = lane + 32 * i;
      const bool ok = f < 2 * NFEAT;
      o1c[i] = ok ? __ldg(P.out1a + rc + f) : 0.f;
      o2c[i] = ok ? __ldg(P.out2a + rc + f) : 0.f;
      o1u[i] = (ok && two) ? __ldg(P.out1a + ru + f) : 0.f;
      o2u[i] = (ok && two) ? __ldg(P.out2a + ru + f) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int f = lane + 32 * i;
      if (f < 2 * NFEAT) {
        const int p = f >= NFEAT, fi = meta[i] >> 8;
        const float wc = wv[meta[i] & 63];
        const float wu = two ? wv[48 + (meta[i] & 63)] : 0.f;
        const float mc = o2c[i] + wc * (o1c[i] - o2c[i]);
        const float mu = two ? o2u[i] + wu * (o1u[i] - o2u[i]) : 0.f;
        if (P.mix_only == 1 || (P.mix_only == 0 && P.mix_out != nullptr)) {   // out_influenced ("eval" history)
          P.mix_out[rc + f] = mc;
          if (two) P.mix_out[ru + f] = mu;
        }
        if (P.hist_i1 != nullptr) {
          float* hbase = (p ? P.hist_i2 : P.hist_i1) +
                         (size_t)(P.mix_only ? 0 : (P.nsteps - 1 - step)) * ((size_t)P.ncfg * B * T * NFEAT);
          hbase[((size_t)b * T + t) * NFEAT + fi] = wc;
          if (two) hbase[((size_t)(B + b) * T + t) * NFEAT + fi] = wu;
        }
        if (P.mix_only != 1) {
          const float g = s_cfg * mc + s_unc * mu;
          if (P.mix_only == 2) {
            P.mix_out[rc + f] = g;
          } else {
            crow[f] = g;
            if (meta[i] & 64) {
              if (p) ym1 = fminf(ym1, g); else ym0 = fminf(ym0, g);
            }
            if (final_step) P.final_out[rc + f] = g;   // rows of the un-doubled batch coincide with the cond half
          }
        }
      }
    }
    __syncwarp();
  }
  if (final_step || P.mix_only) {   // t == 0: nothing is centred or normalised, the guided prediction is the result
    cluster.sync();
    return;
  }
  ym0 = warp_min(ym0);
  ym1 = warp_min(ym1);
  if (lane == 0) { ymin_w[warp * 2] = ym0; ymin_w[warp * 2 + 1] = ym1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float m = INFINITY;
    for (int w = 0; w < MIX_WARPS; ++w) m = fminf(m, ymin_w[w * 2 + threadIdx.x]);
    ex->ymin[threadIdx.x] = m;
  }
  if (rank == 0 && threadIdx.x < 18) {
    const int p = threadIdx.x / 9, e = threadIdx.x % 9;
    ex->f0[p][e] = cfg[p * NFEAT + e];
  }
  cluster.sync();
  CenterXf cx[2];
  {
    const MixExport* e0 = cluster.map_shared_rank(ex, 0);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      float m = INFINITY;
      for (int r = 0; r < MIX_CLUSTER; ++r) m = fminf(m, cluster.map_shared_rank(ex, r)->ymin[p]);
      const float* f0 = e0->f0[p];
      // FACE_JOINT_INDX[:2] = (r_hip=2, l_hip=1)  (utils/paramUtil.py:89)
      cx[p] = make_center(v3(f0[0], f0[1] - m, f0[2]), v3(f0[6], f0[7], f0[8]), v3(f0[3], f0[4], f0[5]), m);
    }
  }

  // ---- phase 2: process_xstart + DDIM on both state streams -----------------------------------------
  const float c_recip = P.coef[step * 4 + 0], c_inv_recipm1 = 1.0f / P.coef[step * 4 + 1];
  const float c_ap = P.coef[step * 4 + 2], c_eps = P.coef[step * 4 + 3];
  for (int fl = warp; fl < nfr; fl += MIX_WARPS) {
    const int t = t_begin + fl;
    float* crow = cfg + (size_t)fl * 2 * NFEAT;
    const size_t rx = ((size_t)b * T + t) * (2 * NFEAT);
    // both state rows are requested up front; they arrive while the centring / rotation math below runs
    float xs1[NIT], xs2[NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int f = lane + 32 * i;
      const bool ok = f < 2 * NFEAT;
      xs2[i] = ok ? P.x2[rx + f] : 0.f;
      xs1[i] = ok ? P.x1[rx + f] : 0.f;
    }
    // stream 2 (interaction space): normalise the raw guided prediction
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int f = lane + 32 * i;
      if (f < 2 * NFEAT) {
        const int fi = meta[i] >> 8;
        const float pred = (crow[f] - tab[2 * NFEAT + fi]) * tab[3 * NFEAT + fi];
        const float eps = (c_recip * xs2[i] - pred) * c_inv_recipm1;
        P.x2[rx + f] = pred * c_ap + c_eps * eps;
      }
    }
    __syncwarp();
    // stream 1 (individual space): ih_to_smpl -> center_motion -> smpl_to_ih per person, then normalise
    if (P.center) {
      if (lane < NJ) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          float* pp = crow + p * NFEAT + lane * 3;
          const V3 q = center_pos(cx[p], v3(pp[0], pp[1], pp[2]));
          pp[0] = q.x; pp[1] = q.y; pp[2] = q.z;
          float* vp = crow + p * NFEAT + OFF_VEL + lane * 3;
          const V3 v = qrot(cx[p].q, v3(vp[0], vp[1], vp[2]));
          vp[0] = v.x; vp[1] = v.y; vp[2] = v.z;
        }
      }
      for (int idx = lane; idx < 2 * NROT; idx += 32) {
        const int p = idx / NROT, j = idx - p * NROT;
        float* rp = crow + p * NFEAT + OFF_ROT + j * 6;
        float o[6];
        rot6d_roundtrip(rp, o);
#pragma unroll
        for (int e = 0; e < 6; ++e) rp[e] = o[e];
      }
      if (lane < 8) crow[(lane >> 2) * NFEAT + OFF_FC + (lane & 3)] = 0.f;
      __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int f = lane + 32 * i;
      if (f < 2 * NFEAT) {
        const int fi = meta[i] >> 8;
        const float pred = (crow[f] - tab[fi]) * tab[NFEAT + fi];
        const float eps = (c_recip * xs1[i] - pred) * c_inv_recipm1;
        P.x1[rx + f] = pred * c_ap + c_eps * eps;
      }
    }
  }
  cluster.sync();   // keep my shared memory alive until every CTA of the cluster has read the exports
}

constexpr int MIX_SMEM_MAX = 227 * 1024;
int motion_init() {
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(mixstep_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MIX_SMEM_MAX));
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(mixstep_kernel<23>, cudaFuncAttributeMaxDynamicSharedMemorySize, MIX_SMEM_MAX));
  return 0;
}

int launch_mixstep(const MixStepParams& P, cudaStream_t stream) {
  static int inited = motion_init();
  MMDM_REQUIRE(inited == 0, "mixstep: kernel attribute setup failed");
  MMDM_REQUIRE(P.mode >= 1 && P.mode <= 4, "Mixing mode not recognized (%d)", P.mode);
  MMDM_REQUIRE(P.T >= 1 && P.T <= 320 && P.B >= 1, "mixstep: bad shape B=%d T=%d", P.B, P.T);
  MMDM_REQUIRE(P.weights != nullptr, "mixstep: Mixer weights missing");
  const int NW = (P.mode <= 2) ? 1 : 23;
  const int FR = (P.T + MIX_CLUSTER - 1) / MIX_CLUSTER;
  const size_t smem = mix_smem_bytes(FR);
  MMDM_REQUIRE(smem <= (size_t)MIX_SMEM_MAX, "mixstep: shared memory budget exceeded (%zu)", smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(P.B * MIX_CLUSTER));
  cfg.blockDim = dim3(MIX_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MIX_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (NW == 1) MMDM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mixstep_kernel<1>, P));
  else MMDM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mixstep_kernel<23>, P));
  return 0;
}

// mean over the T tokens of every sequence: (S*T, D) bf16 -> (S, D) bf16   (influence.py:116-118, modes 1 and 3)
__global__ void pool_mean_kernel(const __nv_bfloat16* __restrict__ h, int T, int D, __nv_bfloat16* __restrict__ out) {
  const int s = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc += __bfloat162float(h[((size_t)s * T + t) * D + d]);
    out[(size_t)s * D + d] = __float2bfloat16_rn(acc / (float)T);
  }
}
int launch_pool_mean(const __nv_bfloat16* h, int S, int T, int D, __nv_bfloat16* out, cudaStream_t stream) {
  pool_mean_kernel<<<S, 256, 0, stream>>>(h, T, D, out);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mmdm
