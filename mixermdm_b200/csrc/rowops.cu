// Row-wise / small kernels around the GEMMs: LayerNorm + AdaLN modulation (emitting bf16 GEMM operands),
// conditioning-vector preparation, small fp32 linears (text / timestep embedders), state packing.
#include "rowops.cuh"

namespace mmdm {

// ------------------------------------------------------------------------------------------------
// LN(h) * (1 + scale[s]) + shift[s]  (reference: models/utils/layers.py:15-25, eps = 1e-6, no affine)
// One warp per row; the row lives in registers (D = 128 * NV floats). Optionally a second modulation
// of the SAME normalised row is emitted (block input feeds SA of its own stream and the CA-KV of the
// partner stream: models/utils/blocks.py:50,57 with models/in2in.py:439-440).
// ------------------------------------------------------------------------------------------------
// A warp owns a run of consecutive rows of ONE sequence: the sequence's (1 + scale) and shift vectors are staged
// once per run in the warp's shared-memory slice (registers would cost occupancy), and the next row is already in
// flight while the current one is reduced and written, so the kernel streams h like a copy (one fp32 read, one
// bf16 write per element) instead of stalling on a load -> reduce -> load(modulation) -> store chain per row.
template <int NV, bool DUAL>
__global__ void __launch_bounds__(128)
ln_modulate_kernel(const float* __restrict__ h, int S, int T, int runs_per_seq, int run_len, int rev,
                   const float* __restrict__ mod0, int ld_mod0, __nv_bfloat16* __restrict__ out0,
                   const float* __restrict__ mod1, int ld_mod1, __nv_bfloat16* __restrict__ out1) {
  constexpr int D = NV * 128;
  constexpr int NM = DUAL ? 2 : 1;
  extern __shared__ float4 ln_smem[];   // [4 warps][NM][scale+1 | shift][D/4]
  int task = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (task >= S * runs_per_seq) return;
  if (rev) task = S * runs_per_seq - 1 - task;   // last rows first: they are the ones the producer left in L2
  const int s = task / runs_per_seq;
  const int t0 = (task - s * runs_per_seq) * run_len;
  const int t1 = min(T, t0 + run_len);
  if (t0 >= t1) return;

  const float4* src = reinterpret_cast<const float4*>(h + ((size_t)s * T + t0) * D);
  float4 v[NV], nx[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = src[i * 32 + lane];

  float4* ms = ln_smem + (size_t)(threadIdx.x >> 5) * (NM * 2 * (D / 4));
#pragma unroll
  for (int m = 0; m < NM; ++m) {
    const float* mod = m == 0 ? mod0 : mod1;
    const int ld = m == 0 ? ld_mod0 : ld_mod1;
    const float4* psc = reinterpret_cast<const float4*>(mod + (size_t)s * ld);
    const float4* psh = reinterpret_cast<const float4*>(mod + (size_t)s * ld + D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 a = __ldg(psc + i * 32 + lane);
      a.x += 1.f; a.y += 1.f; a.z += 1.f; a.w += 1.f;
      ms[(m * 2) * (D / 4) + i * 32 + lane] = a;
      ms[(m * 2 + 1) * (D / 4) + i * 32 + lane] = __ldg(psh + i * 32 + lane);
    }
  }
  __syncwarp();

  for (int t = t0; t < t1; ++t) {
    if (t + 1 < t1) {
#pragma unroll
      for (int i = 0; i < NV; ++i) nx[i] = src[(size_t)(t + 1 - t0) * (D / 4) + i * 32 + lane];
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(sum) * (1.0f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + 1e-6f);
    const size_t row = (size_t)s * T + t;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;
    }
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      uint2* dst = reinterpret_cast<uint2*>((m == 0 ? out0 : out1) + row * D);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 a = ms[(m * 2) * (D / 4) + i * 32 + lane];
        const float4 b = ms[(m * 2 + 1) * (D / 4) + i * 32 + lane];
        uint2 pk;
        pk.x = pack_bf16x2(fmaf(v[i].x, a.x, b.x), fmaf(v[i].y, a.y, b.y));
        pk.y = pack_bf16x2(fmaf(v[i].z, a.z, b.z), fmaf(v[i].w, a.w, b.w));
        dst[i * 32 + lane] = pk;
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = nx[i];
  }
}

template <int NV>
static void launch_ln_t(const float* h, int S, int T, const float* mod0, int ld_mod0, __nv_bfloat16* out0,
                        const float* mod1, int ld_mod1, __nv_bfloat16* out1, int rev, cudaStream_t stream) {
  // runs of ~24 rows: long enough to amortise the modulation vectors, short enough to fill the SMs (S*T/24 warps)
  const int runs = (T + 23) / 24;
  const int run_len = (T + runs - 1) / runs;
  const int tasks = S * runs;
  const int blocks = (tasks + 3) / 4;   // 4 warps per block, 8 KB (16 KB dual) of modulation vectors per warp at D=1024
  const size_t smem1 = (size_t)4 * 2 * NV * 128 * sizeof(float);
  static PerDeviceFlag attr;   // per device; the dual variant needs 64 KB at D = 1024
  if (attr.first())
    cudaFuncSetAttribute(ln_modulate_kernel<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * smem1));
  if (out1 != nullptr)
    ln_modulate_kernel<NV, true><<<blocks, 128, 2 * smem1, stream>>>(h, S, T, runs, run_len, rev, mod0, ld_mod0, out0, mod1, ld_mod1, out1);
  else
    ln_modulate_kernel<NV, false><<<blocks, 128, smem1, stream>>>(h, S, T, runs, run_len, rev, mod0, ld_mod0, out0, nullptr, 0, nullptr);
}

int launch_ln_modulate(const float* h, int M, int D, int T, const float* mod0, int ld_mod0,
                       __nv_bfloat16* out0, const float* mod1, int ld_mod1, __nv_bfloat16* out1,
                       cudaStream_t stream, int rev) {
  MMDM_REQUIRE(D == 128 || D == 256 || D == 512 || D == 1024, "ln_modulate: unsupported latent dim %d", D);
  MMDM_REQUIRE((ld_mod0 % 4) == 0 && (out1 == nullptr || (ld_mod1 % 4) == 0), "ln_modulate: mod pitch %% 4");
  MMDM_REQUIRE(T > 0 && M % T == 0, "ln_modulate: M=%d is not a whole number of T=%d row sequences", M, T);
  const int S = M / T;
  switch (D) {
    case 128: launch_ln_t<1>(h, S, T, mod0, ld_mod0, out0, mod1, ld_mod1, out1, rev, stream); break;
    case 256: launch_ln_t<2>(h, S, T, mod0, ld_mod0, out0, mod1, ld_mod1, out1, rev, stream); break;
    case 512: launch_ln_t<4>(h, S, T, mod0, ld_mod0, out0, mod1, ld_mod1, out1, rev, stream); break;
    default: launch_ln_t<8>(h, S, T, mod0, ld_mod0, out0, mod1, ld_mod1, out1, rev, stream); break;
  }
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// LN fold, first block of a network: the rows come from the embedding GEMM, whose epilogue knows nothing of the fold,
// so this pass emits what the residual GEMMs emit for the later blocks (gemm.cuh): out_m = bf16(h * (1 + scale_m[s])) for
// one or two AdaLN sites and the per-row (sum h, sum h^2) of every 64-column slot. Same traffic as ln_modulate.
// ------------------------------------------------------------------------------------------------
template <int NV, bool DUAL>
__global__ void __launch_bounds__(128)
ln_fold_first_kernel(const float* __restrict__ h, int S, int T, int runs_per_seq, int run_len, int rev,
                     const float* __restrict__ sc0, int ld0, __nv_bfloat16* __restrict__ out0,
                     const float* __restrict__ sc1, int ld1, __nv_bfloat16* __restrict__ out1,
                     float2* __restrict__ stats, long long pitch, uint4* __restrict__ rowa, float* __restrict__ moff_out) {
  constexpr int D = NV * 128;
  constexpr int NM = DUAL ? 2 : 1;
  extern __shared__ float4 ln_smem[];   // [4 warps][NM][D/4]: 1 + scale
  int task = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (task >= S * runs_per_seq) return;
  if (rev) task = S * runs_per_seq - 1 - task;
  const int s = task / runs_per_seq;
  const int t0 = (task - s * runs_per_seq) * run_len;
  const int t1 = min(T, t0 + run_len);
  if (t0 >= t1) return;
  float4* ms = ln_smem + (size_t)(threadIdx.x >> 5) * (NM * (D / 4));
#pragma unroll
  for (int m = 0; m < NM; ++m) {
    const float4* psc = reinterpret_cast<const float4*>((m == 0 ? sc0 : sc1) + (size_t)s * (m == 0 ? ld0 : ld1));
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 a = __ldg(psc + i * 32 + lane);
      a.x += 1.f; a.y += 1.f; a.z += 1.f; a.w += 1.f;
      ms[m * (D / 4) + i * 32 + lane] = a;
    }
  }
  __syncwarp();
  const float4* src = reinterpret_cast<const float4*>(h + ((size_t)s * T + t0) * D);
  float4 v[NV], nx[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = src[i * 32 + lane];
  for (int t = t0; t < t1; ++t) {
    if (t + 1 < t1) {
#pragma unroll
      for (int i = 0; i < NV; ++i) nx[i] = src[(size_t)(t + 1 - t0) * (D / 4) + i * 32 + lane];
    }
    const size_t row = (size_t)s * T + t;
    // This pass sees the whole row: operand and row sums are those of the exactly centred row (GemmEpi::lnp_moff_in:
    // the offset m_r is the row mean itself here), the mean goes to moff_out for the next site's producer.
    float rsum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) rsum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(rsum) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    }
    if (moff_out != nullptr && lane == 0) moff_out[row] = mean;
    float t1 = 0.f, t2 = 0.f;   // whole-row sums, slots added in order
#pragma unroll
    for (int i = 0; i < NV; ++i) {   // lanes 0-15 hold slot 2i, lanes 16-31 slot 2i + 1
      float a = (v[i].x + v[i].y) + (v[i].z + v[i].w);
      float b = fmaf(v[i].x, v[i].x, v[i].y * v[i].y) + fmaf(v[i].z, v[i].z, v[i].w * v[i].w);
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if ((lane & 15) == 0) stats[(size_t)(2 * i + (lane >> 4)) * pitch + row] = make_float2(a, b);
      const float a_hi = __shfl_sync(0xffffffffu, a, 16), b_hi = __shfl_sync(0xffffffffu, b, 16);
      t1 = (t1 + a) + a_hi;   // (meaningful in lane 0: slot 2i, then slot 2i + 1)
      t2 = (t2 + b) + b_hi;
    }
    if (rowa != nullptr && lane == 0) {   // the finished per-row values (gemm.cuh: GemmEpi::lnp_rowa)
      const float mu = t1 * (1.0f / D);
      const float var = fmaxf(fmaf(-mu, mu, t2 * (1.0f / D)), 0.f) + 1e-6f;
      const float rho = rsqrtf(var), sg = var * rho;
      const __nv_bfloat16 mh = __float2bfloat16_rn(-mu), ml = __float2bfloat16_rn(-mu - __bfloat162float(mh));
      const __nv_bfloat16 sh = __float2bfloat16_rn(sg), sl = __float2bfloat16_rn(sg - __bfloat162float(sh));
      const uint32_t umh = __bfloat16_as_ushort(mh), uml = __bfloat16_as_ushort(ml), ush = __bfloat16_as_ushort(sh),
                     usl = __bfloat16_as_ushort(sl);
      rowa[row] = make_uint4(umh | (umh << 16), uml | (ush << 16), ush | (usl << 16), __float_as_uint(rho));
    }
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      uint2* dst = reinterpret_cast<uint2*>((m == 0 ? out0 : out1) + row * D);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 a = ms[m * (D / 4) + i * 32 + lane];
        uint2 pk;
        pk.x = pack_bf16x2(v[i].x * a.x, v[i].y * a.y);
        pk.y = pack_bf16x2(v[i].z * a.z, v[i].w * a.w);
        dst[i * 32 + lane] = pk;
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = nx[i];
  }
}

template <int NV>
static void launch_ln_fold_first_t(const float* h, int S, int T, const float* sc0, int ld0, __nv_bfloat16* out0,
                                   const float* sc1, int ld1, __nv_bfloat16* out1, float2* stats, long long pitch,
                                   uint4* rowa, float* moff_out, int rev, cudaStream_t stream) {
  const int runs = (T + 23) / 24;
  const int run_len = (T + runs - 1) / runs;
  const int tasks = S * runs;
  const int blocks = (tasks + 3) / 4;
  const size_t smem1 = (size_t)4 * NV * 128 * sizeof(float);
  if (out1 != nullptr)
    ln_fold_first_kernel<NV, true><<<blocks, 128, 2 * smem1, stream>>>(h, S, T, runs, run_len, rev, sc0, ld0, out0, sc1, ld1, out1, stats, pitch, rowa, moff_out);
  else
    ln_fold_first_kernel<NV, false><<<blocks, 128, smem1, stream>>>(h, S, T, runs, run_len, rev, sc0, ld0, out0, nullptr, 0, nullptr, stats, pitch, rowa, moff_out);
}

int launch_ln_fold_first(const float* h, int M, int D, int T, const float* sc0, int ld0, __nv_bfloat16* out0,
                         const float* sc1, int ld1, __nv_bfloat16* out1, float2* stats, long long pitch, uint4* rowa,
                         float* moff_out, cudaStream_t stream, int rev) {
  MMDM_REQUIRE(D == 128 || D == 256 || D == 512 || D == 1024, "ln_fold_first: unsupported latent dim %d", D);
  MMDM_REQUIRE((ld0 % 4) == 0 && (out1 == nullptr || (ld1 % 4) == 0), "ln_fold_first: scale pitch %% 4");
  MMDM_REQUIRE(T > 0 && M % T == 0 && pitch >= M && stats != nullptr, "ln_fold_first: M=%d is not a whole number of T=%d row sequences", M, T);
  const int S = M / T;
  switch (D) {   // (32 KB of shared memory at most: no attribute needed)
    case 128: launch_ln_fold_first_t<1>(h, S, T, sc0, ld0, out0, sc1, ld1, out1, stats, pitch, rowa, moff_out, rev, stream); break;
    case 256: launch_ln_fold_first_t<2>(h, S, T, sc0, ld0, out0, sc1, ld1, out1, stats, pitch, rowa, moff_out, rev, stream); break;
    case 512: launch_ln_fold_first_t<4>(h, S, T, sc0, ld0, out0, sc1, ld1, out1, stats, pitch, rowa, moff_out, rev, stream); break;
    default: launch_ln_fold_first_t<8>(h, S, T, sc0, ld0, out0, sc1, ld1, out1, stats, pitch, rowa, moff_out, rev, stream); break;
  }
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// E[s, :] = bf16( SiLU( temb[step, :] + text[text_row[s], :] ) )    (AdaLN input: layers.py:9-11 with
// in2in.py:414-424 / mixermdm.py:677-682). `step` is read from device memory so the launch can be
// replayed from a CUDA graph for every DDIM step.
// ------------------------------------------------------------------------------------------------
__global__ void cond_silu_kernel(const float* __restrict__ temb, const int* __restrict__ temb_idx,
                                 const int* __restrict__ step_ptr, const float* __restrict__ text,
                                 const int* __restrict__ text_row, int S, int D, __nv_bfloat16* __restrict__ E) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * D) return;
  const int s = idx / D, d = idx - s * D;
  const int trow = step_ptr ? *step_ptr : temb_idx[s];
  const float x = temb[(size_t)trow * D + d] + text[(size_t)text_row[s] * D + d];
  const __nv_bfloat16 e = __float2bfloat16_rn(x / (1.0f + expf(-x)));
  E[(size_t)s * 2 * D + d] = e;       // rows are [e | e], pitch 2D (see engine.cu run_conditioning)
  E[(size_t)s * 2 * D + D + d] = e;
}

int launch_cond_silu(const float* temb, const int* temb_idx, const int* step_ptr, const float* text,
                     const int* text_row, int S, int D, __nv_bfloat16* E, cudaStream_t stream) {
  const int n = S * D;
  cond_silu_kernel<<<(n + 255) / 256, 256, 0, stream>>>(temb, temb_idx, step_ptr, text, text_row, S, D, E);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// sequence -> row maps for the conditioning tables (see engine.cu for the table layouts)
__global__ void build_maps_kernel(int* map0, int* map1, int* tmap, int S, int persons, int Bt, int dup,
                                  int nslices, int k0_base) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int bx = s / persons, p = s - bx * persons;
  const int c = dup ? bx / Bt : 0;
  const int b = dup ? bx - c * Bt : bx;
  map0[s] = c ? nslices * Bt : (k0_base + p) * Bt + b;
  map1[s] = c ? nslices * Bt : b;
  tmap[s] = bx;
}
int launch_build_maps(int* map0, int* map1, int* tmap, int S, int persons, int Bt, int dup, int nslices,
                      int k0_base, cudaStream_t stream) {
  build_maps_kernel<<<(S + 127) / 128, 128, 0, stream>>>(map0, map1, tmap, S, persons, Bt, dup, nslices, k0_base);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// out[r, :] = table[idx[r], :]   (pe[timesteps], models/utils/utils.py:54-55)
__global__ void gather_rows_kernel(const float* __restrict__ table, int rows, const long long* __restrict__ idx,
                                   const int* __restrict__ idx32, int R, int D, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * D) return;
  const int r = i / D, d = i - r * D;
  long long t = idx ? idx[r] : (long long)idx32[r];
  t = t < 0 ? 0 : (t >= rows ? rows - 1 : t);
  out[i] = table[(size_t)t * D + d];
}
int launch_gather_rows(const float* table, int rows, const long long* idx64, const int* idx32, int R, int D,
                       float* out, cudaStream_t stream) {
  const int n = R * D;
  gather_rows_kernel<<<(n + 255) / 256, 256, 0, stream>>>(table, rows, idx64, idx32, R, D, out);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// out[r, n] = act( sum_k a[r, k] * w[n, k] + b[n] )  in fp32; one warp per output element.
// Used once per batch / per weight load (text_embed, time_embed MLP), never per token.
// ------------------------------------------------------------------------------------------------
__global__ void small_linear_kernel(const float* __restrict__ a, int lda, const float* __restrict__ w,
                                    const float* __restrict__ b, float* __restrict__ out, int ldo, int R, int N,
                                    int K, int act) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)R * N) return;
  const int r = (int)(warp / N), n = (int)(warp - (long long)r * N);
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc += a[(size_t)r * lda + k] * w[(size_t)n * K + k];
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc + (b ? b[n] : 0.f);
    if (act == 1) v = v / (1.0f + expf(-v));
    out[(size_t)r * ldo + n] = v;
  }
}

int launch_small_linear(const float* a, int lda, const float* w, const float* b, float* out, int ldo, int R,
                        int N, int K, int act, cudaStream_t stream) {
  const long long warps = (long long)R * N;
  const int threads = 256;
  const long long blocks = (warps * 32 + threads - 1) / threads;
  small_linear_kernel<<<(unsigned)blocks, threads, 0, stream>>>(a, lda, w, b, out, ldo, R, N, K, act);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// state (B, T, persons*nf) fp32 -> bf16 GEMM operand rows ((b*persons+p)*T + t, kpad), zero padded to kpad.
// ------------------------------------------------------------------------------------------------
// One thread per 8 output elements (one 16-byte store); kpad % 8 == 0, person blocks start 8-byte aligned (nf even).
__global__ void pack_state_kernel(const float* __restrict__ x, int B, int T, int nf, int kpad, int persons,
                                  int row_pitch, __nv_bfloat16* __restrict__ out) {
  const int chunks = kpad >> 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = B * persons * T * chunks;
  if (idx >= total) return;
  const int row = idx / chunks, k0 = (idx - row * chunks) << 3;
  const int sp = row / T, t = row - sp * T;
  const int b = sp / persons, p = sp - b * persons;
  const float* src = x + ((size_t)b * T + t) * row_pitch + p * nf + k0;
  float v[8];
  if (k0 + 8 <= nf && ((nf & 1) == 0) && ((row_pitch & 1) == 0)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __ldg(reinterpret_cast<const float2*>(src) + i);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = k0 + i < nf ? __ldg(src + i) : 0.f;
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(out + (size_t)row * kpad + k0) = o;
}

int launch_pack_state(const float* x, int B, int T, int nf, int kpad, int persons, __nv_bfloat16* out,
                      cudaStream_t stream, int row_pitch) {
  if (row_pitch <= 0) row_pitch = persons * nf;
  MMDM_REQUIRE((kpad % 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0,
               "pack_state: kpad must be a multiple of 8 and the buffers aligned");
  const long long total = (long long)B * persons * T * (kpad / 8);
  MMDM_REQUIRE(total < (1LL << 31), "pack_state: problem too large");
  pack_state_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, B, T, nf, kpad, persons, row_pitch, out);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// step counter helpers (device-resident so that graph replays advance through the DDIM schedule)
__global__ void set_int_kernel(int* p, int v) { *p = v; }
__global__ void add_int_kernel(int* p, int v) { *p += v; }
int launch_set_int(int* p, int v, cudaStream_t stream) {
  set_int_kernel<<<1, 1, 0, stream>>>(p, v);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_add_int(int* p, int v, cudaStream_t stream) {
  add_int_kernel<<<1, 1, 0, stream>>>(p, v);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// x1 += scale * n1, x2 += scale * n2 on the interleaved state rows (x1 | x2, 2 * rowf floats per frame): the noise term of
// an ancestral (DDPM) update, reference GaussianDiffusion.p_sample (gaussian_diffusion.py:683-691). float4 per thread.
__global__ void add_state_noise_kernel(float* __restrict__ x1, float* __restrict__ x2, const float* __restrict__ n1,
                                       const float* __restrict__ n2, long long nrows, int row4, float scale) {
  const long long total = nrows * row4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / row4;
    const int c = (int)(i - r * row4);
    const float4 a = reinterpret_cast<const float4*>(n1)[i], b = reinterpret_cast<const float4*>(n2)[i];
    float4* p1 = reinterpret_cast<float4*>(x1) + r * 2 * row4 + c;
    float4* p2 = reinterpret_cast<float4*>(x2) + r * 2 * row4 + c;
    float4 u = *p1, v = *p2;
    u.x = fmaf(scale, a.x, u.x); u.y = fmaf(scale, a.y, u.y); u.z = fmaf(scale, a.z, u.z); u.w = fmaf(scale, a.w, u.w);
    v.x = fmaf(scale, b.x, v.x); v.y = fmaf(scale, b.y, v.y); v.z = fmaf(scale, b.z, v.z); v.w = fmaf(scale, b.w, v.w);
    *p1 = u;
    *p2 = v;
  }
}
int launch_add_state_noise(float* x1, float* x2, const float* n1, const float* n2, long long nrows, int rowf, float scale,
                           cudaStream_t stream) {
  MMDM_REQUIRE(rowf % 4 == 0 && ((uintptr_t)x1 % 16 == 0) && ((uintptr_t)x2 % 16 == 0) && ((uintptr_t)n1 % 16 == 0) &&
               ((uintptr_t)n2 % 16 == 0), "add_state_noise: rows must be float4-aligned");
  const long long total = nrows * (rowf / 4);
  const long long want = (total + 255) / 256, cap = (long long)device_sm_count() * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  add_state_noise_kernel<<<grid, 256, 0, stream>>>(x1, x2, n1, n2, nrows, rowf / 4, scale);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mmdm
