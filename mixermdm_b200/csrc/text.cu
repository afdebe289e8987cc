// Text-conditioning path: the step in front of the denoising loop.
//   reference: MixerMDM.text_process / generate_cond  models/mixermdm.py:283-356
//              in2IN.text_process                      models/in2in.py:109-135
// Both run a frozen CLIP ViT-L/14 text tower (third-party: OpenAI `clip` 1.0, clip/model.py - 12 pre-LN residual
// attention blocks of width 768 with 12 heads, causal mask, QuickGELU MLP, LayerNorm eps 1e-5, ln_final), then a
// 2-layer post-norm nn.TransformerEncoder (d_model 768, 8 heads, ff 2048, exact GELU) and a LayerNorm, and keep the
// row of the end-of-text token (the arg-max token id). Here: the linears run on the tcgen05 GEMM with fused
// epilogues, the 77-token attention (head dims 64 / 96, optionally causal, no zero-attention key) on a small
// shared-memory kernel, LayerNorm and the gathers on row kernels. One tower pass feeds any number of encoders.
#include "../../include/mixermdm_b200.h"
#include "common.cuh"
#include "gemm.cuh"

namespace mmdm {

typedef __nv_bfloat16 bf16;

// h[r, :] = tok_emb[tokens[r], :] + pos_emb[r % ctx, :]     (models/mixermdm.py:303-304)
__global__ void text_embed_kernel(const int32_t* __restrict__ tokens, const float* __restrict__ tok_emb, int vocab,
                                  const float* __restrict__ pos_emb, int ctx, int rows, int D, float* __restrict__ h) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int D4 = D >> 2;
  if (i >= rows * D4) return;
  const int r = i / D4, c = i - r * D4;
  int tok = tokens[r];
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
  const float4 a = __ldg(reinterpret_cast<const float4*>(tok_emb + (size_t)tok * D) + c);
  const float4 b = __ldg(reinterpret_cast<const float4*>(pos_emb + (size_t)(r % ctx) * D) + c);
  reinterpret_cast<float4*>(h + (size_t)r * D)[c] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// LayerNorm with affine parameters over rows of D = 32 * NV * 4 floats, one warp per row:
//   y = (x - mean) * rsqrt(var + eps) * gamma + beta  ->  bf16 copy (GEMM operand) and / or fp32 (residual stream of
//   the post-norm encoder; may alias x).
template <int NV>
__global__ void __launch_bounds__(128)
ln_affine_kernel(const float* __restrict__ x, int rows, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, bf16* __restrict__ out_bf16, float* out_f32) {
  constexpr int D = NV * 128;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(x + (size_t)row * D);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = src[i * 32 + lane];
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
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
    const float4 y = make_float4(fmaf(v[i].x * rstd, g.x, b.x), fmaf(v[i].y * rstd, g.y, b.y),
                                 fmaf(v[i].z * rstd, g.z, b.z), fmaf(v[i].w * rstd, g.w, b.w));
    if (out_f32 != nullptr) reinterpret_cast<float4*>(out_f32 + (size_t)row * D)[i * 32 + lane] = y;
    if (out_bf16 != nullptr) {
      uint2 pk;
      pk.x = pack_bf16x2(y.x, y.y);
      pk.y = pack_bf16x2(y.z, y.w);
      reinterpret_cast<uint2*>(out_bf16 + (size_t)row * D)[i * 32 + lane] = pk;
    }
  }
}

static int launch_ln_affine(const float* x, int rows, int D, const float* gamma, const float* beta, float eps,
                            bf16* out_bf16, float* out_f32, cudaStream_t st) {
  MMDM_REQUIRE(D == 768 || D == 512 || D == 1024, "text LayerNorm: unsupported width %d", D);
  const int blocks = (rows + 3) / 4;
  if (D == 768) ln_affine_kernel<6><<<blocks, 128, 0, st>>>(x, rows, gamma, beta, eps, out_bf16, out_f32);
  else if (D == 512) ln_affine_kernel<4><<<blocks, 128, 0, st>>>(x, rows, gamma, beta, eps, out_bf16, out_f32);
  else ln_affine_kernel<8><<<blocks, 128, 0, st>>>(x, rows, gamma, beta, eps, out_bf16, out_f32);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Attention over one short sequence (ctx <= 96 tokens) per CTA and head: qkv rows hold [q | k | v] of width 3 * D,
// head h at columns h * hd. softmax(q k^T / sqrt(hd) [+ causal mask]) v, fp32 arithmetic, bf16 in / out.
// nn.MultiheadAttention semantics without zero-attention key (CLIP blocks: causal attn_mask, clip/model.py
// build_attention_mask; encoder layers: no mask - padding tokens take part, as in the reference).
constexpr int TEXT_MAX_CTX = 96;
template <int HD>
__global__ void __launch_bounds__(128)
text_attention_kernel(const bf16* __restrict__ qkv, int ctx, int D, int H, int causal, float scale,
                      bf16* __restrict__ out) {
  constexpr int PITCH = HD + 2;   // odd number of 32-bit words per row: the per-key dot products are conflict-free
  __shared__ bf16 Ks[TEXT_MAX_CTX * PITCH];
  __shared__ bf16 Vs[TEXT_MAX_CTX * PITCH];
  __shared__ float Qs[4][HD];
  const int s = blockIdx.x / H, h = blockIdx.x - s * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* base = qkv + (size_t)s * ctx * (3 * D) + h * HD;
  for (int i = threadIdx.x; i < ctx * (HD / 2); i += blockDim.x) {
    const int t = i / (HD / 2), c = i - t * (HD / 2);
    const __nv_bfloat162 kk = *reinterpret_cast<const __nv_bfloat162*>(base + (size_t)t * (3 * D) + D + 2 * c);
    const __nv_bfloat162 vv = *reinterpret_cast<const __nv_bfloat162*>(base + (size_t)t * (3 * D) + 2 * D + 2 * c);
    *reinterpret_cast<__nv_bfloat162*>(&Ks[t * PITCH + 2 * c]) = kk;
    *reinterpret_cast<__nv_bfloat162*>(&Vs[t * PITCH + 2 * c]) = vv;
  }
  __syncthreads();
  for (int i = warp; i < ctx; i += 4) {
    for (int d = lane; d < HD; d += 32) Qs[warp][d] = __bfloat162float(base[(size_t)i * (3 * D) + d]) * scale;
    __syncwarp();
    const int nk = causal ? i + 1 : ctx;
    float sc[3];
    float mx = -INFINITY;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int j = lane + 32 * r;
      float acc = -INFINITY;
      if (j < nk) {
        acc = 0.f;
#pragma unroll 8
        for (int d = 0; d < HD; d += 2) {
          const float2 kf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&Ks[j * PITCH + d]));
          acc = fmaf(Qs[warp][d], kf.x, acc);
          acc = fmaf(Qs[warp][d + 1], kf.y, acc);
        }
      }
      sc[r] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      sc[r] = (lane + 32 * r < nk) ? __expf(sc[r] - mx) : 0.f;
      sum += sc[r];
    }
    const float inv = 1.0f / warp_sum(sum);
    float o[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < nk; ++j) {
      const float pj = __shfl_sync(0xffffffffu, sc[j >> 5], j & 31);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int d = lane + 32 * r;
        if (d < HD) o[r] = fmaf(pj, __bfloat162float(Vs[j * PITCH + d]), o[r]);
      }
    }
    bf16* dst = out + ((size_t)s * ctx + i) * D + h * HD;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int d = lane + 32 * r;
      if (d < HD) dst[d] = __float2bfloat16_rn(o[r] * inv);
    }
    __syncwarp();
  }
}

static int launch_text_attention(const bf16* qkv, int nseq, int ctx, int D, int H, int causal, bf16* out,
                                 cudaStream_t st) {
  const int hd = D / H;
  MMDM_REQUIRE(ctx >= 1 && ctx <= TEXT_MAX_CTX, "text attention: context of %d tokens (max %d)", ctx, TEXT_MAX_CTX);
  MMDM_REQUIRE(hd * H == D && (hd == 64 || hd == 96), "text attention: head dim %d not in {64, 96}", hd);
  const float scale = 1.0f / sqrtf((float)hd);
  if (hd == 64) text_attention_kernel<64><<<nseq * H, 128, 0, st>>>(qkv, ctx, D, H, causal, scale, out);
  else text_attention_kernel<96><<<nseq * H, 128, 0, st>>>(qkv, ctx, D, H, causal, scale, out);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// out[s, :] = x[s * ctx + idx[s], :]      (the end-of-text row: models/mixermdm.py:315)
__global__ void text_pick_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, int ctx, int D,
                                      float* __restrict__ out, int ldo, int nseq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseq * D) return;
  const int s = i / D, d = i - s * D;
  int t = idx[s];
  t = t < 0 ? 0 : (t >= ctx ? ctx - 1 : t);
  out[(size_t)s * ldo + d] = x[((size_t)s * ctx + t) * D + d];
}

struct TextWs {
  float *h, *c32, *e32;
  bf16 *xn, *cb, *qkv, *att, *mid;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t carve_text(int nseq, int ctx, int D, int Fmax, void* base, TextWs* w) {
  const size_t M = (size_t)nseq * ctx;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = align_up(off, 1024);
    void* r = p ? p + off : nullptr;
    off += bytes;
    return r;
  };
  TextWs ws;
  ws.h = (float*)take(M * D * 4);
  ws.c32 = (float*)take(M * D * 4);
  ws.e32 = (float*)take(M * D * 4);
  ws.xn = (bf16*)take(M * D * 2);
  ws.cb = (bf16*)take(M * D * 2);
  ws.qkv = (bf16*)take(M * 3 * D * 2);
  ws.att = (bf16*)take(M * D * 2);
  ws.mid = (bf16*)take(M * Fmax * 2);
  if (w) *w = ws;
  return align_up(off, 1024);
}

static int text_gemm(const bf16* A, int lda, const void* W, int M, int N, int K, int mode, const float* bias, void* out,
                     int ldo, const float* resid, cudaStream_t st) {
  GemmEpi ep;
  ep.mode = mode; ep.bias = bias; ep.out = out; ep.ldo = ldo; ep.resid = resid;
  return launch_gemm(A, lda, reinterpret_cast<const bf16*>(W), K, M, N, K, ep, st);
}

}  // namespace mmdm

using namespace mmdm;

extern "C" {

size_t mmdm_text_workspace_bytes(int nseq, int ctx, int width, int ff_max) {
  if (nseq < 1 || ctx < 1 || width < 1 || ff_max < 1) return 0;
  return carve_text(nseq, ctx, width, ff_max, nullptr, nullptr);
}

int mmdm_text_encode(const mmdm_text_tower* tw, const mmdm_text_encoder* encs, int n_enc, const int32_t* tokens,
                     const int32_t* eot, int nseq, float* const* cond_out, int cond_ld, void* ws, size_t ws_bytes,
                     void* stream) {
  MMDM_REQUIRE(tw && encs && tokens && eot && cond_out && ws, "mmdm_text_encode: null argument");
  MMDM_REQUIRE(nseq >= 1 && n_enc >= 1, "mmdm_text_encode: nothing to encode");
  const int D = tw->width, ctx = tw->ctx;
  MMDM_REQUIRE(D == 768 && tw->heads >= 1 && D % tw->heads == 0 && tw->layers >= 1 && tw->blocks,
               "mmdm_text_encode: unsupported tower geometry (width %d)", D);
  int Fmax = 4 * D;
  for (int e = 0; e < n_enc; ++e) {
    MMDM_REQUIRE(encs[e].layers_ && encs[e].layers >= 1 && encs[e].heads >= 1 && D % encs[e].heads == 0,
                 "mmdm_text_encode: encoder %d is incomplete", e);
    Fmax = encs[e].ff > Fmax ? encs[e].ff : Fmax;
    MMDM_REQUIRE(cond_out[e] != nullptr && cond_ld >= D, "mmdm_text_encode: output %d missing / pitch too small", e);
  }
  MMDM_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "text workspace must be 1024-byte aligned");
  MMDM_REQUIRE(ws_bytes >= carve_text(nseq, ctx, D, Fmax, nullptr, nullptr), "text workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TextWs w;
  carve_text(nseq, ctx, D, Fmax, ws, &w);
  const int M = nseq * ctx;

  // ---- CLIP text tower (frozen): x = tok + pos; 12 x { x += attn(ln_1 x) ; x += mlp(ln_2 x) }; ln_final ---------
  {
    const int n = M * (D / 4);
    text_embed_kernel<<<(n + 255) / 256, 256, 0, st>>>(tokens, tw->tok_emb, tw->vocab, tw->pos_emb, ctx, M, D, w.h);
    MMDM_CHECK_CUDA(cudaGetLastError());
  }
  for (int l = 0; l < tw->layers; ++l) {
    const mmdm_clip_block& b = tw->blocks[l];
    MMDM_TRY(launch_ln_affine(w.h, M, D, b.ln1_g, b.ln1_b, 1e-5f, w.xn, nullptr, st));
    MMDM_TRY(text_gemm(w.xn, D, b.attn_in_w, M, 3 * D, D, EPI_BF16, b.attn_in_b, w.qkv, 3 * D, nullptr, st));
    MMDM_TRY(launch_text_attention(w.qkv, nseq, ctx, D, tw->heads, 1, w.att, st));
    MMDM_TRY(text_gemm(w.att, D, b.attn_out_w, M, D, D, EPI_RESID, b.attn_out_b, w.h, D, w.h, st));
    MMDM_TRY(launch_ln_affine(w.h, M, D, b.ln2_g, b.ln2_b, 1e-5f, w.xn, nullptr, st));
    MMDM_TRY(text_gemm(w.xn, D, b.fc_w, M, 4 * D, D, EPI_QGELU_BF16, b.fc_b, w.mid, 4 * D, nullptr, st));
    MMDM_TRY(text_gemm(w.mid, 4 * D, b.proj_w, M, D, 4 * D, EPI_RESID, b.proj_b, w.h, D, w.h, st));
  }
  MMDM_TRY(launch_ln_affine(w.h, M, D, tw->lnf_g, tw->lnf_b, 1e-5f, w.cb, w.c32, st));

  // ---- per encoder: 2 x post-norm nn.TransformerEncoderLayer, LayerNorm, end-of-text row ----------------------
  for (int e = 0; e < n_enc; ++e) {
    const mmdm_text_encoder& en = encs[e];
    MMDM_CHECK_CUDA(cudaMemcpyAsync(w.e32, w.c32, (size_t)M * D * 4, cudaMemcpyDeviceToDevice, st));
    const bf16* xb = w.cb;
    for (int l = 0; l < en.layers; ++l) {
      const mmdm_enc_layer& b = en.layers_[l];
      MMDM_TRY(text_gemm(xb, D, b.in_w, M, 3 * D, D, EPI_BF16, b.in_b, w.qkv, 3 * D, nullptr, st));
      MMDM_TRY(launch_text_attention(w.qkv, nseq, ctx, D, en.heads, 0, w.att, st));
      MMDM_TRY(text_gemm(w.att, D, b.out_w, M, D, D, EPI_RESID, b.out_b, w.e32, D, w.e32, st));
      MMDM_TRY(launch_ln_affine(w.e32, M, D, b.norm1_g, b.norm1_b, 1e-5f, w.xn, w.e32, st));
      MMDM_TRY(text_gemm(w.xn, D, b.lin1_w, M, en.ff, D, EPI_GELU_BF16, b.lin1_b, w.mid, en.ff, nullptr, st));
      MMDM_TRY(text_gemm(w.mid, en.ff, b.lin2_w, M, D, en.ff, EPI_RESID, b.lin2_b, w.e32, D, w.e32, st));
      MMDM_TRY(launch_ln_affine(w.e32, M, D, b.norm2_g, b.norm2_b, 1e-5f, w.xn, w.e32, st));
      xb = w.xn;
    }
    MMDM_TRY(launch_ln_affine(w.e32, M, D, en.ln_g, en.ln_b, 1e-5f, nullptr, w.e32, st));
    {
      const int n = nseq * D;
      text_pick_rows_kernel<<<(n + 255) / 256, 256, 0, st>>>(w.e32, eot, ctx, D, cond_out[e], cond_ld, nseq);
      MMDM_CHECK_CUDA(cudaGetLastError());
    }
  }
  return 0;
}

}  // extern "C"
