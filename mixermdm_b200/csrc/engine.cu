// Engine: orchestration of one Mixer step / the DDIM loop over the kernels in this directory.
// Sequence layout used everywhere: a batch row bx (CFG-doubled in the sampler: rows [0,B) conditional,
// [B,2B) unconditional) contributes two token sequences s = bx*2 + p (p = person / stream), T tokens each,
// so all three networks see GEMMs with M = 2*Bx*T rows and the two streams of the interaction denoiser are
// batched along M (they share weights: models/in2in.py:439-440), as are the two persons of the individual
// denoiser (models/mixermdm.py:685-686) and of the Influence network (:735-736).
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mixermdm_b200.h"
#include "attention.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "mix_stream.cuh"
#include "motion.cuh"
#include "rowops.cuh"

namespace mmdm {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------
// profiler: CUDA events around every launch, only when enabled and not capturing
// ------------------------------------------------------------------------------------------------
struct ProfRec {
  std::string name;
  long long launches = 0;
  double ms = 0, flops = 0, bytes = 0;
};
struct ProfPending {
  int rec;
  cudaEvent_t a, b;
};
struct Profiler {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<ProfPending> pending;
  std::vector<cudaEvent_t> pool;
  int find(const char* name) {
    for (size_t i = 0; i < recs.size(); ++i)
      if (recs[i].name == name) return (int)i;
    ProfRec r;
    r.name = name;
    recs.push_back(r);
    return (int)recs.size() - 1;
  }
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  void flush() {
    for (auto& p : pending) {
      cudaEventSynchronize(p.b);
      float ms = 0;
      cudaEventElapsedTime(&ms, p.a, p.b);
      recs[p.rec].ms += ms;
      pool.push_back(p.a);
      pool.push_back(p.b);
    }
    pending.clear();
  }
};

struct Net {
  mmdm_net_weights w;
  std::vector<mmdm_block_weights> blocks;
  bool present = false;
  bool fold = false;   // AdaLN may be folded into the neighbouring GEMMs: composite conditioning tables given (gemm.cuh)
  bool fold_on = false;   // ... and is, for the bound frame count (set by mmdm_bind_workspace)
  int hd() const { return w.D / w.H; }
};

struct Workspace {
  // shared by the three networks (they run one after another)
  float* h = nullptr;       // residual stream (M, Dmax)           [influence: (2M, 512) = h_i | m_I]
  bf16* xn = nullptr;       // LN+modulate output / bf16 copy of the last state (M, Dmax)
  bf16* xf = nullptr;       // xf_norm output (M, Dmax)
  bf16* qk = nullptr;       // Q | K (M, 2*Dmax)
  bf16* vt = nullptr;       // V^T (S*Dmax, Tp)
  bf16* att = nullptr;      // attention output (M, Dmax)
  bf16* mid = nullptr;      // FFN hidden (M, Fmax)
  bf16* xin = nullptr;      // packed embed operand (2*M, kpad)
  bf16 *E0 = nullptr, *E1 = nullptr;   // SiLU(emb), each row twice: (S, 2 Dmax) = [e | e] (K = 2D operand of the composite GEMM)
  float *mod0 = nullptr, *mod1 = nullptr;  // AdaLN outputs (S, ada_n max)
  float *gb0 = nullptr, *gb1 = nullptr;    // LN fold: G_s | Bc_s rows of every AdaLN site (S, comp_n max)
  float2* stat[3] = {nullptr, nullptr, nullptr};   // LN fold: row sums per 64-column slot [Dmax / 64][M]; block input /
                                                   // after self attention / after cross attention
  uint4* rowa[3] = {nullptr, nullptr, nullptr};    // ... and the finished per-row values (M)
  float* moff[3] = {nullptr, nullptr, nullptr};    // ... and the row means (M): the next site's per-row operand offsets
  int *map0 = nullptr, *map1 = nullptr, *tmap = nullptr;
  float* text[3] = {nullptr, nullptr, nullptr};   // per net: (nslices*Bx + 1, D)
  float* temb_tab[3] = {nullptr, nullptr, nullptr};  // per net: (nsteps, D)
  float* temb_dyn = nullptr;   // (Bx, Dmax) per-sample timestep embeddings (explicit-timestep entry points)
  float* tmp_rows = nullptr;   // (max(Bx,nsteps), Dmax) scratch for the time-embed MLP
  float* zero_row = nullptr;   // (1, text_dim)
  float *out1 = nullptr, *out2 = nullptr, *out1a = nullptr, *out2a = nullptr;  // (Bx, T, 524)
  bf16* hfin = nullptr;     // Influence output, bf16 (M, 512)
  bf16* hpool = nullptr;    // its mean over time (S, 512) for the pooled mixing modes
  float* wmix = nullptr;    // Mixer weights sigmoid(head): (M, 24) or (S, 24) fp32
  float *x1 = nullptr, *x2 = nullptr;   // sampler state (B, T, 524)
  float* final = nullptr;               // result of the last step (B, T, 524); copied to the caller's buffer
  int* step = nullptr;
  int* ts_idx = nullptr;    // (nsteps) int: timestep_map on device
  unsigned long long* kmask1 = nullptr;   // key-padding words, one sequence per batch row   [Bx][NB]
  unsigned long long* kmask2 = nullptr;   // ... two sequences (persons) per batch row        [2 Bx][NB]
};

}  // namespace mmdm

using namespace mmdm;

struct mmdm_handle {
  int device = 0;            // ordinal the handle was created on; every entry point makes it current
  std::string last_error;    // message of the last failed call on this handle (mmdm_last_error(h))
  int step_host = -1;        // host mirror of the device-side DDIM index (-1: no state set)
  bool mask_on = false;      // key-padding words in ws.kmask1 / kmask2 are live (mmdm_set_key_mask)
  int seq_persons = 2;       // sequences per batch row of the network pass being enqueued (selects kmask1 / kmask2)
  bool cond_ready = false;   // the text-embedding tables hold the sampler's condition (mmdm_set_condition); the
                             // explicit-batch entry points (mixer / denoiser forward) and a re-bind overwrite them
  mmdm_config cfg;
  Net net[3];
  std::vector<int32_t> timestep_map;
  Workspace ws;
  void* ws_base = nullptr;
  size_t ws_bytes = 0;
  int B = 0, T = 0, Bx = 0;   // bound un-doubled batch, frames, rows fed to the networks (2B)
  int Tp = 0;
  bool temb_ready = false;
  // One captured step graph per (workspace, batch, frames, history buffers): a caller that alternates between motion
  // lengths (the evaluation generators: T = motion_lens[0] per prompt) re-binds the workspace without re-capturing.
  struct GraphEntry {
    void* ws_base; int B, T; float *h1, *h2; bool masked;
    cudaGraphExec_t exec; long long step_launches; long long stamp;
  };
  std::vector<GraphEntry> graphs;
  long long graph_stamp = 0, graph_captures = 0;
  cudaStream_t cap_stream = nullptr;   // capture happens here: the legacy default stream cannot be captured
  Profiler prof;
  long long launches = 0;
  // Serpentine traversal: every big launch (LN, GEMM, attention) walks its rows in the direction opposite to the
  // launch before it, so a consumer starts on the rows its producer wrote last - the part of a 80-160 MB
  // activation that is still in the 126 MB L2. MMDM_SERPENTINE=0 turns it off (A/B measurements).
  bool serpentine = true;
  int rev = 0;
  bool fold_any_T = false;   // MMDM_LN_FOLD=2
};

namespace mmdm {

struct Scope {   // profiling scope around one launch
  mmdm_handle* h;
  int rec = -1;
  cudaEvent_t a = nullptr;
  cudaStream_t st;
  Scope(mmdm_handle* h_, const char* name, cudaStream_t st_, double flops = 0, double bytes = 0) : h(h_), st(st_) {
    h->launches++;
    if (!h->prof.on) return;
    rec = h->prof.find(name);
    h->prof.recs[rec].launches++;
    h->prof.recs[rec].flops += flops;
    h->prof.recs[rec].bytes += bytes;
    a = h->prof.get();
    cudaEventRecord(a, st);
  }
  ~Scope() {
    if (rec < 0) return;
    cudaEvent_t b = h->prof.get();
    cudaEventRecord(b, st);
    h->prof.pending.push_back({rec, a, b});
  }
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* b) : base(reinterpret_cast<uint8_t*>(b)) {}
  template <typename Tt>
  Tt* take(size_t n) {
    off = align_up(off, 1024);
    Tt* p = base ? reinterpret_cast<Tt*>(base + off) : nullptr;
    off += n * sizeof(Tt);
    return p;
  }
};

static void carve(const mmdm_handle* h, int B, int T, void* base, Workspace* w, size_t* total) {
  const int Bx = 2 * B, S = 2 * Bx;
  const size_t M = (size_t)S * T;
  int Dmax = 0, Fmax = 0, adamax = 0, compmax = 0;
  for (int i = 0; i < 3; ++i) {
    Dmax = std::max(Dmax, h->net[i].w.D);
    Fmax = std::max(Fmax, h->net[i].w.F);
    adamax = std::max(adamax, std::max(h->net[i].w.ada0_n, h->net[i].w.ada1_n));
    if (h->net[i].fold) compmax = std::max(compmax, std::max(h->net[i].w.comp0_n, h->net[i].w.comp1_n));
  }
  const int Tp = (T + 7) / 8 * 8;
  const int nf2 = 2 * h->cfg.nfeats;
  Carver c(base);
  Workspace ws;
  ws.h = c.take<float>(std::max(M * Dmax, 2 * M * (size_t)h->net[2].w.D));
  ws.xn = c.take<bf16>(M * Dmax);
  ws.xf = c.take<bf16>(M * Dmax);
  ws.qk = c.take<bf16>(M * 2 * Dmax);
  ws.vt = c.take<bf16>((size_t)S * Dmax * Tp);
  ws.att = c.take<bf16>(M * Dmax);
  ws.mid = c.take<bf16>(M * Fmax);
  ws.xin = c.take<bf16>(2 * M * h->cfg.kpad);
  ws.E0 = c.take<bf16>((size_t)S * 2 * Dmax);
  ws.E1 = c.take<bf16>((size_t)S * 2 * Dmax);
  if (compmax > 0) {
    ws.gb0 = c.take<float>((size_t)S * compmax);
    ws.gb1 = c.take<float>((size_t)S * compmax);
    for (int i = 0; i < 3; ++i) ws.stat[i] = c.take<float2>((size_t)(Dmax / 64) * M);
    for (int i = 0; i < 3; ++i) ws.rowa[i] = c.take<uint4>(M);
    for (int i = 0; i < 3; ++i) ws.moff[i] = c.take<float>(M);
  }
  ws.mod0 = c.take<float>((size_t)S * adamax);
  ws.mod1 = c.take<float>((size_t)S * adamax);
  ws.map0 = c.take<int>(S);
  ws.map1 = c.take<int>(S);
  ws.tmap = c.take<int>(S);
  for (int i = 0; i < 3; ++i) {
    ws.text[i] = c.take<float>((size_t)(3 * Bx + 1) * h->net[i].w.D);
    ws.temb_tab[i] = c.take<float>((size_t)h->cfg.nsteps * h->net[i].w.D);
  }
  ws.temb_dyn = c.take<float>((size_t)Bx * Dmax);
  ws.tmp_rows = c.take<float>((size_t)std::max(Bx, h->cfg.nsteps) * Dmax * 2);
  ws.zero_row = c.take<float>(h->cfg.text_dim);
  ws.out1 = c.take<float>((size_t)Bx * T * nf2);
  ws.out2 = c.take<float>((size_t)Bx * T * nf2);
  // aligned predictions and sampler state: the two buffers of each pair are interleaved row by row
  // ([out1a row | out2a row], [x1 row | x2 row]), so the fused mix kernel fetches a pair with one 4192-byte bulk copy
  ws.out1a = c.take<float>((size_t)2 * Bx * T * nf2);
  ws.out2a = ws.out1a ? ws.out1a + nf2 : nullptr;
  ws.hfin = c.take<bf16>(M * h->net[2].w.D);
  ws.hpool = c.take<bf16>((size_t)S * h->net[2].w.D);
  ws.wmix = c.take<float>(M * 24);
  ws.x1 = c.take<float>((size_t)2 * B * T * nf2);
  ws.x2 = ws.x1 ? ws.x1 + nf2 : nullptr;
  ws.final = c.take<float>((size_t)B * T * nf2);
  ws.step = c.take<int>(4);
  ws.ts_idx = c.take<int>(h->cfg.nsteps);
  {
    const size_t NB = (size_t)(T + 1 + 63) / 64;
    ws.kmask1 = c.take<unsigned long long>((size_t)Bx * NB);
    ws.kmask2 = c.take<unsigned long long>((size_t)2 * Bx * NB);
  }
  if (w) *w = ws;
  if (total) *total = align_up(c.off, 1024);
}

// ------------------------------------------------------------------------------------------------
// building blocks
// ------------------------------------------------------------------------------------------------
static int next_dir(mmdm_handle* h, bool flip = true) {
  if (flip) h->rev ^= 1;
  return h->serpentine ? h->rev : 0;
}

static int gemm(mmdm_handle* h, const char* tag, const bf16* A, int lda, const void* W, int ldw, int M, int N, int K,
                const GemmEpi& ep, cudaStream_t st, bool flip = true) {
  // algorithmic bytes: both operands once + what the epilogue reads and writes (for the HBM-bound residual GEMMs)
  double bytes = 2.0 * M * K + 2.0 * N * K;
  const double mn = (double)M * N;
  switch (ep.mode) {
    case EPI_RESID: bytes += 8.0 * mn + (ep.out_bf16 ? 2.0 * mn : 0.0); break;
    case EPI_RESID_LN: bytes += 8.0 * mn + 2.0 * mn * (ep.lnp_out1 ? 2 : 1); break;
    case EPI_EMBED_PE: case EPI_EMBED_PE_LN: bytes += 4.0 * mn * (ep.dup_row_offset ? 2 : 1); break;
    case EPI_F32: case EPI_FINAL: case EPI_SIGMOID_F32: bytes += 4.0 * mn; break;
    default: bytes += 2.0 * mn; break;
  }
  Scope sc(h, tag, st, 2.0 * M * N * K, bytes);
  GemmEpi e = ep;
  e.rev = next_dir(h, flip);
  return launch_gemm(A, lda, reinterpret_cast<const bf16*>(W), ldw, M, N, K, e, st);
}

struct Mods {   // AdaLN outputs of one block: pointers to the (scale|shift) rows of sequence 0 + pitches
  const float *sa, *ffn, *caq, *xf;
  int ld_sa, ld_ffn, ld_caq, ld_xf;
};

static Mods block_mods(const Net& n, const Workspace& ws, int l) {
  const int D = n.w.D;
  Mods m = {};
  switch (n.w.kind) {
    case MMDM_NET_INDIVIDUAL:
      m.sa = ws.mod0 + (size_t)l * 4 * D; m.ffn = m.sa + 2 * D;
      m.ld_sa = m.ld_ffn = n.w.ada0_n;
      break;
    case MMDM_NET_INTERACTION:
      m.sa = ws.mod0 + (size_t)l * 4 * D; m.ffn = m.sa + 2 * D;
      m.caq = ws.mod1 + (size_t)l * 4 * D; m.xf = m.caq + 2 * D;
      m.ld_sa = m.ld_ffn = n.w.ada0_n;
      m.ld_caq = m.ld_xf = n.w.ada1_n;
      break;
    default:  // influence
      m.sa = ws.mod0 + (size_t)l * 2 * D;
      m.caq = ws.mod1 + (size_t)l * 6 * D; m.xf = m.caq + 2 * D; m.ffn = m.caq + 4 * D;
      m.ld_sa = n.w.ada0_n;
      m.ld_caq = m.ld_xf = m.ld_ffn = n.w.ada1_n;
      break;
  }
  return m;
}

// conditioning of one network: E = SiLU(temb + text) -> all AdaLN (scale|shift) vectors in one GEMM per table
static int run_conditioning(mmdm_handle* h, int ni, int S, const float* temb, const int* step_ptr, cudaStream_t st) {
  Net& n = h->net[ni];
  Workspace& ws = h->ws;
  const int D = n.w.D;
  // E rows are written twice, [e | e] (pitch 2D): the composite weights of the LN fold are bf16 pairs [hi | lo]
  // along K, so one GEMM over K = 2D adds both halves in fp32
  {
    Scope sc(h, "cond_silu", st);
    MMDM_TRY(launch_cond_silu(temb, ws.tmap, step_ptr, ws.text[ni], ws.map0, S, D, ws.E0, st));
  }
  GemmEpi ep;
  ep.mode = EPI_F32;
  ep.bias = n.w.ada0_b; ep.out = ws.mod0; ep.ldo = n.w.ada0_n;
  MMDM_TRY(gemm(h, "gemm_adaln", ws.E0, 2 * D, n.w.ada0_w, D, S, n.w.ada0_n, D, ep, st));
  if (n.fold_on) {
    ep.bias = n.w.comp0_b; ep.out = ws.gb0; ep.ldo = n.w.comp0_n;
    MMDM_TRY(gemm(h, "gemm_adaln", ws.E0, 2 * D, n.w.comp0_w, 2 * D, S, n.w.comp0_n, 2 * D, ep, st));
  }
  if (n.w.ada1_n > 0) {
    {
      Scope sc(h, "cond_silu", st);
      MMDM_TRY(launch_cond_silu(temb, ws.tmap, step_ptr, ws.text[ni], ws.map1, S, D, ws.E1, st));
    }
    ep.bias = n.w.ada1_b; ep.out = ws.mod1; ep.ldo = n.w.ada1_n;
    MMDM_TRY(gemm(h, "gemm_adaln", ws.E1, 2 * D, n.w.ada1_w, D, S, n.w.ada1_n, D, ep, st));
    if (n.fold_on) {
      ep.bias = n.w.comp1_b; ep.out = ws.gb1; ep.ldo = n.w.comp1_n;
      MMDM_TRY(gemm(h, "gemm_adaln", ws.E1, 2 * D, n.w.comp1_w, 2 * D, S, n.w.comp1_n, 2 * D, ep, st));
    }
  }
  return 0;
}

static int attention(mmdm_handle* h, const Net& n, int S, int T, int kv_xor, cudaStream_t st) {
  Workspace& ws = h->ws;
  const int D = n.w.D;
  Scope sc(h, "attention", st, 4.0 * S * n.w.H * (double)T * (T + 1) * n.hd());
  const unsigned long long* km = h->mask_on ? (h->seq_persons == 2 ? ws.kmask2 : ws.kmask1) : nullptr;
  return launch_attention(ws.qk, 2 * D, ws.qk + D, 2 * D, ws.vt, h->Tp, ws.att, D, S, T, n.w.H, n.hd(), kv_xor, st,
                          next_dir(h), km);
}

static int ln(mmdm_handle* h, const float* src, int M, int D, int T, const float* m0, int ld0, bf16* o0,
              const float* m1, int ld1, bf16* o1, cudaStream_t st) {
  Scope sc(h, "ln_modulate", st);
  return launch_ln_modulate(src, M, D, T, m0, ld0, o0, m1, ld1, o1, st, next_dir(h));
}

// ---- LN fold (gemm.cuh): where the G_s | Bc_s rows of a block's AdaLN sites sit in gb0 / gb1 -----------------------
// Column layout of the composite tables per block l (mirrors the ada0 / ada1 row order, every site = [G (N) | Bc (N)]
// for the Linear of N outputs that consumes the normalised rows):
//   INDIVIDUAL : comp0 = [sa: in_proj 3D, ffn: linear1 F]
//   INTERACTION: comp0 = [sa: in_proj 3D, ffn: linear1 F]      comp1 = [ca.norm: q D, ca.xf_norm: kv 2D]
//   INFLUENCE  : comp0 = [sa: in_proj 3D]                      comp1 = [ca.norm: q D, ffn: linear1 F]
//                (its xf_norm normalises m_I, which no GEMM of the block writes: that site keeps the LN kernel)
struct FoldSite {
  const float *g, *b;
  int ld;
};
struct Folds {
  FoldSite sa, ffn, caq, xf;
};
static Folds block_folds(const Net& n, const Workspace& ws, int l) {
  const int D = n.w.D, F = n.w.F;
  Folds f = {};
  auto site = [](const float* base, int N, int ld) { return FoldSite{base, base + N, ld}; };
  if (n.w.kind == MMDM_NET_INFLUENCE) {
    f.sa = site(ws.gb0 + (size_t)l * 6 * D, 3 * D, n.w.comp0_n);
    const float* p1 = ws.gb1 + (size_t)l * 2 * (D + F);
    f.caq = site(p1, D, n.w.comp1_n);
    f.ffn = site(p1 + 2 * D, F, n.w.comp1_n);
  } else {
    const float* p0 = ws.gb0 + (size_t)l * 2 * (3 * D + F);
    f.sa = site(p0, 3 * D, n.w.comp0_n);
    f.ffn = site(p0 + 6 * D, F, n.w.comp0_n);
    if (n.w.kind == MMDM_NET_INTERACTION) {
      const float* p1 = ws.gb1 + (size_t)l * 6 * D;
      f.caq = site(p1, D, n.w.comp1_n);
      f.xf = site(p1 + 2 * D, 2 * D, n.w.comp1_n);
    }
  }
  return f;
}
static size_t comp_cols(const mmdm_net_weights& w, int table) {   // expected width of comp0 / comp1
  const size_t D = w.D, F = w.F, L = w.L;
  if (w.kind == MMDM_NET_INFLUENCE) return table == 0 ? L * 6 * D : L * 2 * (D + F);
  if (table == 0) return L * 2 * (3 * D + F);
  return w.kind == MMDM_NET_INTERACTION ? L * 6 * D : 0;
}

static void as_consumer(GemmEpi& e, int mode, const FoldSite& fs, const Workspace& ws, int si, int M, int D, int T) {
  e.mode = mode;
  e.bias = nullptr;   // part of Bc
  e.lnc_stats = ws.stat[si]; e.lnc_rowa = ws.rowa[si]; e.lnc_pitch = M; e.lnc_slots = D / 64;
  e.lnc_g = fs.g; e.lnc_b = fs.b; e.lnc_ld = fs.ld;
  e.T = T;
}
// si: the AdaLN site the GEMM produces for (statistics set); sp: the site before it on the same rows, whose row means
// serve as the operand offsets
static void as_producer(GemmEpi& e, int mode, const Workspace& ws, int si, int sp, int M, int T, const float* sc0, int ld0, bf16* o0,
                        const float* sc1 = nullptr, int ld1 = 0, bf16* o1 = nullptr) {
  e.mode = mode;
  e.lnp_moff_in = ws.moff[sp]; e.lnp_moff_out = ws.moff[si];
  e.lnp_stats = ws.stat[si]; e.lnp_rowa = ws.rowa[si]; e.lnp_pitch = M; e.lnp_T = T;
  e.lnp_scale0 = sc0; e.lnp_ld0 = ld0; e.lnp_out0 = o0;
  e.lnp_scale1 = sc1; e.lnp_ld1 = ld1; e.lnp_out1 = o1;
}

// one transformer block on the residual stream hres (M, D), in place.
//   kvsrc == nullptr : cross attention (if any) reads the partner stream's block input (s ^ 1)
//   kvsrc != nullptr : cross attention reads LN(kvsrc) of the same sequence index (Influence: m_I)
// With the LN fold (n.fold_on) the block expects bf16(h (1 + scale_sa)) in ws.xn (and bf16(h (1 + scale_xf)) in ws.xf for
// the partner-stream cross attention) plus the row sums in ws.stat[0] on entry - the previous block's FFN-2 (or the
// embedding GEMM) wrote them - and leaves the same for block l + 1.
static int run_block(mmdm_handle* h, int ni, int l, float* hres, const float* kvsrc, int S, int T, bf16* bf16_copy,
                     cudaStream_t st) {
  Net& n = h->net[ni];
  Workspace& ws = h->ws;
  const mmdm_block_weights& b = n.blocks[l];
  const int D = n.w.D, F = n.w.F, H = n.w.H, hd = n.hd();
  const int M = S * T;
  const Mods m = block_mods(n, ws, l);
  const bool has_ca = b.ca_q_w != nullptr;
  const bool fold = n.fold_on;
  const Folds f = fold ? block_folds(n, ws, l) : Folds{};

  // ---- self attention -------------------------------------------------------------------------
  if (!fold) {
    if (has_ca && kvsrc == nullptr)
      MMDM_TRY(ln(h, hres, M, D, T, m.sa, m.ld_sa, ws.xn, m.xf, m.ld_xf, ws.xf, st));
    else
      MMDM_TRY(ln(h, hres, M, D, T, m.sa, m.ld_sa, ws.xn, nullptr, 0, nullptr, st));
  }
  GemmEpi ep;
  ep.mode = EPI_BF16; ep.bias = b.sa_in_b; ep.out = ws.qk; ep.ldo = 2 * D;
  ep.vt_col0 = 2 * D; ep.vt = ws.vt; ep.T = T; ep.H = H; ep.hd = hd; ep.Tp = h->Tp;
  if (fold) as_consumer(ep, EPI_BF16_LN, f.sa, ws, 0, M, D, T);
  MMDM_TRY(gemm(h, "gemm_qkv", ws.xn, D, b.sa_in_w, D, M, 3 * D, D, ep, st));
  MMDM_TRY(attention(h, n, S, T, 0, st));
  GemmEpi er;
  er.mode = EPI_RESID; er.out = hres; er.resid = hres; er.ldo = D;
  er.bias = b.sa_out_b;
  if (fold) {   // next AdaLN site on these rows: the cross attention's query norm, or the FFN norm
    if (has_ca) as_producer(er, EPI_RESID_LN, ws, 1, 0, M, T, m.caq, m.ld_caq, ws.xn);
    else as_producer(er, EPI_RESID_LN, ws, 1, 0, M, T, m.ffn, m.ld_ffn, ws.xn);
  }
  MMDM_TRY(gemm(h, "gemm_proj", ws.att, D, b.sa_out_w, D, M, D, D, er, st));

  // ---- cross attention --------------------------------------------------------------------------
  if (has_ca) {
    if (kvsrc != nullptr) MMDM_TRY(ln(h, kvsrc, M, D, T, m.xf, m.ld_xf, ws.xf, nullptr, 0, nullptr, st));
    if (!fold) MMDM_TRY(ln(h, hres, M, D, T, m.caq, m.ld_caq, ws.xn, nullptr, 0, nullptr, st));
    GemmEpi eq;
    eq.mode = EPI_BF16; eq.bias = b.ca_q_b; eq.out = ws.qk; eq.ldo = 2 * D;
    if (fold) as_consumer(eq, EPI_BF16_LN, f.caq, ws, 1, M, D, T);
    MMDM_TRY(gemm(h, "gemm_qkv", ws.xn, D, b.ca_q_w, D, M, D, D, eq, st));
    GemmEpi ek;
    ek.mode = EPI_BF16; ek.bias = b.ca_kv_b; ek.out = ws.qk + D; ek.ldo = 2 * D;
    ek.vt_col0 = D; ek.vt = ws.vt; ek.T = T; ek.H = H; ek.hd = hd; ek.Tp = h->Tp;
    if (fold && kvsrc == nullptr) as_consumer(ek, EPI_BF16_LN, f.xf, ws, 0, M, D, T);   // statistics of the block input
    MMDM_TRY(gemm(h, "gemm_qkv", ws.xf, D, b.ca_kv_w, D, M, 2 * D, D, ek, st, /*flip=*/false));   // same way as the q GEMM
    MMDM_TRY(attention(h, n, S, T, kvsrc == nullptr ? 1 : 0, st));
    er.bias = b.ca_out_b;
    if (fold) as_producer(er, EPI_RESID_LN, ws, 2, 1, M, T, m.ffn, m.ld_ffn, ws.xn);
    MMDM_TRY(gemm(h, "gemm_proj", ws.att, D, b.ca_out_w, D, M, D, D, er, st));
  }

  // ---- feed forward -------------------------------------------------------------------------------
  if (!fold) MMDM_TRY(ln(h, hres, M, D, T, m.ffn, m.ld_ffn, ws.xn, nullptr, 0, nullptr, st));
  GemmEpi e1;
  e1.mode = EPI_GELU_BF16; e1.bias = b.ffn_b1; e1.out = ws.mid; e1.ldo = F;
  if (fold) as_consumer(e1, EPI_GELU_BF16_LN, f.ffn, ws, has_ca ? 2 : 1, M, D, T);
  MMDM_TRY(gemm(h, "gemm_ffn1", ws.xn, D, b.ffn_w1, D, M, F, D, e1, st));
  er.bias = b.ffn_b2;
  er.mode = EPI_RESID;
  er.out_bf16 = bf16_copy;
  if (fold && l + 1 < n.w.L) {   // operands of the next block
    const Mods mn = block_mods(n, ws, l + 1);
    const int sp = has_ca ? 2 : 1;
    if (has_ca && kvsrc == nullptr) as_producer(er, EPI_RESID_LN, ws, 0, sp, M, T, mn.sa, mn.ld_sa, ws.xn, mn.xf, mn.ld_xf, ws.xf);
    else as_producer(er, EPI_RESID_LN, ws, 0, sp, M, T, mn.sa, mn.ld_sa, ws.xn);
  }
  MMDM_TRY(gemm(h, "gemm_ffn2", ws.mid, F, b.ffn_w2, F, M, D, F, er, st));
  return 0;
}

// motion_embed + positional encoding -> residual stream; with the LN fold also the first block's operands (a row pass
// of the LayerNorm kernel's cost: the embedding GEMM is all epilogue already and runs three times per step).
// rows: rows of the GEMM; ln_rows: rows that enter the blocks (Influence: the h_i half, m_I follows);
// dup: the rows are written a second time `rows` further down for the unconditional CFG half.
static int run_embed(mmdm_handle* h, int ni, int rows, int ln_rows, int dup, int S, int T, cudaStream_t st) {
  Net& n = h->net[ni];
  Workspace& ws = h->ws;
  const int D = n.w.D, kpad = h->cfg.kpad;
  GemmEpi ee;
  ee.mode = EPI_EMBED_PE; ee.bias = n.w.embed_b; ee.out = ws.h; ee.ldo = D; ee.pe = n.w.pe; ee.pe_ld = D; ee.T = T;
  ee.dup_row_offset = dup ? (long long)rows : 0;
  MMDM_TRY(gemm(h, "gemm_embed", ws.xin, kpad, n.w.embed_w, kpad, rows, D, kpad, ee, st));
  if (n.fold_on) {
    const Mods m0 = block_mods(n, ws, 0);
    const int Mb = dup ? 2 * ln_rows : ln_rows;   // rows entering the blocks = S * T
    Scope sc(h, "ln_modulate", st);
    if (n.w.kind == MMDM_NET_INTERACTION)
      MMDM_TRY(launch_ln_fold_first(ws.h, Mb, D, T, m0.sa, m0.ld_sa, ws.xn, m0.xf, m0.ld_xf, ws.xf, ws.stat[0], S * T, ws.rowa[0], ws.moff[0], st, next_dir(h)));
    else
      MMDM_TRY(launch_ln_fold_first(ws.h, Mb, D, T, m0.sa, m0.ld_sa, ws.xn, nullptr, 0, nullptr, ws.stat[0], S * T, ws.rowa[0], ws.moff[0], st, next_dir(h)));
  }
  return 0;
}

// in2INDenoiser.forward body. x: (Bu, T, persons*nfeats) fp32 where Bu rows are embedded once and, when
// dup != 0, shared by the two CFG halves (Bx = 2*Bu). out: (Bx, T, persons*nfeats), de-normalised when
// std/mean are given.
static int run_denoiser(mmdm_handle* h, int ni, const float* x, int Bu, int dup, int persons, const float* temb,
                        const int* step_ptr, float* out, const float* stdv, const float* meanv, cudaStream_t st,
                        int x_pitch = 0) {
  Net& n = h->net[ni];
  Workspace& ws = h->ws;
  const int T = h->T, D = n.w.D, nf = h->cfg.nfeats, kpad = h->cfg.kpad;
  const int Bx = dup ? 2 * Bu : Bu;
  const int S = Bx * persons;
  const int Mu = Bu * persons * T, M = S * T;
  h->seq_persons = persons;
  MMDM_TRY(run_conditioning(h, ni, S, temb, step_ptr, st));
  {
    Scope sc(h, "pack_state", st);
    MMDM_TRY(launch_pack_state(x, Bu, T, nf, kpad, persons, ws.xin, st, x_pitch));
  }
  MMDM_TRY(run_embed(h, ni, Mu, Mu, dup, S, T, st));
  for (int l = 0; l < n.w.L; ++l)
    MMDM_TRY(run_block(h, ni, l, ws.h, nullptr, S, T, l == n.w.L - 1 ? ws.xn : nullptr, st));
  GemmEpi ef;
  ef.mode = EPI_FINAL; ef.bias = n.w.out_b; ef.out = out; ef.T = T; ef.nfeats = nf; ef.persons_per_row = persons;
  ef.std = stdv; ef.mean = meanv;
  MMDM_TRY(gemm(h, "gemm_final", ws.xn, D, n.w.out_w, D, M, nf, D, ef, st));
  return 0;
}

// text_embed of `nsl` slices starting at column col0 of cond (Bt, cond_ld) -> text[ni] rows k*Bt + b; the
// unconditional row (text_embed(0) = bias) lands at row nsl*Bt.
static int run_text(mmdm_handle* h, int ni, const float* cond, int cond_ld, int Bt, int col0, int nsl, cudaStream_t st) {
  Net& n = h->net[ni];
  const int D = n.w.D, K = h->cfg.text_dim;
  Scope sc(h, "text_embed", st);
  for (int k = 0; k < nsl; ++k)
    MMDM_TRY(launch_small_linear(cond + col0 + (size_t)k * K, cond_ld, n.w.text_w, n.w.text_b,
                                 h->ws.text[ni] + (size_t)k * Bt * D, D, Bt, D, K, 0, st));
  MMDM_TRY(launch_small_linear(h->ws.zero_row, K, n.w.text_w, n.w.text_b, h->ws.text[ni] + (size_t)nsl * Bt * D, D,
                               1, D, K, 0, st));
  return 0;
}

// time_embed(pe[idx]) for R indices -> dst (R, D)   (models/utils/utils.py:41-55)
static int run_time_embed(mmdm_handle* h, int ni, const long long* idx64, const int* idx32, int R, float* dst,
                          cudaStream_t st) {
  Net& n = h->net[ni];
  const int D = n.w.D;
  Scope sc(h, "time_embed", st);
  float* g = h->ws.tmp_rows;
  float* mid = g + (size_t)R * D;
  MMDM_TRY(launch_gather_rows(n.w.pe, n.w.pe_rows, idx64, idx32, R, D, g, st));
  MMDM_TRY(launch_small_linear(g, D, n.w.time_w0, n.w.time_b0, mid, D, R, D, D, 1, st));
  MMDM_TRY(launch_small_linear(mid, D, n.w.time_w2, n.w.time_b2, dst, D, R, D, D, 0, st));
  return 0;
}

// condition column layout of the (B, 6144) vector (models/mixermdm.py:672-682)
static const int COL_INTER = 0, COL_IND = 3 * 768, COL_MIX = 5 * 768;

// the Mixer body after conditioning tables are in place. Bu unique rows (+dup) as in run_denoiser.
static int run_mixer_nets(mmdm_handle* h, const float* x1, const float* x2, int Bu, int dup, const float* const temb[3],
                          const int* step_ptr, cudaStream_t st, int x_pitch = 0) {
  Workspace& ws = h->ws;
  const int T = h->T, nf = h->cfg.nfeats, kpad = h->cfg.kpad;
  const int Bx = dup ? 2 * Bu : Bu, S = 2 * Bx, M = S * T;
  const int Bt = Bu;
  MMDM_REQUIRE(!(h->mask_on && h->cfg.align),
               "frame masks need Mixer(align=False): the reference's masked trajectory alignment does not run "
               "(utils/alignment.py:86-89 indexes one frame per batch row with a (B, 2) length tensor)");
  // individual denoiser on both persons (text slices: cond1_1, cond1_2)
  MMDM_TRY(launch_build_maps(ws.map0, ws.map1, ws.tmap, S, 2, Bt, dup, 2, 0, st));
  MMDM_TRY(run_denoiser(h, 0, x1, Bu, dup, 2, temb[0], step_ptr, ws.out1, h->cfg.std1, h->cfg.mean1, st, x_pitch));
  // interaction denoiser (text slices: emb, emb_individual1, emb_individual2)
  MMDM_TRY(launch_build_maps(ws.map0, ws.map1, ws.tmap, S, 2, Bt, dup, 3, 1, st));
  MMDM_TRY(run_denoiser(h, 1, x2, Bu, dup, 2, temb[1], step_ptr, ws.out2, h->cfg.std2, h->cfg.mean2, st, x_pitch));
  // alignment + Mixer embedding
  {
    Scope sc(h, "align", st, 0, (double)Bx * T * (2.0 * nf) * 4 * 4 + 2.0 * M * kpad * 2);
    MMDM_TRY(launch_align(ws.out1, ws.out2, ws.out1a, ws.out2a, ws.xin, ws.xin + (size_t)M * kpad, Bx, T, kpad,
                          h->cfg.align, st, 2 * 2 * nf));
  }
  Net& n = h->net[2];
  const int D = n.w.D;
  h->seq_persons = 2;
  MMDM_TRY(launch_build_maps(ws.map0, ws.map1, ws.tmap, S, 2, Bt, dup, 3, 1, st));
  MMDM_TRY(run_conditioning(h, 2, S, temb[2], step_ptr, st));
  MMDM_TRY(run_embed(h, 2, 2 * M, M, 0, S, T, st));
  const float* mI = ws.h + (size_t)M * D;
  for (int l = 0; l < n.w.L; ++l)
    MMDM_TRY(run_block(h, 2, l, ws.h, mI, S, T, l == n.w.L - 1 ? ws.hfin : nullptr, st));
  return 0;
}

// Influence head: weights = sigmoid(Linear(512 -> 1|23)(h)) (influence.py:116-125); pooled over time first in modes 1/3
static int run_head(mmdm_handle* h, int S, int T, cudaStream_t st) {
  Workspace& ws = h->ws;
  const int D = h->net[2].w.D;
  const int nw = (h->cfg.mixing_mode <= 2) ? 1 : 23;
  const bool pooled = (h->cfg.mixing_mode == 1 || h->cfg.mixing_mode == 3);
  const bf16* src = ws.hfin;
  int rows = S * T;
  if (pooled) {
    Scope sc(h, "pool_mean", st);
    MMDM_TRY(launch_pool_mean(ws.hfin, S, T, D, ws.hpool, st));
    src = ws.hpool;
    rows = S;
  }
  GemmEpi ep;
  ep.mode = EPI_SIGMOID_F32; ep.bias = h->cfg.head_b; ep.out = ws.wmix; ep.ldo = 24;
  return gemm(h, "gemm_head", src, D, h->cfg.head_w, D, rows, nw, D, ep, st);
}

static MixStepParams mix_params(mmdm_handle* h) {
  MixStepParams P = {};
  P.B = h->B; P.T = h->T; P.mode = h->cfg.mixing_mode; P.ncfg = 2; P.mix_only = 0; P.mix_out = nullptr;
  P.center = h->cfg.center; P.has_force = h->cfg.has_force; P.force_val = h->cfg.force_val;
  P.cfg_scale = h->cfg.cfg_scale; P.step_ptr = h->ws.step; P.nsteps = h->cfg.nsteps; P.coef = h->cfg.coef;
  P.weights = h->ws.wmix;
  P.out1a = h->ws.out1a; P.out2a = h->ws.out2a; P.x1 = h->ws.x1; P.x2 = h->ws.x2; P.final_out = nullptr;
  P.pred_pitch = P.state_pitch = 2 * 2 * h->cfg.nfeats;
  P.mean1 = h->cfg.mean1; P.std1 = h->cfg.std1; P.mean2 = h->cfg.mean2; P.std2 = h->cfg.std2;
  P.hist_i1 = nullptr; P.hist_i2 = nullptr;
  return P;
}

static double mix_bytes(const mmdm_handle* h) {
  // SURVEY 8(d), variant with materialised weights: per sample-step
  //   T*(2*2*524*4 [aligned preds, cond+uncond] + 2*524*4 [x, x2 read] + 2*524*4 [x', x2' write]) + weights 2*2*nw*4*(T or 1)
  const int nw = (h->cfg.mixing_mode <= 2) ? 1 : 23;
  const bool pooled = (h->cfg.mixing_mode == 1 || h->cfg.mixing_mode == 3);
  return (double)h->B * (h->T * (2.0 * 2 * 524 * 4 + 2.0 * 524 * 4 + 2.0 * 524 * 4) + 2.0 * 2 * nw * 4 * (pooled ? 1 : h->T));
}

struct EvalHist {   // optional per-step outputs of Mixer mode "eval" (mixermdm.py:805-809), each (nsteps, 2B, T, 524)
  float* out1 = nullptr;
  float* out2 = nullptr;
  float* mix = nullptr;
  int slot = 0;
};

static int step_body(mmdm_handle* h, float* final_out, float* h1, float* h2, cudaStream_t st,
                     const EvalHist* eh = nullptr) {
  const float* temb[3] = {h->ws.temb_tab[0], h->ws.temb_tab[1], h->ws.temb_tab[2]};
  h->rev = 0;
  MMDM_TRY(run_mixer_nets(h, h->ws.x1, h->ws.x2, h->B, 1, temb, h->ws.step, st, 2 * 2 * h->cfg.nfeats));
  MMDM_TRY(run_head(h, 4 * h->B, h->T, st));
  MixStepParams P = mix_params(h);
  P.final_out = final_out ? final_out : h->ws.final;
  P.hist_i1 = h1; P.hist_i2 = h2;
  if (eh != nullptr && eh->out1 != nullptr) {
    const size_t n = (size_t)2 * h->B * h->T * 2 * h->cfg.nfeats;
    const size_t rowb = (size_t)2 * h->cfg.nfeats * sizeof(float), nrows = (size_t)2 * h->B * h->T;
    MMDM_CHECK_CUDA(cudaMemcpy2DAsync(eh->out1 + n * eh->slot, rowb, h->ws.out1a, 2 * rowb, rowb, nrows, cudaMemcpyDeviceToDevice, st));
    MMDM_CHECK_CUDA(cudaMemcpy2DAsync(eh->out2 + n * eh->slot, rowb, h->ws.out2a, 2 * rowb, rowb, nrows, cudaMemcpyDeviceToDevice, st));
    P.mix_out = eh->mix + n * eh->slot;
  }
  {
    Scope sc(h, "mix_cfg_ddim", st, 0, mix_bytes(h));
    MMDM_TRY(launch_mixstep(P, st));
  }
  {
    Scope sc(h, "step_counter", st);
    MMDM_TRY(launch_add_int(h->ws.step, -1, st));
  }
  return 0;
}

static int ensure_temb(mmdm_handle* h, cudaStream_t st) {
  if (h->temb_ready) return 0;
  MMDM_CHECK_CUDA(cudaMemcpyAsync(h->ws.ts_idx, h->timestep_map.data(), sizeof(int) * h->cfg.nsteps,
                                  cudaMemcpyHostToDevice, st));
  MMDM_CHECK_CUDA(cudaMemsetAsync(h->ws.zero_row, 0, sizeof(float) * h->cfg.text_dim, st));
  for (int i = 0; i < 3; ++i)
    if (h->net[i].present) MMDM_TRY(run_time_embed(h, i, nullptr, h->ws.ts_idx, h->cfg.nsteps, h->ws.temb_tab[i], st));
  h->temb_ready = true;
  return 0;
}

constexpr size_t kMaxCachedGraphs = 8;
static void drop_graphs(mmdm_handle* h) {
  for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
}

const char* get_last_error();

// Every handle entry point runs through here: the handle's device is made current for the call (the caller may
// sit on another GPU of the process) and a failure's message is kept per handle.
template <typename F>
static int guarded(mmdm_handle* h, F&& body) {
  if (!h) {
    set_last_error("null handle");
    return -1;
  }
  DeviceGuard dg(h->device);
  const int rc = body();
  if (rc != 0) h->last_error = get_last_error();
  return rc;
}

}  // namespace mmdm

// ------------------------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------------------------
extern "C" {

int mmdm_create(const mmdm_config* cfg, const mmdm_net_weights nets[3], mmdm_handle** out) {
  MMDM_REQUIRE(cfg && nets && out, "mmdm_create: null argument");
  MMDM_REQUIRE(cfg->mixing_mode >= 1 && cfg->mixing_mode <= 4, "Mixing mode not recognized");
  MMDM_REQUIRE(cfg->nfeats == 262 && cfg->kpad >= cfg->nfeats && cfg->kpad % 8 == 0, "mmdm_create: nfeats/kpad");
  MMDM_REQUIRE(cfg->nsteps >= 1 && cfg->timestep_map && cfg->coef, "mmdm_create: schedule tables missing");
  MMDM_TRY(gemm_init());
  MMDM_TRY(attention_init());
  MMDM_TRY(motion_init());
  mmdm_handle* h = new mmdm_handle();
  h->device = current_device();
  h->cfg = *cfg;
  {
    const char* e = getenv("MMDM_SERPENTINE");
    h->serpentine = !(e && e[0] == '0');
  }
  h->timestep_map.assign(cfg->timestep_map, cfg->timestep_map + cfg->nsteps);
  h->cfg.timestep_map = nullptr;
  for (int i = 0; i < 3; ++i) {
    const mmdm_net_weights& w = nets[i];
    h->net[i].w = w;
    h->net[i].w.blocks = nullptr;
    if (w.L == 0 || w.blocks == nullptr) {   // absent network (single-denoiser handles)
      h->net[i].w.D = h->net[i].w.F = h->net[i].w.ada0_n = h->net[i].w.ada1_n = 0;
      h->net[i].w.H = 1;
      continue;
    }
    if (w.kind != i || w.D % 128 != 0 || w.D > 1024 || w.H <= 0 || (w.D / w.H != 64 && w.D / w.H != 128) || w.L < 1) {
      set_last_error("mmdm_create: net %d has unsupported geometry (kind=%d D=%d H=%d L=%d)", i, w.kind, w.D, w.H, w.L);
      delete h;
      return -1;
    }
    h->net[i].present = true;
    h->net[i].blocks.assign(w.blocks, w.blocks + w.L);
    // LN fold: needs the composite conditioning tables (see include/mixermdm_b200.h) in the expected layout
    if (w.comp0_w != nullptr) {
      const bool ok = (size_t)w.comp0_n == comp_cols(w, 0) && w.comp0_b != nullptr && w.D % 64 == 0 &&
                      (comp_cols(w, 1) == 0 || (w.comp1_w != nullptr && w.comp1_b != nullptr && (size_t)w.comp1_n == comp_cols(w, 1)));
      if (!ok) {
        set_last_error("mmdm_create: net %d: composite AdaLN tables do not match the network (comp0_n=%d, expected %zu; "
                       "comp1_n=%d, expected %zu)", i, w.comp0_n, comp_cols(w, 0), w.comp1_n, comp_cols(w, 1));
        delete h;
        return -1;
      }
      // MMDM_LN_FOLD: 0 keeps the stand-alone LayerNorm kernel even when the tables are given; 2 folds at every frame
      // count (tests); default: from 256 frames on, where the consumers' aux k step applies (gemm.cu) - below that the
      // folded GEMMs would run as single-CTA kernels and lose more than the LayerNorm pass costs
      const char* e = getenv("MMDM_LN_FOLD");
      h->net[i].fold = !(e && e[0] == '0');
      h->fold_any_T = e && e[0] == '2';
    }
  }
  if (h->net[2].present && h->net[2].w.D != 512) {
    set_last_error("mmdm_create: the fused mix kernel expects a 512-wide Mixer (got %d)", h->net[2].w.D);
    delete h;
    return -1;
  }
  *out = h;
  return 0;
}

void mmdm_destroy(mmdm_handle* h) {
  if (!h) return;
  DeviceGuard dg(h->device);
  drop_graphs(h);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  h->prof.flush();
  for (auto e : h->prof.pool) cudaEventDestroy(e);
  delete h;
}

size_t mmdm_workspace_bytes(const mmdm_handle* h, int B, int T) {
  if (!h || B < 1 || T < 1) return 0;
  size_t total = 0;
  carve(h, B, T, nullptr, nullptr, &total);
  return total;
}

int mmdm_bind_workspace(mmdm_handle* h, void* ws, size_t bytes, int B, int T) {
  return guarded(h, [&]() -> int {
    MMDM_REQUIRE(h && ws, "mmdm_bind_workspace: null argument");
    int tmax = 1 << 30;   // frames index the positional-encoding tables (reference: 5000 rows, models/utils/utils.py:25)
    for (int i = 0; i < 3; ++i)
      if (h->net[i].present) tmax = std::min(tmax, h->net[i].w.pe_rows);
    MMDM_REQUIRE(T >= 1 && T <= tmax, "mmdm_bind_workspace: T=%d outside [1,%d] (positional-encoding table)", T, tmax);
    MMDM_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "workspace must be 1024-byte aligned");
    size_t need = mmdm_workspace_bytes(h, B, T);
    MMDM_REQUIRE(bytes >= need, "workspace too small: %zu < %zu", bytes, need);
    if (ws != h->ws_base) drop_graphs(h);   // cached graphs hold pointers into the old workspace
    carve(h, B, T, ws, &h->ws, nullptr);
    h->ws_base = ws; h->ws_bytes = bytes; h->B = B; h->T = T; h->Bx = 2 * B; h->Tp = (T + 7) / 8 * 8;
    h->temb_ready = false;
    h->cond_ready = false;
    h->mask_on = false;
    for (int i = 0; i < 3; ++i) h->net[i].fold_on = h->net[i].fold && (T >= 256 || h->fold_any_T);
    return 0;
  });
}

int mmdm_set_condition(mmdm_handle* h, const float* cond, void* stream) {
  return guarded(h, [&]() -> int {
    MMDM_REQUIRE(h && h->ws_base && cond, "mmdm_set_condition: handle not bound / null cond");
    MMDM_REQUIRE(h->net[0].present && h->net[1].present && h->net[2].present, "mmdm_set_condition: handle holds a single network");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    MMDM_TRY(ensure_temb(h, st));
    const int ld = 8 * h->cfg.text_dim;
    MMDM_TRY(run_text(h, 0, cond, ld, h->B, COL_IND, 2, st));
    MMDM_TRY(run_text(h, 1, cond, ld, h->B, COL_INTER, 3, st));
    MMDM_TRY(run_text(h, 2, cond, ld, h->B, COL_MIX, 3, st));
    h->cond_ready = true;
    return 0;
  });
}

int mmdm_set_key_mask(mmdm_handle* h, const float* mask, int rows, void* stream) {
  return guarded(h, [&]() -> int {
    MMDM_REQUIRE(h && h->ws_base, "mmdm_set_key_mask: handle not bound");
    if (mask == nullptr) {
      h->mask_on = false;
      return 0;
    }
    MMDM_REQUIRE(rows >= 1 && rows <= h->Bx, "mmdm_set_key_mask: %d mask rows for a workspace of %d batch rows", rows, h->Bx);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    MMDM_TRY(build_key_mask(mask, rows, h->Bx, h->T, 1, h->ws.kmask1, st));
    MMDM_TRY(build_key_mask(mask, rows, 2 * h->Bx, h->T, 2, h->ws.kmask2, st));
    h->launches += 2;
    h->mask_on = true;
    return 0;
  });
}

int mmdm_set_state(mmdm_handle* h, const float* x1, const float* x2, int step_index, void* stream) {
  return guarded(h, [&]() -> int {
    MMDM_REQUIRE(h && h->ws_base && x1 && x2, "mmdm_set_state: handle not bound / null state");
    MMDM_REQUIRE(step_index >= 0 && step_index < h->cfg.nsteps, "mmdm_set_state: step index %d out of range", step_index);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)h->B * h->T * 2 * h->cfg.nfeats * sizeof(float);
    const size_t rowb = (size_t)2 * h->cfg.nfeats * sizeof(float), nrows = (size_t)h->B * h->T;   // (x1 | x2 interleaved rows)
    (void)bytes;
    MMDM_CHECK_CUDA(cudaMemcpy2DAsync(h->ws.x1, 2 * rowb, x1, rowb, rowb, nrows, cudaMemcpyDeviceToDevice, st));
    MMDM_CHECK_CUDA(cudaMemcpy2DAsync(h->ws.x2, 2 * rowb, x2, rowb, rowb, nrows, cudaMemcpyDeviceToDevice, st));
    MMDM_TRY(launch_set_int(h->ws.step, step_index, st));
    h->step_host = step_index;
    return 0;
  });
}

int mmdm_get_state(mmdm_handle* h, float* x1, float* x2, void* stream) {
  return guarded(h, [&]() -> int {
    MMDM_REQUIRE(h && h->ws_base, "mmdm_get_state: handle not bound");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)h->B * h->T * 2 * h->cfg.nfeats * sizeof(float);
    const size_t rowb = (size_t)2 * h->cfg.nfeats * sizeof(float), nrows = (size_t)h->B * h->T;
    (void)bytes;
    if (x1) MMDM_CHECK_CUDA(cudaMemcpy2DAsync(x1, rowb, h->ws.x1, 2 * rowb, rowb, nrows, cudaMemcpyDeviceToDevice, st));
    if (x2) MMDM_CHECK_CUDA(cudaMemcpy2DAsync(x2, rowb, h->ws.x2, 2 * rowb, rowb, nrows, cudaMemcpyDeviceToDevice, st));
    return 0;
  });
}

int mmdm_add_state_noise(mmdm_handle* h, const float* noise1, const float* noise2, float scale, void* stream) {
  return guarded(h, [&]() -> int {
    MMDM_REQUIRE(h && h->ws_base && noise1 && noise2, "mmdm_add_state_noise: handle not bound / null noise");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Scope sc(h, "state_noise", st);
    MMDM_TRY(launch_add_state_noise(h->ws.x1, h->ws.x2, noise1, noise2, (long long)h->B * h->T, 2 * h->cfg.nfeats, scale, st));
    return 0;
  });
}

int mmdm_step(mmdm_handle* h, float* final_out, void* stream) {
  return guarded(h, [&]() -> int {
    MMDM_REQUIRE(h && h->ws_base, "mmdm_step: handle not bound");
    MMDM_REQUIRE(h->net[0].present && h->net[1].present && h->net[2].present, "mmdm_step: handle holds a single network");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // the DDIM index lives on the device (graph replays); the host mirror only guards against stepping past index 0,
    // which would read the coefficient / embedding tables out of bounds
    MMDM_REQUIRE(h->step_host >= 0, "mmdm_step: no step left (set a state first; the DDIM index ran past 0)");
    MMDM_REQUIRE(h->cond_ready, "mmdm_step: no condition set (mmdm_set_condition; a mixer / denoiser forward or a re-bind "
                                "in between overwrites the conditioning tables)");
    MMDM_TRY(ensure_temb(h, st));
    MMDM_TRY(step_body(h, final_out, nullptr, nullptr, st));
    h->step_host -= 1;
    return 0;
  });
}

int mmdm_sample(mmdm_handle* h, const float* cond, const float* noise, float* out, float* hist_i1, float* hist_i2,
                float* hist_out1, float* hist_out2, float* hist_mix, int use_graph, void* stream) {
  return guarded(h, [&]() -> int {
    MMDM_REQUIRE(h && h->ws_base && cond && noise && out, "mmdm_sample: handle not bound / null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    MMDM_TRY(mmdm_set_condition(h, cond, stream));
    MMDM_TRY(mmdm_set_state(h, noise, noise, h->cfg.nsteps - 1, stream));
    const size_t out_bytes = (size_t)h->B * h->T * 2 * h->cfg.nfeats * sizeof(float);
    const bool eval_hist = hist_out1 || hist_out2 || hist_mix;
    if (eval_hist) MMDM_REQUIRE(hist_out1 && hist_out2 && hist_mix, "mmdm_sample: the three eval-mode history buffers come together");
    if (use_graph && !h->prof.on && !eval_hist) {
      mmdm_handle::GraphEntry* ge = nullptr;
      for (auto& g : h->graphs)
        if (g.ws_base == h->ws_base && g.B == h->B && g.T == h->T && g.h1 == hist_i1 && g.h2 == hist_i2 &&
            g.masked == h->mask_on)
          ge = &g;
      if (!ge) {
        if (h->graphs.size() >= kMaxCachedGraphs) {   // evict the least recently used entry
          size_t lru = 0;
          for (size_t i = 1; i < h->graphs.size(); ++i)
            if (h->graphs[i].stamp < h->graphs[lru].stamp) lru = i;
          cudaGraphExecDestroy(h->graphs[lru].exec);
          h->graphs.erase(h->graphs.begin() + lru);
        }
        cudaGraph_t g = nullptr;
        if (!h->cap_stream) MMDM_CHECK_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
        MMDM_CHECK_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
        const long long before = h->launches;
        int rc = step_body(h, nullptr, hist_i1, hist_i2, h->cap_stream);
        const long long step_launches = h->launches - before;
        h->launches = before;   // capturing enqueues nothing; replays are counted below
        cudaError_t e = cudaStreamEndCapture(h->cap_stream, &g);
        if (rc != 0) {
          if (g) cudaGraphDestroy(g);
          return rc;
        }
        MMDM_CHECK_CUDA(e);
        cudaGraphExec_t exec = nullptr;
        MMDM_CHECK_CUDA(cudaGraphInstantiate(&exec, g, 0));
        cudaGraphDestroy(g);
        h->graphs.push_back({h->ws_base, h->B, h->T, hist_i1, hist_i2, h->mask_on, exec, step_launches, 0});
        h->graph_captures++;
        ge = &h->graphs.back();
      }
      ge->stamp = ++h->graph_stamp;
      for (int i = 0; i < h->cfg.nsteps; ++i) MMDM_CHECK_CUDA(cudaGraphLaunch(ge->exec, st));
      h->launches += ge->step_launches * h->cfg.nsteps;
    } else {
      EvalHist eh;
      eh.out1 = hist_out1; eh.out2 = hist_out2; eh.mix = hist_mix;
      for (int i = 0; i < h->cfg.nsteps; ++i) {
        eh.slot = i;
        MMDM_TRY(step_body(h, nullptr, hist_i1, hist_i2, st, eval_hist ? &eh : nullptr));
      }
    }
    h->step_host = -1;   // the loop consumed every index
    MMDM_CHECK_CUDA(cudaMemcpyAsync(out, h->ws.final, out_bytes, cudaMemcpyDeviceToDevice, st));
    return 0;
  });
}

int mmdm_mixer_forward(mmdm_handle* h, const float* x1, const float* x2, const int64_t* timesteps, const float* cond,
                       float* out_influenced, float* out1, float* out2, float* infl_i1, float* infl_i2, int Bx,
                       int cfg_combine, float cfg_scale, void* stream) {
  return guarded(h, [&]() -> int {
    MMDM_REQUIRE(h && h->ws_base && x1 && x2 && timesteps && cond && out_influenced, "mmdm_mixer_forward: null argument");
    MMDM_REQUIRE(Bx >= 1 && Bx <= h->Bx, "mmdm_mixer_forward: Bx=%d exceeds the bound workspace (%d rows)", Bx, h->Bx);
    MMDM_REQUIRE(h->net[0].present && h->net[1].present && h->net[2].present, "mmdm_mixer_forward: handle holds a single network");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    MMDM_TRY(ensure_temb(h, st));
    Workspace& ws = h->ws;
    const int ld = 8 * h->cfg.text_dim;
    h->cond_ready = false;
    MMDM_TRY(run_text(h, 0, cond, ld, Bx, COL_IND, 2, st));
    MMDM_TRY(run_text(h, 1, cond, ld, Bx, COL_INTER, 3, st));
    MMDM_TRY(run_text(h, 2, cond, ld, Bx, COL_MIX, 3, st));
    // per-sample timestep embeddings are (re)computed per network right before that network runs
    MMDM_TRY(run_time_embed(h, 0, reinterpret_cast<const long long*>(timesteps), nullptr, Bx, ws.temb_dyn, st));
    const int T = h->T, nf = h->cfg.nfeats, kpad = h->cfg.kpad;
    const int S = 2 * Bx, M = S * T;
    MMDM_TRY(launch_build_maps(ws.map0, ws.map1, ws.tmap, S, 2, Bx, 0, 2, 0, st));
    MMDM_TRY(run_denoiser(h, 0, x1, Bx, 0, 2, ws.temb_dyn, nullptr, ws.out1, h->cfg.std1, h->cfg.mean1, st));
    MMDM_TRY(run_time_embed(h, 1, reinterpret_cast<const long long*>(timesteps), nullptr, Bx, ws.temb_dyn, st));
    MMDM_TRY(launch_build_maps(ws.map0, ws.map1, ws.tmap, S, 2, Bx, 0, 3, 1, st));
    MMDM_TRY(run_denoiser(h, 1, x2, Bx, 0, 2, ws.temb_dyn, nullptr, ws.out2, h->cfg.std2, h->cfg.mean2, st));
    MMDM_TRY(launch_align(ws.out1, ws.out2, ws.out1a, ws.out2a, ws.xin, ws.xin + (size_t)M * kpad, Bx, T, kpad,
                          h->cfg.align, st, 2 * 2 * nf));
    Net& n = h->net[2];
    const int D = n.w.D;
    MMDM_TRY(run_time_embed(h, 2, reinterpret_cast<const long long*>(timesteps), nullptr, Bx, ws.temb_dyn, st));
    MMDM_TRY(launch_build_maps(ws.map0, ws.map1, ws.tmap, S, 2, Bx, 0, 3, 1, st));
    MMDM_TRY(run_conditioning(h, 2, S, ws.temb_dyn, nullptr, st));
    MMDM_TRY(run_embed(h, 2, 2 * M, M, 0, S, T, st));
    const float* mI = ws.h + (size_t)M * D;
    for (int l = 0; l < n.w.L; ++l)
      MMDM_TRY(run_block(h, 2, l, ws.h, mI, S, T, l == n.w.L - 1 ? ws.hfin : nullptr, st));
    MMDM_TRY(run_head(h, S, T, st));
    MixStepParams P = mix_params(h);
    if (cfg_combine) {
      MMDM_REQUIRE(Bx % 2 == 0, "mmdm_mixer_forward: cfg_combine needs an even number of rows");
      P.B = Bx / 2; P.ncfg = 2; P.mix_only = 2; P.cfg_scale = cfg_scale;
    } else {
      P.B = Bx; P.ncfg = 1; P.mix_only = 1;
    }
    P.mix_out = out_influenced;
    P.hist_i1 = infl_i1; P.hist_i2 = infl_i2;
    MMDM_TRY(launch_mixstep(P, st));
    const size_t bytes = (size_t)Bx * T * 2 * nf * sizeof(float);
    const size_t rowb = (size_t)2 * nf * sizeof(float), nrows = (size_t)Bx * T;
    (void)bytes;
    if (out1) MMDM_CHECK_CUDA(cudaMemcpy2DAsync(out1, rowb, ws.out1a, 2 * rowb, rowb, nrows, cudaMemcpyDeviceToDevice, st));
    if (out2) MMDM_CHECK_CUDA(cudaMemcpy2DAsync(out2, rowb, ws.out2a, 2 * rowb, rowb, nrows, cudaMemcpyDeviceToDevice, st));
    return 0;
  });
}

int mmdm_denoiser_forward(mmdm_handle* h, int which, const float* x, const int64_t* timesteps, const float* cond,
                          float* out, int Bx, void* stream) {
  return guarded(h, [&]() -> int {
    MMDM_REQUIRE(h && h->ws_base && x && timesteps && cond && out, "mmdm_denoiser_forward: null argument");
    MMDM_REQUIRE(which == MMDM_NET_INDIVIDUAL || which == MMDM_NET_INTERACTION, "Mode not recognized");
    MMDM_REQUIRE(h->net[which].present, "mmdm_denoiser_forward: network %d is not loaded in this handle", which);
    const int persons = which == MMDM_NET_INTERACTION ? 2 : 1;
    MMDM_REQUIRE(Bx >= 1 && Bx <= h->Bx, "mmdm_denoiser_forward: Bx=%d exceeds the bound workspace (%d rows)", Bx, h->Bx);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    MMDM_TRY(ensure_temb(h, st));
    Workspace& ws = h->ws;
    const int nsl = which == MMDM_NET_INTERACTION ? 3 : 1;
    h->cond_ready = false;
    MMDM_TRY(run_text(h, which, cond, nsl * h->cfg.text_dim, Bx, 0, nsl, st));
    MMDM_TRY(run_time_embed(h, which, reinterpret_cast<const long long*>(timesteps), nullptr, Bx, ws.temb_dyn, st));
    MMDM_TRY(launch_build_maps(ws.map0, ws.map1, ws.tmap, Bx * persons, persons, Bx, 0, nsl, which == MMDM_NET_INTERACTION ? 1 : 0, st));
    return run_denoiser(h, which, x, Bx, 0, persons, ws.temb_dyn, nullptr, out, nullptr, nullptr, st);
  });
}

int mmdm_align_motions(const float* out1, const float* out2, float* out1a, float* out2a, void* emb_i, void* emb_I,
                       int Be, int T, int kpad, int do_align, void* stream) {
  MMDM_REQUIRE(out1 && out2 && out1a && out2a && emb_i && emb_I, "mmdm_align_motions: null argument");
  return launch_align(out1, out2, out1a, out2a, reinterpret_cast<bf16*>(emb_i), reinterpret_cast<bf16*>(emb_I), Be, T,
                      kpad, do_align, reinterpret_cast<cudaStream_t>(stream));
}

int mmdm_mix_cfg_ddim(const mmdm_mixstep_args* a, void* stream) {
  MMDM_REQUIRE(a && a->weights && a->out1a && a->out2a && a->step_ptr, "mmdm_mix_cfg_ddim: null argument");
  MixStepParams P = {};
  P.B = a->B; P.T = a->T; P.mode = a->mixing_mode; P.ncfg = a->ncfg; P.mix_only = a->mix_only; P.mix_out = a->mix_out;
  P.center = a->center; P.has_force = a->has_force; P.force_val = a->force_val; P.cfg_scale = a->cfg_scale;
  P.step_ptr = a->step_ptr; P.nsteps = a->nsteps; P.coef = a->coef;
  P.weights = a->weights;
  P.out1a = a->out1a; P.out2a = a->out2a; P.x1 = a->x1; P.x2 = a->x2; P.final_out = a->final_out;
  // dense (.,T,524) buffers, or - recognised by out2a == out1a + 524 (x2 == x1 + 524), which dense buffers cannot
  // satisfy - the two interleaved row by row ([out1a row | out2a row], pitch 1048): one bulk copy per pair of rows
  P.pred_pitch = (a->out2a == a->out1a + 524) ? 1048 : 524;
  P.state_pitch = (a->x1 != nullptr && a->x2 == a->x1 + 524) ? 1048 : 524;
  P.mean1 = a->mean1; P.std1 = a->std1; P.mean2 = a->mean2; P.std2 = a->std2;
  P.hist_i1 = a->hist_i1; P.hist_i2 = a->hist_i2;
  if (!P.mix_only)
    MMDM_REQUIRE(P.x1 && P.x2 && P.final_out && P.coef && P.mean1 && P.std1 && P.mean2 && P.std2 && P.ncfg == 2,
                 "mmdm_mix_cfg_ddim: sampler mode needs state, tables and both CFG halves");
  return launch_mixstep(P, reinterpret_cast<cudaStream_t>(stream));
}

int mmdm_smooth_motion(const float* x, float* out, int B, int T, int F, const float* weights, int radius, void* stream) {
  MMDM_REQUIRE(x && out && weights, "mmdm_smooth_motion: null argument");
  return launch_smooth_time(x, out, B, T, F, weights, radius, reinterpret_cast<cudaStream_t>(stream));
}

int mmdm_format_motion(const float* x, float* out, int B, int T, int max_length, int nfeats, const float* mean,
                       const float* std_, void* stream) {
  MMDM_REQUIRE(x && out, "mmdm_format_motion: null argument");
  return launch_format_motion(x, out, B, T, max_length, nfeats, mean, std_, reinterpret_cast<cudaStream_t>(stream));
}

int mmdm_debug_mix_stream(int S, int nfr, int has_phase2, int ring_slots, int item, int32_t* out) {
  MMDM_REQUIRE(out && S >= 1 && nfr >= 1 && ring_slots >= 1 && ring_slots <= nfr, "mmdm_debug_mix_stream: bad argument");
  MixStream st;
  st.S = S; st.nfr = nfr; st.has2 = has_phase2 ? 1 : 0;
  MMDM_REQUIRE(item >= 0 && item < mix_stream_items(st), "mmdm_debug_mix_stream: item %d outside the stream", item);
  bool p1;
  int n, fl, slot, lap;
  mix_item_decode(st, item, p1, n, fl);
  mix_ring_pos(item, ring_slots, mix_ring_rcp(ring_slots), slot, lap);
  out[0] = p1 ? 1 : 2; out[1] = n; out[2] = fl; out[3] = slot; out[4] = lap;
  out[5] = p1 ? mix_base_p1(st, n) + fl : mix_base_p2(st, n) + fl;   // must give `item` back
  bool q1 = p1;
  int qn = n, qf = fl;
  const bool ok = mix_item_after(st, ring_slots, q1, qn, qf);
  out[6] = ok ? 1 : 0; out[7] = q1 ? 1 : 2; out[8] = qn; out[9] = qf;
  return 0;
}

int mmdm_profile_enable(mmdm_handle* h, int on) {
  return guarded(h, [&]() -> int {
    MMDM_REQUIRE(h, "mmdm_profile_enable: null handle");
    h->prof.flush();
    h->prof.on = on != 0;
    if (on) h->prof.recs.clear();
    return 0;
  });
}

int mmdm_profile_read(mmdm_handle* h, mmdm_profile_entry* entries, int max_entries) {
  if (!h || !entries) return 0;
  DeviceGuard dg(h->device);
  h->prof.flush();
  int n = 0;
  for (auto& r : h->prof.recs) {
    if (n >= max_entries) break;
    memset(&entries[n], 0, sizeof(entries[n]));
    strncpy(entries[n].name, r.name.c_str(), sizeof(entries[n].name) - 1);
    entries[n].launches = r.launches;
    entries[n].ms = r.ms;
    entries[n].flops = r.flops;
    entries[n].bytes = r.bytes;
    ++n;
  }
  return n;
}

int64_t mmdm_launch_count(const mmdm_handle* h) { return h ? h->launches : 0; }
int64_t mmdm_graph_captures(const mmdm_handle* h) { return h ? h->graph_captures : 0; }

const char* mmdm_last_error(const mmdm_handle* h) {
  return (h && !h->last_error.empty()) ? h->last_error.c_str() : mmdm::get_last_error();
}

}  // extern "C"
