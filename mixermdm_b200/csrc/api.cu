// extern "C" surface declared in include/mixermdm_b200.h — operator-level entry points.
#include "../../include/mixermdm_b200.h"

#include "attention.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "rowops.cuh"

namespace mmdm {
const char* get_last_error();

GemmEpi to_internal(const mmdm_epilogue& e) {
  GemmEpi g;
  g.mode = e.mode;
  g.bias = e.bias;
  g.out = e.out;
  g.ldo = e.ldo;
  g.vt_col0 = e.vt_col0;
  g.vt = reinterpret_cast<__nv_bfloat16*>(e.vt);
  g.T = e.T > 0 ? e.T : 1;
  g.H = e.H > 0 ? e.H : 1;
  g.hd = e.hd > 0 ? e.hd : 1;
  g.Tp = e.Tp > 0 ? e.Tp : 1;
  g.resid = e.resid;
  g.out_bf16 = reinterpret_cast<__nv_bfloat16*>(e.out_bf16);
  g.pe = e.pe;
  g.pe_ld = e.pe_ld;
  g.dup_row_offset = e.dup_row_offset;
  g.std = e.std_;
  g.mean = e.mean;
  g.nfeats = e.nfeats;
  g.persons_per_row = e.persons_per_row > 0 ? e.persons_per_row : 2;
  g.lnp_scale0 = e.lnp_scale0; g.lnp_ld0 = e.lnp_ld0; g.lnp_out0 = reinterpret_cast<__nv_bfloat16*>(e.lnp_out0);
  g.lnp_scale1 = e.lnp_scale1; g.lnp_ld1 = e.lnp_ld1; g.lnp_out1 = reinterpret_cast<__nv_bfloat16*>(e.lnp_out1);
  g.lnp_stats = reinterpret_cast<float2*>(e.lnp_stats);
  g.lnp_pitch = e.lnp_pitch; g.lnp_T = e.lnp_T > 0 ? e.lnp_T : 1; g.lnp_rows = e.lnp_rows; g.lnp_dup_seq = e.lnp_dup_seq;
  g.lnc_stats = reinterpret_cast<const float2*>(e.lnc_stats);
  g.lnc_pitch = e.lnc_pitch; g.lnc_slots = 0;   // K / 64, set by launch_gemm's caller below
  g.lnc_g = e.lnc_g; g.lnc_b = e.lnc_b; g.lnc_ld = e.lnc_ld;
  g.lnp_rowa = reinterpret_cast<uint4*>(e.lnp_rowa);
  g.lnc_rowa = reinterpret_cast<const uint4*>(e.lnc_rowa);
  g.lnp_moff_in = e.lnp_moff_in; g.lnp_moff_out = e.lnp_moff_out;
  return g;
}
}  // namespace mmdm

using namespace mmdm;

extern "C" {

int mmdm_abi_version(void) { return MMDM_ABI_VERSION; }
void mmdm_debug_gemm_simt(int on) { mmdm::gemm_set_debug_simt(on); }

int mmdm_linear_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                     const mmdm_epilogue* ep, void* stream) {
  MMDM_REQUIRE(A && W && ep && ep->out, "mmdm_linear_bf16: null argument");
  GemmEpi g = to_internal(*ep);
  g.lnc_slots = K / kLnSlotCols;
  return launch_gemm(reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(W), ldw,
                     M, N, K, g, reinterpret_cast<cudaStream_t>(stream));
}

int mmdm_ln_modulate(const float* h, int M, int D, int T, const float* mod0, int ld_mod0, void* out0,
                     const float* mod1, int ld_mod1, void* out1, void* stream) {
  MMDM_REQUIRE(h && mod0 && out0, "mmdm_ln_modulate: null argument");
  return launch_ln_modulate(h, M, D, T, mod0, ld_mod0, reinterpret_cast<__nv_bfloat16*>(out0), mod1, ld_mod1,
                            reinterpret_cast<__nv_bfloat16*>(out1), reinterpret_cast<cudaStream_t>(stream));
}

int mmdm_attention(const void* q, int ldq, const void* k, int ldk, const void* vt, int Tp, void* out, int ldo,
                   int S, int T, int H, int hd, int kv_xor, void* stream) {
  MMDM_REQUIRE(q && k && vt && out, "mmdm_attention: null argument");
  return launch_attention(reinterpret_cast<const __nv_bfloat16*>(q), ldq, reinterpret_cast<const __nv_bfloat16*>(k),
                          ldk, reinterpret_cast<const __nv_bfloat16*>(vt), Tp, reinterpret_cast<__nv_bfloat16*>(out),
                          ldo, S, T, H, hd, kv_xor, reinterpret_cast<cudaStream_t>(stream));
}

int mmdm_attention_masked(const void* q, int ldq, const void* k, int ldk, const void* vt, int Tp, void* out, int ldo,
                          int S, int T, int H, int hd, int kv_xor, const uint64_t* key_mask_words, void* stream) {
  MMDM_REQUIRE(q && k && vt && out, "mmdm_attention_masked: null argument");
  return launch_attention(reinterpret_cast<const __nv_bfloat16*>(q), ldq, reinterpret_cast<const __nv_bfloat16*>(k),
                          ldk, reinterpret_cast<const __nv_bfloat16*>(vt), Tp, reinterpret_cast<__nv_bfloat16*>(out),
                          ldo, S, T, H, hd, kv_xor, reinterpret_cast<cudaStream_t>(stream), 0,
                          reinterpret_cast<const unsigned long long*>(key_mask_words));
}

int mmdm_build_key_mask(const float* mask, int rows, int nseq, int T, int persons, uint64_t* words, void* stream) {
  return build_key_mask(mask, rows, nseq, T, persons, reinterpret_cast<unsigned long long*>(words),
                        reinterpret_cast<cudaStream_t>(stream));
}

int mmdm_small_linear(const float* a, int lda, const float* w, const float* b, float* out, int ldo, int R,
                      int N, int K, int act, void* stream) {
  MMDM_REQUIRE(a && w && out, "mmdm_small_linear: null argument");
  return launch_small_linear(a, lda, w, b, out, ldo, R, N, K, act, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
