#pragma once
#include "common.cuh"

namespace mmdm {

// q: rows (s*T+t) with pitch ldq, head h at columns [h*hd,(h+1)*hd); k likewise with pitch ldk;
// vt: [S*H][hd][Tp] (V transposed per sequence/head); out: rows (s*T+t), pitch ldo.
// Keys/values of query sequence s come from sequence (s ^ kv_xor).
int launch_attention(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* vt,
                     int Tp, __nv_bfloat16* out, int ldo, int S, int T, int H, int hd, int kv_xor,
                     cudaStream_t stream, int rev = 0);

int attention_init();

}  // namespace mmdm
