#pragma once
#include "common.cuh"

namespace mmdm {

// q: rows (s*T+t) with pitch ldq, head h at columns [h*hd,(h+1)*hd); k likewise with pitch ldk;
// vt: [S*H][hd][Tp] (V transposed per sequence/head); out: rows (s*T+t), pitch ldo.
// Keys/values of query sequence s come from sequence (s ^ kv_xor).
// key_mask (optional): [S][ceil((T+1)/64)] words, bit j of word kb set = key kb*64+j of that sequence is attendable;
// the caller sets bit T (the zero-attention key) and clears the bits beyond it (build_key_mask does).
int launch_attention(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* vt,
                     int Tp, __nv_bfloat16* out, int ldo, int S, int T, int H, int hd, int kv_xor,
                     cudaStream_t stream, int rev = 0, const unsigned long long* key_mask = nullptr);

// mask: (rows, T) fp32, > 0.5 = frame is real (reference: key_padding_mask = ~(mask > 0.5), in2in.py:433-435); sequence s
// of nseq reads row (s / persons) % rows. words: [nseq][ceil((T+1)/64)]
int build_key_mask(const float* mask, int rows, int nseq, int T, int persons, unsigned long long* words,
                   cudaStream_t stream);

int attention_init();

}  // namespace mmdm
