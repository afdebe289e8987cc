// Index arithmetic of the fused mix kernel's frame stream (host + device, so that the CPU tests can walk it):
// which frame of which phase of which sample is item c of a CTA's stream, where it sits in the shared-memory ring,
// and which item takes over a ring slot when its frame has been consumed.
//
// A CTA streams S samples of nfr frames each. Phase 1 (P1) of a sample reads the aligned predictions of a frame,
// phase 2 (P2, absent at the last DDIM step and in Mixer.forward) reads its state rows and guided prediction.
// The stream is software-pipelined across samples:   P1(0), [P1(k), P2(k-1)] for k = 1 .. S-1, P2(S-1)
// (without a phase 2 just P1(0), ..., P1(S-1)). Item c lives in ring slot c % NS during its "lap" c / NS.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MIX_HD __host__ __device__ __forceinline__
#else
#define MIX_HD inline
#endif

namespace mmdm {

struct MixStream {
  int S;      // samples this CTA's cluster works on
  int nfr;    // frames per sample owned by the CTA
  int has2;   // phase 2 present
};

MIX_HD int mix_stream_items(const MixStream& s) { return s.S * s.nfr * (s.has2 ? 2 : 1); }

// first item of P1(k) / of P2(n)
MIX_HD int mix_base_p1(const MixStream& s, int k) {
  return k * s.nfr + (s.has2 && k > 0 ? (k - 1) * s.nfr : 0);
}
MIX_HD int mix_base_p2(const MixStream& s, int n) {   // P2(n) follows P1(n+1) (or, for the last sample, P2(n-1))
  const int p1_before = n + 2 < s.S ? n + 2 : s.S;
  return p1_before * s.nfr + n * s.nfr;
}

// item c -> (phase, sample, frame)
MIX_HD void mix_item_decode(const MixStream& s, int c, bool& p1, int& n, int& fl) {
  const int r1 = s.nfr, r2 = s.has2 ? s.nfr : 0;
  if (c < r1) { p1 = true; n = 0; fl = c; return; }
  const int cp = c - r1, k = 1 + cp / (r1 + r2), rem = cp - (k - 1) * (r1 + r2);
  if (k < s.S && rem < r1) { p1 = true; n = k; fl = rem; }
  else { p1 = false; n = k - 1; fl = k < s.S ? rem - r1 : rem; }
}

// The item NS positions after frame f of (phase p1, sample n), NS <= nfr: it lies in the same segment or in the next
// one. Updates (p1, n, f); returns false past the end of the stream.
MIX_HD bool mix_item_after(const MixStream& s, int NS, bool& p1, int& n, int& f) {
  int t = f + NS;
  if (t >= s.nfr) {
    t -= s.nfr;
    if (!s.has2) { n = n + 1; }                        // P1(n) -> P1(n+1)
    else if (p1) {
      if (n >= 1) { p1 = false; n = n - 1; }           // P1(k) -> P2(k-1)
      else if (s.S > 1) { n = 1; }                     // P1(0) -> P1(1)
      else { p1 = false; n = 0; }                      // P1(0) -> P2(0)   (one sample)
    } else {
      if (n + 2 < s.S) { p1 = true; n = n + 2; }       // P2(j) -> P1(j+2)
      else { n = n + 1; }                              // P2(j) -> P2(j+1)
    }
    if (n >= s.S) return false;
  }
  f = t;
  return true;
}

// ring position of item c: slot c % NS, lap c / NS, with the quotient from a reciprocal multiply
// (ns_rcp = ceil(2^32 / NS), exact for the < 2^20 items a CTA streams; NS == 1 would need 2^32 and is special-cased)
MIX_HD uint32_t mix_ring_rcp(int NS) { return (uint32_t)((0x100000000ULL + (uint64_t)NS - 1) / (uint64_t)NS); }
MIX_HD void mix_ring_pos(int c, int NS, uint32_t ns_rcp, int& slot, int& lap) {
#if defined(__CUDA_ARCH__)
  lap = NS == 1 ? c : (int)__umulhi((uint32_t)c, ns_rcp);
#else
  lap = NS == 1 ? c : (int)(((uint64_t)(uint32_t)c * (uint64_t)ns_rcp) >> 32);
#endif
  slot = c - lap * NS;
}

}  // namespace mmdm
