// Motion-space kernels of the Mixer step (HBM-bound, fp32):
//   align_kernel   : ih_to_smpl -> align_motions -> smpl_to_ih on the four de-normalised predictions and
//                    bf16 packing of the Mixer's motion_embed operand (models/mixermdm.py:702-726).
//   mixstep_kernel : expansion of the Influence head's weights (models/utils/influence.py:120-125, sigmoid(head) comes
//                    from the head GEMM's epilogue) + mix (models/mixermdm.py:739-801), classifier-free guidance
//                    (models/utils/cfg_sampler.py:52-55), process_xstart (centre + both normalisers,
//                    models/utils/gaussian_diffusion.py:2031-2062) and the DDIM update (:1936-1965) in one launch.
//                    Persistent 4-CTA clusters, one sample at a time; all HBM reads stream through a shared-memory
//                    ring filled by cp.async.bulk; the per-sample reductions centring needs (min-y over all frames,
//                    frame-0 root and hips) are exchanged through distributed shared memory; the guided prediction
//                    is handed from the mix phase to the update phase through L2, never through HBM.
#include "motion.cuh"

#include <cooperative_groups.h>
#include <stdlib.h>

#include "mix_stream.cuh"
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
             int Be, int T, int kpad, int do_align, int pitch_a) {
  __shared__ __align__(16) float buf[ALIGN_WARPS][4][NFEAT + 2];  // [model*2 + person][feature]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * ALIGN_WARPS + warp;
  if (item >= (long long)Be * T) return;
  const int be = (int)(item / T), t = (int)(item % T);
  const float* r1 = out1 + ((size_t)be * T + t) * (2 * NFEAT);
  const float* r2 = out2 + ((size_t)be * T + t) * (2 * NFEAT);
  float(*b)[NFEAT + 2] = buf[warp];
  // rows are 8-byte aligned and a person's block is an even number of floats: 262 float2 per row, 131 per person
  for (int j = lane; j < NFEAT; j += 32) {
    const int p = j >= NFEAT / 2, fi = 2 * j - p * NFEAT;
    const float2 v1 = __ldg(reinterpret_cast<const float2*>(r1) + j), v2 = __ldg(reinterpret_cast<const float2*>(r2) + j);
    *reinterpret_cast<float2*>(&b[p][fi]) = v1;
    *reinterpret_cast<float2*>(&b[2 + p][fi]) = v2;
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
  float* w1 = out1a + ((size_t)be * T + t) * pitch_a;
  float* w2 = out2a + ((size_t)be * T + t) * pitch_a;
  for (int j = lane; j < NFEAT; j += 32) {
    const int p = j >= NFEAT / 2, fi = 2 * j - p * NFEAT;
    reinterpret_cast<float2*>(w1)[j] = *reinterpret_cast<const float2*>(&b[p][fi]);
    reinterpret_cast<float2*>(w2)[j] = *reinterpret_cast<const float2*>(&b[2 + p][fi]);
  }
  // bf16 operand rows (sequence index be*2+p), eight values = one 16-byte store per lane and round
  const int chunks = kpad >> 3;
  for (int idx = lane; idx < 4 * chunks; idx += 32) {
    const int m = idx / chunks, k0 = (idx - m * chunks) << 3;   // m = model * 2 + person
    const int p = m & 1;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = k0 + i < NFEAT ? b[m][k0 + i] : 0.f;
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    __nv_bfloat16* dst = (m < 2 ? emb_i : emb_I) + ((size_t)(be * 2 + p) * T + t) * kpad + k0;
    *reinterpret_cast<uint4*>(dst) = o;
  }
}

int launch_align(const float* out1, const float* out2, float* out1a, float* out2a, __nv_bfloat16* emb_i,
                 __nv_bfloat16* emb_I, int Be, int T, int kpad, int do_align, cudaStream_t stream, int pitch_a) {
  MMDM_REQUIRE((kpad % 8) == 0 && (pitch_a % 2) == 0 &&
                   ((reinterpret_cast<uintptr_t>(out1) | reinterpret_cast<uintptr_t>(out2) | reinterpret_cast<uintptr_t>(out1a) |
                     reinterpret_cast<uintptr_t>(out2a)) & 7) == 0 &&
                   ((reinterpret_cast<uintptr_t>(emb_i) | reinterpret_cast<uintptr_t>(emb_I)) & 15) == 0,
               "align: buffers must be 8-byte (fp32) / 16-byte (bf16 rows, kpad %% 8 == 0) aligned");
  const long long items = (long long)Be * T;
  const unsigned blocks = (unsigned)((items + ALIGN_WARPS - 1) / ALIGN_WARPS);
  align_kernel<<<blocks, ALIGN_WARPS * 32, 0, stream>>>(out1, out2, out1a, out2a, emb_i, emb_I, Be, T, kpad, do_align, pitch_a);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// mixstep_kernel
//
// Persistent 4-CTA clusters, one sample at a time per cluster; CTA `rank` owns frames [rank*FR, rank*FR+FR).
// HBM latency under load is ~2 us, so an SM needs >= 100 KB of reads in flight to stream at its share of the
// bandwidth (tools/micro/bulk_ring.cu). Shared memory is therefore ONE ring of 25 frame slots (4 rows of 2096 B,
// one mbarrier per slot) filled with cp.async.bulk. There are no producer warps: the 16 warps take frames from a
// ticket counter in the order the ring is filled, and the warp that has just emptied a slot sends the copies of the
// frame 25 positions further down the stream into it (lane 0 arms the barrier, lanes 0..3 send one row each), so a
// slot is back in flight the moment its data is no longer needed:
//   phase 1 of a sample, per frame: rows out1a, out2a of the cond and of the uncond half  -> mix, guidance,
//           floor search; the guided prediction g goes to global memory (P.final_out doubles as that scratch:
//           it is only a result at the last step) where it stays in L2,
//   phase 2, per frame: rows x1, x2 and g -> process_xstart (centring in place in the slot) + DDIM update
//           (normalisers and DDIM coefficients folded into two per-feature tables: two FMAs per value).
// The frame stream is software-pipelined across samples - P1(0), P1(1), P2(0), P1(2), P2(1), ... - so the cluster
// exchange centring needs (min-y over all frames, frame-0 root and hips, through DSMEM and remote mbarrier
// arrives) and the visibility of g have a whole phase to complete and the ring never drains at a phase boundary.
// Streaming reads carry an L2 evict_first policy so that they do not push g out of L2.
// ------------------------------------------------------------------------------------------------
constexpr int MIX_CONSUMERS = 16;
constexpr int MIX_THREADS = MIX_CONSUMERS * 32;       // no producer warps: the warp that empties a slot refills it
constexpr int MIX_CLUSTER = 4;                         // 4 SMs per sample: 33 clusters resident (8-CTA clusters: 15)
constexpr int ROWF = 2 * NFEAT;                        // 524 floats per (sample, frame) row
constexpr int ROWB = ROWF * 4;                         // 2096 bytes
constexpr int ROW4 = ROWF / 4;                         // 131 float4
constexpr int NIT4 = (ROW4 + 31) / 32;                 // 5 float4 per lane per row
constexpr int SLOT_ROWS = 4;                           // a ring slot holds the rows of one frame of one phase
constexpr int SLOTB = SLOT_ROWS * ROWB;                // 8384 bytes
constexpr int MIX_MAX_SLOTS = 26;
constexpr int MIX_SMEM_MAX = 227 * 1024;

struct MixExport {        // per-CTA data other CTAs of the cluster read through DSMEM
  float ymin[2];
  float f0[2][9];         // frame-0 joints 0,1,2 (xyz) per person (valid on cluster rank 0)
};

struct MixLayout {
  int tab_off, wv_off, ymin_off, ex_off, cx_off, bar_off, ctr_off, ring_off, total, nslots;
};
__host__ __device__ inline MixLayout mix_layout() {
  MixLayout L;
  int o = 0;
  L.tab_off = o;   o += 4 * ROWB;                       // DDIM / normaliser tables TA1 | TB1 | TA2 | TB2, row layout
  L.wv_off = o;    o += MIX_CONSUMERS * 96 * 4;         // per-warp Mixer weights of the current frame [4][24]
  L.ymin_off = o;  o += 2 * MIX_CONSUMERS * 2 * 4;      // [sample parity][warp][person]
  L.ex_off = o;    o += 2 * (int)sizeof(MixExport);     // [sample parity]
  L.cx_off = o;    o += 2 * 2 * 8 * 4;                  // centring transforms [sample parity][person]
  o = (o + 15) / 16 * 16;
  L.bar_off = o;   o += (MIX_MAX_SLOTS + 1) * 8;        // full[] | mid
  L.ctr_off = o;   o += 16 + MIX_MAX_SLOTS * 4;         // claims[2], slot_lap[]
  o = (o + 15) / 16 * 16;
  L.ring_off = o;
  int n = (MIX_SMEM_MAX - o) / SLOTB;
  L.nslots = n > MIX_MAX_SLOTS ? MIX_MAX_SLOTS : n;
  L.total = o + L.nslots * SLOTB;
  return L;
}
static_assert(sizeof(MixExport) % 16 == 0, "MixExport is laid out in 16-byte units");
static_assert(2 * MIX_CONSUMERS * 32 >= ROWF, "two table entries per consumer thread cover a row");

__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(MIX_CONSUMERS * 32) : "memory");
}
__device__ __forceinline__ float4 lds128g(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// cluster-scope mbarrier signalling (remote arrive / acquiring wait)
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// (mbar_wait_cluster: common.cuh)
// Monotonic counters, plain volatile shared-memory words (a release store compiles to MEMBAR.ALL.CTA):
//   slot_lap[s]: how often slot s has been armed. An mbarrier parity wait is only meaningful within one
//              phase of the barrier's current one and the consumers look further ahead than that, so a warp first
//              waits until the slot has been armed for its frame - its previous use is then over, the slot's
//              barrier is in that frame's phase - and only then on the barrier.
__device__ __forceinline__ void st_volatile_smem(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile_smem(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until(const int* ctr, int want, const char* what) {
  if (ld_volatile_smem(ctr) >= want) return;
  const long long t0 = clock64();
  while (ld_volatile_smem(ctr) < want) {
    __nanosleep(32);
    if (clock64() - t0 > 6000000000LL) {
      printf("mmdm mixstep: stalled waiting for %s >= %d (block %d thread %d)\n", what, want, (int)blockIdx.x,
             (int)threadIdx.x);
      __trap();
    }
  }
}

// mbarrier wait for the ring. try_wait carries a suspend-time hint: the hardware parks the thread until the phase
// completes (or the hint expires) instead of having it poll - waiting warps share issue slots with working ones,
// and in this kernel somebody is always waiting (the producers on free slots, by design).
__device__ __forceinline__ void mix_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t it = 0;; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(200000u)
        : "memory");
    if (ok) return;
    if (it > 20000u) {   // ~4 s of 200 us naps
      printf("mmdm mixstep: mbarrier wait timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

// Ring position of item c of the CTA's stream (mix_stream.cuh)
struct RowPos { int slot, lap; };
__device__ __forceinline__ RowPos locate_row(int c, int NS, uint32_t ns_rcp) {
  RowPos r;
  mix_ring_pos(c, NS, ns_rcp, r.slot, r.lap);
  return r;
}

// Optional timeline of block 0 (tools/mix_trace.py): [warp][event][4] = tag, three clock64 stamps. Null in production.
__device__ long long* g_mix_trace = nullptr;
constexpr int MIX_TRACE_EVENTS = 512;
__device__ __forceinline__ void mix_trace(long long* tr, int warp, int& n, long long tag, long long t0, long long t1,
                                          long long t2) {
  if (tr != nullptr && n < MIX_TRACE_EVENTS) {
    long long* e = tr + ((size_t)warp * MIX_TRACE_EVENTS + n) * 4;
    e[0] = tag; e[1] = t0; e[2] = t1; e[3] = t2;
    ++n;
  }
}

template <int NW>
__global__ void __launch_bounds__(MIX_THREADS, 1)
mixstep_kernel(const MixStepParams P) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / MIX_CLUSTER, nclusters = gridDim.x / MIX_CLUSTER;
  const int T = P.T, B = P.B;
  const int FR = (T + MIX_CLUSTER - 1) / MIX_CLUSTER;
  const int t_begin = rank * FR;
  const int nfr = max(0, min(FR, T - t_begin));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool pooled = (P.mode == 1 || P.mode == 3);
  const int ncfg = P.ncfg;
  long long* const tr = (blockIdx.x == 0 && lane == 0) ? g_mix_trace : nullptr;
  int ntr = 0;

  extern __shared__ __align__(16) uint8_t smem[];
  const MixLayout L = mix_layout();
  float* tab = reinterpret_cast<float*>(smem + L.tab_off);
  float* ymin_w = reinterpret_cast<float*>(smem + L.ymin_off);
  MixExport* ex = reinterpret_cast<MixExport*>(smem + L.ex_off);
  float* cxs = reinterpret_cast<float*>(smem + L.cx_off);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* mid_bar = full + MIX_MAX_SLOTS;
  int* claims = reinterpret_cast<int*>(smem + L.ctr_off);   // frame tickets handed out so far, per phase
  int* slot_lap = claims + 4;                  // per slot: number of times it has been armed
  uint8_t* ring = smem + L.ring_off;
  const int NS = min(L.nslots, max(nfr, 1));   // (a slot is refilled by the frame NS items later, which must not be a
                                               //  later frame's phase 2 than the one being finished: NS <= nfr)
  const uint32_t ns_rcp = mix_ring_rcp(NS);

  const int step = P.mix_only ? 0 : *P.step_ptr;
  const bool final_step = (step == 0) && !P.mix_only;
  const bool phase2 = !final_step && !P.mix_only;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full[i], 1);
    }
    mbar_init(mid_bar, MIX_CLUSTER);
    for (int i = 0; i < NS; ++i) slot_lap[i] = 0;
    claims[0] = 0;
    claims[1] = 0;
    fence_barrier_init();
  }
  cluster_sync_all();   // every CTA's barriers exist before the first remote arrive

  // Item stream of this CTA: one item = the rows of one frame in one phase (phase 1: out1a, out2a of each cfg half;
  // phase 2: x1, x2, g). S samples. Order: P1(0), then for k = 1..S: [P1(k) if k < S], P2(k-1).
  const int S = cluster_id < B ? (B - cluster_id + nclusters - 1) / nclusters : 0;
  const int R1 = 2 * ncfg, R2 = 3;
  const int r1 = nfr, r2 = phase2 ? nfr : 0;   // items per sample and phase
  const int total_items = S * (r1 + r2);
  MixStream stream;
  stream.S = S; stream.nfr = nfr; stream.has2 = phase2 ? 1 : 0;
  auto base_p1 = [&](int k) { return mix_base_p1(stream, k); };
  auto base_p2 = [&](int k) { return mix_base_p2(stream, k - 1); };   // P2(k-1), processed in stage k

  {
    // ---- every warp is a consumer (and refills the slots it empties) --------------------------------------------------------------------------
    // Frames are handed out dynamically (one shared-memory ticket counter per phase) so that the ring is consumed
    // in the order it is filled - a warp held up by the cluster exchange or a late row does not sit on slots the
    // others are waiting behind. Every warp holds the ticket of its NEXT frame while it works on the current one:
    // the Mixer weights of that frame are fetched a whole frame ahead.
    const bool two = (ncfg == 2);
    const f32x2 s_cfg2 = pack2(P.cfg_scale, P.cfg_scale), s_unc2 = pack2(1.f - P.cfg_scale, 1.f - P.cfg_scale);
    const f32x2 neg1 = pack2(-1.f, -1.f);
    float* wv = reinterpret_cast<float*>(smem + L.wv_off) + warp * 96;
    // A lane always handles the same 20 features of a row (float4 j = lane + 32 i): the index of the Mixer
    // weight that multiplies each of them and whether it is a joint height (floor search) are computed once.
    // Heights only occur in j < 17 (person 0, i = 0) and 65 < j < 82 (person 1, i = 2).
    int woff[NIT4 * 4];
    uint32_t yflags = 0;   // bit e: element e of iteration 0, bit 4 + e: element e of iteration 2
#pragma unroll
    for (int i = 0; i < NIT4 * 4; ++i) {
      const int f = (lane + 32 * (i >> 2)) * 4 + (i & 3);
      const int p = f >= NFEAT, fi = f - p * NFEAT;
      const int wi = (NW == 1) ? 0 : weight_index(fi);
      woff[i] = p * 24 + wi;
      if (fi < OFF_VEL && (fi % 3) == 1 && f < ROWF) {
        if ((i >> 2) == 0) yflags |= 1u << (i & 3);
        if ((i >> 2) == 2) yflags |= 1u << (4 + (i & 3));
      }
    }
    // Mixer weights of frame t for (cfg half, person) = 4 rows of 24: three values per lane. Which row / column a
    // lane fetches never changes: element index of (sequence 0 of the cfg half and person, column) and validity.
    int wsrc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int idx = lane + 32 * k, cp = idx / 24, kk = idx - cp * 24;
      wsrc[k] = (cp < 2 * ncfg && kk < NW) ? ((cp >> 1) * B * 2 + (cp & 1)) : -1;
    }
    const int wkk0 = lane % 24, wkk1 = (lane + 32) % 24, wkk2 = (lane + 64) % 24;
    auto load_w = [&](int b, int t, float* w3) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float v = 0.f;
        if (P.has_force) {
          v = P.force_val;
        } else if (wsrc[k] >= 0) {
          const int seq = wsrc[k] + 2 * b;
          v = __ldg(P.weights + (pooled ? seq : seq * T + t) * 24 + (k == 0 ? wkk0 : k == 1 ? wkk1 : wkk2));
        }
        w3[k] = v;
      }
    };
    auto acquire_item = [&](RowPos rp) -> float* {   // wait for item rp of the stream; returns its slot
      if (lane == 0) spin_until(slot_lap + rp.slot, rp.lap + 1, "slot_lap");
      __syncwarp();
      mix_wait(&full[rp.slot], (uint32_t)rp.lap & 1u);
      return reinterpret_cast<float*>(ring + (size_t)rp.slot * SLOTB);
    };
    auto claim = [&](int which) -> int {   // next ticket of phase `which`: frame q % nfr of sample q / nfr
      int q = 0;
      if (lane == 0)
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(q) : "r"(smem_u32(claims + which)) : "memory");
      return __shfl_sync(0xffffffffu, q, 0);
    };
    // DDIM update with the normalisation folded in: x' = A pred + Bc x, pred = (g - mean) / std
    //   (gaussian_diffusion.py:1936-1965 with eta = 0: x' = pred sqrt(ab_prev) + sqrt(1 - ab_prev) eps,
    //    eps = (sqrt(1/ab) x - pred) / sqrt(1/ab - 1))  ->  x' = TA g + TB + Bc x with per-feature tables TA, TB
    // Arm slot (c % NS) with item c of the stream: lane 0 arms the barrier with the bytes of all rows and publishes
    // the acquisition, lanes 0..3 send one row copy each. Called by the warp that has just emptied the slot (item
    // c - NS), so a slot is back in flight the moment its data is no longer needed, and at start-up for the first
    // NS items. The g row of a phase-2 item waits for the flag of its frame (set by the warp that wrote the row in
    // phase 1: an earlier ticket, so that warp does not depend on this one).
    uint64_t pol_stream;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    // send the rows of item (phase p1, sample n, frame fl) into slot rp: lane 0 arms the barrier and publishes the
    // acquisition, lanes 0..3 send one row copy each
    // When out1a|out2a (x1|x2) are interleaved row by row, a pair of rows is one contiguous 4192-byte copy.
    const bool pred_pairs = P.out2a == P.out1a + ROWF && P.pred_pitch == 2 * ROWF;
    const bool state_pairs = P.x2 == P.x1 + ROWF && P.state_pitch == 2 * ROWF;
    auto arm_item = [&](bool p1, int n, int fl, RowPos rp, bool defer_g) {
      const int rows = p1 ? R1 : R2;
      if (lane == 0) {
        mbar_arrive_expect_tx(&full[rp.slot], (uint32_t)rows * ROWB);
        st_volatile_smem(slot_lap + rp.slot, rp.lap + 1);
      }
      __syncwarp();
      if (lane < rows) {
        const int r = lane;
        const int b = cluster_id + n * nclusters;
        const uint32_t dst = smem_u32(ring + (size_t)rp.slot * SLOTB + r * ROWB), bar = smem_u32(&full[rp.slot]);
        if (p1) {
          const float* row = P.out1a + ((size_t)((r >> 1) * B + b) * T + t_begin + fl) * P.pred_pitch;
          if (pred_pairs) { if (!(r & 1)) bulk_g2s_hint(dst, row, 2 * ROWB, bar, pol_stream); }
          else bulk_g2s_hint(dst, (r & 1) ? row + (P.out2a - P.out1a) : row, ROWB, bar, pol_stream);
        } else if (r < 2) {
          const float* row = P.x1 + ((size_t)b * T + t_begin + fl) * P.state_pitch;
          if (state_pairs) { if (r == 0) bulk_g2s_hint(dst, row, 2 * ROWB, bar, pol_stream); }
          else bulk_g2s_hint(dst, r ? row + (P.x2 - P.x1) : row, ROWB, bar, pol_stream);
        } else if (!defer_g) {
          bulk_g2s_hint(dst, P.final_out + ((size_t)b * T + t_begin + fl) * ROWF, ROWB, bar, pol_stream);
        }
      }
      __syncwarp();
    };
    // Refill the slot a warp has just emptied with the item NS positions further down the stream (same slot, next
    // lap), which lies in the current segment or the one after it (NS <= nfr). Segments: P1(0), then
    // [P1(k), P2(k-1)] for k = 1..S-1, then P2(S-1); without a phase 2 just P1(0..S-1).
    // The g row of a phase-2 item was written (and fenced for the async proxy) before the consumer barrier that ended
    // phase 1 of its sample. That barrier lies behind this warp unless the item is armed from that very phase 1 (only
    // with one sample per cluster): those rows are sent right after the barrier (send_deferred_g).
    auto rearm = [&](bool cur_p1, int n_cur, int f, RowPos cur) {
      bool p1 = cur_p1;
      int n = n_cur, t = f;
      if (!mix_item_after(stream, NS, p1, n, t)) return;   // end of the stream
      RowPos rp = cur;
      ++rp.lap;
      arm_item(p1, n, t, rp, cur_p1 && S == 1);
    };
    auto send_deferred_g = [&]() {   // S == 1: g rows of the phase-2 frames 0 .. NS-1, whose slots were armed during phase 1
      for (int fl = warp; fl < min(NS, nfr); fl += MIX_CONSUMERS) {
        if (lane == 0) {
          const RowPos rp = locate_row(r1 + fl, NS, ns_rcp);
          bulk_g2s_hint(smem_u32(ring + (size_t)rp.slot * SLOTB + 2 * ROWB),
                        P.final_out + ((size_t)cluster_id * T + t_begin + fl) * ROWF, ROWB, smem_u32(&full[rp.slot]), pol_stream);
        }
      }
    };
    for (int c = warp; c < min(NS, total_items); c += MIX_CONSUMERS) {   // start-up: the first NS items are P1(0) frames
      RowPos rp; rp.slot = c; rp.lap = 0;
      arm_item(true, 0, c, rp, false);
    }

    const int tickets = S * nfr;
    int next1 = nfr > 0 ? claim(0) : tickets, next2 = (nfr > 0 && phase2) ? claim(1) : tickets;
    int nk1 = 0;   // sample of ticket next1 (tickets only grow: no division)
    auto prefetch_w = [&](float* w3) {
      if (next1 < tickets) {
        while (next1 >= (nk1 + 1) * nfr) ++nk1;
        load_w(cluster_id + nk1 * nclusters, t_begin + next1 - nk1 * nfr, w3);
      }
    };
    float w3[3] = {0.f, 0.f, 0.f};
    prefetch_w(w3);   // in flight while the tables below are built
    f32x2 bc2 = pack2(0.f, 0.f);
    if (phase2) {
      // (first read in phase 2 of sample 0, i.e. behind the consumer barrier that ends its phase 1)
      const float c_recip = P.coef[step * 4 + 0], c_recipm1 = P.coef[step * 4 + 1];
      const float c_ap = P.coef[step * 4 + 2], c_eps = P.coef[step * 4 + 3];
      float sd1[2], sd2[2], mn1[2], mn2[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {   // all loads first: one memory latency for the whole prologue
        const int i = threadIdx.x + q * MIX_CONSUMERS * 32;
        const int fi = i >= NFEAT ? i - NFEAT : i;
        const bool ok = i < ROWF;
        sd1[q] = ok ? P.std1[fi] : 1.f; sd2[q] = ok ? P.std2[fi] : 1.f;
        mn1[q] = ok ? P.mean1[fi] : 0.f; mn2[q] = ok ? P.mean2[fi] : 0.f;
      }
      const float A = (float)((double)c_ap - (double)c_eps / (double)c_recipm1);
      const float Bc = (float)((double)c_eps / (double)c_recipm1 * (double)c_recip);
      bc2 = pack2(Bc, Bc);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i = threadIdx.x + q * MIX_CONSUMERS * 32;
        if (i < ROWF) {
          const float r1s = 1.0f / sd1[q], r2s = 1.0f / sd2[q];
          tab[i] = A * r1s;                          // TA, stream 1 (individual-model normaliser)
          tab[ROWF + i] = -A * mn1[q] * r1s;         // TB
          tab[2 * ROWF + i] = A * r2s;               // TA, stream 2 (interaction-model normaliser)
          tab[3 * ROWF + i] = -A * mn2[q] * r2s;     // TB
        }
      }
    }

    for (int k = 0; k <= S; ++k) {
      if (k < S) {
        // ---- phase 1 of sample k: mix + guidance for my frames -------------------------------------
        const int b = cluster_id + k * nclusters;
        const int cb = base_p1(k);
        float ym0 = INFINITY, ym1 = INFINITY;
        while (next1 < (k + 1) * nfr) {
          const int fl = next1 - k * nfr;
          const int t = t_begin + fl;
          __syncwarp();   // the previous frame's readers of wv are done
#pragma unroll
          for (int q = 0; q < 3; ++q) wv[lane + 32 * q] = w3[q];
          next1 = claim(0);
          prefetch_w(w3);
          const long long tc0 = tr ? clock64() : 0;
          const RowPos ip = locate_row(cb + fl, NS, ns_rcp);
          const float* o1c = acquire_item(ip);       // (the __syncwarp inside also orders the wv stores)
          const float* o2c = o1c + ROWF;
          const float* o1u = two ? o1c + 2 * ROWF : o1c;
          const float* o2u = two ? o1c + 3 * ROWF : o2c;
          const size_t rc = ((size_t)b * T + t) * ROWF;            // cond half row
          const size_t ru = ((size_t)(B + b) * T + t) * ROWF;      // uncond half row
          float4* const gp = reinterpret_cast<float4*>((P.mix_only == 2 ? P.mix_out : P.final_out) + rc) + lane;
          const long long tc1 = tr ? clock64() : 0;
#pragma unroll
          for (int i = 0; i < NIT4; ++i) {
            const int j = lane + 32 * i;
            if (j < ROW4) {
              const float4 a4 = lds128g(o1c + 4 * j), b4 = lds128g(o2c + 4 * j);
              float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = c4;
              if (two) { c4 = lds128g(o1u + 4 * j); d4 = lds128g(o2u + 4 * j); }
              float g[4], mcv[4], muv[4], wcv[4], wuv[4];
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {   // two features per packed fp32x2 instruction
                const int e = 2 * h2;
                wcv[e] = wv[woff[i * 4 + e]];          // (plain loads: the five iterations are independent and the
                wcv[e + 1] = wv[woff[i * 4 + e + 1]];  //  compiler may interleave them)
                const f32x2 av = h2 ? pack2(a4.z, a4.w) : pack2(a4.x, a4.y);
                const f32x2 bv = h2 ? pack2(b4.z, b4.w) : pack2(b4.x, b4.y);
                const f32x2 mc = fma2(pack2(wcv[e], wcv[e + 1]), fma2(bv, neg1, av), bv);   // b + w (a - b)
                f32x2 gg = mul2(s_cfg2, mc);
                unpack2(mc, mcv[e], mcv[e + 1]);
                muv[e] = muv[e + 1] = wuv[e] = wuv[e + 1] = 0.f;
                if (two) {
                  wuv[e] = wv[48 + woff[i * 4 + e]];
                  wuv[e + 1] = wv[48 + woff[i * 4 + e + 1]];
                  const f32x2 cv = h2 ? pack2(c4.z, c4.w) : pack2(c4.x, c4.y);
                  const f32x2 dv = h2 ? pack2(d4.z, d4.w) : pack2(d4.x, d4.y);
                  const f32x2 mu = fma2(pack2(wuv[e], wuv[e + 1]), fma2(dv, neg1, cv), dv);
                  gg = fma2(s_unc2, mu, gg);
                  unpack2(mu, muv[e], muv[e + 1]);
                }
                unpack2(gg, g[e], g[e + 1]);
              }
              if (i == 0 || i == 2) {   // joint heights of person 0 / person 1
                const uint32_t yf = i == 0 ? yflags : yflags >> 4;
                float& ym = i == 0 ? ym0 : ym1;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (yf & (1u << e)) ym = fminf(ym, g[e]);
              }
              if (P.hist_i1 != nullptr) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int f = 4 * j + e, p = f >= NFEAT, fi = f - p * NFEAT;
                  float* hbase = (p ? P.hist_i2 : P.hist_i1) +
                                 (size_t)(P.mix_only ? 0 : (P.nsteps - 1 - step)) * ((size_t)ncfg * B * T * NFEAT);
                  hbase[((size_t)b * T + t) * NFEAT + fi] = wcv[e];
                  if (two) hbase[((size_t)(B + b) * T + t) * NFEAT + fi] = wuv[e];
                }
              }
              if (P.mix_only == 1 || (P.mix_only == 0 && P.mix_out != nullptr)) {   // out_influenced ("eval" history)
                *reinterpret_cast<float4*>(P.mix_out + rc + 4 * j) = make_float4(mcv[0], mcv[1], mcv[2], mcv[3]);
                if (two) *reinterpret_cast<float4*>(P.mix_out + ru + 4 * j) = make_float4(muv[0], muv[1], muv[2], muv[3]);
              }
              if (P.mix_only != 1) {
                // the guided prediction: Mixer.forward's CFG output, the sampler's result at the last step (rows of
                // the un-doubled batch coincide with the cond half), phase 2's input at every other step
                gp[32 * i] = make_float4(g[0], g[1], g[2], g[3]);
              }
            }
          }
          __syncwarp();
          rearm(true, k, fl, ip);   // the slot goes straight back into flight with the frame NS items ahead
          mix_trace(tr, warp, ntr, (1LL << 32) | (k * nfr + fl), tc0, tc1, tr ? clock64() : 0);
        }
        if (phase2) {
          ym0 = warp_min(ym0);
          ym1 = warp_min(ym1);
          float* yw = ymin_w + (k & 1) * (MIX_CONSUMERS * 2);
          if (lane == 0) { yw[warp * 2] = ym0; yw[warp * 2 + 1] = ym1; }
          fence_proxy_async_all();   // my g stores are ordered before the bulk reads of those rows in phase 2
          const long long ts0 = tr ? clock64() : 0;
          consumer_sync();
          if (S == 1) send_deferred_g();
          mix_trace(tr, warp, ntr, (4LL << 32) | k, ts0, tr ? clock64() : 0, 0);
          if (warp == 0) {
            MixExport* exk = ex + (k & 1);
            if (lane < 2) {
              float m = INFINITY;
              for (int w = 0; w < MIX_CONSUMERS; ++w) m = fminf(m, yw[w * 2 + lane]);
              exk->ymin[lane] = m;
            }
            if (rank == 0 && lane < 18) {   // frame 0 of the sample is mine: joints 0, 1, 2 of both persons
              const int p = lane / 9, e = lane - p * 9;
              exk->f0[p][e] = __ldcg(P.final_out + (size_t)b * T * ROWF + p * NFEAT + e);
            }
            __syncwarp();
            if (lane < MIX_CLUSTER) mbar_arrive_remote_release(mapa_u32(smem_u32(mid_bar), (uint32_t)lane));
            mbar_wait_cluster(mid_bar, (uint32_t)k & 1u);
            if (lane < 2) {
              const int p = lane;
              const MixExport* e0 = cluster.map_shared_rank(exk, 0);
              float m = INFINITY;
              for (int r = 0; r < MIX_CLUSTER; ++r) m = fminf(m, cluster.map_shared_rank(exk, r)->ymin[p]);
              const float* f0 = e0->f0[p];
              // FACE_JOINT_INDX[:2] = (r_hip=2, l_hip=1)  (utils/paramUtil.py:89)
              const CenterXf c = make_center(v3(f0[0], f0[1] - m, f0[2]), v3(f0[6], f0[7], f0[8]), v3(f0[3], f0[4], f0[5]), m);
              float* o = cxs + ((k & 1) * 2 + p) * 8;
              o[0] = c.q.w; o[1] = c.q.x; o[2] = c.q.y; o[3] = c.q.z; o[4] = c.floor_y; o[5] = c.sx; o[6] = c.sz;
            }
            __syncwarp();
          }
        }
      }
      if (k >= 1 && phase2) {
        // ---- phase 2 of sample k-1: process_xstart + DDIM on both state streams ----------------------
        const int n = k - 1;
        const int b = cluster_id + n * nclusters;
        const int cb = base_p2(k);
        const long long ts1 = tr ? clock64() : 0;
        consumer_sync();   // warp 0 has published the sample's centring transforms
        mix_trace(tr, warp, ntr, (5LL << 32) | n, ts1, tr ? clock64() : 0, 0);
        CenterXf cx[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const float* o = cxs + ((n & 1) * 2 + p) * 8;
          cx[p].q.w = o[0]; cx[p].q.x = o[1]; cx[p].q.y = o[2]; cx[p].q.z = o[3];
          cx[p].floor_y = o[4]; cx[p].sx = o[5]; cx[p].sz = o[6];
        }
        while (next2 < (n + 1) * nfr) {
          const int fl = next2 - n * nfr;
          next2 = claim(1);
          const int t = t_begin + fl;
          const size_t rx = ((size_t)b * T + t) * P.state_pitch;
          float4* const x1p = reinterpret_cast<float4*>(P.x1 + rx) + lane;
          float4* const x2p = reinterpret_cast<float4*>(P.x2 + rx) + lane;
          const long long tc0 = tr ? clock64() : 0;
          const RowPos ip = locate_row(cb + fl, NS, ns_rcp);
          const float* xr1 = acquire_item(ip);
          const float* xr2 = xr1 + ROWF;
          float* crow = const_cast<float*>(xr1) + 2 * ROWF;
          const long long tc1 = tr ? clock64() : 0;
          // stream 2 (interaction space): the raw guided prediction, normalised inside the tables
#pragma unroll
          for (int i = 0; i < NIT4; ++i) {
            const int j = lane + 32 * i;
            if (j < ROW4) {
              const float4 x4 = lds128g(xr2 + 4 * j), g4 = lds128g(crow + 4 * j);
              const float4 ta = lds128g(tab + 2 * ROWF + 4 * j), tb = lds128g(tab + 3 * ROWF + 4 * j);
              float o[4];
              unpack2(fma2(pack2(ta.x, ta.y), pack2(g4.x, g4.y), fma2(bc2, pack2(x4.x, x4.y), pack2(tb.x, tb.y))), o[0], o[1]);
              unpack2(fma2(pack2(ta.z, ta.w), pack2(g4.z, g4.w), fma2(bc2, pack2(x4.z, x4.w), pack2(tb.z, tb.w))), o[2], o[3]);
              x2p[32 * i] = make_float4(o[0], o[1], o[2], o[3]);
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
          for (int i = 0; i < NIT4; ++i) {
            const int j = lane + 32 * i;
            if (j < ROW4) {
              const float4 x4 = lds128g(xr1 + 4 * j), g4 = lds128g(crow + 4 * j);
              const float4 ta = lds128g(tab + 4 * j), tb = lds128g(tab + ROWF + 4 * j);
              float o[4];
              unpack2(fma2(pack2(ta.x, ta.y), pack2(g4.x, g4.y), fma2(bc2, pack2(x4.x, x4.y), pack2(tb.x, tb.y))), o[0], o[1]);
              unpack2(fma2(pack2(ta.z, ta.w), pack2(g4.z, g4.w), fma2(bc2, pack2(x4.z, x4.w), pack2(tb.z, tb.w))), o[2], o[3]);
              x1p[32 * i] = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
          fence_proxy_async_smem();   // the centring wrote into the slot: order it before the next bulk copy into it
          __syncwarp();
          rearm(false, n, fl, ip);
          mix_trace(tr, warp, ntr, (2LL << 32) | (n * nfr + fl), tc0, tc1, tr ? clock64() : 0);
        }
      }
    }
  }
  cluster_sync_all();   // no CTA retires while another one may still read its exports or signal its barriers
}

static int g_mix_clusters[kMaxDevices][2] = {};   // resident clusters per device and instantiation (0 = not queried yet)

extern "C" int mmdm_debug_mix_trace(long long* dev_ptr) {
  return cudaMemcpyToSymbol(g_mix_trace, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? 0 : -2;
}

int motion_init() {   // kernel attributes are per device: done once for each device that launches
  static PerDeviceFlag attr;
  if (!attr.first()) return 0;
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(mixstep_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MIX_SMEM_MAX));
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(mixstep_kernel<23>, cudaFuncAttributeMaxDynamicSharedMemorySize, MIX_SMEM_MAX));
  return 0;
}

int launch_mixstep(const MixStepParams& P, cudaStream_t stream) {
  MMDM_REQUIRE(motion_init() == 0, "mixstep: kernel attribute setup failed");
  MMDM_REQUIRE(P.mode >= 1 && P.mode <= 4, "Mixing mode not recognized (%d)", P.mode);
  MMDM_REQUIRE(P.T >= 1 && P.B >= 1 && (long long)P.B * P.T * 12 < (1LL << 31), "mixstep: bad shape B=%d T=%d", P.B, P.T);
  MMDM_REQUIRE(P.ncfg == 1 || P.ncfg == 2, "mixstep: ncfg must be 1 or 2 (got %d)", P.ncfg);
  MMDM_REQUIRE(P.pred_pitch >= 524 && (P.pred_pitch % 4) == 0 && (P.mix_only || (P.state_pitch >= 524 && (P.state_pitch % 4) == 0)),
               "mixstep: row pitches must be multiples of 4 floats, at least 524 (got %d, %d)", P.pred_pitch, P.state_pitch);
  MMDM_REQUIRE(P.weights != nullptr, "mixstep: Mixer weights missing");
  MMDM_REQUIRE(((reinterpret_cast<uintptr_t>(P.out1a) | reinterpret_cast<uintptr_t>(P.out2a) |
                 reinterpret_cast<uintptr_t>(P.weights) | reinterpret_cast<uintptr_t>(P.x1) |
                 reinterpret_cast<uintptr_t>(P.x2) | reinterpret_cast<uintptr_t>(P.mix_out) |
                 reinterpret_cast<uintptr_t>(P.final_out)) & 15) == 0,
               "mixstep: buffers must be 16-byte aligned");
  const int NW = (P.mode <= 2) ? 1 : 23;
  const MixLayout L = mix_layout();
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(MIX_THREADS);
  cfg.dynamicSmemBytes = (size_t)L.total;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MIX_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int& resident = g_mix_clusters[current_device()][NW == 1 ? 0 : 1];
  if (resident == 0) {   // how many clusters the device holds at once (the kernel is persistent over the samples)
    cfg.gridDim = dim3((unsigned)(MIX_CLUSTER * 64));
    cfg.dynamicSmemBytes = MIX_SMEM_MAX;
    int n = 0;
    cudaError_t e = NW == 1 ? cudaOccupancyMaxActiveClusters(&n, mixstep_kernel<1>, &cfg)
                            : cudaOccupancyMaxActiveClusters(&n, mixstep_kernel<23>, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    resident = n > 0 ? n : -1;
    cfg.dynamicSmemBytes = (size_t)L.total;
  }
  const int nclusters = resident > 0 && resident < P.B ? resident : P.B;
  cfg.gridDim = dim3((unsigned)(nclusters * MIX_CLUSTER));
  if (NW == 1) MMDM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mixstep_kernel<1>, P));
  else MMDM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mixstep_kernel<23>, P));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// smooth_time_kernel: temporal Gaussian blur of a generated motion, the post-processing step of the reference's
// inference script (scripts/infer/mixermdm.py:127-132: gaussian_filter1d(motion, sigma = 1, axis = 0,
// mode = 'nearest') on the (T, 2, 262) view of output[0]). out[b,t,:] = sum_k w[k] x[b, clamp(t+k-R, 0, T-1), :].
// HBM-bound (one read, one write per value): a thread owns one float4 column of one sample for a run of frames
// and slides a register window of 2R+1 rows over it, so every input row is read once per run (+2R halo rows).
// ------------------------------------------------------------------------------------------------
constexpr int SMOOTH_MAX_R = 8;
constexpr int SMOOTH_RUN = 32;   // frames per thread

struct SmoothParams {
  const float* x;
  float* out;
  int B, T, F4;   // F4 = features / 4
  int R;
  float w[2 * SMOOTH_MAX_R + 1];
};

template <int R>
__global__ void __launch_bounds__(128) smooth_time_kernel(const SmoothParams P) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= P.F4) return;
  const int b = blockIdx.z, t0 = blockIdx.y * SMOOTH_RUN;
  const int t1 = min(P.T, t0 + SMOOTH_RUN);
  const float4* src = reinterpret_cast<const float4*>(P.x) + (size_t)b * P.T * P.F4 + col;
  float4* dst = reinterpret_cast<float4*>(P.out) + (size_t)b * P.T * P.F4 + col;
  auto row = [&](int t) { return __ldg(src + (size_t)min(max(t, 0), P.T - 1) * P.F4); };
  float4 win[2 * R + 1];
#pragma unroll
  for (int k = 0; k < 2 * R; ++k) win[k + 1] = row(t0 - R + k);
  for (int t = t0; t < t1; ++t) {
#pragma unroll
    for (int k = 0; k < 2 * R; ++k) win[k] = win[k + 1];
    win[2 * R] = row(t + R);
    // symmetric taps summed pairwise around the centre, like scipy's correlate1d does for a symmetric kernel
    float4 acc = make_float4(P.w[R] * win[R].x, P.w[R] * win[R].y, P.w[R] * win[R].z, P.w[R] * win[R].w);
#pragma unroll
    for (int k = 1; k <= R; ++k) {
      const float wk = P.w[R + k];
      acc.x = fmaf(wk, win[R - k].x + win[R + k].x, acc.x);
      acc.y = fmaf(wk, win[R - k].y + win[R + k].y, acc.y);
      acc.z = fmaf(wk, win[R - k].z + win[R + k].z, acc.z);
      acc.w = fmaf(wk, win[R - k].w + win[R + k].w, acc.w);
    }
    dst[(size_t)t * P.F4] = acc;
  }
}

int launch_smooth_time(const float* x, float* out, int B, int T, int F, const float* weights, int R,
                       cudaStream_t stream) {
  MMDM_REQUIRE(B >= 1 && T >= 1 && F >= 4 && (F % 4) == 0, "smooth_motion: bad shape B=%d T=%d F=%d", B, T, F);
  MMDM_REQUIRE(R >= 0 && R <= SMOOTH_MAX_R, "smooth_motion: kernel radius %d outside [0, %d]", R, SMOOTH_MAX_R);
  MMDM_REQUIRE(x != out, "smooth_motion: in-place filtering is not supported");
  MMDM_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               "smooth_motion: buffers must be 16-byte aligned");
  SmoothParams P = {};
  P.x = x; P.out = out; P.B = B; P.T = T; P.F4 = F / 4; P.R = R;
  for (int k = 0; k < 2 * R + 1; ++k) P.w[k] = weights[k];
  const dim3 grid((unsigned)((P.F4 + 127) / 128), (unsigned)((T + SMOOTH_RUN - 1) / SMOOTH_RUN), (unsigned)B);
  switch (R) {
    case 0: smooth_time_kernel<0><<<grid, 128, 0, stream>>>(P); break;
    case 1: smooth_time_kernel<1><<<grid, 128, 0, stream>>>(P); break;
    case 2: smooth_time_kernel<2><<<grid, 128, 0, stream>>>(P); break;
    case 3: smooth_time_kernel<3><<<grid, 128, 0, stream>>>(P); break;
    case 4: smooth_time_kernel<4><<<grid, 128, 0, stream>>>(P); break;
    case 5: smooth_time_kernel<5><<<grid, 128, 0, stream>>>(P); break;
    case 6: smooth_time_kernel<6><<<grid, 128, 0, stream>>>(P); break;
    case 7: smooth_time_kernel<7><<<grid, 128, 0, stream>>>(P); break;
    default: smooth_time_kernel<8><<<grid, 128, 0, stream>>>(P); break;
  }
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// format_motion_kernel: output format of the evaluation generators (evaluation/datasets.py:121-134): the sampled
// (B, T, 2*nf) motion is de-normalised per person (MotionNormalizer.backward: x * std + mean, optional) and padded
// with zero frames to the data set's max_length -> (B, Tmax, 2, nf). One pass, one thread per output value pair.
// ------------------------------------------------------------------------------------------------
__global__ void format_motion_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int T, int Tmax,
                                     int nf, const float* __restrict__ mean, const float* __restrict__ stdv) {
  const int F = 2 * nf;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // pair index
  const long long total = (long long)B * Tmax * (F / 2);
  if (idx >= total) return;
  const int f = (int)(idx % (F / 2)) * 2;
  const long long bt = idx / (F / 2);
  const int t = (int)(bt % Tmax), b = (int)(bt / Tmax);
  float2 v = make_float2(0.f, 0.f);
  if (t < T) {
    v = *reinterpret_cast<const float2*>(x + ((size_t)b * T + t) * F + f);
    if (mean != nullptr) {
      const int f0 = f >= nf ? f - nf : f, f1 = f + 1 >= nf ? f + 1 - nf : f + 1;
      v.x = v.x * stdv[f0] + mean[f0];
      v.y = v.y * stdv[f1] + mean[f1];
    }
  }
  *reinterpret_cast<float2*>(out + ((size_t)b * Tmax + t) * F + f) = v;
}

int launch_format_motion(const float* x, float* out, int B, int T, int Tmax, int nf, const float* mean,
                         const float* stdv, cudaStream_t stream) {
  MMDM_REQUIRE(B >= 1 && T >= 1 && Tmax >= T && nf >= 2 && (nf % 2) == 0,
               "format_motion: bad shape B=%d T=%d max_length=%d nfeats=%d", B, T, Tmax, nf);
  MMDM_REQUIRE((mean == nullptr) == (stdv == nullptr), "format_motion: mean and std come together");
  MMDM_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 7) == 0,
               "format_motion: buffers must be 8-byte aligned");
  const long long total = (long long)B * Tmax * nf;
  format_motion_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, out, B, T, Tmax, nf, mean, stdv);
  MMDM_CHECK_CUDA(cudaGetLastError());
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
