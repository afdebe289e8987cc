// Fused attention for motion sequences on sm_100a, with the reference's zero-attention sink:
// nn.MultiheadAttention(add_zero_attn=True) appends one key with logit 0 and value 0
// (reference: models/utils/layers.py:33-34,74-75; torch F.multi_head_attention_forward).
//
// Round-2 design: 64-key blocks with an online (lazily rescaled) softmax and TWO query tiles in flight per SM.
//   S_g[128 x 64]  = Q_g K_kb^T     tcgen05.mma (SS), fp32 in TMEM, one buffer per softmax warpgroup g
//   P_g            = exp2((S_g - m) c)   warpgroup g: thread <-> query row, 64 logits in registers, no shuffles
//   O_g[128 x hd] += P_g V_kb       tcgen05.mma with the A operand (P, packed bf16) read from TMEM (double buffered),
//                                   V consumed K-major from the transposed V^T buffer the QKV GEMM writes.
// A warpgroup never waits for the tensor pipe in steady state: S_g(kb+1) is issued as soon as the warpgroup has
// pulled S_g(kb) into registers, P has its own double-buffered TMEM columns, so the exponentials (MUFU, the
// bound of this kernel next to the MMAs: 128 x 304 / 16 per clock = 2432 cycles per tile for either) of the two
// warpgroups interleave on the SM sub-partitions while the other one loads, reduces or stores.
// Key row T is zero-filled by TMA out-of-bounds handling, which *is* the zero-attention key (logit 0, value 0);
// keys > T are masked. The running maximum starts at 0 (the sink's logit) and is only raised when a block exceeds it
// by more than 8 in the exponent (factor 256): the rescale of O is rare and done in place by the warpgroup itself.
#include "attention.cuh"

#include <stdlib.h>

namespace mmdm {

struct AttnParams {
  int S, T, H;
  int NB;        // 64-key blocks per query tile: ceil((T + 1) / 64)
  int QT;        // 128-row query tiles per (sequence, head): ceil(T / 128)
  int GS;        // tiles that share one pass over the K / V blocks (see "groups" below)
  int kv_xor, rev;
  int ldo, wide_store;   // output row pitch (elements); 32-byte stores allowed
  int stagger;           // cycles the second warpgroup starts late
  __nv_bfloat16* out;
  const unsigned long long* kmask;   // optional key-padding mask: [S][NB] words, bit j of word kb = key kb*64+j attendable
  float scale_log2, tau_raw;
};

template <int HD>
struct AttnCfg {
  static constexpr int KA = HD / 64;                       // 64-wide K atoms along the head dim
  static constexpr int NSLOT = 5;                          // K / V block ring slots (5 x 64 keys hold T <= 319 whole)
  static constexpr uint32_t Q_BYTES = KA * 128 * 128;      // one Q tile  [KA][128 rows][128 B]
  static constexpr uint32_t KB_BYTES = KA * 64 * 128;      // one key block   [KA][64 keys][128 B]
  static constexpr uint32_t VB_BYTES = HD * 128;           // one V^T block   [HD rows][64 keys = 128 B]
  static constexpr uint32_t NBAR = 18 + 4 * NSLOT;
  static constexpr uint32_t USED_BYTES = 2 * Q_BYTES + NSLOT * (KB_BYTES + VB_BYTES) + NBAR * 8 + 16;
  static constexpr size_t SMEM_BYTES = USED_BYTES + 1024;
  // TMEM columns: warpgroup g owns [g*128, +128) as two 64-column S buffers (block m -> buffer m & 1); P(m), packed
  // bf16, overwrites the first 32 columns of its own S buffer once the row's logits are in registers; O_g at
  // [256 + g*HD, +HD)
  static constexpr uint32_t SP_COL = 0, O_COL = 256;
  static constexpr int THREADS = 352;   // 8 softmax warps, 2 MMA warps (one per warpgroup), loader
};

#ifdef MMDM_ATTN_TRACE
// clock64 timeline of CTA 0 (tools/attn_trace.py): [12 warps][64 blocks][16 events]
__device__ long long* g_attn_trace = nullptr;
#define TR(w, n, ev)                                                                                   \
  do {                                                                                                 \
    if (trace_ptr != nullptr && (n) < 64) trace_ptr[(((w) * 64) + (n)) * 16 + (ev)] = clock64();         \
  } while (0)
#else
#define TR(w, n, ev) do {} while (0)
#endif

// Opaque copy: keeps ptxas from hoisting "base + constant" address arithmetic of every call site out of the loops
__device__ __forceinline__ uint32_t launder(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

// Tile t of this CTA (t counts the 128-row query tiles of the CTA's work items in order).
//   item k = (sequence, head), qt = query tile inside the item;
//   group  = the tiles that walk the K / V blocks of an item together. If the whole key range fits the ring
//            (NB <= NSLOT) the group is the item: blocks are loaded once and stay resident for all its tiles.
//            Otherwise (long sequences) a group is a pair of tiles that streams the blocks through the ring.
//   wg     = softmax warpgroup that owns the tile.
struct TileInfo {
  int k, qt, grp, gsize, pos, wg;
};
__device__ __forceinline__ TileInfo tile_info(const AttnParams& p, int t) {
  TileInfo ti;
  ti.k = t / p.QT;
  ti.qt = t - ti.k * p.QT;
  const int ng = (p.QT + p.GS - 1) / p.GS;
  const int gi = ti.qt / p.GS;
  ti.grp = ti.k * ng + gi;
  ti.pos = ti.qt - gi * p.GS;
  ti.gsize = min(p.GS, p.QT - gi * p.GS);
  ti.wg = (p.GS == p.QT) ? (t & 1) : ti.pos;   // resident: alternate over the CTA's tiles; streaming: position in the pair
  return ti;
}

// Ring position of K / V block number bidx (blocks are numbered in load order: group * NB + kb)
struct RingPos {
  int slot, ph;
};
template <int NSLOT>
__device__ __forceinline__ RingPos ring_pos(int bidx) {
  RingPos r;
  r.ph = bidx / NSLOT;
  r.slot = bidx - r.ph * NSLOT;
  return r;
}
template <int NSLOT>
__device__ __forceinline__ void ring_next(RingPos& r) {
  if (++r.slot == NSLOT) {
    r.slot = 0;
    r.ph++;
  }
}

// Barrier protocol (one CTA; g = softmax warpgroup, m = running block number of the warpgroup). A barrier may run ONE
// phase ahead of its waiter, never two (parity waits alias).
//   sfull[g,b] MMA commit      S(m) complete in buffer b = m & 1. A commit covers ALL earlier MMAs of the issuing thread
//                              and the issue order is S(0), S(1), then P V(m), S(m+2) for every m: S runs up to two blocks
//                              ahead of the softmax (its round trip through the MMA warp and the in-order pipe, 1300-1900
//                              cycles, is off the critical path), S(m+2) complete => P V(m) complete, i.e. the V block of
//                              m is free - the softmax warps release K and V ring slots from that. S(m+2) may overwrite
//                              buffer b because it is issued behind P V(m), the reader of P(m), and P(m) was published
//                              after the warpgroup pulled S(m) out.
//   pfull[g,b] 4 softmax warps P(m) written over the head of S buffer b = m & 1
//   flush[g]   MMA commit      once per tile behind P V(NB - 2): what S(m + 1) is for earlier blocks, for the rescale
//                              path of the last block
//   ofull[g]   MMA commit      last P V of the tile complete;   ofree[g]  4 warps: O_g read out
//   qfull[g]   TMA             Q tile landed;                   qfree[g]  MMA commit: last S of the tile complete
//   kfull/vfull[slot] TMA;     kfree/vfree[slot]  one arrival per tile of the group (a softmax warp that has seen the
//                              covering S complete, or an MMA commit for the last two V blocks of a tile)
template <int HD, bool MASKED>
__global__ void __launch_bounds__(AttnCfg<HD>::THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = AttnCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;                                   // [2 warpgroups][Q_BYTES]
  uint8_t* Ks = Qs + 2 * C::Q_BYTES;                    // [NSLOT][KB_BYTES]
  uint8_t* Vs = Ks + C::NSLOT * C::KB_BYTES;            // [NSLOT][VB_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + C::NSLOT * C::VB_BYTES);
  uint64_t* qfull = bars;          // [2]
  uint64_t* qfree = bars + 2;      // [2]
  uint64_t* sfull = bars + 4;      // [2 warpgroups][2 buffers]
  uint64_t* pfull = bars + 8;      // [2][2]
  uint64_t* ofull = bars + 12;     // [2]
  uint64_t* ofree = bars + 14;     // [2]
  uint64_t* flush = bars + 16;     // [2]
  uint64_t* kfull = bars + 18;                   // [NSLOT]
  uint64_t* kfree = kfull + C::NSLOT;            // [NSLOT]
  uint64_t* vfull = kfree + C::NSLOT;            // [NSLOT]
  uint64_t* vfree = vfull + C::NSLOT;            // [NSLOT]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NBAR);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
#ifdef MMDM_ATTN_TRACE
  long long* const trace_ptr = (blockIdx.x == 0 && lane == 0) ? g_attn_trace : nullptr;
#endif
  const int total_items = p.S * p.H;
  const int nitems = (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int ntiles = nitems * p.QT;
  const int NB = p.NB;

  if (tid == 0) {
    if ((size_t)(smem - smem_raw) + C::USED_BYTES > C::SMEM_BYTES) {
      printf("mmdm attention: dynamic shared memory window is not aligned enough\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&qfull[g], 1);
      mbar_init(&qfree[g], 1);
      mbar_init(&sfull[g * 2], 1);
      mbar_init(&sfull[g * 2 + 1], 1);
      mbar_init(&ofull[g], 1);
      mbar_init(&ofree[g], 4);
      mbar_init(&flush[g], 1);
      for (int b = 0; b < 2; ++b) mbar_init(&pfull[g * 2 + b], 4);
    }
    for (int s = 0; s < C::NSLOT; ++s) {
      mbar_init(&kfull[s], 1);
      mbar_init(&kfree[s], (uint32_t)p.GS);
      mbar_init(&vfull[s], 1);
      mbar_init(&vfree[s], (uint32_t)p.GS);
    }
    fence_barrier_init();
  }
  pdl_trigger();   // (common.cuh: programmatic dependent launch)
  if (warp == 8) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32(smem_u32(tmem_slot));
  pdl_wait();

  auto item_sh = [&](int k, int& s, int& h) {
    int item = (int)blockIdx.x + k * (int)gridDim.x;
    if (p.rev) item = total_items - 1 - item;   // last sequences first (L2 reuse of the projection GEMM's tail)
    s = item / p.H;
    h = item - s * p.H;
  };

  if (warp == 10) {
    // ---- loader: K / V blocks through the ring, Q tiles into the warpgroups' slots -----------------------------
    if (elect_one()) {
      int nq[2] = {0, 0};   // Q tiles loaded per warpgroup
      auto load_q = [&](int t) {
        const TileInfo ti = tile_info(p, t);
        const int g = ti.wg;
        int s, h;
        item_sh(ti.k, s, h);
        if (nq[g] > 0) mbar_wait(&qfree[g], (nq[g] - 1) & 1);   // the slot's previous tile has issued (and finished) its S MMAs
        mbar_arrive_expect_tx(&qfull[g], C::Q_BYTES);
        for (int a = 0; a < C::KA; ++a)
          tma_load_3d(Qs + g * C::Q_BYTES + a * (128 * 128), &tmQ, &qfull[g], h * HD + a * 64, ti.qt * 128, s);
        nq[g]++;
      };
      int t = 0;
      while (t < ntiles) {
        const TileInfo t0 = tile_info(p, t);
        int s, h;
        item_sh(t0.k, s, h);
        const int ks = s ^ p.kv_xor;
        // K runs a block ahead of V: the ring's previous occupant releases a V block two key blocks after its K block,
        // so waiting for V(kb) in front of K(kb + 1) would hold that K load back for no reason
        RingPos rk = ring_pos<C::NSLOT>(t0.grp * NB), rv = rk;
        for (int i = 0; i < NB + 1; ++i) {
          if (i < NB) {
            if (rk.ph > 0) mbar_wait(&kfree[rk.slot], (rk.ph - 1) & 1);
            mbar_arrive_expect_tx(&kfull[rk.slot], C::KB_BYTES);
            for (int a = 0; a < C::KA; ++a)
              tma_load_3d(Ks + rk.slot * C::KB_BYTES + a * (64 * 128), &tmK, &kfull[rk.slot], h * HD + a * 64, i * 64, ks);
            ring_next<C::NSLOT>(rk);
            if (i == 0) {   // Q of the group's first two tiles (different warpgroups, their slots free independently)
              load_q(t);
              if (t0.gsize > 1) load_q(t + 1);
            }
          }
          const int kv = NB == 1 ? i : i - 1;   // (a single block: V right behind K)
          if (kv >= 0 && kv < NB && (NB > 1 || i < NB)) {
            if (rv.ph > 0) mbar_wait(&vfree[rv.slot], (rv.ph - 1) & 1);
            mbar_arrive_expect_tx(&vfull[rv.slot], C::VB_BYTES);
            tma_load_3d(Vs + rv.slot * C::VB_BYTES, &tmV, &vfull[rv.slot], kv * 64, 0, ks * p.H + h);
            ring_next<C::NSLOT>(rv);
          }
        }
        // later tiles of the group reuse a warpgroup's slot: they wait for an earlier tile of this same group, which
        // needs every block above, so they come last
        for (int i = 2; i < t0.gsize; ++i) load_q(t + i);
        t += t0.gsize;
        // pull the next group's operands towards L2 (HBM latency would otherwise sit in front of its first block)
        if (t < ntiles) {
          const TileInfo t1 = tile_info(p, t);
          int s1, h1;
          item_sh(t1.k, s1, h1);
          const int ks1 = s1 ^ p.kv_xor;
          if (t1.qt == 0) {
            const int nb_pf = NB < 2 * C::NSLOT ? NB : 2 * C::NSLOT;
            for (int kb = 0; kb < nb_pf; ++kb) {
              for (int a = 0; a < C::KA; ++a) tma_prefetch_3d(&tmK, h1 * HD + a * 64, kb * 64, ks1);
              tma_prefetch_3d(&tmV, kb * 64, 0, ks1 * p.H + h1);
            }
          }
          for (int i = 0; i < t1.gsize; ++i)
            for (int a = 0; a < C::KA; ++a) tma_prefetch_3d(&tmQ, h1 * HD + a * 64, (t1.qt + i) * 128, s1);
        }
      }
    }
  } else if (warp == 8 || warp == 9) {
    // ---- MMA issuers, one per warpgroup -----------------------------------------------------------------------
    // Issue order per tile: S(0), then for every block kb: S(kb + 1) (as soon as S(kb) sits in the softmax warps'
    // registers), P V(kb) (when P(kb) is published).
    const int g = warp - 8;
    if (elect_one()) {
      const uint32_t ids = umma_idesc_bf16(128, 64);
      const uint32_t ido = umma_idesc_bf16(128, HD);
      const uint64_t dq0 = umma_desc_k_sw128(smem_u32(Qs) + g * C::Q_BYTES);
      const uint64_t dk0 = umma_desc_k_sw128(smem_u32(Ks));
      const uint64_t dv0 = umma_desc_k_sw128(smem_u32(Vs));
      const uint32_t scol = tmem_base + C::SP_COL + (uint32_t)g * 128u;
      const uint32_t ocol = tmem_base + C::O_COL + (uint32_t)g * HD;
      int m0 = 0;   // blocks of this warpgroup before the current tile
      int nt = 0;   // tiles of this warpgroup so far
      for (int t = 0; t < ntiles; ++t) {
        const TileInfo ti = tile_info(p, t);
        if (ti.wg != g) continue;
        // blocks this thread has already seen land (an earlier tile of the same resident group): no barrier query
        const bool seen = (p.GS == p.QT) && ti.pos >= 2;
        RingPos rs = ring_pos<C::NSLOT>(ti.grp * NB);   // ring position of the next S block / next P V block
        RingPos rv = rs;
        auto issue_s = [&](int kb) {
          const int m = m0 + kb;
          TR(warp, m, 0);
          if (!seen) mbar_wait(&kfull[rs.slot], rs.ph & 1);
          tc_fence_after();
          TR(warp, m, 1);
          uint64_t da = dq0;
          uint64_t db = dk0 + (uint64_t)launder(((uint32_t)rs.slot * C::KB_BYTES) >> 4);
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {   // 16 elements = 32 B per step, next 64-wide atom after four steps
            umma_bf16_ss(scol + (uint32_t)(m & 1) * 64u, da, db, ids, kk > 0 ? 1u : 0u);
            const bool wrap = (kk & 3) == 3;
            da += wrap ? (uint64_t)(128 * 128 / 16 - 6) : 2u;
            db += wrap ? (uint64_t)(64 * 128 / 16 - 6) : 2u;
          }
          umma_commit(&sfull[g * 2 + (m & 1)]);
          if (kb == NB - 1) umma_commit(&qfree[g]);
          ring_next<C::NSLOT>(rs);
          TR(warp, m, 2);
        };
        mbar_wait(&qfull[g], nt & 1);
        issue_s(0);
        if (NB > 1) issue_s(1);
        for (int kb = 0; kb < NB; ++kb) {
          const int m = m0 + kb, b = m & 1, u = m >> 1;
          TR(warp, m, 3);
          if (!seen) mbar_wait(&vfull[rv.slot], rv.ph & 1);
          if (kb == 0 && nt > 0) mbar_wait(&ofree[g], (nt - 1) & 1);   // the previous tile's O has been read out
          mbar_wait(&pfull[g * 2 + b], u & 1);
          tc_fence_after();
          TR(warp, m, 5);
          uint32_t pa = scol + (uint32_t)b * 64u;
          uint64_t db = dv0 + (uint64_t)launder(((uint32_t)rv.slot * C::VB_BYTES) >> 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {   // 16 keys per step
            umma_bf16_ts(ocol, pa, db, ido, (kb | i) != 0 ? 1u : 0u);
            pa += 8u;
            db += 2u;
          }
          // the V blocks of the tile's last two key blocks are released from here (the softmax warps see the
          // completion of P V(kb) through S(kb + 2), which these blocks do not have)
          if (kb + 2 >= NB)
            for (int r = 0; r <= p.GS - ti.gsize; ++r) umma_commit(&vfree[rv.slot]);
          if (kb == NB - 1) umma_commit(&ofull[g]);
          ring_next<C::NSLOT>(rv);
          TR(warp, m, 6);
          if (kb + 2 < NB) issue_s(kb + 2);                      // into the buffer whose P this P V has just consumed
          else if (kb + 2 == NB) umma_commit(&flush[g]);   // covers P V(NB - 2) for the rescale path of the last block
        }
        m0 += NB;
        nt++;
      }
    }
  } else if (warp < 8) {
    // ---- softmax warpgroups: thread <-> query row of the warpgroup's tile -----------------------------------
    // Measured on B200 (tools/micro/tmem_shapes.cu, xu_rates.cu): pulling a 32-lane x 64-column fp32 block out of TMEM
    // costs the issuing warp ~300 cycles whatever the shape and does not overlap with its own arithmetic, one warp
    // alone can saturate the SM sub-partition's MUFU (8 cycles per ex2 instruction), mbarrier.try_wait on a completed
    // phase costs ~90 cycles. So a block is: wait for S (complete long ago) - load - exponentials - publish, and the
    // overlap comes from the OTHER warpgroup's warp on the same sub-partition, whose tile is in another phase.
    const int g = warp >> 2, q = warp & 3;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t spcol = trow + C::SP_COL + (uint32_t)g * 128u;
    const uint32_t ocol = trow + C::O_COL + (uint32_t)g * HD;
    const float c = p.scale_log2;
    int m0 = 0, nt = 0;   // blocks / tiles of this warpgroup so far
    // epilogue of the previous tile, deferred behind the first block of the next one (its last P V is long complete
    // by then, so nobody waits for the tensor pipe): 1 / row sum, output row pointer, V slots still to release
    float ep_inv = 0.f;
    __nv_bfloat16* ep_dst = nullptr;
    bool ep_pending = false, ep_valid = false;
    auto epilogue = [&]() {
      TR(warp, m0, 5);
      mbar_wait(&ofull[g], (nt - 1) & 1);
      tc_fence_after();
      TR(warp, m0, 6);
      if (ep_valid) {
#pragma unroll
        for (int cc = 0; cc < HD / 32; ++cc) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(ocol + (uint32_t)(cc * 32), r);
          tmem_ld_wait();
          uint4 o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            o[i].x = pack_bf16x2(__uint_as_float(r[8 * i]) * ep_inv, __uint_as_float(r[8 * i + 1]) * ep_inv);
            o[i].y = pack_bf16x2(__uint_as_float(r[8 * i + 2]) * ep_inv, __uint_as_float(r[8 * i + 3]) * ep_inv);
            o[i].z = pack_bf16x2(__uint_as_float(r[8 * i + 4]) * ep_inv, __uint_as_float(r[8 * i + 5]) * ep_inv);
            o[i].w = pack_bf16x2(__uint_as_float(r[8 * i + 6]) * ep_inv, __uint_as_float(r[8 * i + 7]) * ep_inv);
          }
          if (ep_dst != nullptr) {
            if (p.wide_store) {
              stg256(ep_dst + cc * 32, o[0], o[1]);
              stg256(ep_dst + cc * 32 + 16, o[2], o[3]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(ep_dst + cc * 32 + 8 * i) = o[i];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ofree[g]);     // the next tile's first P V may overwrite O_g
      ep_pending = false;
      TR(warp, m0, 7);
    };

    for (int t = 0; t < ntiles; ++t) {
      const TileInfo ti = tile_info(p, t);
      if (ti.wg != g) continue;
      const int row = ti.qt * 128 + q * 32 + lane;          // query row inside the sequence
      const bool rows_valid = ti.qt * 128 + q * 32 < p.T;   // warp-uniform: row groups past the sequence end idle
      const int narr = 1 + p.GS - ti.gsize;                 // ring arrivals owed (missing partners of a short group)
      RingPos rk = ring_pos<C::NSLOT>(ti.grp * NB);         // key block whose S completion is observed next
      RingPos rv = rk;                                      // V block whose P V completion is observed next
      float m_used = 0.f;    // the zero-attention key contributes logit 0: the maximum never drops below it
      float lsum = 0.f;
      const unsigned long long* kmrow = nullptr;   // key-padding words of the sequence the keys come from
      if constexpr (MASKED) {
        int s_, h_;
        item_sh(ti.k, s_, h_);
        kmrow = p.kmask + (size_t)(s_ ^ p.kv_xor) * NB;
      }
      for (int kb = 0; kb < NB; ++kb) {
        const int m = m0 + kb, b = m & 1;
        TR(warp, m, 0);
        // S(kb) has completed: its key block is free; it was issued behind P V(kb - 2), so that V block is free too
        mbar_wait(&sfull[g * 2 + b], (m >> 1) & 1);
        tc_fence_after();
        if (kb == 0 && g == 1 && p.stagger > 0) {
          // The two warpgroups start in lock step (their first key block lands at the same moment) and, with equal work
          // per block, stay there: both load from TMEM together, then both queue for the MUFU together. Half a block of
          // delay, once, puts one warpgroup's exponentials under the other's TMEM reads and barrier traffic.
          const long long t0 = clock64();
          while (clock64() - t0 < p.stagger) {}
        }
        if (lane == 0 && q == 0) {
          for (int r = 0; r < narr; ++r) mbar_arrive(&kfree[rk.slot]);
          if (kb >= 2)
            for (int r = 0; r < narr; ++r) mbar_arrive(&vfree[rv.slot]);
        }
        ring_next<C::NSLOT>(rk);
        if (kb >= 2) ring_next<C::NSLOT>(rv);
        TR(warp, m, 1);
        if (rows_valid) {
          const uint32_t sbuf = spcol + (uint32_t)b * 64u;   // S(kb); P(kb) goes over its first 32 columns
          const uint32_t pcol = sbuf;
          // attendable keys of the block: keys 0..T-1, key T = the zero-attention key, nothing beyond. MASKED: minus
          // the caller's padding mask, as a bit word (the zero key stays attendable: torch pads the mask with False)
          const int key0 = kb * 64;
          unsigned long long kw = 0ull;
          bool full;
          if constexpr (MASKED) {
            kw = __ldg(kmrow + kb);
            full = kw == ~0ull;
          } else {
            full = key0 + 64 <= p.T + 1;
          }
          auto attendable = [&](int j) -> bool {   // j: key inside the block (compile-time after unrolling)
            if constexpr (MASKED) return ((kw >> j) & 1ull) != 0ull;
            else return key0 + j <= p.T;
          };
          uint32_t sv[64];
          tmem_ld_32x32b_x32(sbuf, sv);
          tmem_ld_32x32b_x32(sbuf + 32u, sv + 32);
          tmem_ld_wait();
          TR(warp, m, 2);
          // running maximum of the row from this block's logits (keys beyond T are zero rows of K: logit 0, harmless
          // below a maximum that is >= 0)
          auto raise_max = [&]() -> float {
            float m4[4] = {m_used, 0.f, 0.f, 0.f};
            if (!MASKED || full) {   // (keys beyond T are zero rows of K: logit 0, harmless below a maximum >= 0)
#pragma unroll
              for (int i = 0; i < 32; ++i)
                m4[i & 3] = fmaxf(m4[i & 3], fmaxf(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])));
            } else {   // masked keys must not set the reference: everything attendable could underflow below them
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float a0 = attendable(2 * i) ? __uint_as_float(sv[2 * i]) : 0.f;
                const float a1 = attendable(2 * i + 1) ? __uint_as_float(sv[2 * i + 1]) : 0.f;
                m4[i & 3] = fmaxf(m4[i & 3], fmaxf(a0, a1));
              }
            }
            const float bm = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            const float f = ex2_approx((m_used - bm) * c);
            m_used = bm;
            return f;
          };
          uint32_t pk[32];
          float bsum;
          auto exps = [&]() {   // pk <- packed bf16 probabilities relative to the running maximum, bsum <- their sum
            const float moff = m_used * c;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (full) {
#pragma unroll
                for (int i = 8 * j; i < 8 * j + 8; ++i) {
                  const float e0 = ex2_approx(fmaf(__uint_as_float(sv[2 * i]), c, -moff));
                  const float e1 = ex2_approx(fmaf(__uint_as_float(sv[2 * i + 1]), c, -moff));
                  s0 += e0;
                  s1 += e1;
                  pk[i] = pack_bf16x2(e0, e1);
                }
              } else if (MASKED ? (((uint32_t)(kw >> (16 * j)) & 0xffffu) != 0u) : (key0 + 16 * j <= p.T)) {
                // some keys of the 16-key group attendable
#pragma unroll
                for (int i = 8 * j; i < 8 * j + 8; ++i) {
                  float e0 = ex2_approx(fmaf(__uint_as_float(sv[2 * i]), c, -moff));
                  float e1 = ex2_approx(fmaf(__uint_as_float(sv[2 * i + 1]), c, -moff));
                  e0 = attendable(2 * i) ? e0 : 0.f;
                  e1 = attendable(2 * i + 1) ? e1 : 0.f;
                  s0 += e0;
                  s1 += e1;
                  pk[i] = pack_bf16x2(e0, e1);
                }
              } else {   // a 16-key group without attendable keys (past T, or padding): no MUFU work
#pragma unroll
                for (int i = 8 * j; i < 8 * j + 8; ++i) pk[i] = 0u;
              }
            }
            bsum = s0 + s1;
          };
          // The first block sets the running maximum (nothing is accumulated yet, so adopting it is free). Later blocks
          // take no maximum pass: exp2 has 127 binades of headroom, a block that really outgrows the reference is
          // caught by its sum.
          if (kb == 0) raise_max();
          exps();
          // rare: the block outgrew the reference by more than 2^40 (or overflowed). Take its true maximum, rescale
          // this row's O and sum in place and redo the exponentials. P V(kb - 1) is complete once S(kb + 1), issued
          // behind it, is (the flush commit for the last block); nothing else touches O_g until P(kb) is published.
          const bool grow = kb > 0 && !(bsum < 1.0995116e12f);
          if (__any_sync(0xffffffffu, grow)) {
            if (kb + 1 == NB) mbar_wait(&flush[g], nt & 1);
            else mbar_wait(&sfull[g * 2 + (b ^ 1)], ((m + 1) >> 1) & 1);
            tc_fence_after();
            float f = 1.0f;
            if (grow) {   // (register arithmetic only: the warp-collective TMEM accesses stay outside the branch)
              f = raise_max();
              exps();
            }
#pragma unroll 1
            for (int cc = 0; cc < HD / 16; ++cc) {
              uint32_t o[16];
              tmem_ld_32x32b_x16(ocol + (uint32_t)(cc * 16), o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
              tmem_st_32x32b_x16(ocol + (uint32_t)(cc * 16), o);
            }
            tmem_st_wait();
            lsum *= f;
          }
          lsum += bsum;
          TR(warp, m, 3);
          tmem_st_32x32b_x16(pcol, pk);
          tmem_st_32x32b_x16(pcol + 16u, pk + 16);
          tmem_st_wait();
        }   // (row groups past the sequence end only keep the barrier counts)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pfull[g * 2 + b]);
        TR(warp, m, 4);
        if (kb == 0 && ep_pending) epilogue();   // the previous tile's output, behind this tile's first block
      }
      // hand the tile's epilogue to the next tile of this warpgroup (or to the code below)
      ep_inv = 1.0f / lsum;
      ep_valid = rows_valid;
      {
        int s, h;
        item_sh(ti.k, s, h);
        ep_dst = (rows_valid && row < p.T) ? p.out + ((size_t)s * p.T + row) * p.ldo + h * HD : nullptr;
      }
      ep_pending = true;
      m0 += NB;
      nt++;
    }
    if (ep_pending) epilogue();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem_base);
}

template <int HD>
static int launch_attention_t(const unsigned long long* kmask, const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk,
                              const __nv_bfloat16* vt, int Tp, __nv_bfloat16* out, int ldo, int S, int T, int H,
                              int kv_xor, int rev, cudaStream_t stream) {
  using C = AttnCfg<HD>;
  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)T, (uint64_t)S};
    uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)T * ldq * 2};
    uint32_t box[3] = {64, 128, 1};
    MMDM_TRY(make_tensor_map(&tmQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)T, (uint64_t)S};
    uint64_t str[2] = {(uint64_t)ldk * 2, (uint64_t)T * ldk * 2};
    uint32_t box[3] = {64, 64, 1};
    MMDM_TRY(make_tensor_map(&tmK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    uint64_t dims[3] = {(uint64_t)T, (uint64_t)HD, (uint64_t)S * H};
    uint64_t str[2] = {(uint64_t)Tp * 2, (uint64_t)HD * Tp * 2};
    uint32_t box[3] = {64, (uint32_t)HD, 1};
    MMDM_TRY(make_tensor_map(&tmV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, vt, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  AttnParams p;
  p.S = S; p.T = T; p.H = H;
  p.NB = (T + 1 + 63) / 64;
  p.QT = (T + 127) / 128;
  p.GS = p.NB <= C::NSLOT ? p.QT : 2;   // whole item resident in the ring, or pairs of tiles streaming through it
  p.kv_xor = kv_xor;
  p.rev = rev;
  p.out = out;
  p.kmask = kmask;
  p.ldo = ldo;
  p.wide_store = ((reinterpret_cast<uintptr_t>(out) & 31) == 0 && (ldo % 16) == 0) ? 1 : 0;
  {
    static int stagger = -1;
    if (stagger < 0) {
      const char* e = getenv("MMDM_ATTN_STAGGER");   // A/B measurements
      stagger = e ? atoi(e) : 0;   // measured: no effect (94.3 - 95.5 us for 0 .. 1400 cycles), off by default
    }
    p.stagger = stagger;
  }
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  p.tau_raw = 8.0f / p.scale_log2;
  const int items = S * H;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  if (kmask != nullptr)
    MMDM_CHECK_CUDA(launch_maybe_pdl(attention_kernel<HD, true>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, stream, tmQ, tmK, tmV, p));
  else
    MMDM_CHECK_CUDA(launch_maybe_pdl(attention_kernel<HD, false>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, stream, tmQ, tmK, tmV, p));
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#ifdef MMDM_ATTN_TRACE
extern "C" int mmdm_debug_attn_trace(long long* dev_ptr) {
  return cudaMemcpyToSymbol(g_attn_trace, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? 0 : -2;
}
#endif

int attention_init() {   // kernel attributes are per device
  static PerDeviceFlag attr;
  if (!attr.first()) return 0;
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)AttnCfg<128>::SMEM_BYTES));
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)AttnCfg<64>::SMEM_BYTES));
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)AttnCfg<128>::SMEM_BYTES));
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)AttnCfg<64>::SMEM_BYTES));
  return 0;
}

// one thread per (sequence, word): 64 mask values -> bits; bit T = zero-attention key, nothing beyond
__global__ void build_key_mask_kernel(const float* __restrict__ mask, int rows, int nseq, int T, int persons, int NB,
                                      unsigned long long* __restrict__ words) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseq * NB) return;
  const int s = i / NB, kb = i - s * NB;
  const float* row = mask + (size_t)((s / persons) % rows) * T;
  unsigned long long w = 0ull;
  for (int j = 0; j < 64; ++j) {
    const int key = kb * 64 + j;
    const bool on = key < T ? row[key] > 0.5f : key == T;
    w |= (unsigned long long)on << j;
  }
  words[i] = w;
}

int build_key_mask(const float* mask, int rows, int nseq, int T, int persons, unsigned long long* words,
                   cudaStream_t stream) {
  MMDM_REQUIRE(mask && words && rows >= 1 && nseq >= 1 && T >= 1 && persons >= 1, "build_key_mask: bad argument");
  const int NB = (T + 1 + 63) / 64, n = nseq * NB;
  build_key_mask_kernel<<<(n + 127) / 128, 128, 0, stream>>>(mask, rows, nseq, T, persons, NB, words);
  MMDM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_attention(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* vt,
                     int Tp, __nv_bfloat16* out, int ldo, int S, int T, int H, int hd, int kv_xor,
                     cudaStream_t stream, int rev, const unsigned long long* key_mask) {
  MMDM_REQUIRE(attention_init() == 0, "attention: kernel attribute setup failed");
  MMDM_REQUIRE(T >= 1 && S >= 1 && H >= 1, "attention: empty problem S=%d T=%d H=%d", S, T, H);
  MMDM_REQUIRE(hd == 64 || hd == 128, "attention: head dim %d not in {64,128}", hd);
  MMDM_REQUIRE((ldq % 8) == 0 && (ldk % 8) == 0 && (Tp % 8) == 0 && (ldo % 8) == 0 && Tp >= T,
               "attention: pitches must be multiples of 8 elements");
  if (hd == 128) return launch_attention_t<128>(key_mask, q, ldq, k, ldk, vt, Tp, out, ldo, S, T, H, kv_xor, rev, stream);
  return launch_attention_t<64>(key_mask, q, ldq, k, ldk, vt, Tp, out, ldo, S, T, H, kv_xor, rev, stream);
}

}  // namespace mmdm
