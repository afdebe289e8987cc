// Fused attention for short motion sequences (T <= 303) on sm_100a, with the reference's zero-attention
// sink: nn.MultiheadAttention(add_zero_attn=True) appends one key with logit 0 and value 0
// (reference: models/utils/layers.py:33-34,74-75; torch F.multi_head_attention_forward).
//
// The whole key range fits one tile, so there is no online-softmax rescaling:
//   S[128 x NK] = Q K^T        tcgen05.mma, fp32 in TMEM (NK = round_up(T+1,16), issued as two key halves)
//   P = exp2((S - max) * c)    8 softmax warps: thread <-> (query row, key half)
//   O[128 x hd] = P V          tcgen05.mma with the A operand (P, packed bf16) read from TMEM,
//                              V consumed K-major from the transposed V^T buffer the QKV GEMM writes.
// Key row T is zero-filled by TMA out-of-bounds handling, which *is* the zero-attention key; key columns
// > T are masked. Rows/keys beyond the sequence never touch HBM.
#include "attention.cuh"

namespace mmdm {

struct AttnParams {
  int S, T, H, NK, n1, n2, kv_xor, rev;
  float scale_log2;
};

template <int HD>
struct AttnCfg {
  static constexpr int KA = HD / 64;                       // 64-wide K atoms along the head dim
  static constexpr uint32_t Q_BYTES = KA * 128 * 128;      // one Q tile  [KA][128 rows][128 B]
  static constexpr int NQ = HD == 128 ? 1 : 2;             // Q slots (hd=128 fills the SM: one slot + the O staging)
  static constexpr int OC = HD / 2;                        // output columns per softmax thread (row half)
  static constexpr uint32_t OROW_BYTES = OC * 2;           // one staged output row per thread: 128 B / 64 B
  static constexpr uint32_t O_BYTES = 8 * 32 * OROW_BYTES; // O staging: per softmax warp [32 rows][OROW_BYTES], swizzled
  static constexpr uint32_t K_BYTES = KA * 320 * 128;      // keys        [KA][320 rows][128 B]
  static constexpr uint32_t V_BYTES = 5 * HD * 128;        // V^T         [5][HD rows][128 B]
  static constexpr uint32_t XCH_BYTES = 2 * 2 * 128 * 4;   // row max / row sum exchange between the two key halves
  static constexpr uint32_t USED_BYTES = NQ * Q_BYTES + O_BYTES + K_BYTES + V_BYTES + XCH_BYTES + 256;
  // hd=128 fills the SM: the request is the 227 KB maximum and the 1024-byte alignment slack is whatever is left
  // (768 B; the kernel traps if the dynamic shared window is less aligned than that, it is 1024-aligned in practice)
  static constexpr size_t SMEM_BYTES = (USED_BYTES + 1024 > 232448) ? 232448 : USED_BYTES + 1024;
  // TMEM columns (512): S [0,304) | O [304,304+HD) | P of the lower key half (80 columns, 2 keys per column).
  // hd=128: 304 + 128 + 80 = 512, which is where the T <= 303 limit comes from. P of the upper key half (72
  // columns) is written over the logits of keys [224,304): S is therefore issued in two parts, keys [0,224)
  // ("early", free as soon as the softmax warps hold the logits) and keys [224,304) ("late", behind P V).
  static constexpr uint32_t HALF_KEYS = 160;               // softmax warps split the keys at 160
  static constexpr uint32_t EARLY_KEYS = 224;
  static constexpr uint32_t PHI_COL = 224;
  static constexpr uint32_t O_COL = 304;
  static constexpr uint32_t PLO_COL = O_COL + HD;
  static constexpr int THREADS = 384;                      // 8 softmax warps + TMA, MMA, L2-prefetch warps + 1 idle warp
  // (register reallocation is per warpgroup: the third warpgroup shrinks to 72 registers, exactly what lets the
  //  two softmax warpgroups grow from the launch-time 168 to 216: 128*(168-72) = 256*(216-168). An unbalanced pair
  //  deadlocks in setmaxnreg.inc; the bounded mbarrier waits of the other warps then trap the kernel.)
};

#ifdef MMDM_ATTN_TRACE
__device__ long long* g_attn_trace = nullptr;
#define TR(w, j, ev)                                                                      \
  do {                                                                                    \
    if (trace_ptr != nullptr && (j) < 24) trace_ptr[(((w) * 24) + (j)) * 16 + (ev)] = clock64(); \
  } while (0)
#else
#define TR(w, j, ev) do {} while (0)
#endif

// Opaque copy: keeps ptxas from hoisting "base + constant" address arithmetic of every call site out of the tile
// loop, which on the 56-register MMA warp turned into local-memory spills (an L1 miss in front of an MMA issue).
__device__ __forceinline__ uint32_t launder(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

__device__ __forceinline__ void pair_barrier(int id) {
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

// exp2 of 32 logits of one row -> 16 packed bf16 pairs; returns their (unrounded) sum. MASK: keys > T give 0.
template <bool MASK>
__device__ __forceinline__ float exp_chunk(const uint32_t* sv, uint32_t* pk, int n, float scale_log2, float moff,
                                           int key0, int T) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (2 * i < n) {
      float e0 = ex2_approx(fmaf(__uint_as_float(sv[2 * i]), scale_log2, -moff));
      float e1 = ex2_approx(fmaf(__uint_as_float(sv[2 * i + 1]), scale_log2, -moff));
      if (MASK) {   // key T is the zero-attention key, keys > T are padding
        e0 = (key0 + 2 * i <= T) ? e0 : 0.f;
        e1 = (key0 + 2 * i + 1 <= T) ? e1 : 0.f;
      }
      s0 += e0;
      s1 += e1;
      pk[i] = pack_bf16x2(e0, e1);
    }
  }
  return s0 + s1;
}

// Persistent kernel. Work item = (sequence, head); its K / V^T tiles stay in shared memory while the
// ceil(T/128) query tiles stream through the Q slot(s).
//   warp 8 : TMA loads (Q tiles, K, V^T), gated by "free" barriers the MMA warp commits
//   warp 10: L2 prefetch of the next work item
//   warp 9 : tcgen05.mma issue.  S = Q K^T in two parts: keys [0,224) of tile j+1 are issued while the softmax
//            warps still work on tile j (its columns are free once the logits sit in registers); keys [224,NK)
//            share their columns with P(j) and are issued right behind the last P V (j) instruction - tcgen05.mma
//            executes in issue order, so no completion wait is needed. O = P V in six key groups as the softmax
//            warps publish P.
//   warps 0-7 : softmax. Warps w and w+4 share TMEM lane quarter w&3 (32 query rows) and split the keys in
//            halves. Each thread keeps 128 logits in registers; the last chunk of its half (<= 32 keys) is folded
//            into the row max first and read from TMEM a second time for its exponentials (it would not fit the
//            register file). The halves exchange row max / row sum through shared memory. The normalised output
//            is staged in shared memory and leaves through one bulk tensor store per warp.
template <int HD>
__global__ void __launch_bounds__(AttnCfg<HD>::THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                 const AttnParams p) {
  using C = AttnCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;                       // NQ slots
  uint8_t* Os = Qs + C::NQ * C::Q_BYTES;    // output staging, one 32-row slab per softmax warp
  uint8_t* Ks = Os + C::O_BYTES;
  uint8_t* Vs = Ks + C::K_BYTES;
  float* xmax = reinterpret_cast<float*>(Vs + C::V_BYTES);   // [2][128]
  float* xsum = xmax + 2 * 128;                              // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xsum + 2 * 128);
  uint64_t* qfull = bars;         // [2]
  uint64_t* qfree = bars + 2;     // [2]
  uint64_t* kfull = bars + 4;
  uint64_t* kfree = bars + 5;
  uint64_t* vfull = bars + 6;
  uint64_t* vfree = bars + 7;
  uint64_t* sfull = bars + 8;     // [2] early / late S part written
  uint64_t* sfree = bars + 10;    // [2] early / late S part no longer needed in TMEM (8 / 4 warps arrive)
  uint64_t* pfull = bars + 12;    // [6] P key groups per half: chunks {tail,0}, {1,2}, {3} (chunk = 32 keys)
  uint64_t* ofull = bars + 18;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
#ifdef MMDM_ATTN_TRACE
  long long* const trace_ptr = blockIdx.x == 0 ? g_attn_trace : nullptr;
#endif
  const int QT = (p.T + 127) / 128;
  const int total_items = p.S * p.H;
  const int nitems = (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int njobs = nitems * QT;
  const int kgroups = (p.NK + 63) / 64;

  if (tid == 0) {
    if ((size_t)(smem - smem_raw) + C::USED_BYTES > C::SMEM_BYTES) {
      printf("mmdm attention: dynamic shared memory window is not aligned enough\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qfull[i], 1);
      mbar_init(&qfree[i], 1);
      mbar_init(&sfull[i], 1);
      mbar_init(&sfree[i], i == 0 ? 8 : 4);
    }
    mbar_init(kfull, 1);
    mbar_init(kfree, 1);
    mbar_init(vfull, 1);
    mbar_init(vfree, 1);
    for (int i = 0; i < 6; ++i) mbar_init(&pfull[i], 4);
    mbar_init(ofull, 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32(smem_u32(tmem_slot));

  auto item_sh = [&](int k, int& s, int& h) {
    int item = (int)blockIdx.x + k * (int)gridDim.x;
    if (p.rev) item = total_items - 1 - item;   // last sequences first (L2 reuse of the projection GEMM's tail)
    s = item / p.H;
    h = item - s * p.H;
  };

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 8) {
    // ---- TMA producer --------------------------------------------------------------------------------
    if (elect_one()) {
      for (int j = 0; j < njobs; ++j) {
        const int k = j / QT, qt = j - k * QT;
        int s, h;
        item_sh(k, s, h);
        const int ks = s ^ p.kv_xor;
        if (qt == 0) {
          if (k > 0) mbar_wait(kfree, (k - 1) & 1);
          TR(8, j, 0);
          mbar_arrive_expect_tx(kfull, (uint32_t)(C::KA * kgroups) * (64 * 128));
          for (int a = 0; a < C::KA; ++a)
            for (int r = 0; r < kgroups; ++r)
              tma_load_3d(Ks + a * (320 * 128) + r * (64 * 128), &tmK, kfull, h * HD + a * 64, r * 64, ks);
        }
        const int b = j % C::NQ;
        if (j >= C::NQ) mbar_wait(&qfree[b], ((j / C::NQ) - 1) & 1);
        TR(8, j, 1);
        mbar_arrive_expect_tx(&qfull[b], C::Q_BYTES);
        for (int a = 0; a < C::KA; ++a)
          tma_load_3d(Qs + b * C::Q_BYTES + a * (128 * 128), &tmQ, &qfull[b], h * HD + a * 64, qt * 128, s);
        if (qt == 0) {   // V^T last: it is needed only once the first P group of the item's first tile exists
          if (k > 0) mbar_wait(vfree, (k - 1) & 1);
          TR(8, j, 2);
          mbar_arrive_expect_tx(vfull, (uint32_t)kgroups * (HD * 128));
          for (int kt = 0; kt < kgroups; ++kt)
            tma_load_3d(Vs + kt * (HD * 128), &tmV, vfull, kt * 64, 0, ks * p.H + h);
        }
      }
    }
    } else if (warp == 10) {
    // ---- L2 prefetcher -------------------------------------------------------------------------------
    // The shared-memory slots turn over only when their consumer is done, so the TMA loads start late and their
    // latency (HBM: the projections of a 38400-token batch do not fit L2) would be exposed. When item k's keys have
    // landed, pull all of item k+1 (Q, K, V^T: ~230 KB) into L2; issuing the prefetches costs ~100 cycles each,
    // which is why they are kept off the TMA warp.
    if (elect_one()) {
      for (int k = 0; k + 1 < nitems; ++k) {
        mbar_wait(kfull, k & 1);
        int s1, h1;
        item_sh(k + 1, s1, h1);
        const int ks1 = s1 ^ p.kv_xor;
        for (int a = 0; a < C::KA; ++a)
          for (int r = 0; r < kgroups; ++r) tma_prefetch_3d(&tmK, h1 * HD + a * 64, r * 64, ks1);
        for (int qt = 0; qt < QT; ++qt)
          for (int a = 0; a < C::KA; ++a) tma_prefetch_3d(&tmQ, h1 * HD + a * 64, qt * 128, s1);
        for (int kt = 0; kt < kgroups; ++kt) tma_prefetch_3d(&tmV, kt * 64, 0, ks1 * p.H + h1);
      }
    }
    } else if (warp == 9) {
    // ---- MMA issuer ----------------------------------------------------------------------------------
    if (elect_one()) {
      const uint32_t idL = umma_idesc_bf16(128, p.n1);   // n1 = keys in the early part, n2 = in the late part
      const uint32_t idU = umma_idesc_bf16(128, p.n2 > 0 ? p.n2 : 16);
      const uint32_t ido = umma_idesc_bf16(128, HD);
      // This warp runs on 56 registers: descriptors are rebuilt from three base values inside rolled loops (a
      // descriptor differs from the base only by the 16-byte-granular start address in its low word) instead of
      // being hoisted per k-step, which spilled to local memory and put an L1-miss latency in front of every MMA.
      const uint64_t dq0 = umma_desc_k_sw128(smem_u32(Qs));
      const uint64_t dk0 = umma_desc_k_sw128(smem_u32(Ks));
      const uint64_t dv0 = umma_desc_k_sw128(smem_u32(Vs));
      const int ksteps = p.NK / 16;
      auto issue_s = [&](int j, int late) {
        const uint32_t dcol = launder(tmem_base) + (late ? C::EARLY_KEYS : 0u);
        const uint32_t idesc = late ? idU : idL;
        uint64_t da = dq0 + (uint64_t)launder(((uint32_t)(j % C::NQ) * C::Q_BYTES) >> 4);
        uint64_t db = dk0 + (uint64_t)launder((late ? C::EARLY_KEYS * 128u : 0u) >> 4);
#pragma unroll 1
        for (int kk = 0; kk < HD / 16; ++kk) {   // 16 elements = 32 B per step, next 64-wide atom after four steps
          umma_bf16_ss(dcol, da, db, idesc, kk > 0 ? 1u : 0u);
          const bool wrap = (kk & 3) == 3;
          da += wrap ? (uint64_t)(128 * 128 / 16 - 6) : 2u;
          db += wrap ? (uint64_t)(320 * 128 / 16 - 6) : 2u;
        }
      };
      // S(j) keys [0,224): needs the operands and the softmax warps done with those columns of S(j-1)
      auto early_ready = [&](int j) -> bool {
        const int k = j / QT, qt = j - k * QT;
        if (qt == 0 && !mbar_try_wait(kfull, k & 1)) return false;
        if (!mbar_try_wait(&qfull[j % C::NQ], (j / C::NQ) & 1)) return false;
        return j == 0 || mbar_try_wait(&sfree[0], (j - 1) & 1);
      };
      auto issue_early = [&](int j) {
        const int k = j / QT, qt = j - k * QT;
        TR(9, j, 0);
        if (qt == 0) mbar_wait(kfull, k & 1);
        mbar_wait(&qfull[j % C::NQ], (j / C::NQ) & 1);
        if (j > 0) mbar_wait(&sfree[0], (j - 1) & 1);
        tc_fence_after();
        TR(9, j, 1);
        issue_s(j, 0);
        umma_commit(&sfull[0]);
        TR(9, j, 2);
      };
      // S(j) keys [224,NK): overwrites P(j-1); ordered behind P V (j-1) by the in-order execution of tcgen05.mma
      auto issue_late = [&](int j) {
        const int k = j / QT, qt = j - k * QT;
        if (j > 0) mbar_wait(&sfree[1], (j - 1) & 1);
        tc_fence_after();
        if (p.n2 > 0) issue_s(j, 1);
        umma_commit(&sfull[1]);
        umma_commit(&qfree[j % C::NQ]);
        if (qt == QT - 1) umma_commit(kfree);
        TR(9, j, 9);
      };
      // P V for 16-key steps [i0, i1) of one key half; P lower lives in its own columns, P upper over S late
      auto issue_pv = [&](int half, int i0, int i1, uint32_t& acc) {
        i0 += half * 10;
        i1 = min(i1 + half * 10, ksteps);
        const uint32_t tb = launder(tmem_base);
        uint32_t pa = tb + (half ? C::PHI_COL - 80u : C::PLO_COL) + (uint32_t)(i0 * 8);
        uint64_t db = dv0 + (uint64_t)launder((uint32_t)(i0 >> 2) * (HD * 128 / 16) + (uint32_t)(i0 & 3) * 2);
#pragma unroll 1
        for (int i = i0; i < i1; ++i) {
          umma_bf16_ts(tb + C::O_COL, pa, db, ido, acc);
          acc = 1u;
          pa += 8u;
          db += (i & 3) == 3 ? (uint64_t)(HD * 128 / 16 - 6) : 2u;
        }
      };
      if (njobs > 0) {
        issue_early(0);
        issue_late(0);
      }
      for (int j = 0; j < njobs; ++j) {
        const int k = j / QT, qt = j - k * QT;
        // S(j+1) early part goes into the gap before the first P group of tile j arrives (the tensor pipe is idle
        // there); once P is waiting, P V has priority and S(j+1) early follows at the end of the tile.
        bool early_done = j + 1 >= njobs;
        while (!early_done) {
          if (early_ready(j + 1)) {
            issue_early(j + 1);
            early_done = true;
          } else if (mbar_try_wait(&pfull[0], j & 1) && mbar_try_wait(&pfull[3], j & 1)) {
            break;
          }
        }
        if (qt == 0) mbar_wait(vfull, k & 1);
        uint32_t acc = 0u;
        // the first P V instruction overwrites O: every softmax warp has published its first group by then, i.e.
        // has finished reading O of the previous tile
        mbar_wait(&pfull[0], j & 1);
        mbar_wait(&pfull[3], j & 1);
        tc_fence_after();
        TR(9, j, 3);
        issue_pv(0, 8, 10, acc);   // tail chunk + chunk 0
        issue_pv(0, 0, 2, acc);
        issue_pv(1, 8, 10, acc);
        issue_pv(1, 0, 2, acc);
        if (!early_done && early_ready(j + 1)) {   // missed the gap (operands late): still cheaper here than at the end
          issue_early(j + 1);
          early_done = true;
        }
        mbar_wait(&pfull[1], j & 1);
        mbar_wait(&pfull[4], j & 1);
        tc_fence_after();
        TR(9, j, 4);
        issue_pv(0, 2, 6, acc);    // chunks 1, 2
        issue_pv(1, 2, 6, acc);
        mbar_wait(&pfull[2], j & 1);
        mbar_wait(&pfull[5], j & 1);
        tc_fence_after();
        TR(9, j, 5);
        issue_pv(0, 6, 8, acc);    // chunk 3
        issue_pv(1, 6, 8, acc);
        umma_commit(ofull);
        if (qt == QT - 1) umma_commit(vfree);
        TR(9, j, 6);
        if (j + 1 < njobs) {
          if (!early_done) issue_early(j + 1);
          issue_late(j + 1);
        }
      }
    }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ---- softmax / output warps: thread <-> (query row, key half) --------------------------------------
    const int q = warp & 3, hf = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int kbase = hf * (int)C::HALF_KEYS;
    const uint32_t scol = trow + (uint32_t)kbase;                          // this thread's logits
    const uint32_t pcol = trow + (hf ? C::PHI_COL : C::PLO_COL);           // and where its P goes
    const int keyt = kbase + 128;                                          // first key of the tail chunk
    const int ntail = min(32, p.NK - keyt);                                // <= 0: no tail chunk
    for (int j = 0; j < njobs; ++j) {
      const int k = j / QT, qt = j - k * QT;
      const bool rows_valid = qt * 128 + q * 32 < p.T;   // warp-uniform: tiles past the sequence end are skipped
      const bool tail = rows_valid && ntail > 0;
      if (lane == 0) TR(warp, j, 0);
      mbar_wait(&sfull[0], j & 1);
      if (hf) mbar_wait(&sfull[1], j & 1);
      tc_fence_after();
      if (lane == 0) TR(warp, j, 1);
      // One register block serves the logits, the tail chunk and the output tile: ptxas gives every 32-register
      // tcgen05.ld tuple its own physical range and does not overlap tuples of different arrays, which spilled.
      uint32_t sv[4][32];
      float mx = 0.f;   // the zero-attention key contributes logit 0
      if (rows_valid) {
        float m4[4] = {0.f, 0.f, 0.f, 0.f};   // independent chains: the dependent fmax latency is not the bound
        if (ntail > 0) {
          uint32_t* t4 = sv[3];   // columns past NK hold other data (O / P); they are masked by key index below
          tmem_ld_32x32b_x32(scol + 128u, t4);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (keyt + i < p.T) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(t4[i]));
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (kbase + c * 32 < p.NK) tmem_ld_32x32b_x32(scol + (uint32_t)(c * 32), sv[c]);
        tmem_ld_wait();
      }
      if (hf) {   // the upper warps' early columns (keys [160,224)) are in registers now
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sfree[0]);
      }
      if (rows_valid) {
        float m4[4] = {mx, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int key0 = kbase + c * 32;
          if (key0 < p.NK) {
            if (key0 + 32 <= p.T) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                m4[i & 3] = fmax3(m4[i & 3], __uint_as_float(sv[c][2 * i]), __uint_as_float(sv[c][2 * i + 1]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (key0 + i < p.T) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(sv[c][i]));
            }
          }
        }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
      if (lane == 0) TR(warp, j, 2);
      xmax[hf * 128 + row] = mx;
      pair_barrier(1 + q);
      mx = fmaxf(mx, xmax[(hf ^ 1) * 128 + row]);
      const float moff = mx * p.scale_log2;
      float lsum = 0.f;
      // tail chunk first: it is the only part of S(j) still needed in TMEM, and S(j+1) waits for it
      uint32_t t4[32];
      if (tail) {
        tmem_ld_32x32b_x32(scol + 128u, t4);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sfree[hf]);   // lower warps: early part (with the upper warps' arrive above)
      if (lane == 0) TR(warp, j, 3);
      if (tail) {
        uint32_t pk[16];
        lsum += exp_chunk<true>(t4, pk, 32, p.scale_log2, moff, keyt, p.T);
        tmem_st_32x32b_x8(pcol + 64u, pk);
        if (ntail > 16) tmem_st_32x32b_x8(pcol + 72u, pk + 8);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int key0 = kbase + c * 32;
        if (rows_valid && key0 < p.NK) {
          uint32_t pk[16];
          if (key0 + 32 <= p.T) lsum += exp_chunk<false>(sv[c], pk, 32, p.scale_log2, moff, key0, p.T);
          else lsum += exp_chunk<true>(sv[c], pk, 32, p.scale_log2, moff, key0, p.T);
          tmem_st_32x32b_x16(pcol + (uint32_t)(c * 16), pk);
        }
        if (c != 1) {   // groups {tail, 0}, {1, 2}, {3}
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&pfull[hf * 3 + (c + 1) / 2]);
          if (lane == 0) TR(warp, j, 4 + (c + 1) / 2);
        }
      }
      xsum[hf * 128 + row] = lsum;
      pair_barrier(1 + q);
      const float inv = 1.0f / (lsum + xsum[(hf ^ 1) * 128 + row]);

      if (lane == 0) TR(warp, j, 10);
      mbar_wait(ofull, j & 1);
      tc_fence_after();
      if (lane == 0) TR(warp, j, 7);
      if (rows_valid) {
        constexpr int OC = C::OC;
        uint32_t* r = &sv[0][0];
#pragma unroll
        for (int c = 0; c < OC / 32; ++c) tmem_ld_32x32b_x32(trow + C::O_COL + (uint32_t)(hf * OC + c * 32), r + c * 32);
        tmem_ld_wait();
        // Stage the warp's 32 x OC bf16 block in shared memory (TMA swizzle pattern, conflict-free 16-byte stores)
        // and hand it to one bulk tensor store: rows >= T are clipped by the tensor map, HBM sees full lines.
        if (lane == 0) bulk_wait_group_read0();   // the previous store of this warp has drained its staging slab
        __syncwarp();
        const uint32_t slab = smem_u32(Os) + (uint32_t)warp * (32 * C::OROW_BYTES);
        const uint32_t rowb = (uint32_t)lane * C::OROW_BYTES;
        const uint32_t sw = (rowb >> 7) & (C::OROW_BYTES / 16 - 1);
#pragma unroll
        for (int i = 0; i < OC / 8; ++i) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(r[8 * i]) * inv, __uint_as_float(r[8 * i + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv);
          sts128u(slab + rowb + (((uint32_t)i ^ sw) << 4), o);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          int s, h;
          item_sh(k, s, h);
          tma_store_3d(&tmO, slab, h * HD + hf * OC, qt * 128 + q * 32, s);
          bulk_commit_group();
        }
      }
      if (lane == 0) TR(warp, j, 8);
      tc_fence_before();   // O reads retire before this warp's next P arrive lets P V (j+1) overwrite O
    }
  }
  if (warp < 8 && lane == 0) bulk_wait_group0();   // output stores have left shared memory before the CTA retires
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

static int num_sms_attn() { return device_sm_count(); }

template <int HD>
static int launch_attention_t(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk,
                              const __nv_bfloat16* vt, int Tp, __nv_bfloat16* out, int ldo, int S, int T, int H,
                              int kv_xor, int rev, cudaStream_t stream) {
  using C = AttnCfg<HD>;
  CUtensorMap tmQ, tmK, tmV, tmO;
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
  {
    uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)T, (uint64_t)S};
    uint64_t str[2] = {(uint64_t)ldo * 2, (uint64_t)T * ldo * 2};
    uint32_t box[3] = {(uint32_t)C::OC, 32, 1};
    MMDM_TRY(make_tensor_map(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, dims, str, box,
                             HD == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B));
  }
  AttnParams p;
  p.S = S; p.T = T; p.H = H;
  p.NK = ((T + 1 + 15) / 16) * 16;
  p.n1 = p.NK < (int)C::EARLY_KEYS ? p.NK : (int)C::EARLY_KEYS;   // keys [0,224) -> early S part, the rest -> late
  p.n2 = p.NK - p.n1;
  p.kv_xor = kv_xor;
  p.rev = rev;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  const int items = S * H;
  const int grid = items < num_sms_attn() ? items : num_sms_attn();
  attention_kernel<HD><<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
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
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)AttnCfg<128>::SMEM_BYTES));
  MMDM_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)AttnCfg<64>::SMEM_BYTES));
  return 0;
}

int launch_attention(const __nv_bfloat16* q, int ldq, const __nv_bfloat16* k, int ldk, const __nv_bfloat16* vt,
                     int Tp, __nv_bfloat16* out, int ldo, int S, int T, int H, int hd, int kv_xor,
                     cudaStream_t stream, int rev) {
  MMDM_REQUIRE(attention_init() == 0, "attention: kernel attribute setup failed");
  MMDM_REQUIRE(T >= 1 && T <= 303, "attention: T=%d outside the single-tile range [1,303]", T);
  MMDM_REQUIRE(hd == 64 || hd == 128, "attention: head dim %d not in {64,128}", hd);
  MMDM_REQUIRE((ldq % 8) == 0 && (ldk % 8) == 0 && (Tp % 8) == 0 && (ldo % 8) == 0 && Tp >= T,
               "attention: pitches must be multiples of 8 elements");
  if (hd == 128) return launch_attention_t<128>(q, ldq, k, ldk, vt, Tp, out, ldo, S, T, H, kv_xor, rev, stream);
  return launch_attention_t<64>(q, ldq, k, ldk, vt, Tp, out, ldo, S, T, H, kv_xor, rev, stream);
}

}  // namespace mmdm
