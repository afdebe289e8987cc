// Shared helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM wrappers (inline PTX),
// UMMA descriptor construction, small math utilities. No CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace mmdm {

// ----------------------------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
#define MMDM_CHECK_CUDA(expr)                                                                     \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      mmdm::set_last_error("%s:%d CUDA error in %s: %s", __FILE__, __LINE__, #expr,               \
                           cudaGetErrorString(_e));                                               \
      return -2;                                                                                  \
    }                                                                                             \
  } while (0)
#define MMDM_REQUIRE(cond, ...)                                                                   \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      mmdm::set_last_error(__VA_ARGS__);                                                          \
      return -1;                                                                                  \
    }                                                                                             \
  } while (0)
#define MMDM_TRY(expr)                                                                            \
  do {                                                                                            \
    int _rc = (expr);                                                                             \
    if (_rc != 0) return _rc;                                                                     \
  } while (0)

// Builds a CUtensorMap (driver entry point resolved at run time, no -lcuda link dependency).
// dims/strides innermost first; strides in bytes for dims 1..rank-1. Returns 0 or negative.
int make_tensor_map(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base,
                    const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                    CUtensorMapSwizzle swizzle);

// Per-device host state. Kernel attributes (cudaFuncSetAttribute) and SM counts belong to a device, not to the
// process: a second GPU driven from the same process needs its own. kMaxDevices bounds the small tables below.
constexpr int kMaxDevices = 64;
int current_device();   // ordinal of the calling thread's current device, clamped to [0, kMaxDevices)
int device_sm_count();  // multiprocessors of the current device (cached per device)
struct PerDeviceFlag {  // "done once per device" latch; first(): true exactly once for the current device
  bool done[kMaxDevices] = {};
  bool first() {
    const int d = current_device();
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
// makes `device` current for the lifetime of the object (no-op when it already is)
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int device) {
    int cur = -1;
    if (device >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != device && cudaSetDevice(device) == cudaSuccess) prev = cur;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Programmatic dependent launch: a kernel launched with it may start - block by block, as the kernel in front of it in
// the stream frees SMs - before that kernel has finished; it runs its prologue (barrier init, TMEM allocation,
// descriptor prefetch) and then waits in pdl_wait() for the predecessor's completion and memory flush. The predecessor
// allows this by calling pdl_trigger() (all blocks, at their start). Opt-in with MMDM_PDL=1; otherwise such kernels are
// launched the ordinary way (both device calls are no-ops then). Used by the CTA-pair GEMM and the attention kernel, which follow each other
// directly (33.91 / 34.00 against 33.86 / 33.86 samples/s in one job); extending it to the single-CTA GEMM and the row
// kernels made the step 3 % SLOWER (32.0 against 33.0): their early-resident blocks sit in the wait on SMs the next
// persistent kernel needs whole.
bool pdl_enabled();
#if defined(__CUDACC__)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_maybe_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                           Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Non-blocking probe. (try_wait is NOT one: it suspends the thread for a system-defined time, ~2000 cycles measured on
// B200, before it reports "not yet".)
__device__ __forceinline__ uint32_t mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a descriptor/protocol bug must trap, never hang the GPU box. ~2-4 s budget.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 6000000000LL) {
      printf("mmdm: mbarrier wait timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

// ---- proxies / fences -------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMA --------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}

// L2 prefetch of a tensor-map box (no shared memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// bulk tensor store shared -> global (bulk async-group completion); out-of-bounds box elements are not written
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_wait_group_read1() {
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------------
// Called by one full warp. Writes the TMEM base address into *slot (shared memory).
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  static_assert(kCols == 32 || kCols == 64 || kCols == 128 || kCols == 256 || kCols == 512, "TMEM cols");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)),
               "r"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread i <-> lane base+i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns <- 16 registers per thread
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- UMMA (tcgen05.mma) -------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle, rows of 64 bf16 (128 B):
//   8-row core groups are 1024 B apart (SBO); LBO unused for swizzled K-major (set to 1 like CUTLASS);
//   version = 1 (Blackwell); layout type 2 = SWIZZLE_128B. Tile base must be 1024-byte aligned;
//   advancing along K by 16 elements inside the 128-byte atom = +32 bytes on the start address.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                           // LBO            [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // SBO            [32,46)
  d |= static_cast<uint64_t>(1) << 46;                           // version        [46,48)
  d |= static_cast<uint64_t>(2) << 61;                           // SWIZZLE_128B   [61,64)
  return d;
}
// Instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, dense.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4)                       // c_format = F32
         | (1u << 7)                     // a_format = BF16
         | (1u << 10)                    // b_format = BF16
         | (0u << 15) | (0u << 16)       // a_major, b_major = K
         | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : A is read from tensor memory (lane = row, 32-bit column = 2 bf16 along K)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}


// ---- CTA pairs (cta_group::2): cluster rank / sync, peer-addressed barriers, 2-SM TMA / MMA / commit -------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// Relaxed on purpose: the default .release at cluster scope compiles to MEMBAR.ALL.GPU + ERRBAR, i.e. the arriving
// warp first waits for all of its global stores to become visible GPU-wide. The only thing the waiter depends on
// are this warp's TMEM reads, which tcgen05.wait::ld + tcgen05.fence::before_thread_sync have already retired.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Release / acquire pair at cluster scope: hands shared-memory contents written by the arriving CTA over to the waiter
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > 6000000000LL) {
      printf("mmdm: cluster mbarrier wait timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
// TMA load into this CTA's shared memory whose completion bytes are signalled on a barrier that may live in the
// peer CTA of the pair (bar_cluster_addr is a shared::cluster address)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)),
               "r"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}
// D[tmem, both CTAs] (+)= A * B^T over the CTA pair: M = 256 (128 rows per CTA), each CTA supplies half of B's rows.
// Issued by ONE thread of the leader CTA (rank 0).
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once all previously issued MMAs of this thread have completed) on the barrier at this shared-memory
// offset in every CTA of cta_mask.
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ---- explicit shared-space vector access (pointers re-derived from an aligned base lose their address
// space and would otherwise compile to generic LD/ST) ------------------------------------------------
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// 32-byte global store (STG.256, sm_100+): one full sector per thread
__device__ __forceinline__ void stg256(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z),
               "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// ---- misc math ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// GELU with the exact-erf definition (reference: nn.GELU() in models/utils/layers.py:97), erf evaluated with
// Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7): 0.5 x (1 + erf(x/sqrt2)) = x - g for x >= 0 and g for x < 0, where
// g = 0.5 x poly(t) exp(-x^2/2), t = 1 / (1 + p |x| / sqrt2). 14 instructions, 2 of them on the XU pipe, against
// ~30 for erff(): the FFN-1 epilogue was issue-bound by it. Max abs deviation from the erff() form 4e-7.
__device__ __forceinline__ float gelu_fast(float x);
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {   // one FMNMX3
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx(z * z * -1.4426950408889634f);
  const float g = 0.5f * x * p * e;
  return x >= 0.f ? x - g : g;
}
// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2: two lanes per issue slot) ----------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// gelu_fast() on two values at once (same formula, constants folded: d = 1 + (p/sqrt2) |x|, poly pre-scaled by
// -0.5, exp argument -(log2e/2) x^2): ~20 instructions per PAIR, 4 of them MUFU. Returns the packed bf16 pair.
__device__ __forceinline__ uint32_t gelu_fast2_bf16(f32x2 x) {
  float x0, x1;
  unpack2(x, x0, x1);
  const f32x2 ax = pack2(fabsf(x0), fabsf(x1));
  const f32x2 d = fma2(ax, pack2(0.23164189f, 0.23164189f), pack2(1.0f, 1.0f));   // 0.3275911 / sqrt(2)
  float d0, d1;
  unpack2(d, d0, d1);
  const f32x2 t = pack2(rcp_approx(d0), rcp_approx(d1));
  f32x2 p = fma2(t, pack2(-0.5307027145f, -0.5307027145f), pack2(0.7265760135f, 0.7265760135f));   // -0.5 * a5, -0.5 * a4
  p = fma2(p, t, pack2(-0.7107068705f, -0.7107068705f));                                            // -0.5 * a3
  p = fma2(p, t, pack2(0.142248368f, 0.142248368f));                                                // -0.5 * a2
  p = fma2(p, t, pack2(-0.127414796f, -0.127414796f));                                              // -0.5 * a1
  const f32x2 q = mul2(mul2(p, t), ax);                              // -0.5 |x| poly(t)
  const f32x2 arg = mul2(mul2(ax, ax), pack2(-0.72134752f, -0.72134752f));   // -(x^2 / 2) log2(e)
  float a0, a1;
  unpack2(arg, a0, a1);
  const f32x2 e = pack2(ex2_approx(a0), ex2_approx(a1));
  float r0, r1;
  unpack2(fma2(q, e, pack2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), r0, r1);   // relu(x) - 0.5 |x| poly exp
  return pack_bf16x2(r0, r1);
}
// GELU through one MUFU: x Phi(x) = 0.5 x (1 + tanh(w(x))) holds exactly for w = atanh(erf(x / sqrt 2)), an odd function
// that a degree-5 odd polynomial fits to |GELU error| < 5.4e-5 over the whole real line (the classic 0.044715 cubic:
// 4.7e-4); tanh.approx adds at most 2^-11 relative on tanh, i.e. < 2.5e-4 |x| on the result - an eighth of the bf16
// rounding of the stored value. 6 packed fp32x2 operations + 2 MUFU + the bf16 pack per PAIR (gelu_fast2_bf16: ~20 + 4).
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t gelu_tanh2_bf16(f32x2 x) {
  const f32x2 x2 = mul2(x, x);
  f32x2 p = fma2(x2, pack2(-3.81889112e-04f, -3.81889112e-04f), pack2(3.72153111e-02f, 3.72153111e-02f));
  p = fma2(p, x2, pack2(7.97237410e-01f, 7.97237410e-01f));
  float w0, w1;
  unpack2(mul2(p, x), w0, w1);
  const f32x2 t = pack2(tanh_approx(w0), tanh_approx(w1));
  const f32x2 hx = mul2(x, pack2(0.5f, 0.5f));
  float r0, r1;
  unpack2(fma2(hx, t, hx), r0, r1);
  return pack_bf16x2(r0, r1);
}
__device__ __forceinline__ float gelu_tanh5(float x) {   // the same arithmetic on one value (edge tiles)
  const float x2 = x * x;
  float p = fmaf(x2, -3.81889112e-04f, 3.72153111e-02f);
  p = fmaf(p, x2, 7.97237410e-01f);
  const float t = tanh_approx(p * x);
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }
// CLIP's QuickGELU: x * sigmoid(1.702 x) (OpenAI clip/model.py, class QuickGELU)
__device__ __forceinline__ float quick_gelu(float x) { return x / (1.0f + __expf(-1.702f * x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#endif  // __CUDACC__

}  // namespace mmdm
