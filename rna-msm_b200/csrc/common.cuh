// Shared host/device helpers for the rnamsm_b200 CUDA library (sm_100a only).
//
// Everything that touches Blackwell-specific hardware goes through the thin inline-PTX
// wrappers below: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit /
// ld) and the UMMA shared-memory / instruction descriptors.  No CUTLASS types.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace rnamsm {

// ---------------------------------------------------------------------------------------------
// Host-side error plumbing (C ABI returns int status; message via rnamsm_last_error()).
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define RNAMSM_CHECK_CUDA(expr)                                                                   \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::rnamsm::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));  \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

#define RNAMSM_REQUIRE(cond, ...)                                                                 \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      ::rnamsm::set_error(__VA_ARGS__);                                                           \
      return 2;                                                                                   \
    }                                                                                             \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

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

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, %1;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}

// ---- programmatic dependent launch -------------------------------------------------------------
// With RNAMSM_PDL=1 (off by default: measured no gain, see api.cu) the kernels of the forward are launched with the
// programmatic-stream-serialization attribute (launch_pdl below): a kernel may START (barrier init, TMEM allocation, descriptor prefetch) while its predecessor in the stream drains,
// and blocks in pdl_wait() until the predecessor has completed and its writes are visible.  Rule: no global-memory
// access before pdl_wait(); pdl_launch_dependents() once this CTA holds every resource it needs.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug (wrong expect_tx byte count, missed commit) must never hang the
// GPU.  try_wait sleeps in hardware for a bounded time per call, so ~2^26 failed probes is
// many seconds; then we trap, which surfaces as a CUDA error on the host instead of a hang.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("rnamsm: mbarrier wait timeout (block %d thread %d bar %p parity %u)\n", blockIdx.x,
             threadIdx.x, (void*)bar, parity);
      __trap();
    }
  }
}

// The same on a precomputed shared-memory address (the generic -> shared conversion of a pointer is re-derived from
// %cluster_ctaid at every use otherwise: measurable in per-step barrier traffic).
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_u32(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe (try_wait may suspend the warp for a while when the phase is still open: measured, a probe issued
// early inside the softmax step cost 21 % of the kernel).
__device__ __forceinline__ bool mbar_test_wait_u32(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_quiet_u32(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_u32(bar, parity))
    if (++spins > (1u << 26)) __trap();
}

// Same bound, no message: for waits inside register-tight loops (the printf call site costs live registers).
__device__ __forceinline__ void mbar_wait_quiet(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 26)) __trap();
}

// Wait used by single-thread roles (TMA producer, MMA issuer) that share a scheduler with compute
// warps: try_wait with a suspend-time hint, so a failed probe parks the thread in hardware until the
// phase flips (or ~16 us pass) instead of re-probing every few hundred cycles -- measured: the probe
// loops of 5 role threads were 33 % of all instructions issued by the column-attention kernel.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, 16384u)) {
    if (++spins > (1u << 22)) {
      printf("rnamsm: mbarrier wait timeout (block %d thread %d bar %p parity %u)\n", blockIdx.x, threadIdx.x,
             (void*)bar, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// all state spaces: orders async-proxy accesses (TMA stores / reductions to global memory) with this thread's
// generic-proxy accesses
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- TMA -----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- tcgen05 / TMEM ------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (base_lane + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}


// registers -> 32 lanes x 32 consecutive fp32 columns (inverse of tmem_ld_32x32).
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A * B, single CTA, 16-bit operand type selected by the instruction descriptor.
__device__ __forceinline__ void umma_16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (M = 128 rows = lanes, 16-bit elements packed two per
// 32-bit column) comes from tensor memory -- used for P V with P written by tcgen05.st.
__device__ __forceinline__ void umma_16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- thread-block clusters / CTA pairs (cta_group::2) -----------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive / arrive.expect_tx on an mbarrier given by its shared::cluster address (may be remote).
// Default (.release.cta) semantics on purpose: the .release.cluster forms compile to
// MEMBAR.ALL.GPU + ERRBAR per arrive, and .acquire.cluster waits to CCTL.IVALL -- measured 3x slower
// mainloop.  The data these barriers order moves through the async proxy (TMA, tcgen05) and is
// ordered by complete_tx / tcgen05.commit / tcgen05.fence, exactly as in single-CTA kernels.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr),
               "r"(bytes)
               : "memory");
}
// TMA load issued by either CTA of a pair: data lands in the issuing CTA's smem, the byte count is
// signalled on the mbarrier at `bar_cluster_addr` (the leader CTA's barrier).
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B over the CTA pair (M = 256: 128 rows per CTA; B halves from both CTAs).
__device__ __forceinline__ void umma_16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with tf32 operands (fp32 words in shared memory, the low 13 mantissa bits ignored), K = 8 per instruction.
__device__ __forceinline__ void umma_tf32_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on the mbarrier at this smem offset in BOTH CTAs of the pair once all prior MMAs completed.
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// ---- TMA stores (bulk async-group completion) ----------------------------------------------------
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// global[box] += smem[box] (element type from the tensor map; fp32 here), performed at L2.
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// all but the most recent kPending bulk groups of this thread have COMPLETED (writes performed)
template <int kPending>
__device__ __forceinline__ void bulk_wait_pending() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kPending) : "memory");
}

// ---- UMMA descriptors ----------------------------------------------------------------------
// Shared-memory matrix descriptor (PTX "matrix descriptor", sm_100 version field = 1).
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version (1 on Blackwell)
//   bits [49,52) base offset (0: tiles are 1024 B aligned)   bits [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor, kind::f16 with f16 (fmt 0) or bf16 (fmt 1) A/B and fp32 D:
//   [4,6) D fmt (1=f32)  [7,10) A fmt  [10,13) B fmt  [15] A major (0=K,1=MN)  [16] B major  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_16(int M, int N, int fp16, int a_mn_major, int b_mn_major) {
  const uint32_t fmt = fp16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::tf32: A / B format 2 (tf32), fp32 accumulate, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FADD2: one issue slot for two lanes of work) ------------
__device__ __forceinline__ uint64_t f32x2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f32x2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f32x2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f32x2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__device__ __forceinline__ uint64_t f32x2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ---- warp-level LayerNorm ------------------------------------------------------------------------
// A row held as nv float4 per lane (features lane*4 + i*128 ..).  Two-pass (mean, then centred variance) in
// registers: matches nn.LayerNorm's biased variance.  ONE definition shared by the stand-alone LayerNorm kernels
// and the GEMM's fused residual + LayerNorm epilogue, so both produce bit-identical rows.
template <int kMaxV>
__device__ __forceinline__ void warp_layernorm(float4 (&v)[kMaxV], int nv, int D, float eps,
                                               const float* __restrict__ w, const float* __restrict__ b, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxV; ++i)
    if (i < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxV; ++i)
    if (i < nv) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      // explicit rounding steps: every instantiation (stand-alone kernels, GEMM epilogue) contracts the same way
      q = __fadd_rn(q, __fadd_rn(__fmaf_rn(v[i].x, v[i].x, __fmul_rn(v[i].y, v[i].y)),
                                 __fmaf_rn(v[i].z, v[i].z, __fmul_rn(v[i].w, v[i].w))));
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
  for (int i = 0; i < kMaxV; ++i)
    if (i < nv) {
      const int f = lane * 4 + i * 128;
      const float4 ww = *reinterpret_cast<const float4*>(w + f);
      const float4 bb = *reinterpret_cast<const float4*>(b + f);
      v[i].x = __fmaf_rn(__fmul_rn(v[i].x, rstd), ww.x, bb.x);
      v[i].y = __fmaf_rn(__fmul_rn(v[i].y, rstd), ww.y, bb.y);
      v[i].z = __fmaf_rn(__fmul_rn(v[i].z, rstd), ww.z, bb.z);
      v[i].w = __fmaf_rn(__fmul_rn(v[i].w, rstd), ww.w, bb.w);
    }
}

// Two rows at once (their shuffle reductions interleave, so the latency of one hides behind the other); per row
// exactly the operations of warp_layernorm, hence the same bits.
template <int kMaxV>
__device__ __forceinline__ void warp_layernorm2(float4 (&a)[kMaxV], float4 (&c)[kMaxV], int nv, int D, float eps,
                                                const float* __restrict__ w, const float* __restrict__ b, int lane) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxV; ++i)
    if (i < nv) {
      s0 += (a[i].x + a[i].y) + (a[i].z + a[i].w);
      s1 += (c[i].x + c[i].y) + (c[i].z + c[i].w);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  const float m0 = s0 / (float)D, m1 = s1 / (float)D;
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxV; ++i)
    if (i < nv) {
      a[i].x -= m0; a[i].y -= m0; a[i].z -= m0; a[i].w -= m0;
      c[i].x -= m1; c[i].y -= m1; c[i].z -= m1; c[i].w -= m1;
      q0 = __fadd_rn(q0, __fadd_rn(__fmaf_rn(a[i].x, a[i].x, __fmul_rn(a[i].y, a[i].y)),
                                   __fmaf_rn(a[i].z, a[i].z, __fmul_rn(a[i].w, a[i].w))));
      q1 = __fadd_rn(q1, __fadd_rn(__fmaf_rn(c[i].x, c[i].x, __fmul_rn(c[i].y, c[i].y)),
                                   __fmaf_rn(c[i].z, c[i].z, __fmul_rn(c[i].w, c[i].w))));
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    q0 += __shfl_xor_sync(0xffffffffu, q0, o);
    q1 += __shfl_xor_sync(0xffffffffu, q1, o);
  }
  const float r0 = rsqrtf(q0 / (float)D + eps), r1 = rsqrtf(q1 / (float)D + eps);
#pragma unroll
  for (int i = 0; i < kMaxV; ++i)
    if (i < nv) {
      const int f = lane * 4 + i * 128;
      const float4 ww = *reinterpret_cast<const float4*>(w + f);
      const float4 bb = *reinterpret_cast<const float4*>(b + f);
      a[i].x = __fmaf_rn(__fmul_rn(a[i].x, r0), ww.x, bb.x);
      a[i].y = __fmaf_rn(__fmul_rn(a[i].y, r0), ww.y, bb.y);
      a[i].z = __fmaf_rn(__fmul_rn(a[i].z, r0), ww.z, bb.z);
      a[i].w = __fmaf_rn(__fmul_rn(a[i].w, r0), ww.w, bb.w);
      c[i].x = __fmaf_rn(__fmul_rn(c[i].x, r1), ww.x, bb.x);
      c[i].y = __fmaf_rn(__fmul_rn(c[i].y, r1), ww.y, bb.y);
      c[i].z = __fmaf_rn(__fmul_rn(c[i].z, r1), ww.z, bb.z);
      c[i].w = __fmaf_rn(__fmul_rn(c[i].w, r1), ww.w, bb.w);
    }
}

// ---- misc math -----------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

#endif  // __CUDACC__

// Host: launch, with the programmatic-dependent-launch attribute when RNAMSM_PDL=1.
bool pdl_enabled();
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// Host: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed).
enum { TMAP_BF16 = 0, TMAP_F16 = 1, TMAP_F32 = 2 };
int encode_tmap(CUtensorMap* map, int elem /* TMAP_* */, const void* base, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes /* rank-1 entries, dims 1.. */, const uint32_t* box);

}  // namespace rnamsm
