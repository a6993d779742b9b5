// sm_100a building blocks for the tensor-core path: mbarrier, TMA (cp.async.bulk.tensor),
// TMEM allocation, tcgen05.mma / .commit / .ld, UMMA shared-memory + instruction descriptors,
// and host-side tensor-map encoding.  Raw inline PTX; no CUTLASS.
#pragma once
#include <cuda.h>
#include <cstdio>
#include "common.cuh"

namespace plk {
namespace tc {

// ------------------------------------------------------------------------------------------
// geometry shared by every tcgen05 kernel in this library
// ------------------------------------------------------------------------------------------
constexpr int kTileRows = 128;              // UMMA M (rows of the resident operand per CTA)
constexpr int kChunkK = 64;                 // bf16 elements per 128-byte swizzle row
constexpr int kChunkBytes = kTileRows * kChunkK * 2;  // one [128 x 64] bf16 TMA box = 16 KiB
constexpr int kUmmaK = 16;                  // K per tcgen05.mma for 16-bit inputs
constexpr int kMaxSmem = 232448;            // 227 KiB opt-in dynamic shared memory per CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of this cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// The same without the cluster-scope release: `.release.cluster` costs MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR per
// arrive (~1000 cycles with loads in flight; measured in infonce_grad_tc5: 3300-cycle epilogues against
// 1900).  Enough when what the arrive publishes is already ordered by its own fence -- tcgen05.fence::
// before_thread_sync after a tcgen05.ld / st, fence.proxy.async after shared-memory stores that the tensor
// core of THIS SM will read -- the form CUTLASS's ClusterBarrier::arrive(cta) uses for the same hand-offs.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (reported as a CUDA error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > 4000000000LL) {
      printf("plk: mbarrier wait timed out (block %d,%d thread %d bar@%u parity %u)\n", blockIdx.x,
             blockIdx.y, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// the same wait with cluster-scope acquire: pairs with mbar_arrive_cluster from the OTHER CTA of a pair
__device__ __forceinline__ bool mbar_try_wait_cl(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cl(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_cl(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > 4000000000LL) {
      printf("plk: cluster mbarrier wait timed out (block %d,%d thread %d bar@%u parity %u)\n", blockIdx.x,
             blockIdx.y, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// generic-proxy writes to shared memory -> visible to the async proxy (TMA / tcgen05 operands)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: box lands at `dst` (shared), completion bytes are signalled on `bar`.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 3-D tiled store shared -> global (bulk async group), and the fences around it
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// Only until the staged shared memory has been READ: the CTA may then reuse it or exit; the global writes
// finish on their own and are complete, like every other write of the grid, when the grid is.
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Multicast variant: the box is written at the same CTA-relative offset in every CTA of `mask`
// and complete_tx is signalled on the mbarrier at the same offset in each of them.
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                               int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "h"(mask)
      : "memory");
}

// ------------------------------------------------------------------------------------------
// programmatic dependent launch
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void griddep_wait() { pdl_wait(); }   // see common.cuh

// ------------------------------------------------------------------------------------------
// thread-block clusters
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Execution barrier only (arrive.relaxed: no MEMBAR.ALL.GPU / ERRBAR, ~1000 cycles less per use).  Enough (a)
// after mbarrier.init + fence.mbarrier_init.release.cluster, which is what publishes the barriers to the
// peers, and (b) before exit, where the point is only that no CTA leaves while a peer can still write
// into its shared memory or barriers.
__device__ __forceinline__ void cluster_sync_exec() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// One lane of a CONVERGED warp (elect.sync).  The MMA-issuing warps run their loops with all 32 lanes
// and predicate only the tcgen05 instructions on this: inside `if (lane == 0) { loop }` the control flow
// is divergent, ptxas keeps every descriptor in per-lane registers and moves it to the uniform
// registers the instruction needs through an elect / R2UR.BROADCAST / branch sequence -- ~12 SASS
// instructions and ~95 cycles per tcgen05.mma, more than the 64 (32) cycles a 128x128x16 (128x64x16)
// MMA occupies the tensor pipe.  In warp-uniform control flow the address arithmetic stays on the
// uniform datapath.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------
// TMEM + tcgen05
// ------------------------------------------------------------------------------------------
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, 16-bit inputs, fp32 accumulate; issued by ONE thread.
// The two shared-memory descriptors are given as (low word, shared high word): the issuing
// thread is a single serial instruction stream, so everything loop-invariant is hoisted and a
// K-step costs two integer adds (see umma_desc_lo / kUmmaDescHi).
__device__ __forceinline__ void umma_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(0x40004040u /* kUmmaDescHi */), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]: the A operand (bf16, K-major, two K elements per 32-bit column,
// row m in TMEM lane m) is read straight from tensor memory -- used for G . V in the backward,
// where G was just produced from the logits tile by the epilogue warps (tcgen05.st).
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(0x40004040u /* kUmmaDescHi */), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
// ---- cta_group::2: one MMA over a PAIR of CTAs (M = 256: each CTA's 128 rows; B split along N between the
// two CTAs' shared memories at the same offset; D rows in each CTA's own TMEM).  Issued by the leader
// CTA (cluster rank 0) only; allocation / deallocation by one warp of EACH CTA.
template <int kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem) {  // one full warp, in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma2_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(0x40004040u /* kUmmaDescHi */), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(0x40004040u /* kUmmaDescHi */), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the pair's previously issued MMAs retired
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}

// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// Same, arriving on the barrier at this offset in every CTA of `mask` (multicast ring release).
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp reads TMEM lane (base_lane + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
// 16 lanes x 32 consecutive fp32 columns as four 16x256b fragments (the mma accumulator layout):
//   r[4n + 2h + b] = element (lane base_lane + t/4 + 8h, column 8n + 2(t%4) + b),  n = 0..3, h, b = 0..1
// A thread gets 2 rows x 16 columns of the block: sums along rows AND along columns are mostly in-thread.
// (Tried for the forward's column sums -- 7 shuffles per 32x32 block instead of 31 -- and measured SLOWER
// than the 32x32b row-per-thread epilogue: 2500 vs 2030 cycles per 128x128 tile, two loads per warp and
// four rows of bucket bookkeeping per thread; kept for reference, not used by the shipped kernels.)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns, registers -> TMEM (thread t writes lane base_lane + t)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),
        "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: PTX ISA "tcgen05 shared memory descriptor" / "instruction descriptor")
// ------------------------------------------------------------------------------------------
// Operand tile in shared memory = rows of 128 bytes (64 bf16), 128-byte swizzle, 8-row groups
// of 1024 bytes (what a TMA box {64, rows} with CU_TENSOR_MAP_SWIZZLE_128B produces).
//   K-major view  : row = M/N index, the 64 elements of a row run along K.
//   MN-major view : row = K index,   the 64 elements of a row run along M/N.
// In both views the stride between 8-row groups (SBO) is 1024 bytes.  LBO is the stride between
// 64-element column blocks in the MN-major view (unused when the operand is one block wide).
// 64-bit descriptor: [0,14) start address >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, [46,48) version = 1
// (sm_100), [61,64) layout = 2 (SWIZZLE_128B).  Split in words so the loop-invariant half is a constant:
// high word = SBO 1024 B (>>4 = 64), version 1 (bit 46 -> 14),
// SWIZZLE_128B (bits 61..63 -> 29..31); low word = start address >> 4 | (LBO >> 4) << 16.
constexpr uint32_t kUmmaDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
static_assert(kUmmaDescHi == 0x40004040u, "descriptor high word");
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t umma_idesc_16(int m, int n, int a_mn_major, int b_mn_major,
                                                     int f16 /* 0: bf16 operands, 1: fp16 operands */) {
  return (1u << 4)                       // [4,6)   D format  = F32
         | ((f16 ? 0u : 1u) << 7)        // [7,10)  A format  (0 = F16, 1 = BF16)
         | ((f16 ? 0u : 1u) << 10)       // [10,13) B format
         | ((uint32_t)a_mn_major << 15)  // [15]    A major   (0 = K, 1 = MN)
         | ((uint32_t)b_mn_major << 16)  // [16]    B major
         | ((uint32_t)(n >> 3) << 17)    // [17,23) N >> 3
         | ((uint32_t)(m >> 4) << 24);   // [24,29) M >> 4
}

// ------------------------------------------------------------------------------------------
// streamed-operand ring shared by a cluster of CS CTAs that sweep the SAME [128 x 64] chunks
// (different resident rows): every CTA fetches 1/CS of each chunk and multicasts it to all, so
// the L2 -> SM traffic of the streamed operand drops by CS.  A slot may be refilled only after
// every CTA's MMAs have read it: the `empty` barriers count CS arrivals (one multicast
// tcgen05.commit per CTA); the `full` barriers count the local producer + 16 KiB of tx bytes.
// ------------------------------------------------------------------------------------------
template <int CS>
__device__ __forceinline__ void ring_load(uint8_t* slot, const CUtensorMap* tm_full,
                                          const CUtensorMap* tm_part, uint64_t* full_bar, int c0,
                                          int row0, uint32_t cta_rank) {
  mbar_expect_tx(full_bar, kChunkBytes);
  if constexpr (CS == 1) {
    tma_load_2d(slot, tm_full, full_bar, c0, row0);
  } else {
    tma_load_2d_mc(slot + cta_rank * (kChunkBytes / CS), tm_part, full_bar, c0,
                   row0 + (int)cta_rank * (kTileRows / CS), (uint16_t)((1u << CS) - 1));
  }
}
// one chunk of a multi-chunk transaction (the caller armed `full_bar` with the total byte count)
template <int CS>
__device__ __forceinline__ void chunk_load(uint8_t* slot, const CUtensorMap* tm_full,
                                           const CUtensorMap* tm_part, uint64_t* full_bar, int c0,
                                           int row0, uint32_t cta_rank) {
  if constexpr (CS == 1) {
    tma_load_2d(slot, tm_full, full_bar, c0, row0);
  } else {
    tma_load_2d_mc(slot + cta_rank * (kChunkBytes / CS), tm_part, full_bar, c0,
                   row0 + (int)cta_rank * (kTileRows / CS), (uint16_t)((1u << CS) - 1));
  }
}
template <int CS>
__device__ __forceinline__ void ring_release(uint64_t* empty_bar) {
  if constexpr (CS == 1) umma_commit(empty_bar);
  else umma_commit_mc(empty_bar, (uint16_t)((1u << CS) - 1));
}

// ------------------------------------------------------------------------------------------
// misc device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2 -- one FMA-pipe issue slot for two lanes' worth of work).  A
// three-register FFMA occupies a scheduler's FMA pipe for two cycles, and the epilogues spend most of their
// issue slots there, so the element-wise parts run on pairs.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t f2_pack_u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 2^x for a PAIR on the FMA pipe (no MUFU): x = n + f with n = round(x), |f| <= 0.5 (magic-number rounding),
// 2^f by a degree-5 minimax polynomial (max relative error 2.3e-7 in fp32 Horner form, i.e. the accuracy of
// MUFU.EX2's 2 ulp), times 2^n built from the integer bits.  x is clamped at -127, where 2^n becomes the bit
// pattern 0: results below the normal range are 0 like ex2.approx.ftz; callers guarantee x < 128.
// The InfoNCE epilogues need one exponential per logit: 16384 per 128 x 128 tile = 1024 cycles of the SM's 16
// MUFU lanes, the longest stage of every tile.  Moving one pair in four here (-DPLK_EXP_POLY_EVERY=4) trades 256
// of those cycles for 13 packed fp32 instructions per pair.  MEASURED (B200, parity green, 166 GPU tests): it
// does not pay -- backward 98.3 -> 106.4 us at B = 8192, 104.5 -> 112.7 us at d = 512, step 67.6 -> 69.6 us,
// forward unchanged: the packed instructions occupy the fp32 pipe for two cycles each and the four epilogue
// warps per scheduler are short of issue slots in exactly the phase that was MUFU-bound.  Default: off.
__device__ __forceinline__ uint64_t ex2_poly2(uint64_t x2) {
  float x0, x1;
  f2_unpack(x2, x0, x1);
  const uint64_t xc = f2_pack(fmaxf(x0, -127.f), fmaxf(x1, -127.f));
  const uint64_t r = f2_add(xc, f2_pack(12582912.f, 12582912.f));           // 1.5 * 2^23: integer in the low bits
  const uint64_t n = f2_add(r, f2_pack(-12582912.f, -12582912.f));
  const uint64_t f = f2_fma(n, f2_pack(-1.f, -1.f), xc);
  uint64_t p = f2_fma(f, f2_pack(0.0013276456f, 0.0013276456f), f2_pack(0.009675541f, 0.009675541f));
  p = f2_fma(p, f, f2_pack(0.055507135f, 0.055507135f));
  p = f2_fma(p, f, f2_pack(0.2402212f, 0.2402212f));
  p = f2_fma(p, f, f2_pack(0.69314694f, 0.69314694f));
  p = f2_fma(p, f, f2_pack(1.0000001f, 1.0000001f));
  const uint32_t s0 = ((uint32_t)r << 23) + 0x3F800000u;                    // 2^n: (n + 127) << 23
  const uint32_t s1 = ((uint32_t)(r >> 32) << 23) + 0x3F800000u;
  return f2_mul(p, f2_pack_u(s0, s1));
}
// which pairs of a 32-column chunk (pair index 0..15) take the polynomial: every PLK_EXP_POLY_EVERY-th; 0 = none
#ifndef PLK_EXP_POLY_EVERY
#define PLK_EXP_POLY_EVERY 0
#endif
__device__ __forceinline__ constexpr bool exp_pair_on_fma(int pair) {
  return PLK_EXP_POLY_EVERY > 0 && (pair % (PLK_EXP_POLY_EVERY > 0 ? PLK_EXP_POLY_EVERY : 1)) == (PLK_EXP_POLY_EVERY - 1);
}

// Logistic pieces for the SigLIP epilogues; zl = z * log2(e).
//   sigmoid(z)  = 1 / (1 + e^-z)                 (2 MUFU; relative accuracy also for z << 0)
//   softplus(z) = max(z, 0) + log1p(e^-|z|)      (2 MUFU; series below t = 0.01 where 1 + t loses t)
__device__ __forceinline__ float sigmoid_l2(float zl) {
  const float e = ex2_approx(-fabsf(zl));
  const float r = rcp_approx(1.0f + e);
  return zl >= 0.f ? r : e * r;
}
__device__ __forceinline__ float softplus_l2(float zl) {
  const float t = ex2_approx(-fabsf(zl));
  const float series = t * fmaf(t, fmaf(t, 0.33333334f, -0.5f), 1.0f);
  const float l1p = t < 0.01f ? series : lg2_approx(1.0f + t) * 0.6931471805599453f;
  return fmaf(fmaxf(zl, 0.f), 0.6931471805599453f, l1p);
}

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <bool F16>
__device__ __forceinline__ uint32_t pack_16x2(float lo, float hi) {
  if constexpr (F16) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}

// ------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------
// 16-bit (bf16 or fp16) row-major [rows, ld] matrix viewed through boxes of {64 columns, box_rows rows},
// 128-byte swizzle, out-of-bounds elements read as zero.
int make_tmap_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld,
                   int box_rows);
int make_tmap_f32_slabs(CUtensorMap* out, const void* base, int64_t slabs, int64_t rows, int64_t cols);

// launch with an optional {1, cluster_y, 1} thread-block cluster; `overlap_prev` adds programmatic
// stream serialization: the grid may start while the previous kernel in the stream is still
// running (once that kernel has executed griddepcontrol.launch_dependents, or finished), and must
// execute griddep_wait() before touching anything the previous kernel produces.
template <typename... KArgs, typename... Args>
int launch_kernel_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                     int cluster_y, bool overlap_prev, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster_y > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = (unsigned)cluster_y;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (overlap_prev) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  PLK_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
  return PLK_OK;
}
template <typename... KArgs, typename... Args>
int launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                  int cluster_y, Args... args) {
  return launch_kernel_ex(kern, grid, block, smem, st, cluster_y, false, args...);
}

}  // namespace tc
}  // namespace plk
