// Thin inline-PTX layer for sm_100a: mbarrier, TMA (tiled + im2col), tcgen05 (alloc/mma/commit/ld).
// Every wait is bounded by a clock64() watchdog: a kernel that deadlocks records an error code in
// global memory and unwinds instead of hanging the GPU.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace tsr {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// Watchdog budget in SM clocks (~1 s at 1.9 GHz).
#ifndef TSR_WATCHDOG_CYCLES
#define TSR_WATCHDOG_CYCLES (1ll << 31)
#endif

// Returns false on timeout after writing `code` to *err (if non-null).
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > TSR_WATCHDOG_CYCLES) {
      if (err) atomicExch(err, code);
      return false;
    }
  }
  return true;
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// Tiled 4-D load on an NHWC tensor map {C, W, H, N}: the whole box starting at (c, w, h, n); coordinates may be
// negative / run past the tensor (the TMA unit zero-fills out-of-bounds elements and still delivers the full box).
__device__ __forceinline__ void tma_load_tile_4d(uint32_t dst, const void* map, uint32_t bar, int c, int w, int h, int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}
// im2col-mode load on an NHWC tensor map {C, W, H, N}: `pixelsPerColumn` pixels starting at (w, h, n),
// walking W then H then N inside the map's bounding box, each shifted by the filter offsets (off_w, off_h).
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const void* map, uint32_t bar, int c, int w, int h,
                                                   int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
      "h"(off_h)
      : "memory");
}

// Tiled stores shared -> global (bulk async-group completion): the box leaves shared memory in the tensor map's swizzle
// layout; elements outside the tensor are clipped. The issuing thread must have seen every generic-proxy write of the
// tile (CTA barrier) and those writes must have been made visible to the async proxy (fence_proxy_async_smem).
__device__ __forceinline__ void tma_store_2d(const void* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// The same boxes ADDED to global memory (bf16 add performed by the memory system): y += tile. Used for residual adds
// whose other operand already sits in the output buffer (in-place residual blocks of the inference plans).
__device__ __forceinline__ void tma_reduce_add_2d(const void* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const void* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk groups of this thread have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... have completed (their global writes are performed)
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------ thread-block clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// Cluster-wide barrier (every thread of every CTA of the cluster executes both halves).
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// im2col load whose box lands at the same shared-memory offset of every CTA in `cta_mask` and signals the mbarrier at
// the same offset in each of them (TMA multicast): the CTAs of a cluster that compute different N tiles of the same M
// tile fetch the activation box once per cluster instead of once per CTA.
__device__ __forceinline__ void tma_load_im2col_4d_mc(uint32_t dst, const void* map, uint32_t bar, int c, int w, int h,
                                                      int n, uint16_t off_w, uint16_t off_h, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes.multicast::cluster [%0], "
      "[%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
      "h"(off_h), "h"(cta_mask)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; bf16 inputs, fp32 accumulate. One thread issues for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The four K=16 steps of one 64-element K chunk (K-major bf16 operands: the start-address field of both shared-memory
// descriptors advances by 32 bytes = 2 units per step) followed by the commit that recycles the ring stage: one asm
// block, so the single issuing thread spends as few instructions as possible per K iteration.
__device__ __forceinline__ void umma_bf16_x4_commit(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                    uint32_t accumulate, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred p, t;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, %4, %4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %6, %7, %3, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %8, %9, %3, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %10, %11, %3, t;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(bar), "l"(adesc + 2), "l"(bdesc + 2),
        "l"(adesc + 4), "l"(bdesc + 4), "l"(adesc + 6), "l"(bdesc + 6)
      : "memory");
}
// Two K=16 steps + commit: 32-element K chunks (64-byte swizzled rows), e.g. the 27 -> 32 column row expansions of the
// 9x9 layers.
__device__ __forceinline__ void umma_bf16_x2_commit(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                    uint32_t accumulate, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred p, t;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, %4, %4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %6, %7, %3, t;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(bar), "l"(adesc + 2), "l"(bdesc + 2)
      : "memory");
}
// The same with the commit arriving on the barrier at this offset in EVERY CTA of `cta_mask`: a ring stage that was
// filled by multicast loads may only be refilled once all receiving CTAs have consumed it.
__device__ __forceinline__ void umma_bf16_x4_commit_mc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                       uint32_t accumulate, uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "{\n\t.reg .pred p, t;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, %4, %4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %6, %7, %3, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %8, %9, %3, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %10, %11, %3, t;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%5], %12;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(bar), "l"(adesc + 2), "l"(bdesc + 2),
        "l"(adesc + 4), "l"(bdesc + 4), "l"(adesc + 6), "l"(bdesc + 6), "h"(cta_mask)
      : "memory");
}
// Same four steps without the commit (halo mode issues nine of these per stage, one per filter tap).
__device__ __forceinline__ void umma_bf16_x4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, t;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, %4, %4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %5, %6, %3, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %7, %8, %3, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %9, %10, %3, t;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "l"(adesc + 2), "l"(bdesc + 2),
        "l"(adesc + 4), "l"(bdesc + 4), "l"(adesc + 6), "l"(bdesc + 6)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp reads TMEM lane (lane_base + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns in one instruction (the staged epilogue reads a warp's whole share of the tile
// before its single wait::ld).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// The mirror image: thread i of the warp writes 16 consecutive fp32 columns of TMEM lane (lane_base + i). Used by the
// fused BatchNorm backward to park dz in the (already consumed) accumulator across the grid barrier.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor (tcgen05), see PTX "matrix descriptor":
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [49,52) base offset | [61,64) swizzle
// swizzle: 0 none, 2 = 128B, 4 = 64B, 6 = 32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type & 7) << 61;
  return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
//   [4,6) D fmt (1 = f32) | [7,10) A fmt (1 = bf16) | [10,13) B fmt | [15] A major (1 = MN) | [16] B major
//   [17,23) N>>3 | [24,29) M>>4
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t layout_type_for_row_bytes(int row_bytes) {
  return row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
}

// One lane of a converged warp (the same lane every time); the others get false.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ programmatic dependent launch
// Every kernel of the library may be launched with cudaLaunchAttributeProgrammaticStreamSerialization (launch.h):
// its CTAs then become resident while the previous kernel of the stream is still running. pdl_wait() returns once
// that kernel has completed and its writes are visible (no-op for a plain launch); everything before it (barrier
// init, TMEM allocation, tensor-map prefetch, loads of data no kernel of the chain writes) overlaps the previous
// kernel's tail. Every thread calls it before touching dependent global memory and before exiting, so completion of
// kernel N implies completion of kernels < N. pdl_trigger() lets the NEXT kernel's CTAs be scheduled; it is issued
// after the wait so that at most one successor is staged at a time.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
#ifdef TSR_PDL_EARLY
  pdl_trigger();   // experiment: lets the successor be staged before this kernel's own dependency has resolved
  pdl_wait();
#else
  pdl_wait();
  pdl_trigger();
#endif
}

// Pulls the 128-byte line holding `p` into L1 (no register destination, no ordering): used by the epilogue warps to
// fetch their residual / statistics operands while the main loop is still running.
__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// ------------------------------------------------------------------ misc
// Pins a loop-invariant kernel parameter in a register: ptxas otherwise re-loads it from the constant bank inside the
// loop (LDC/LDCU, ~35 clk each, in the dependency chain of the single-warp producer / MMA loops).
// (A warp shuffle is the cheapest operation whose result ptxas cannot trace back to the constant bank; call it
// with the whole warp converged.)
__device__ __forceinline__ uint32_t in_reg(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float bf16_bits_to_f32(uint32_t hi16) { return __uint_as_float(hi16 << 16); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t u, float& lo, float& hi) {
  lo = __uint_as_float(u << 16);
  hi = __uint_as_float(u & 0xFFFF0000u);
}

}  // namespace tsr
