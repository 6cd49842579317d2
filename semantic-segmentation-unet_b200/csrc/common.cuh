// Shared device/host helpers for libunetb200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/unetb200.h"

// ---------------------------------------------------------------------------------------------
// error plumbing: every entry point returns an int status and records a thread-local message
// ---------------------------------------------------------------------------------------------
void ub_set_error(const char* fmt, ...);

#define UB_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ub_set_error(__VA_ARGS__);                \
      return UB_ERR_INVALID_ARG;                \
    }                                           \
  } while (0)

#define UB_CHECK_SHAPE(cond, ...)               \
  do {                                          \
    if (!(cond)) {                              \
      ub_set_error(__VA_ARGS__);                \
      return UB_ERR_UNSUPPORTED_SHAPE;          \
    }                                           \
  } while (0)

#define UB_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ub_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return UB_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define UB_LAUNCH_CHECK() UB_CUDA(cudaGetLastError())

int ub_num_sms();

// ---------------------------------------------------------------------------------------------
// TMA descriptor encode (driver entry point fetched at run time; no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------
// NHWC-style 4-D bf16 view: dims (innermost first) {C, W, H, N} with byte strides for W, H, N.
// box = {64 channels, box_w, box_h, 1}, SWIZZLE_128B, OOB elements read as zero / are not written.
int ub_tmap_act4d(CUtensorMap* out, const void* base, int C, int W, int H, int N, long long stride_w_bytes,
                  long long stride_h_bytes, long long stride_n_bytes, int box_w, int box_h);
// 2-D bf16 K-major matrix [rows][K]: box = {64, box_rows}, SWIZZLE_128B.
int ub_tmap_mat2d(CUtensorMap* out, const void* base, long long rows, long long K, int box_rows);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// Bounded spin: a protocol bug traps (cudaErrorLaunchFailure) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
      printf("unetb200: mbarrier wait timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
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
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread = TMEM lane (row), v[j] = column j.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- warp-convergent issue helpers ---------------------------------------------------------------------------------
// The producer / MMA warps run their loops with all 32 lanes in uniform control flow and let ONE elected lane issue
// (elect.sync inside the asm).  Issuing from `if (lane == 0)` instead makes the compiler wrap every UTMALDG / UTCHMMA /
// UTCBAR in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop (their operands must be uniform registers), which cost ~120 clocks
// per MMA in the first version of these kernels and made the 64/128-column tiles issue-bound (profiles/r01_*).
// elect.sync is deterministic for a given member mask, so the MMAs and their tcgen05.commit come from the same thread.
__device__ __forceinline__ void mbar_expect_tx_e(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_e(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_e(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_e(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_e(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Four K = 16 steps of one 64-element K block (K-major operands: +32 bytes = +2 in the descriptor address field per step) with
// ONE elect: the issuing warp's instruction stream, not the tensor pipe, bounds tiles with few columns (profiles/r01g: a
// single warp sustains roughly one dependent instruction per 4-5 clocks; N = 64 MMAs retire every 48).
__device__ __forceinline__ void tc_mma4_bf16_e(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred p, q, t;\n\t"
      ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, 2;\n\t"
      "add.s64 a2, %1, 4;\n\tadd.s64 b2, %2, 4;\n\t"
      "add.s64 a3, %1, 6;\n\tadd.s64 b3, %2, 6;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
// NK consecutive K = 16 steps with ONE elect; the operand descriptors advance by a_step / b_step (in descriptor address units of
// 16 bytes) per step.  Used by the MN-major (weight-gradient) kernels, whose K steps are pixel rows.
template <int NK>
__device__ __forceinline__ void tc_mma_steps_bf16_e(uint32_t d_tmem, uint64_t adesc, uint64_t a_step, uint64_t bdesc, uint64_t b_step,
                                                    uint32_t idesc, uint32_t accumulate_first) {
  static_assert(NK == 4 || NK == 8, "NK");
  if (NK == 8) {
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t"
        ".reg .b64 a, b;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %5, p;\n\t"
        "add.s64 a, %1, %2;\n\tadd.s64 b, %3, %4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.s64 a, a, %2;\n\tadd.s64 b, b, %4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.s64 a, a, %2;\n\tadd.s64 b, b, %4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.s64 a, a, %2;\n\tadd.s64 b, b, %4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.s64 a, a, %2;\n\tadd.s64 b, b, %4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.s64 a, a, %2;\n\tadd.s64 b, b, %4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.s64 a, a, %2;\n\tadd.s64 b, b, %4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(a_step), "l"(bdesc), "l"(b_step), "r"(idesc), "r"(accumulate_first)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t"
        ".reg .b64 a, b;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %5, p;\n\t"
        "add.s64 a, %1, %2;\n\tadd.s64 b, %3, %4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.s64 a, a, %2;\n\tadd.s64 b, b, %4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t"
        "add.s64 a, a, %2;\n\tadd.s64 b, b, %4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, t;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(a_step), "l"(bdesc), "l"(b_step), "r"(idesc), "r"(accumulate_first)
        : "memory");
  }
}
// ---- CTA pair (cta_group::2): two SMs of one TPC issue ONE tcgen05.mma of M = 256; each CTA holds its own 128 rows of A and HALF of
// the B tile, so the shared-memory port of an SM serves 6 KB instead of 8 KB per 128x128x16 MMA step.  Only the leader (cluster rank 0)
// issues MMAs; both CTAs issue TMA loads that signal the LEADER's mbarrier (peer bit 24 of the shared::cluster address cleared), and
// tcgen05.commit multicasts its arrive to the same barrier in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t UB_PEER_MASK = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the even (leader) CTA of the pair
__device__ __forceinline__ void tma_load_4d_2sm_e(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar) & UB_PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm_e(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar) & UB_PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm_e(uint64_t* bar) {          // arrive on `bar` in BOTH CTAs once the MMAs issued so far are done
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b16 m;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "mov.b16 m, 3;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_mma4_bf16_2sm_e(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred p, q, t;\n\t"
      ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, 2;\n\t"
      "add.s64 a2, %1, 4;\n\tadd.s64 b2, %2, 4;\n\t"
      "add.s64 a3, %1, 6;\n\tadd.s64 b3, %2, 6;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, t;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive on the barrier at the same offset in cluster CTA `cta` (the leader's accumulator-empty barrier, from the peer's epilogue warps)
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}

// value known to be identical in all lanes -> tell the compiler (lets it keep descriptors / addresses in uniform registers)
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// Shared-memory matrix descriptor, SWIZZLE_128B, 8-row atoms of 1024 B (version 1 = Blackwell).
//   K-major operand  : rows = M/N index at 128 B pitch, SBO = stride between 8-row groups.
//   MN-major operand : rows = K index at 128 B pitch (64 MN elements per row), SBO = stride between 8-K groups,
//                      LBO = stride between 64-element MN blocks.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;   // descriptor version (sm_100)
  d |= 2ull << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif  // __CUDACC__
