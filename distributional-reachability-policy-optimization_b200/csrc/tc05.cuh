// tcgen05 / TMEM / TMA / mbarrier PTX wrappers.
// Shared by the fused rollout step kernel (rollout_pipe.cuh) and the fused critic kernels (critic_umma.cu).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace drpo {
namespace tc {

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the warp until the phase completes (or ~10 ms pass) instead of
// returning after a few dozen cycles.  Without the hint every waiting warp polls in a hot loop; the profile of the first
// version of this kernel showed 40% of all issued instructions were such polls, issued by the highest-priority warps.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (fails the launch) after ~2^26 polls (seconds) instead of hanging the GPU.  The slow path
// is out of line (every wait site costs two instructions of the small instruction cache) and as lean as possible: a failed
// try_wait returns after only ~40 ns, so a waiting warp re-issues the loop body every ~80 cycles and those instructions
// compete with the working warps of its scheduler (the first profile spent half of all issue slots on wait loops).
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar_addr, uint32_t parity, int* err_flag, int code) {
  long long t0 = 0;
#pragma unroll 1
  for (uint32_t it = 0;; ++it) {
    uint32_t ok;
    // try_wait WITH a suspend-time hint: the hardware parks the warp until the phase completes (or the hint, ~1 ms, expires) and wakes
    // it ~60 cycles after the arrival.  Without the hint a failed try_wait returns after ~40 ns and the waiting warps poll: the ncu
    // profile of the two-tile rollout kernel (profiles/r2_rollout.md) showed 56 % of all issued warp instructions were this loop, issued
    // with scheduler priority over the epilogue warps that were doing the work.
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar_addr), "r"(parity), "r"(1000000u) : "memory");
    if (ok) return;
    if ((it & 7) == 7) {
      // ~0.5 s of SM clocks without progress is a protocol bug; once one wait of the launch failed the others give up after 1 ms
      if (t0 == 0) t0 = clock64();
      const long long dt = clock64() - t0;
      if (dt > 1000000000ll || (dt > 2000000ll && *(volatile int*)err_flag)) break;
    }
  }
  // A protocol bug must fail loudly but must not hang the GPU (and a trap would hide which wait failed): record and report
  // the first failing wait, then let every warp run to completion; the results are garbage and the host checks err_flag.
  if (atomicCAS(err_flag, 0, code) == 0)
    printf("drpo_b200: mbarrier wait timed out (code %d, block %d, warp %d, parity %u, barrier smem 0x%x)\n", code, (int)blockIdx.x,
           (int)(threadIdx.x >> 5), parity, bar_addr);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag, int code) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  if (!ok) mbar_wait_slow(smem_u32(bar), parity, err_flag, code);
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity, int* err_flag, int code) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  if (!ok) mbar_wait_slow(addr, parity, err_flag, code);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one elected lane of a converged warp: unlike `lane == 0`, the compiler knows the branch is single-threaded, keeps the
// tcgen05.mma operands in uniform registers and emits back-to-back UTCHMMA (measured: 9-32 cycles/MMA instead of 63)
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xFFFFFFFF;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
  return pred;
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit_addr(uint32_t addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T      (A: 128 lanes x K bf16 packed two per column; B: K-major canonical layout)
// The descriptor travels as two 32-bit words so that stepping along K is one 32-bit add in the issue loop.
template <bool kAccumulate>
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi, uint32_t idesc) {
  if (kAccumulate)
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 d;\n\tsetp.eq.u32 p, 1, 1;\n\tmov.b64 d, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], d, %4, p;\n\t}\n"
                 ::"r"(d_tmem), "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 d;\n\tsetp.eq.u32 p, 1, 0;\n\tmov.b64 d, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], d, %4, p;\n\t}\n"
                 ::"r"(d_tmem), "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float sqrt_fast(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// softplus / soft clamp with fast intrinsics (bf16-path tolerance)
__device__ __forceinline__ float softplus_fast(float x) { return x > 15.f ? x : __logf(1.f + __expf(x)); }
__device__ __forceinline__ float soft_clamp_fast(float x, float lo, float hi) {
  x = hi - softplus_fast(hi - x);
  return lo + softplus_fast(x - lo);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// packed bf16x2 math for the epilogues (element 2j in the low half, 2j+1 in the high half)
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {      // max(x, 0) fused into the conversion
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t x) {
  uint32_t r; const uint32_t z = 0u;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(z));
  return r;
}
__device__ __forceinline__ uint32_t silu_bf16x2(uint32_t x) {          // x*sigmoid(x) = h + h*tanh(h), h = x/2
  uint32_t h, t, r; const uint32_t half2 = 0x3F003F00u;                  // (0.5, 0.5) in bf16
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(h) : "r"(x), "r"(half2));
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(t) : "r"(h));
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(h), "r"(t), "r"(h));
  return r;
}


// SS-mode MMA: D[tmem] (+)= A[smem] * B[smem]      (both operands through 64-bit shared-memory descriptors)
template <bool kAccumulate>
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  if (kAccumulate)
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.eq.u32 p, 1, 1;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
                 ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.eq.u32 p, 1, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
                 ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}
// run-time accumulate flag (the first MMA of an accumulation chain clears the accumulator)
__device__ __forceinline__ void mma_ss_p(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.u32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
               ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_ts_p(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 d;\n\tsetp.ne.u32 p, %5, 0;\n\tmov.b64 d, {%2, %3};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], d, %4, p;\n\t}\n"
               ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// ---- thread-block clusters: peer barriers and multicast bulk copies ---------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
// tcgen05.commit that arrives on the mbarrier at the same offset in every CTA of `mask` once this CTA's prior MMAs are done
__device__ __forceinline__ void tc_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// global -> shared memory of every CTA in `mask` (same offset in each), completing `bytes` on each CTA's own mbarrier
__device__ __forceinline__ void bulk_g2s_multicast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}

// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma / TMA reads of shared memory)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// UMMA shared-memory descriptor, no swizzle (cute::UMMA::SmemDescriptor, version 1).
//   K-major operand : core matrix = 8 (M/N) rows x 16 B of K; LBO = stride between K octets, SBO = stride between 8-row groups
//   MN-major operand: core matrix = 8 K rows x 16 B of M/N;   LBO = stride between 8-row K groups, SBO = stride between M/N octets
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor: D = f32, A = B = bf16, M = 128; a_mn / b_mn select MN-major operands
__host__ __device__ inline uint32_t make_idesc(int n, int a_mn = 0, int b_mn = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

}  // namespace tc
}  // namespace drpo
