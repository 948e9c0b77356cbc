// DRPO_PREC_BF16 rollout: one persistent, warp-specialised tcgen05 kernel per rollout step that runs the whole
//   policy MLP -> squashed-Gaussian sample -> ensemble-member MLP (trunk + 2 heads) -> Gaussian next-state sample
//   -> env hooks
// chain for a 128-row tile without ever leaving the SM:
//   * every dense layer is a sequence of tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, M = 128 trajectories) issued per
//     <=64-column output chunk into one of three TMEM accumulator buffers, so the epilogue of chunk j overlaps the MMAs of
//     chunk j+1 (and of the next independent layer);
//   * the activations live in TMEM too: the epilogue stores them as packed bf16 and the next layer reads its A operand
//     straight from TMEM (TS-mode MMA) - activations never touch shared or global memory;
//   * weights are packed once per rollout into the UMMA canonical K-major (no-swizzle) layout, one contiguous block per
//     (layer, chunk), and streamed from L2 into a shared-memory ring by TMA bulk copies (cp.async.bulk + mbarrier);
//   * bias rides in the GEMM: every activation tile carries a constant-1 column and the packed weights hold the bias in
//     the matching K slot; the epilogue is activation (packed bf16x2 math) + TMEM store only.
// Warp roles (768 threads, one CTA per SM):
//   warps 0-3   group C: the three narrow output chunks (policy head -> action + model input, diff head, log-var head ->
//               Gaussian sample), the coalesced next-state store, and the NEXT tile's prologue (cp.async prefetch of states and
//               noise, normalisation, policy input), so that the tail of tile i overlaps the first policy layers of tile i+1;
//   warps 4-19  four hidden-layer epilogue groups: group pair (g>>1) takes every other hidden chunk, (g&1) picks the column
//               half, so two chunks are always in their epilogue and the fixed latencies of one (barrier wake-up, tcgen05.ld,
//               tcgen05.st, fence) hide behind the other's MUFU work;
//   warp 20     idle; warp 21 TMA producer (weight chunks -> shared-memory ring);
//   warps 22-23 two MMA issuers (one elected lane each) alternating chunks; warp 23 also owns the TMEM allocation.
// The per-step Gaussian draws (Philox keyed by global trajectory id, or the injected parity tensors) are staged row-compact
// by noise_stage_kernel; the env hooks run fused with the replay-buffer store in hooks_store_kernel.
// Measured design inputs (tools/ubench_*.cu, profiles/): tcgen05.mma issues at the N/2-cycle hardware floor only from an
// elect.sync branch (a `lane == 0` branch costs 63 cycles/MMA); tcgen05.ld moves ~3 KB/clk/SM; MUFU is 4 lanes/clk/SMSP and
// tanh.approx.bf16x2 is two MUFU ops, which makes the SiLU epilogues MUFU-bound.
#include <cuda_bf16.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "common.cuh"
#include "hooks.cuh"
#include "nets.cuh"
#include "rollout.cuh"
#include "umma_api.h"

namespace drpo {
namespace umma {

constexpr int TILE_M = 128;
constexpr int NSLAB = 64;                  // max output columns per chunk / accumulator buffer
constexpr int NACC = 3;                    // accumulator buffers in TMEM
constexpr int MAX_CHUNKS = 32;
constexpr int MAX_LAYERS = 9;
constexpr int GROUP_THREADS = 128;         // one epilogue group = 4 warps = the 128 TMEM lanes
constexpr int N_HID_GROUPS = 4;            // hidden groups: pair (g>>1) takes every other hidden chunk, (g&1) picks the column half
constexpr int EPI_THREADS = (1 + N_HID_GROUPS) * GROUP_THREADS;   // group C (output chunks, prologue) + the hidden groups
constexpr int NUM_THREADS = EPI_THREADS + 128;      // + warp 20 idle, warp 21 TMA producer, warps 22-23 MMA issuers
// a warp reaches TMEM lanes 32*(warp%4)..; the SM's arbiter prefers high warp ids, so the single-lane roles sit on top
constexpr int GROUP_C_WARP0 = 0, GROUP_H_WARP0 = 4;
constexpr int PRODUCER_WARP = NUM_THREADS / 32 - 3, MMA_WARP = NUM_THREADS / 32 - 1, MMA_WARP2 = NUM_THREADS / 32 - 2;
constexpr int MAX_STAGES = 6;
constexpr int NSPECIAL = 3;                // output chunks per tile: policy head, diff head, log-var head
constexpr uint32_t TM_COLS = 512;
// TMEM column map (32-bit columns; bf16 activations take kp/2 columns)
constexpr uint32_t TM_ACC = 0;                                   // 3 x 64 fp32 accumulator columns
constexpr uint32_t TM_PA = 192, TM_PB = 328, TM_XP = 504;        // policy: hidden A (<=136), hidden B (<=136), input (8)
constexpr uint32_t TM_H2 = 192, TM_D1 = 296, TM_L1 = 400;        // model: h2, h1/d1, x_m/l1 (<=104 each)
// TM_XP is outside every model-phase region, so group C can stage tile i+1's policy input while tile i is in its model phase

enum LayerKind { HID_RELU = 0, HID_SILU = 1, OUT_POLICY = 2, OUT_DIFF = 3, OUT_LOGVAR = 4 };

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

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
__device__ __noinline__ void mbar_wait_slow(uint32_t bar_addr, uint32_t parity, int* err_flag, int code) {
  long long t0 = 0;
#pragma unroll 1
  for (uint32_t it = 0;; ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "@!p mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
    if (ok) return;
    // A failed try_wait returns after ~40 ns, so a waiting warp would re-issue this loop every ~80 cycles on the scheduler it
    // shares with working warps (the profile of an earlier version spent 65% of all issued instructions here).  Short waits
    // (the pipeline's hand-offs) poll back to back; long ones (output group between its chunks, idle groups at a layer
    // boundary, the producer) back off with a short sleep.
    if (it >= 8) __nanosleep(64);
    if ((it & 63) == 63) {
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

// K-major, no-swizzle UMMA shared-memory descriptor: core matrix = 8 rows x 16 B, LBO = K-direction stride,
// SBO = 8-row-group stride (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128
__host__ __device__ inline uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------------
// weight image: per layer, per 64-row slab, a contiguous block in canonical layout  [n/8][k/8][8 rows][8 elems]
// ---------------------------------------------------------------------------------------------------------------
struct LayerSpec {
  int n_real, k_real;      // nn.Linear out / in
  int np, kp;              // padded MMA N (x16) and K (x16, includes the bias slot at k_real)
  int kind;                // LayerKind
  int a_col, out_col;      // TMEM column of the A operand / of the activation written by the epilogue
  int first_chunk, n_chunks;
  int dep;                 // layer whose epilogue must be complete before this layer's MMAs may read a_col (-1: tile input)
  int next_kp;             // K (padded) of the consumer of out_col
};
// One MMA group: output columns [n0, n0+nc) of a layer, accumulated into TMEM buffer (chunk index % NACC).
// special = -1 for hidden-layer chunks (epilogue groups A/B, `hid` = ordinal among the tile's hidden chunks), else the
// index of the group-C barrier pair (0 policy head, 1 diff head, 2 log-var head).
struct ChunkSpec { uint32_t offset, bytes; uint16_t n0, nc; uint16_t layer; int8_t special; uint8_t hid; };
struct NetPlan {
  LayerSpec layer[MAX_LAYERS];
  ChunkSpec chunk[MAX_CHUNKS];
  int n_layers, n_chunks;
  uint32_t policy_bytes, model_bytes;     // image sizes; policy chunks index the actor image, model chunks the member image
  int n_policy_chunks;
  uint32_t max_chunk_bytes;
  // Flattened per-chunk records, copied to shared memory at kernel start so that the per-chunk loops of the issuer and of the
  // hidden groups read one 16-byte word instead of chasing the layer/chunk tables through the constant bank.
  //   irec (issuer):  x = idesc, y = descriptor high word (SBO, version), z = a_col | nk << 16 | (special + 1) << 24 | (hid & 1) << 30,
  //                   w = kA | kB << 8 | (bufA + 1) << 16 | (bufB + 1) << 20   (K-sliced start of a layer's first chunk, 0 = none)
  //   erec (hidden):  x = n0 | nc << 16, y = out_col | kind << 16 | (hid & 1) << 24 | hidden << 25, z = n_real | np << 16, w = next_kp
  uint4 irec[MAX_CHUNKS], erec[MAX_CHUNKS];
};

// Weight image of one layer: per chunk a contiguous block in the canonical K-major layout [n/8][k/8][8 rows][8 elems];
// column kp-? : the bias sits in K slot k_real (the activations carry a constant 1 there).
struct PackChunks { int n; int n0[5]; };      // chunk boundaries n0[0..n]
__global__ void pack_layer_kernel(const float* __restrict__ W, const float* __restrict__ b, int n_real, int k_real, int np, int kp,
                                  PackChunks pc, __nv_bfloat16* __restrict__ dst /* start of this layer inside the image */) {
  const int total = np * kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / kp, k = i % kp;
    float v = 0.f;
    if (n < n_real) v = k < k_real ? W[(int64_t)n * k_real + k] : (k == k_real ? b[n] : 0.f);
    int ci = 0;
    while (ci + 1 < pc.n && n >= pc.n0[ci + 1]) ++ci;
    const int nin = n - pc.n0[ci];
    const int64_t idx = (int64_t)pc.n0[ci] * kp + ((int64_t)(nin >> 3) * (kp >> 3) + (k >> 3)) * 64 + (nin & 7) * 8 + (k & 7);
    dst[idx] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the fused step kernel
// ---------------------------------------------------------------------------------------------------------------
struct StepParams {
  NetPlan plan;
  const uint8_t* policy_img; const uint8_t* model_img;
  const float* cur; const int* n_dev; int64_t n_max;
  const float* eps_p;                    // [n][4]   this step's policy noise, row-compact (noise_stage_kernel)
  const float* eps_m;                    // [n][NMG] this step's model noise
  float *actions, *next_states, *rewards;
  const float *norm_mean, *norm_std, *min_lv, *max_lv;
  int S, A, SP, OP, NM, NMG, stages;
  int* err_flag;
  int dump_layer; float* dump_out;       // debug: dump the fp32 accumulator of one layer
};

struct SmemLayout {
  // full_bar: [0..3] hid_full[pair][slot] (accumulator full -> the pair that drains it), [4..6] sp_full[k] (-> group C)
  // free_bar: [0..2] acc_free[b] (hidden chunk drained and stored), [3..5] sp_free[k] (output chunk fully done)
  uint64_t full[MAX_STAGES], empty[MAX_STAGES], full_bar[4 + NSPECIAL], free_bar[NACC + NSPECIAL], tile_ready, token[2];
  uint32_t tmem_base, pad[3];
  // per-dim constants of the member, staged once per CTA: normaliser, and the log-var soft clamp folded into
  //   std = exp(lv/2) = s0 * sqrt(1 + E / (1 + exp(hi - x)))   with s0 = exp(lo/2), E = exp(hi - lo)      (src/dynamics.py:120-121,201)
  float norm_mean[64], norm_inv[64], lv_hi[64], lv_E[64], lv_s0[64];
  uint4 irec[MAX_CHUNKS], erec[MAX_CHUNKS];
};

// debug timing (kDebug build, dump_layer == 100): CTA 0 stamps clock() for its first 4 tiles into dump_out viewed as uint32
// [(tile*32 + chunk)*8 + k], k: 0 epilogue wait begin, 1 accumulator ready, 2 epilogue done, 3 mma deps ok, 4 weights ready,
// 5 mma issued, 6 tile prologue begin (chunk 0 only), 7 tile prologue end
template <bool kDebug>
__device__ __forceinline__ void stamp(const StepParams& p, uint32_t tile_it, int chunk, int k) {
  if (kDebug) {
    if (p.dump_layer == 100 && blockIdx.x == 0 && tile_it < 4)
      reinterpret_cast<uint32_t*>(p.dump_out)[(tile_it * 32 + chunk) * 8 + k] = (uint32_t)clock();
  }
}

// write one input row (k_real values + constant 1) as packed bf16 into a TMEM activation buffer
__device__ __forceinline__ void write_input_row(uint32_t lane_base, uint32_t col, const float* row, int kp) {
  for (int e0 = 0; e0 < kp; e0 += 16) {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(row[e0 + 2 * j], row[e0 + 2 * j + 1]);    // rows are staged with their [.., 1, 0, ..] tail up to kp
    tmem_st8(lane_base + col + (uint32_t)(e0 >> 1), pk);
  }
}

// 16 accumulator columns -> activation -> 8 packed bf16x2 words
template <bool kSilu>
__device__ __forceinline__ void act_pack16(const uint32_t (&r)[16], uint32_t (&pk)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t x = pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
    pk[j] = kSilu ? silu_bf16x2(x) : relu_bf16x2(x);
  }
}
// force element `one` (0..15) of a packed piece to 1.0: the bias slot of the consumer layer
__device__ __forceinline__ void patch_one(uint32_t (&pk)[8], int one) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (one == 2 * j) pk[j] = (pk[j] & 0xFFFF0000u) | 0x00003F80u;
    if (one == 2 * j + 1) pk[j] = (pk[j] & 0x0000FFFFu) | 0x3F800000u;
  }
}

// epilogue of one column half (hn = 16 or 32 columns starting at layer column h0) of a hidden-layer chunk, one row per thread:
// ACC[acc_col + j] -> activation -> packed bf16 -> TMEM out_col + (h0 + j)/2.  Both loads are issued before the single wait.
template <bool kSilu>
__device__ __forceinline__ void hidden_half(uint32_t lane_base, uint32_t acc_col, uint32_t out_col, int n_real, int np, int next_kp, int h0, int hn) {
  uint32_t ra[16], rb[16];
  const uint32_t src = lane_base + acc_col;
  tmem_ld16(src, ra);
  if (hn > 16) tmem_ld16(src + 16, rb);
  tmem_ld_wait();
  const int one = n_real - h0;                         // position of the consumer's bias slot relative to this half
  const uint32_t dst = lane_base + out_col + (uint32_t)(h0 >> 1);
  uint32_t pa[8], pb[8];
  act_pack16<kSilu>(ra, pa);
  if (one >= 0 && one < 16) patch_one(pa, one);          // rare: only the piece that holds column n_real
  tmem_st8(dst, pa);
  if (hn > 16) {
    act_pack16<kSilu>(rb, pb);
    if (one >= 16 && one < 32) patch_one(pb, one - 16);
    tmem_st8(dst + 8, pb);
  }
  if (h0 + hn == np && next_kp > np) {                   // rare: constant tail [np, next_kp): 1 at n_real, else 0
    for (int e0 = np; e0 < next_kp; e0 += 16) {
      uint32_t t8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t8[j] = pack_bf16((e0 + 2 * j) == n_real ? 1.f : 0.f, (e0 + 2 * j + 1) == n_real ? 1.f : 0.f);
      tmem_st8(lane_base + out_col + (uint32_t)(e0 >> 1), t8);
    }
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

template <bool kDebug>
__global__ void __launch_bounds__(NUM_THREADS, 1) rollout_step_umma_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const NetPlan& plan = p.plan;
  // shared memory: [weight ring | raw states x2 | model input / next state rows x2 | policy noise x2 | model noise x2 | barriers, constants]
  uint8_t* ring = smem_raw;
  const uint32_t slot_bytes = (plan.max_chunk_bytes + 1023u) & ~1023u;
  float* st_s = reinterpret_cast<float*>(ring + (size_t)slot_bytes * p.stages);      // [2][128][SP] raw states (fp32)
  float* st_o = st_s + 2 * TILE_M * p.SP;                                            // [2][128][OP] [norm s, a, 1, 0..] then [next state, reward]
  float* st_np = st_o + 2 * TILE_M * p.OP;                                           // [2][128][4]  policy noise
  float* st_nm = st_np + 2 * TILE_M * 4;                                             // [2][128][NM] model noise (NM = 0: read from global)
  SmemLayout* sl = reinterpret_cast<SmemLayout*>(st_nm + 2 * TILE_M * p.NM);

  const int n = (int)min((int64_t)*p.n_dev, p.n_max);
  const int n_tiles = (n + TILE_M - 1) / TILE_M;
  const int S = p.S, A = p.A, O = S + 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&sl->full[s], 1); mbar_init(&sl->empty[s], 1); }
    for (int k = 0; k < 4 + NSPECIAL; ++k) mbar_init(&sl->full_bar[k], 1);
    for (int b = 0; b < NACC; ++b) mbar_init(&sl->free_bar[b], 2 * GROUP_THREADS);
    for (int k = 0; k < NSPECIAL; ++k) mbar_init(&sl->free_bar[NACC + k], GROUP_THREADS);
    mbar_init(&sl->tile_ready, GROUP_THREADS);
    mbar_init(&sl->token[0], 1); mbar_init(&sl->token[1], 1);
    fence_barrier_init();
  }
  if (threadIdx.x < EPI_THREADS) {
    // constants + the constant tails of the staging rows: [.., 1, 0, 0 ..] = bias slot and K padding of the two input layers
    const int et0 = threadIdx.x;
    if (et0 < S) { sl->norm_mean[et0] = p.norm_mean[et0]; sl->norm_inv[et0] = 1.f / (p.norm_std[et0] + 1e-6f); }
    if (et0 <= S) {
      const float lo = p.min_lv[et0], hi = p.max_lv[et0];
      sl->lv_hi[et0] = hi; sl->lv_E[et0] = __expf(hi - lo); sl->lv_s0[et0] = __expf(0.5f * lo);
    }
    if (et0 < MAX_CHUNKS) { sl->irec[et0] = plan.irec[et0]; sl->erec[et0] = plan.erec[et0]; }
    for (int i = et0; i < 2 * TILE_M * p.SP; i += EPI_THREADS) { const int c = i % p.SP; st_s[i] = c == S ? 1.f : 0.f; }
    for (int i = et0; i < 2 * TILE_M * p.OP; i += EPI_THREADS) { const int c = i % p.OP; st_o[i] = c == S + A ? 1.f : 0.f; }
  }
  if (warp == MMA_WARP) tmem_alloc(&sl->tmem_base, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sl->tmem_base;

  if (warp == PRODUCER_WARP) {
    // ===================== TMA producer: stream every weight chunk of every tile through the ring =====================
    if (lane == 0) {
      int s = 0; uint32_t round = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int c = 0; c < plan.n_chunks; ++c) {
          if (round > 0) mbar_wait(&sl->empty[s], (round - 1) & 1, p.err_flag, 1);
          const ChunkSpec& ch = plan.chunk[c];
          const uint8_t* src = (c < plan.n_policy_chunks ? p.policy_img : p.model_img) + ch.offset;
          mbar_expect_tx(&sl->full[s], ch.bytes);
          bulk_g2s(ring + (size_t)s * slot_bytes, src, ch.bytes, &sl->full[s]);
          if (++s == p.stages) { s = 0; ++round; }
        }
      }
    }
  } else if (warp == MMA_WARP || warp == MMA_WARP2) {
    // ===================== two MMA issuers =====================
    // The tensor pipe accepts only a few tcgen05.mma ahead of execution, so an issuing thread is blocked for the duration of its
    // chunk and cannot overlap the per-chunk bookkeeping (barrier waits, record load, commits: ~500 cycles of latency) with it.
    // Two issuers alternate chunks: while one is blocked feeding chunk g, the other clears the waits of chunk g+1.  A token
    // keeps the ISSUE order equal to the chunk order (every in-order argument in this file relies on it), and both threads
    // replay the complete (static) schedule so that each one knows the phase parity of every barrier it waits on.
    if (elect_one()) {
      const uint32_t me = warp == MMA_WARP ? 0u : 1u;
      int s = 0; uint32_t ring_par = 0;      // weight ring slot + parity of its current use
      // Accumulator buffer b = chunk index % NACC (n_chunks is a multiple of NACC).  Who drains the buffer's current contents is
      // own (2 bits per buffer: 0 the hidden groups -> free_bar[b], 1+k group C -> free_bar[NACC+k]).  Each free barrier has at
      // most one phase outstanding: pend bit = a phase was started and not yet observed, par bit = its parity (bit = barrier id).
      // The code of this loop is kept small on purpose (one instantiation of every wait and of the unrolled MMA sequence): it
      // is the pipeline's serial section, and its instruction footprint competes with 23 other warps for a 32 KB L1.5 I-cache.
      uint32_t own = 0, pend = 0, par = 0, hcnt = 0;      // hcnt: per-pair count of hidden chunks issued (8 bits each)
      uint32_t g = 0;                        // global chunk counter: chunk g belongs to issuer (g & 1)
      uint32_t tok_par = 0;                  // parity of the next phase of the OTHER issuer's token
      const uint32_t free_base = smem_u32(&sl->free_bar[0]), full_base = smem_u32(&sl->full_bar[0]);
      auto drain = [&](int b, bool mine) {
        const uint32_t o = (own >> (2 * b)) & 3u;
        const uint32_t id = o == 0 ? (uint32_t)b : NACC - 1 + o;
        if ((pend >> id) & 1u) {
          if (mine) { mbar_wait_addr(free_base + 8u * id, (par >> id) & 1u, p.err_flag, 6); tc_fence_after(); }
          pend ^= 1u << id; par ^= 1u << id;
        }
      };
      const int n_chunks = plan.n_chunks;
      uint32_t tile_it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
        mbar_wait(&sl->tile_ready, tile_it & 1, p.err_flag, 2);
        tc_fence_after();
        int b = 0;
        for (int c = 0; c < n_chunks; ++c, ++g) {
          const uint4 rec = sl->irec[c];
          const bool mine = (g & 1u) == me;
          drain(b, mine);                                                 // accumulator buffer still being read by its last user?
          const int sp = (int)((rec.z >> 24) & 0x3Fu) - 1;
          const uint32_t pr = (rec.z >> 30) & 1u;                         // group pair that drains a hidden chunk
          // First chunk of a layer whose A operand is still being produced: the producer layer's chunks drain out of order across
          // the two group pairs, so its last TWO chunks are tracked (older ones precede them in their pair's program order and
          // their buffers were reused since).  K-sliced start: k-steps that only read columns below chunk (last-1) issue at
          // once, the rest wait for (last-1), then (last).
          const int bufA = (int)((rec.w >> 16) & 0xFu) - 1, bufB = (int)((rec.w >> 20) & 0xFu) - 1;
          if (mine) {
            if (me == 0) stamp<kDebug>(p, tile_it, c, 3);
            mbar_wait(&sl->full[s], ring_par, p.err_flag, 3);
            tc_fence_after();
            const uint32_t b_base = smem_u32(ring + (size_t)s * slot_bytes);
            const uint32_t idesc = rec.x, desc_hi = rec.y;
            const uint32_t desc_lo = ((b_base & 0x3FFFFu) >> 4) | (8u << 16);   // LBO = 128 B (K-direction core-matrix stride)
            const uint32_t a_addr = tmem + (rec.z & 0xFFFFu);
            const uint32_t d_addr = tmem + TM_ACC + (uint32_t)b * NSLAB;
            const int nk = (int)((rec.z >> 16) & 0xFFu);
            const int kB = min((int)((rec.w >> 8) & 0xFFu), nk), kA = min((int)(rec.w & 0xFFu), kB);
            if (g > 0) { mbar_wait(&sl->token[me ^ 1u], tok_par, p.err_flag, 8); tok_par ^= 1u; }   // chunk g-1 has been issued
            if (me == 0) stamp<kDebug>(p, tile_it, c, 4);
            // back-to-back UTCHMMA for k-steps [lo, hi); one k-step = 2 core matrices = 256 B = +16 in the address field,
            // +8 TMEM columns.  Fully unrolled with compile-time-simple uniform predicates: the issue cost per MMA is very
            // sensitive to the code shape (a rolled loop, or run-time range bounds inside a rolled segment loop, issue an order
            // of magnitude slower than this).
            auto issue_range = [&](int lo, int hi) {
#pragma unroll
              for (int k = 0; k < 17; ++k) {
                if (k >= lo && k < hi) {
                  if (k == 0) mma_ts<false>(d_addr, a_addr, desc_lo, desc_hi, idesc);
                  else mma_ts<true>(d_addr, a_addr + 8u * k, desc_lo + 16u * k, desc_hi, idesc);
                }
              }
            };
            if (rec.w == 0) {
              issue_range(0, nk);
            } else {
              issue_range(0, kA);
              if (bufA >= 0) drain(bufA, true);
              issue_range(kA, kB);
              if (bufB >= 0) drain(bufB, true);
              issue_range(kB, nk);
            }
            tc_commit(&sl->empty[s]);                                    // frees the ring slot when these MMAs retire
            // "accumulator full" goes to the pair that drains this chunk: slot = parity of the pair's own chunk count, so each
            // pair observes every phase of its two barriers, in order, and nobody else waits on them (a parity wait is only
            // sound while the waiter is at most one phase behind)
            tc_commit_addr(full_base + 8u * (sp < 0 ? 2u * pr + ((hcnt >> (8 * pr)) & 1u) : 4u + (uint32_t)sp));
            mbar_arrive(&sl->token[me]);                                  // chunk g is issued: the other issuer may issue g+1
            if (me == 0) stamp<kDebug>(p, tile_it, c, 5);
          } else {
            if (bufA >= 0) drain(bufA, false);                            // replay the other issuer's K-sliced waits
            if (bufB >= 0) drain(bufB, false);
          }
          own &= ~(3u << (2 * b));
          if (sp < 0) { pend |= 1u << b; hcnt = (hcnt & ~(0xFFu << (8 * pr))) | ((((hcnt >> (8 * pr)) + 1u) & 0xFFu) << (8 * pr)); }
          else { pend |= 1u << (NACC + sp); own |= (uint32_t)(1 + sp) << (2 * b); }
          if (++b == NACC) b = 0;
          if (++s == p.stages) { s = 0; ring_par ^= 1u; }
        }
      }
    }
  } else if (warp >= GROUP_H_WARP0 && warp < GROUP_H_WARP0 + 4 * N_HID_GROUPS) {
    // ===================== hidden-layer epilogue groups (4 groups of 4 warps) =====================
    const int grp = (warp - GROUP_H_WARP0) >> 2, q = warp & 3;
    const int pair = grp >> 1, half = grp & 1;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const bool lead = kDebug && (threadIdx.x & (GROUP_THREADS - 1)) == 0 && half == 0;
    uint32_t m = 0;                                                       // hidden chunks this pair has taken so far
    const int n_chunks_h = plan.n_chunks;
    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      int b = 0;
      for (int c = 0; c < n_chunks_h; ++c) {
        const uint4 rec = sl->erec[c];
        if ((rec.y >> 25) & 1u) {                                         // hidden-layer chunk
          const bool mine = ((rec.y >> 24) & 1u) == (uint32_t)pair;
          if (mine) {
            if (lead) stamp<kDebug>(p, tile_it, c, 0);
            // the pair's m-th chunk arrives on hid_full[pair][m & 1] as phase m >> 1: this pair sees every phase, in order
            mbar_wait(&sl->full_bar[2 * pair + (m & 1u)], (m >> 1) & 1u, p.err_flag, 4);
            ++m;
            const uint32_t acc_col = TM_ACC + (uint32_t)b * NSLAB;
            tc_fence_after();
            if (lead) stamp<kDebug>(p, tile_it, c, 1);
            const int n0 = (int)(rec.x & 0xFFFFu), nc = (int)(rec.x >> 16);
            const uint32_t out_col = rec.y & 0xFFFFu;
            const int n_real = (int)(rec.z & 0xFFFFu), np = (int)(rec.z >> 16), next_kp = (int)rec.w;
            // column halves: 64 -> 32|32, 48 -> 32|16, 32 -> 16|16, 16 -> 16|0
            const int w0 = (((nc >> 4) + 1) >> 1) << 4;
            const int h0 = n0 + (half ? w0 : 0), hn = half ? nc - w0 : w0;
            if (kDebug && p.dump_layer == (int)plan.chunk[c].layer) {     // debug hook: raw accumulator to global
              const int64_t row0 = (int64_t)tile * TILE_M; const int t = q * 32 + lane;
              for (int c0 = 0; c0 < hn; c0 += 16) {
                uint32_t r[16]; tmem_ld16(lane_base + acc_col + (uint32_t)(h0 - n0 + c0), r); tmem_ld_wait();
                if (t < n - (int)row0) for (int j = 0; j < 16; ++j) if (h0 + c0 + j < n_real) p.dump_out[(row0 + t) * n_real + h0 + c0 + j] = __uint_as_float(r[j]);
              }
            }
            if (hn > 0) {
              if (((rec.y >> 16) & 0xFFu) == HID_SILU) hidden_half<true>(lane_base, acc_col + (uint32_t)(h0 - n0), out_col, n_real, np, next_kp, h0, hn);
              else hidden_half<false>(lane_base, acc_col + (uint32_t)(h0 - n0), out_col, n_real, np, next_kp, h0, hn);
              tmem_st_wait();
            }
            tc_fence_before();
            mbar_arrive(&sl->free_bar[b]);                                // accumulator drained, activations visible
            if (lead) stamp<kDebug>(p, tile_it, c, 2);
          }
        }
        if (++b == NACC) b = 0;
      }
    }
  } else if (warp < GROUP_C_WARP0 + 4) {
    // ===================== group C (warps 0-3): output chunks, next tile's prologue =====================
    const int q = warp & 3;
    const int t = q * 32 + lane;                                          // trajectory row of the tile == TMEM lane
    const int ct = threadIdx.x - GROUP_C_WARP0 * 32;                      // 0..127
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const int kp0 = plan.layer[0].kp;
    const int first_model = plan.n_policy_chunks;                        // chunk index of trunk0's first chunk
    const int kpm = plan.layer[plan.chunk[first_model].layer].kp;
    const uint32_t xm_col = (uint32_t)plan.layer[plan.chunk[first_model].layer].a_col;
    const uint32_t xp_col = (uint32_t)plan.layer[0].a_col;
    // The policy input of tile i+1 is staged while tile i is still in its model phase when it fits the 8 spare TMEM columns
    // (state_dim <= 15); wider inputs share columns with the log-var head's activations and are staged once those are dead.
    const bool early_prologue = xp_col == TM_XP;

    // Program of group C: one prologue per tile (tile 0 before its chunks, tile i+1 inside tile i), three output epilogues per
    // tile, one prefetch per tile.  A single loop with pending-work flags keeps one copy of each piece of code (instruction cache).
    uint32_t tile_it = 0;
    int tile = blockIdx.x;
    int pf_tile = tile < n_tiles ? tile : -1, pf_buf = 0;                 // pending prefetch (cp.async of a tile's states and noise)
    bool need_prologue = tile < n_tiles;                                  // pending prologue of `ptile` into buffer `pbuf`
    int ptile = tile, pbuf = 0; uint32_t pit = 0;
    int c = -1, b = NACC - 1;                                             // chunk cursor of the current tile (-1: before the first chunk)
    while (tile < n_tiles) {
      if (pf_tile >= 0) {
        const int64_t r0 = (int64_t)pf_tile * TILE_M;
        const int rws = min(TILE_M, n - (int)r0);
        float* dst = st_s + pf_buf * TILE_M * p.SP;
        for (int i = ct; i < rws * S; i += GROUP_THREADS) { const int r = i / S, cc = i - r * S; cp_async4(dst + r * p.SP + cc, p.cur + r0 * S + i); }
        if (ct < rws) cp_async16(st_np + (pf_buf * TILE_M + ct) * 4, p.eps_p + (r0 + ct) * 4);
        if (p.NM > 0) {
          const int pieces = p.NMG >> 2;
          for (int i = ct; i < rws * pieces; i += GROUP_THREADS) { const int r = i / pieces, cc = i - r * pieces; cp_async16(st_nm + (pf_buf * TILE_M + r) * p.NM + 4 * cc, p.eps_m + (r0 + r) * p.NMG + 4 * cc); }
        }
        cp_async_commit();
        pf_tile = -1;
      }
      if (need_prologue) {
        // ---- tile prologue: policy input [s, 1] -> TMEM, normalised state -> model-input row ----
        if (ct == 0) stamp<kDebug>(p, pit, 0, 6);
        cp_async_wait_all();
        named_bar_sync(1, GROUP_THREADS);                                 // the tile's states and noise landed (all of group C's copies)
        const float* ps = st_s + pbuf * TILE_M * p.SP + t * p.SP;
        float* po = st_o + pbuf * TILE_M * p.OP + t * p.OP;
        write_input_row(lane_base, xp_col, ps, kp0);
        for (int cc = 0; cc < S; ++cc) po[cc] = (ps[cc] - sl->norm_mean[cc]) * sl->norm_inv[cc];        // src/dynamics.py:113
        tmem_st_wait(); tc_fence_before(); mbar_arrive(&sl->tile_ready);
        if (ct == 0) stamp<kDebug>(p, pit, 0, 7);
        need_prologue = false;
        if (ptile == tile && c < 0) {                                     // that was the first tile's own prologue: prefetch its successor
          const int nt = tile + (int)gridDim.x;
          if (nt < n_tiles) { pf_tile = nt; pf_buf = 1; }
        }
        continue;
      }
      // ---- advance to the tile's next output chunk ----
      do { ++c; if (++b == NACC) b = 0; } while (c < plan.n_chunks && plan.chunk[c].special < 0);
      if (c >= plan.n_chunks) {                                           // tile finished
        tile += gridDim.x; ++tile_it; c = -1; b = NACC - 1;
        const int nt = tile + (int)gridDim.x;
        if (tile < n_tiles && nt < n_tiles) { pf_tile = nt; pf_buf = (tile_it & 1) ^ 1; }   // buffers were last read by the finished tile's epilogues
        continue;
      }
      const int buf = tile_it & 1;
      const int64_t row0 = (int64_t)tile * TILE_M;
      const int rows = min(TILE_M, n - (int)row0);
      const bool valid = t < rows;
      const int64_t row = row0 + t;
      const float* my_s = st_s + buf * TILE_M * p.SP + t * p.SP;
      float* my_o = st_o + buf * TILE_M * p.OP + t * p.OP;
      const int next_tile = tile + (int)gridDim.x;
      const ChunkSpec& ch = plan.chunk[c];
      const LayerSpec& L = plan.layer[ch.layer];
      const uint32_t acc_col = TM_ACC + (uint32_t)b * NSLAB;
      const int sp = ch.special;
      if (ct == 0) stamp<kDebug>(p, tile_it, c, 0);
      mbar_wait(&sl->full_bar[4 + sp], tile_it & 1, p.err_flag, 7);
      tc_fence_after();
      if (ct == 0) stamp<kDebug>(p, tile_it, c, 1);
      if (kDebug && p.dump_layer == (int)ch.layer) {                      // debug hook: raw accumulator to global
        for (int c0 = 0; c0 < ch.nc; c0 += 16) {
          uint32_t r[16]; tmem_ld16(lane_base + acc_col + (uint32_t)c0, r); tmem_ld_wait();
          if (valid) for (int j = 0; j < 16; ++j) if (ch.n0 + c0 + j < L.n_real) p.dump_out[row * L.n_real + ch.n0 + c0 + j] = __uint_as_float(r[j]);
        }
      }
      if (L.kind == OUT_POLICY) {
        // ---- policy head: [mu, raw] -> a = tanh(mu + exp(-6 + 10 sigmoid(raw)) eps)      src/policy.py:89-97 ----
        uint32_t r[16]; tmem_ld16(lane_base + acc_col, r); tmem_ld_wait();
        const float4 e4 = *reinterpret_cast<const float4*>(st_np + (buf * TILE_M + t) * 4);
        const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
        const float mu4[4] = {__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3])};
        float raw4[4] = {0.f, 0.f, 0.f, 0.f};                              // raw_j = out[A + j], statically indexed per A
        if (A == 1) { raw4[0] = __uint_as_float(r[1]); }
        else if (A == 2) { raw4[0] = __uint_as_float(r[2]); raw4[1] = __uint_as_float(r[3]); }
        else if (A == 3) { raw4[0] = __uint_as_float(r[3]); raw4[1] = __uint_as_float(r[4]); raw4[2] = __uint_as_float(r[5]); }
        else { raw4[0] = __uint_as_float(r[4]); raw4[1] = __uint_as_float(r[5]); raw4[2] = __uint_as_float(r[6]); raw4[3] = __uint_as_float(r[7]); }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < A) {
            const float sd = __expf(-6.f + __fdividef(10.f, 1.f + __expf(-raw4[j])));
            const float a = tanh_fast(fmaf(ev[j], sd, mu4[j]));
            my_o[S + j] = a;
            if (valid) p.actions[row * A + j] = a;
          }
        }
        // model input x0 = [(s - mean)/(std + 1e-6), a, 1]: the normalised part was written by the prologue  (src/dynamics.py:113-114)
        write_input_row(lane_base, xm_col, my_o, kpm);
        tmem_st_wait(); tc_fence_before(); mbar_arrive(&sl->free_bar[NACC + sp]);
        if (ct == 0) stamp<kDebug>(p, tile_it, c, 2);
        // off the critical path: the next tile's prologue
        if (early_prologue && next_tile < n_tiles) { need_prologue = true; ptile = next_tile; pbuf = buf ^ 1; pit = tile_it + 1; }
      } else if (L.kind == OUT_DIFF) {
        // ---- diff head: means = diffs + [s, 0]  (kept in shared memory)                   src/dynamics.py:118 ----
        for (int c0 = 0; c0 < ch.nc; c0 += 16) {
          uint32_t r[16]; tmem_ld16(lane_base + acc_col + (uint32_t)c0, r);
          float sv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) sv[j] = (c0 + j < S) ? my_s[c0 + j] : 0.f;
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) if (c0 + j < O) my_o[c0 + j] = __uint_as_float(r[j]) + sv[j];
        }
        tc_fence_before(); mbar_arrive(&sl->free_bar[NACC + sp]);
        if (ct == 0) stamp<kDebug>(p, tile_it, c, 2);
      } else {
        // ---- log-var head + Gaussian sample                                              src/dynamics.py:119-121,201-203 ----
        for (int c0 = 0; c0 < ch.nc; c0 += 16) {
          uint32_t r[16]; tmem_ld16(lane_base + acc_col + (uint32_t)c0, r);
          // loads first, math second, stores last: the shared-memory stores of one column never fence the next column's loads
          float ev[16], res[16];
#pragma unroll
          for (int jg = 0; jg < 4; ++jg) {
            const int cg = (c0 >> 2) + jg;
            float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * cg < O) {
              if (p.NM > 0) e4 = *reinterpret_cast<const float4*>(st_nm + (buf * TILE_M + t) * p.NM + 4 * cg);
              else if (valid) e4 = *reinterpret_cast<const float4*>(p.eps_m + row * p.NMG + 4 * cg);
            }
            ev[4 * jg] = e4.x; ev[4 * jg + 1] = e4.y; ev[4 * jg + 2] = e4.z; ev[4 * jg + 3] = e4.w;
          }
          tmem_ld_wait();
          if (c0 + 16 >= ch.nc) { tc_fence_before(); mbar_arrive(&sl->free_bar[NACC + sp]); }        // accumulator in registers: release it early
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int cc = min(c0 + j, O - 1);
            const float u = __expf(sl->lv_hi[cc] - __uint_as_float(r[j]));
            res[j] = fmaf(sl->lv_s0[cc] * sqrt_fast(1.f + __fdividef(sl->lv_E[cc], 1.f + u)), ev[j], my_o[cc]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) if (c0 + j < O) my_o[c0 + j] = res[j];
        }
        if (valid) p.rewards[row] = my_o[S];
        named_bar_sync(2, GROUP_THREADS);                                // every row of the tile is final in st_o
        const float* so = st_o + buf * TILE_M * p.OP;
        for (int i = ct; i < rows * S; i += GROUP_THREADS) {              // coalesced store of the tile's next states
          const int r = i / S, cc = i - r * S;
          p.next_states[row0 * S + i] = so[r * p.OP + cc];
        }
        named_bar_sync(2, GROUP_THREADS);                                // st_o[buf] is rewritten two prologues from now by other threads
        my_o[S + A] = 1.f;                                               // restore the bias slot if the reward column overwrote it (A == 0 never)
        if (ct == 0) stamp<kDebug>(p, tile_it, c, 2);
        // wide policy inputs: the log-var head's MMAs are complete, its activations are dead -> stage the next tile now
        if (!early_prologue && next_tile < n_tiles) { need_prologue = true; ptile = next_tile; pbuf = buf ^ 1; pit = tile_it + 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem, TM_COLS);
}

// This step's noise, row-compact: Philox(seed, global trajectory id) or the injected draws gathered by trajectory id.
__global__ void __launch_bounds__(256) noise_stage_kernel(const int32_t* __restrict__ ids, const int* n_dev, int64_t n_max, NoiseView np_, NoiseView nm_,
                                                          int A, int O, int NMG, float* __restrict__ eps_p, float* __restrict__ eps_m) {
  const int64_t n = min((int64_t)*n_dev, n_max);
  const int pieces = 1 + (NMG >> 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * pieces; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / pieces; const int c = (int)(i - r * pieces);
    const int64_t id = ids[r];
    if (c == 0) *reinterpret_cast<float4*>(eps_p + r * 4) = noise_get4(np_, id, 0, A);
    else *reinterpret_cast<float4*>(eps_m + r * NMG + 4 * (c - 1)) = noise_get4(nm_, id, c - 1, O);
  }
}

// check_done / check_violation / get_constraint_values of the step's next states (src/smbpo.py:238-240) fused with
// buffer.extend: the step's rows go into the ring at (base + r) % capacity (src/sampling.py:128-145, src/smbpo.py:241-242,248).
__global__ void __launch_bounds__(256) hooks_store_kernel(drpo_env_params env, drpo_buffer buf, const RolloutState* st, const int32_t* n_dev,
                                                          const float* __restrict__ s, const float* __restrict__ a, const float* __restrict__ ns,
                                                          const float* __restrict__ rew, uint8_t* __restrict__ done_out) {
  const int64_t n = *n_dev, base = st->base, cap = buf.capacity;
  const int S = buf.state_dim, A = buf.action_dim, C = buf.con_dim;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = t0; i < n * S; i += stride) {
    const int64_t r = i / S; const int c = (int)(i - r * S); const int64_t slot = (base + r) % cap;
    buf.states[slot * S + c] = s[i];
    buf.next_states[slot * S + c] = ns[i];
  }
  for (int64_t i = t0; i < n * A; i += stride) {
    const int64_t r = i / A; const int c = (int)(i - r * A);
    buf.actions[((base + r) % cap) * A + c] = a[i];
  }
  for (int64_t r = t0; r < n; r += stride) {
    const float* row = ns + r * S;
    HookOut ho;
    eval_hooks(env, [row](int d) { return row[d]; }, ho);
    const int64_t slot = (base + r) % cap;
    buf.rewards[slot] = rew[r]; buf.dones[slot] = ho.done; buf.violations[slot] = ho.viol;
    done_out[r] = ho.done;
#pragma unroll
    for (int cc = 0; cc < DRPO_MAX_CON; ++cc) if (cc < C) buf.constraint_values[slot * C + cc] = ho.cv[cc];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static void add_layer(NetPlan& P, int n_real, int k_real, int kind, int a_col, int out_col, int dep, int special, int& hid, uint32_t& img_off) {
  const int l = P.n_layers++;
  LayerSpec& L = P.layer[l];
  L.n_real = n_real; L.k_real = k_real; L.np = round_up(n_real, 16); L.kp = round_up(k_real + 1, 16);
  L.kind = kind; L.a_col = a_col; L.out_col = out_col; L.first_chunk = P.n_chunks; L.n_chunks = 0; L.dep = dep;
  L.next_kp = (kind == HID_RELU || kind == HID_SILU) ? round_up(n_real + 1, 16) : 0;
  // balanced chunks of <= NSLAB columns in units of 16 (208 -> 64,48,48,48; 256 -> 4 x 64)
  const int units = L.np / 16, nch = (L.np + NSLAB - 1) / NSLAB;
  int n0 = 0;
  for (int i = 0; i < nch; ++i) {
    const int u = units / nch + (i < units % nch ? 1 : 0);
    ChunkSpec& c = P.chunk[P.n_chunks++];
    c.n0 = (uint16_t)n0; c.nc = (uint16_t)(16 * u); c.layer = (uint16_t)l;
    c.special = (int8_t)special; c.hid = (uint8_t)(special < 0 ? hid++ : 0);
    c.offset = img_off + (uint32_t)n0 * L.kp * 2; c.bytes = (uint32_t)c.nc * L.kp * 2;
    P.max_chunk_bytes = std::max(P.max_chunk_bytes, c.bytes);
    ++L.n_chunks;
    n0 += 16 * u;
  }
  img_off += (uint32_t)L.np * L.kp * 2;
}

static int build_plan(const drpo_rollout_args& a, NetPlan& P) {
  memset(&P, 0, sizeof(P));
  const int S = a.ensemble->state_dim, A = a.ensemble->action_dim, Hm = a.ensemble->hidden;
  const int Hp = a.actor->l0.out_dim;
  if (a.actor->l1.out_dim != Hp || round_up(Hp + 1, 16) > 272 || round_up(Hm + 1, 16) > 208 || round_up(S + A + 1, 16) > 64 ||
      2 * A > 16 || A > 4 || round_up(S + 1, 16) > NSLAB) {
    set_error("bf16 rollout: dims outside the TMEM plan (S=%d A=%d actor hidden=%d model hidden=%d)", S, A, Hp, Hm);
    return DRPO_ERR_UNSUPPORTED;
  }
  // policy input: the 8 spare columns when it fits (tile i+1 is staged during tile i's model phase), else the columns that are
  // free during the policy phase only
  const int xp = round_up(S + 1, 16) <= 16 ? (int)TM_XP : 464;
  uint32_t off = 0; int hid = 0;
  add_layer(P, Hp, S, HID_RELU, xp, TM_PA, -1, -1, hid, off);           // 0 actor L0
  add_layer(P, Hp, Hp, HID_RELU, TM_PA, TM_PB, 0, -1, hid, off);        // 1 actor L1
  add_layer(P, 2 * A, Hp, OUT_POLICY, TM_PB, 0, 1, 0, hid, off);        // 2 actor L2 -> policy head (writes the model input)
  P.policy_bytes = off; P.n_policy_chunks = P.n_chunks;
  off = 0;
  add_layer(P, Hm, S + A, HID_SILU, TM_L1, TM_D1, 2, -1, hid, off);     // 3 trunk0: x_m lives in the (still unused) l1 region
  add_layer(P, Hm, Hm, HID_SILU, TM_D1, TM_H2, 3, -1, hid, off);        // 4 trunk1 -> h2, kept for both heads
  add_layer(P, Hm, Hm, HID_SILU, TM_H2, TM_D1, 4, -1, hid, off);        // 5 diff head hidden
  add_layer(P, Hm, Hm, HID_SILU, TM_H2, TM_L1, 4, -1, hid, off);        // 6 log-var head hidden (independent of 5: no bubble)
  add_layer(P, S + 1, Hm, OUT_DIFF, TM_D1, 0, 5, 1, hid, off);          // 7 diffs
  add_layer(P, S + 1, Hm, OUT_LOGVAR, TM_L1, 0, 6, 2, hid, off);        // 8 log-vars
  P.model_bytes = off;
  for (int l = 0; l < P.n_layers; ++l) {
    const LayerSpec& L = P.layer[l];
    for (int c = 0; c < L.n_chunks; ++c) {
      const int ci = L.first_chunk + c; const ChunkSpec& ch = P.chunk[ci];
      const uint32_t sbo = (uint32_t)(L.kp >> 3) * 128u;
      uint4 ir, er;
      ir.x = make_idesc(ch.nc);
      ir.y = ((sbo >> 4) & 0x3FFFu) | (1u << 14);                       // descriptor bits 32..45 = SBO, bit 46 = version 1
      ir.z = (uint32_t)L.a_col | ((uint32_t)(L.kp >> 4) << 16) | ((uint32_t)(ch.special + 1) << 24) | ((uint32_t)(ch.hid & 1) << 30);
      ir.w = 0;
      if (c == 0 && L.dep >= 0) {
        const LayerSpec& D = P.layer[L.dep];
        const int last = D.first_chunk + D.n_chunks - 1;
        uint32_t kA = 0, kB = 0, bufA = 0, bufB = 0;
        if (ci - last <= NACC) { bufB = 1 + last % NACC; kB = P.chunk[last].n0 >> 4; }
        if (D.n_chunks > 1 && ci - (last - 1) <= NACC) { bufA = 1 + (last - 1) % NACC; kA = P.chunk[last - 1].n0 >> 4; }
        if (!bufB) kB = kA;                                               // (cannot happen: last is younger than last-1)
        ir.w = kA | (kB << 8) | (bufA << 16) | (bufB << 20);
      }
      er.x = (uint32_t)ch.n0 | ((uint32_t)ch.nc << 16);
      er.y = (uint32_t)L.out_col | ((uint32_t)L.kind << 16) | ((uint32_t)(ch.hid & 1) << 24) | ((ch.special < 0 ? 1u : 0u) << 25);
      er.z = (uint32_t)L.n_real | ((uint32_t)L.np << 16);
      er.w = (uint32_t)L.next_kp;
      P.irec[ci] = ir; P.erec[ci] = er;
    }
  }
  // the accumulator rotation (chunk % NACC) and the group alternation (hid % 2) must repeat identically every tile, and
  // each output layer must be exactly one chunk
  if (P.n_chunks % NACC != 0 || hid % 2 != 0 || P.layer[2].n_chunks != 1 || P.layer[7].n_chunks != 1 || P.layer[8].n_chunks != 1) {
    set_error("bf16 rollout: chunk plan (%d chunks, %d hidden) does not tile the accumulator rotation (actor hidden=%d model hidden=%d)",
              P.n_chunks, hid, Hp, Hm);
    return DRPO_ERR_UNSUPPORTED;
  }
  return DRPO_OK;
}

// order of drpo_linear's handed to pack_net for the member net: trunk0, trunk1, diff0, lvar0, diff1, lvar1 (= layers 3..8)
static int pack_net(const drpo_linear* lin, const NetPlan& P, int first_layer, int count, uint8_t* img, void* stream) {
  uint32_t off = 0;
  for (int i = 0; i < count; ++i) {
    const LayerSpec& L = P.layer[first_layer + i];
    PackChunks pc; pc.n = L.n_chunks;
    for (int c = 0; c < L.n_chunks; ++c) pc.n0[c] = P.chunk[L.first_chunk + c].n0;
    pc.n0[L.n_chunks] = L.np;
    DRPO_LAUNCH(pack_layer_kernel, grid_for((int64_t)L.np * L.kp), 256, 0, stream, lin[i].w, lin[i].b, L.n_real, L.k_real,
                L.np, L.kp, pc, reinterpret_cast<__nv_bfloat16*>(img + off));
    off += (uint32_t)L.np * L.kp * 2;
  }
  return DRPO_OK;
}

static int smem_bytes_for(const NetPlan& P, int S, int stages, int& SP, int& OP, int& NM, int& NMG) {
  SP = P.layer[0].kp | 1;                                  // state row [s, 1, 0..] padded to the first layer's K
  OP = P.layer[3].kp | 1;                                  // [norm s, a, 1, 0..] padded to trunk0's K; later [next state, reward]
  NMG = round_up(S + 1, 4);                                // row length of the step's model-noise array in global memory
  NM = (S + 1) <= 16 ? NMG + 4 : 0;                        // staged copy (+4: conflict-free float4 rows); wide states read global
  const uint32_t slot = (P.max_chunk_bytes + 1023u) & ~1023u;
  return (int)(slot * stages + (size_t)TILE_M * (2 * SP + 2 * OP + 2 * 4 + 2 * NM) * 4 + sizeof(SmemLayout) + 64);
}

}  // namespace umma

using namespace umma;

// status word of the last bf16 rollout: the device-side error flag is copied at the end of the rollout's stream work into a
// 4-byte pinned host word owned by the library (the only allocation it ever makes), so that a later check never touches a
// caller workspace that may have been freed
static int* g_status_host = nullptr;
static thread_local int g_timing_on = 0;
static thread_local std::vector<std::pair<cudaEvent_t, cudaEvent_t>>* g_timing_events = nullptr;

int umma_kernel_status() {
  if (!g_status_host) return 0;
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("drpo_kernel_status: %s", cudaGetErrorString(cudaGetLastError())); return DRPO_ERR_CUDA; }
  const int code = *(volatile int*)g_status_host;
  if (code) set_error("bf16 rollout kernel: an in-kernel wait timed out (code %d): pipeline protocol bug, results invalid", code);
  return code;
}
void umma_timing_enable(int on) {
  g_timing_on = on;
  if (!g_timing_events) g_timing_events = new std::vector<std::pair<cudaEvent_t, cudaEvent_t>>();
  for (auto& e : *g_timing_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  g_timing_events->clear();
}
int umma_timing_read(double* total_ms, int64_t* launches) {
  double t = 0; int64_t n = 0;
  if (g_timing_events)
    for (auto& e : *g_timing_events) {
      float ms = 0.f;
      if (cudaEventSynchronize(e.second) != cudaSuccess || cudaEventElapsedTime(&ms, e.first, e.second) != cudaSuccess) { set_error("drpo_timing_read: event query failed"); return DRPO_ERR_CUDA; }
      t += ms; ++n;
    }
  *total_ms = t; *launches = n;
  return DRPO_OK;
}

int64_t umma_rollout_ws_bytes(const drpo_rollout_args& a) {
  NetPlan P;
  if (build_plan(a, P) != DRPO_OK) return 0;
  const int64_t noise_bytes = a.batch * (4 + round_up(a.ensemble->state_dim + 1, 4)) * 4 + 1024;
  return rollout_ws_bytes_fp32(a) + noise_bytes + (int64_t)align_up(P.policy_bytes, 1024) +
         (int64_t)a.ensemble->ensemble_size * align_up(P.model_bytes, 1024) + 4096;
}

int umma_rollout_impl(const drpo_rollout_args& a, int dump_layer, float* dump_out) {
  NetPlan P; int rc;
  if ((rc = build_plan(a, P))) return rc;
  const int64_t B = a.batch; const int S = a.ensemble->state_dim, A = a.ensemble->action_dim;
  const int H = dump_layer >= 0 ? 1 : a.horizon; void* stream = a.stream;
  if (dump_layer >= P.n_layers && dump_layer != 100) { set_error("debug dump: layer %d out of range", dump_layer); return DRPO_ERR_ARG; }
  if (a.env->con_dim > DRPO_MAX_CON) { set_error("con_dim %d > %d", a.env->con_dim, DRPO_MAX_CON); return DRPO_ERR_ARG; }
  int dev = 0, sms = 148, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  int SP, OP, NM, NMG, stages = MAX_STAGES;
  while (stages > 2 && smem_bytes_for(P, S, stages, SP, OP, NM, NMG) > max_smem) --stages;
  const int smem = smem_bytes_for(P, S, stages, SP, OP, NM, NMG);
  if (smem > max_smem) { set_error("bf16 rollout: needs %d B of shared memory, device offers %d", smem, max_smem); return DRPO_ERR_UNSUPPORTED; }

  Arena ar(a.workspace, a.workspace_bytes);
  RolloutScratch w;
  w.curA = ar.take<float>(B * S); w.curB = ar.take<float>(B * S); w.actions = ar.take<float>(B * A);
  w.next_states = ar.take<float>(B * S); w.rewards = ar.take<float>(B);
  w.done = ar.take<uint8_t>(B);
  w.idsA = ar.take<int32_t>(B); w.idsB = ar.take<int32_t>(B); w.n_alive = ar.take<int32_t>(a.horizon + 2);
  const int nblocks = (int)((B + CBLK - 1) / CBLK);
  w.block_counts = ar.take<int32_t>(nblocks + 1);
  w.st = ar.take<RolloutState>(1);
  int* err_flag = ar.take<int>(4);
  float* eps_p = ar.take<float>(B * 4);
  float* eps_m = ar.take<float>(B * NMG);
  uint8_t* pol_img = ar.take<uint8_t>(align_up(P.policy_bytes, 1024));
  const int E = a.ensemble->ensemble_size;
  uint8_t* mem_img = ar.take<uint8_t>((int64_t)E * align_up(P.model_bytes, 1024));
  if (!ar.ok()) { set_error("drpo_rollout(bf16): workspace too small (%lld needed, %lld given)", (long long)ar.off, (long long)a.workspace_bytes); return DRPO_ERR_WORKSPACE; }

  // ---- pack the actor and every member this rollout uses into the UMMA layout (bf16, bias folded in) ----
  {
    drpo_linear pl[3] = {a.actor->l0, a.actor->l1, a.actor->l2};
    if ((rc = pack_net(pl, P, 0, 3, pol_img, stream))) return rc;
    bool used[64] = {false};
    for (int t = 0; t < H; ++t) used[a.member_idx_host[t]] = true;
    for (int m = 0; m < E; ++m) {
      if (!used[m]) continue;
      MemberNet mn = member_of(*a.ensemble, m);
      drpo_linear ml[6] = {mn.t0, mn.t1, mn.d0, mn.l0, mn.d1, mn.l1};       // plan order: 3 trunk0, 4 trunk1, 5 diff0, 6 lvar0, 7 diff1, 8 lvar1
      if ((rc = pack_net(ml, P, 3, 6, mem_img + (int64_t)m * align_up(P.model_bytes, 1024), stream))) return rc;
    }
  }
  DRPO_CUDA_OK(cudaFuncSetAttribute(rollout_step_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  DRPO_CUDA_OK(cudaFuncSetAttribute(rollout_step_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));

  DRPO_CUDA_OK(cudaMemsetAsync(err_flag, 0, 16, (cudaStream_t)stream));
  if (!g_status_host) { DRPO_CUDA_OK(cudaMallocHost((void**)&g_status_host, sizeof(int))); *g_status_host = 0; }
  DRPO_CUDA_OK(cudaMemcpyAsync(w.curA, a.initial_states, sizeof(float) * B * S, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  DRPO_LAUNCH(rollout_init_kernel, grid_for(B), 256, 0, stream, w.idsA, B, a.traj_id_offset, w.n_alive, w.st, a.virt.pointer);
  float* cur = w.curA; float* nxt = w.curB; int32_t* ids = w.idsA; int32_t* ids_n = w.idsB;
  const int grid = (int)std::min<int64_t>(sms, (B + TILE_M - 1) / TILE_M);
  for (int t = 0; t < H; ++t) {
    const int* n_dev = w.n_alive + t;
    // this step's Gaussian draws, keyed by global trajectory id (torch.normal in policy.act, randn_like in ensemble.sample)
    NoiseView np_ = make_noise(a.eps_policy ? a.eps_policy + (int64_t)t * a.eps_batch_stride * A : nullptr, A, a.seed, TAG_ROLLOUT_POLICY, (uint32_t)t);
    NoiseView nm_ = make_noise(a.eps_model ? a.eps_model + (int64_t)t * a.eps_batch_stride * (S + 1) : nullptr, S + 1, a.seed, TAG_ROLLOUT_MODEL, (uint32_t)t);
    DRPO_LAUNCH(noise_stage_kernel, grid_for(B * (1 + NMG / 4)), 256, 0, stream, ids, n_dev, B, np_, nm_, A, S + 1, NMG, eps_p, eps_m);
    StepParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.plan = P; sp.policy_img = pol_img; sp.model_img = mem_img + (int64_t)a.member_idx_host[t] * align_up(P.model_bytes, 1024);
    sp.cur = cur; sp.n_dev = n_dev; sp.n_max = B; sp.eps_p = eps_p; sp.eps_m = eps_m;
    sp.actions = w.actions; sp.next_states = w.next_states; sp.rewards = w.rewards;
    sp.norm_mean = a.ensemble->norm_mean; sp.norm_std = a.ensemble->norm_std; sp.min_lv = a.ensemble->min_log_var; sp.max_lv = a.ensemble->max_log_var;
    sp.S = S; sp.A = A; sp.SP = SP; sp.OP = OP; sp.NM = NM; sp.NMG = NMG; sp.stages = stages; sp.err_flag = err_flag;
    sp.dump_layer = dump_layer; sp.dump_out = dump_out;
    std::pair<cudaEvent_t, cudaEvent_t> ev{};
    if (g_timing_on) { cudaEventCreate(&ev.first); cudaEventCreate(&ev.second); cudaEventRecord(ev.first, (cudaStream_t)stream); }
    if (dump_layer >= 0) { DRPO_LAUNCH(rollout_step_umma_kernel<true>, grid, NUM_THREADS, smem, stream, sp); }
    else { DRPO_LAUNCH(rollout_step_umma_kernel<false>, grid, NUM_THREADS, smem, stream, sp); }
    if (g_timing_on) { cudaEventRecord(ev.second, (cudaStream_t)stream); g_timing_events->push_back(ev); }
    DRPO_LAUNCH(hooks_store_kernel, grid_for(B * S), 256, 0, stream, *a.env, a.virt, w.st, n_dev, cur, w.actions, w.next_states, w.rewards, w.done);
    DRPO_LAUNCH(compact_count_kernel, nblocks, CBLK, 0, stream, w.done, n_dev, w.block_counts);
    DRPO_LAUNCH(compact_scan_kernel, 1, CBLK, 0, stream, w.block_counts, nblocks, w.n_alive, t, w.st, a.step_counts);
    DRPO_LAUNCH(compact_scatter_kernel, nblocks, CBLK, 0, stream, w.done, n_dev, w.block_counts, w.next_states, ids, nxt, ids_n, S);
    float* tf = cur; cur = nxt; nxt = tf;
    int32_t* ti = ids; ids = ids_n; ids_n = ti;
  }
  DRPO_LAUNCH(rollout_finish_kernel, 1, 1, 0, stream, a.virt.pointer, w.st, a.step_counts, H);
  DRPO_CUDA_OK(cudaMemcpyAsync(g_status_host, err_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return DRPO_OK;
}

int umma_rollout(const drpo_rollout_args& a) { return umma_rollout_impl(a, -1, nullptr); }
int umma_debug_layer(const drpo_rollout_args& a, int layer, float* out) { return umma_rollout_impl(a, layer, out); }

}  // namespace drpo
