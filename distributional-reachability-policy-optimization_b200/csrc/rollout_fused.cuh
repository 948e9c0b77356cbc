// DRPO_PREC_BF16 rollout step, second generation: one persistent, warp-specialised tcgen05 kernel per rollout step (src/smbpo.py:234-246,
// src/policy.py:89-97, src/dynamics.py:112-122,198-203) that keeps a 128-row tile on the SM for the whole
//   policy MLP -> squashed-Gaussian sample -> ensemble-member MLP (trunk + 2 heads) -> Gaussian next-state sample
// chain.  What changed against the first generation (three rotating 64-column accumulators, 27 chunk hand-offs per tile, tensor pipe
// 25 % busy because every chunk waited for an accumulator to come back from its epilogue):
//   * a hidden layer accumulates ALL its output columns at once (policy: one N = 256 MMA per k-step; member: two MMAs, N = 112 + 96,
//     into three accumulator slots used round-robin), so the tensor pipe never waits for an accumulator buffer inside a layer;
//   * the MMAs of layer l+1 are issued PART by part, a part being the k-steps that read one 64/48-column chunk of layer l's activation:
//     the parts of the first three chunks run while the epilogue of the last chunk is still in flight (K-sliced start), and every
//     output chunk of layer l+1 completes in the last part, where the four epilogue groups drain them concurrently;
//   * the two independent head layers (diff / log-var hidden) are issued back to back, so the second one's MMAs hide the first one's
//     epilogue; its activations overwrite the trunk output in place (the columns still being read are never the ones being written);
//   * the layer inputs ([s,1], [norm s, a, 1]) are staged in shared memory and read by SS-mode MMAs, biases of the 256-wide policy
//     layers ride in one extra k-step against a constant tile of ones with the bias split in (hi, lo) bf16 parts;
//   * epilogue -> issuer signalling uses monotone counters in shared memory (release / acquire), not mbarriers: the issuers wait only
//     where a static plan says so, and a counter cannot be "missed" the way an unobserved mbarrier phase can;
//   * CTA pairs (clusters of 2) fetch every weight block from L2 once and multicast it into both CTAs' rings: the L2 -> SM stream
//     halves (435 KB of weights per 128-row tile would otherwise approach the L2 bandwidth ceiling at the higher tile rate).
// Warp roles (768 threads): warps 0-15 four hidden-epilogue groups (group j owns output chunk j of every hidden layer; warp q of a
// group owns TMEM lanes 32q..32q+31), warps 16-19 output group (policy head -> action, diff / log-var heads -> Gaussian sample,
// coalesced stores, next tile's prologue and Philox draws), warp 20 TMA producer, warp 21 the MMA issuer (one elected lane walks the
// static list of MMA groups; the record of the next group is fetched before the current one is issued, so the bookkeeping between two
// groups fits into the ~4 MMAs the tensor pipe queues - two alternating issuers with a token were measured to leave a 400-600 cycle
// bubble at every one of the 22 block hand-offs of a tile).
#pragma once
#include <cuda_bf16.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "nets.cuh"
#include "tc05.cuh"

namespace drpo {
namespace r2 {
using namespace tc;

constexpr int TILE_M = 128, CLUSTER = 2, N_GROUPS = 4, N_HID = 6;
constexpr int EPI_WARPS = 16, OUT_WARP0 = 16, PRODUCER_WARP = 20, ISSUER_WARP0 = 21;
constexpr int NUM_THREADS = 22 * 32, GROUP_THREADS = 128;
constexpr int MAX_IG = 56, MAX_BLOCKS = 28, MAX_STAGES = 6;
constexpr int N_LAYERS = 9;          // 0 actor L0, 1 actor L1, 2 actor L2 (head), 3 trunk0, 4 trunk1, 5 diff0, 6 lvar0, 7 diff1, 8 lvar1

// monotone counters written by the epilogue / output groups (one word per warp), polled by the issuers
enum { CNT_NONE = 0, CNT_ACT = 1 /* +j */, CNT_DRAIN = 5 /* +j */, CNT_TILE = 9, CNT_XM = 10, CNT_OUT = 11, N_CNT = 12 };
enum { IGF_FIRST = 1, IGF_SS = 2, IGF_SRC_SHIFT = 2 /* 0 xp, 1 xm, 2 ones */, IGF_NEWBLK = 16, IGF_ENDBLK = 32 };
enum { COMMIT_H0 = 1, COMMIT_H1 = 2, COMMIT_OUT0 = 4 /* << k */ };

struct IG {                    // one MMA group: <= 4 k-steps accumulated into one D region (32 bytes)
  uint32_t idesc, b_hi, b_off, d_col;
  uint32_t a_src;              // TS: TMEM column of the first k-step;  SS: offset of the first k-step inside the A tile (16-byte units)
  uint32_t waits;              // 4 x 8 bits: (counter id << 3) | c0 ; expected value = c0 + inc(id) * tile_it
  uint16_t nk; uint8_t flags, commit;
  uint32_t pad;
};
// the issuer's resolved form of an IG (built once per CTA at kernel start): every address is absolute, so that the single issuing
// thread executes ~40 instructions per MMA group (one thread retires a dependent instruction every ~5 cycles; the first version's
// generic decode cost ~600 cycles per group - more than the group's MMAs)
struct IGD { uint4 lo /* idesc, b_hi, b_rel, d */, hi /* a, waits, ctl, commits */; };
enum { CTL_NK = 7, CTL_FIRST = 8, CTL_SS = 16, CTL_XM = 32, CTL_NEWBLK = 64, CTL_ENDBLK = 128 };
struct BlockRec { uint32_t src_off, bytes; uint16_t first_ig, n_ig; uint32_t pad; };
struct EpiRec { uint16_t acc_col, nc, out_col, n0; int16_t one_rel; uint8_t silu, layer; uint16_t n_real, pad; };

struct PlanDev {               // POD part of the plan: travels in the kernel parameters, copied to shared memory at kernel start
  IG ig[MAX_IG];
  BlockRec blk[MAX_BLOCKS];
  EpiRec epi[N_HID * N_GROUPS];
  int n_ig, n_blocks, n_pol_blocks;
  int K0p, K0m, No;            // padded K of the two input layers, padded N of the two output heads
  int xp_one, xm_one;          // position of the constant 1 (bias slot) in the staged inputs, -1: none (bias through the ones tile)
  int head_col, d1_col, v1_col;  // TMEM columns of the three output accumulators
  uint32_t slot_bytes;
};

struct PackSub { int layer; int n0, nc, k0, kc, mode, slot; uint32_t dst_off; };     // mode 0 weights, 1 bias block (hi, lo)

struct Plan {
  PlanDev d;
  uint32_t pol_bytes, mem_bytes;
  std::vector<PackSub> pack_pol, pack_mem;
  int n_real[N_LAYERS], k_real[N_LAYERS];
};

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------------------------
// weight images: per block, per sub-block a contiguous canonical K-major tile  [nc/8][kc/8][8 rows][8 elems]
// ---------------------------------------------------------------------------------------------------------------
struct PackJob { const float* W; const float* b; int n_real, k_real, n0, nc, k0, kc, mode, slot; __nv_bfloat16* dst; };
constexpr int PACK_JOBS = 40;
struct PackTable { PackJob job[PACK_JOBS]; };
static __global__ void pack_sub_kernel(const __grid_constant__ PackTable t) {
  const PackJob& j = t.job[blockIdx.y];
  const int total = j.nc * j.kc;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / j.kc, k = i - n * j.kc;
    const int gn = j.n0 + n, gk = j.k0 + k;
    float v = 0.f;
    if (gn < j.n_real) {
      if (j.mode == 0) v = gk < j.k_real ? j.W[(int64_t)gn * j.k_real + gk] : (gk == j.k_real && j.slot ? j.b[gn] : 0.f);
      else {                                                    // bias block against the ones tile: k = 0 -> bf16(b), k = 1 -> bf16(b - bf16(b))
        const float bb = j.b[gn];
        const float hi = __bfloat162float(__float2bfloat16_rn(bb));
        v = k == 0 ? hi : (k == 1 ? bb - hi : 0.f);
      }
    }
    j.dst[((n >> 3) * (j.kc >> 3) + (k >> 3)) * 64 + (n & 7) * 8 + (k & 7)] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the fused step kernel
// ---------------------------------------------------------------------------------------------------------------
struct StepParams {
  PlanDev plan;
  const uint8_t* policy_img; const uint8_t* model_img;
  const float* cur; const int* n_dev; int64_t n_max;
  const int32_t* ids;                    // [n] global trajectory id of every alive row: the key of its Gaussian draws
  NoiseView noise_p, noise_m;            // policy / model draws of this step: injected tensors (parity) or Philox(seed, id, step)
  float *actions, *next_states, *rewards;
  const float *norm_mean, *norm_std, *min_lv, *max_lv;
  int S, A, SPs, OPs, NM, stages;
  int* err_flag;
  int dump_layer; float* dump_out;       // debug: dump the fp32 accumulator of one layer
  const int32_t* ready_flags; int ready_shift;     // step 0 with streamed start states: flag of every 2^shift-row block (else NULL)
};

struct SmemCtl {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES], half_full[2][2], out_full[3];
  uint4 cnt[N_CNT];
  uint32_t tmem_base, pad[3];
  // per-dim constants of the member, staged once per CTA: normaliser, and the log-var soft clamp folded into
  //   std = exp(lv/2) = s0 * sqrt(1 + E / (1 + exp(hi - x)))   with s0 = exp(lo/2), E = exp(hi - lo)      (src/dynamics.py:120-121,201)
  float norm_mean[64], norm_inv[64], lv_hi[64], lv_E[64], lv_s0[64];
  IGD igd[MAX_IG];
  BlockRec blk[MAX_BLOCKS];
  EpiRec epi[N_HID * N_GROUPS];
};

__device__ __forceinline__ void cnt_publish(uint4* c, int q, uint32_t v) {           // one lane per warp
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(reinterpret_cast<uint32_t*>(c) + q)), "r"(v) : "memory");
}
__device__ __forceinline__ bool cnt_reached(const uint4* c, uint32_t want) {
  uint32_t a, b, cc, d;
  asm volatile("ld.acquire.cta.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(cc), "=r"(d) : "r"(smem_u32(c)) : "memory");
  return a >= want && b >= want && cc >= want && d >= want;
}
// bounded like every other wait of this library: a protocol bug is reported through err_flag, never a hang
static __device__ __noinline__ void cnt_wait_slow(const uint4* c, uint32_t want, int* err_flag, int code) {
  long long t0 = 0;
#pragma unroll 1
  for (uint32_t it = 0;; ++it) {
    if (cnt_reached(c, want)) return;
    if (it >= 16) __nanosleep(32);
    if ((it & 63) == 63) {
      if (t0 == 0) t0 = clock64();
      const long long dt = clock64() - t0;
      if (dt > 1000000000ll || (dt > 2000000ll && *(volatile int*)err_flag)) break;
    }
  }
  if (atomicCAS(err_flag, 0, code) == 0)
    printf("drpo_b200: counter wait timed out (code %d, block %d, warp %d, want %u)\n", code, (int)blockIdx.x, (int)(threadIdx.x >> 5), want);
}
__device__ __forceinline__ void cnt_wait(const uint4* c, uint32_t want, int* err_flag, int code) {
  if (!cnt_reached(c, want)) cnt_wait_slow(c, want, err_flag, code);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// Streamed start states: block until the row blocks that contain rows [r0, r1] have landed (their flags are written by the
// copy engine right after the rows, in stream order).  Bounded like every other wait of this kernel.
static __device__ __noinline__ void wait_rows_ready(const int32_t* flags, int shift, int r0, int r1, int* err_flag) {
  const int c0 = r0 >> shift, c1 = r1 >> shift;
  long long t0 = 0;
  for (int c = c0; c <= c1; ++c) {
    for (uint32_t it = 0;; ++it) {
      int v;
      asm volatile("ld.acquire.sys.global.b32 %0, [%1];" : "=r"(v) : "l"(flags + c) : "memory");
      if (v) break;
      __nanosleep(256);
      if ((it & 255) == 255) {
        if (t0 == 0) t0 = clock64();
        if (clock64() - t0 > 4000000000ll) {                        // ~2 s: the transfer never arrived
          if (atomicCAS(err_flag, 0, 77) == 0) printf("drpo_b200: start-state block %d never became ready (block %d)\n", c, (int)blockIdx.x);
          return;
        }
      }
    }
  }
}

// tcgen05.ld of 16 columns into the low half of a 32-register array (keeps the array in registers: no aliasing casts)
__device__ __forceinline__ void tmem_ld16_lo(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
// 16 or 32 accumulator columns (already in registers) -> activation -> packed bf16 -> TMEM (A operand of the next layer)
template <int W, bool kSilu>
__device__ __forceinline__ void act_store(const uint32_t (&r)[32], uint32_t dst, int one) {
  uint32_t pk[W / 2];
#pragma unroll
  for (int j = 0; j < W / 2; ++j) {
    if (kSilu) pk[j] = silu_bf16x2(pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])));
    else pk[j] = pack_bf16_relu(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
  }
  if (one >= 0 && one < W) {                                    // rare: the piece that holds the consumer's bias slot (constant 1)
#pragma unroll
    for (int j = 0; j < W / 2; ++j) {
      if (one == 2 * j) pk[j] = (pk[j] & 0xFFFF0000u) | 0x00003F80u;
      if (one == 2 * j + 1) pk[j] = (pk[j] & 0x0000FFFFu) | 0x3F800000u;
    }
  }
  if constexpr (W == 32) tmem_st16(dst, pk); else tmem_st8(dst, pk);
}
// one output chunk (NC = 16..64 columns) of a hidden layer, one row per thread, in pieces of <= 32 columns (80 registers per thread: the
// register file of an SM sub-partition holds its 6 warps x 32 lanes x 85).  NC is a template parameter: with a run-time width the
// differently sized loads share registers through branches and the compiler spills.
template <int NC, bool kSilu>
__device__ __forceinline__ void hidden_chunk(uint32_t src, uint32_t dst, int one, uint4* drain, int q, uint32_t stamp_v, int lane) {
  uint32_t ra[32];
  if constexpr (NC >= 32) tmem_ld32(src, ra); else tmem_ld16_lo(src, ra);
  tmem_ld_wait();
  if constexpr (NC <= 32) { tc_fence_before(); if (lane == 0) cnt_publish(drain, q, stamp_v); }
  if constexpr (NC >= 32) act_store<32, kSilu>(ra, dst, one); else act_store<16, kSilu>(ra, dst, one);
  if constexpr (NC > 32) {
    if constexpr (NC == 64) tmem_ld32(src + 32, ra); else tmem_ld16_lo(src + 32, ra);
    tmem_ld_wait();
    tc_fence_before();
    if (lane == 0) cnt_publish(drain, q, stamp_v);              // the chunk's accumulator columns are in registers: the slot may be reused
    if constexpr (NC == 64) act_store<32, kSilu>(ra, dst + 16, one - 32); else act_store<16, kSilu>(ra, dst + 16, one - 32);
  }
}
template <bool kSilu>
__device__ __forceinline__ void hidden_chunk_any(uint32_t src, uint32_t dst, int nc, int one, uint4* drain, int q, uint32_t stamp_v, int lane) {
  switch (nc) {
    case 64: hidden_chunk<64, kSilu>(src, dst, one, drain, q, stamp_v, lane); break;
    case 48: hidden_chunk<48, kSilu>(src, dst, one, drain, q, stamp_v, lane); break;
    case 32: hidden_chunk<32, kSilu>(src, dst, one, drain, q, stamp_v, lane); break;
    case 16: hidden_chunk<16, kSilu>(src, dst, one, drain, q, stamp_v, lane); break;
    default: if (lane == 0) cnt_publish(drain, q, stamp_v); break;      // empty chunk (narrow layers): protocol only
  }
}

// octets [o0, o1) of one row of a layer input -> bf16 K-major canonical tile in shared memory: element (row, k) at
// (k/8) * 2048 + row * 16 + (k%8) * 2  (A operand of an SS-mode MMA: LBO = 2048, SBO = 128); `val(k)` yields element k
template <typename F>
__device__ __forceinline__ void write_input_octets(uint8_t* tile, int row, int o0, int o1, F val) {
  for (int o = o0; o < o1; ++o) {
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = val(8 * o + e);
    uint4 v;
    v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]); v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(tile + o * 2048 + row * 16) = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The issue program of the STANDARD structure (actor hidden 256, member hidden 200: every reference config) as a compile-time table.
// The generic issuer loop decodes a record per MMA group at run time: ~120 instructions with a dozen R2UR and predicate conversions,
// retired by ONE thread at a dependent instruction every ~5 cycles, plus shared-memory round trips that take ~250 cycles while the
// tensor core streams its B operand from shared memory - measured ~850 cycles per weight block against 310-510 cycles of MMAs.
// With the structure known at compile time the loop over the 40 groups of a tile unrolls into straight-line code: operands come
// from the kernel parameters (constant bank -> uniform registers), predicates and instruction forms are static, and the only
// shared-memory accesses left are the ring-stage probe (issued one block ahead) and the counter waits at the start of a layer.
// The host checks the plan it built against this table (std_structure_matches) and falls back to the generic loop otherwise.
// ---------------------------------------------------------------------------------------------------------------
enum { SRC_TS = 0, SRC_XP = 1, SRC_XM = 2, SRC_ONES = 3 };
enum { CM_NONE = 0, CM_H0 = 1, CM_H1 = 2, CM_BOTH = 3, CM_OUT0 = 4, CM_OUT1 = 5, CM_OUT2 = 6 };
struct StdIG { int nk /* 0 = run time */, src, first, newblk, endblk, waits, commit, lh; };
constexpr int N_STD_IG = 40;
__device__ constexpr StdIG kStd[N_STD_IG] = {
    {0, SRC_XP, 1, 1, 1, 1, CM_BOTH, 0},                                                      //  0 actor L0
    {1, SRC_ONES, 1, 1, 1, 1, CM_NONE, 1},                                                    //  1 actor L1 bias
    {4, SRC_TS, 0, 1, 1, 1, CM_NONE, 1}, {4, SRC_TS, 0, 1, 1, 0, CM_NONE, 1},                 //  2-5 actor L1 parts
    {4, SRC_TS, 0, 1, 1, 0, CM_NONE, 1}, {4, SRC_TS, 0, 1, 1, 0, CM_BOTH, 1},
    {1, SRC_ONES, 1, 1, 0, 1, CM_NONE, 1},                                                    //  6 head bias
    {4, SRC_TS, 0, 0, 0, 1, CM_NONE, 1}, {4, SRC_TS, 0, 0, 0, 0, CM_NONE, 1},                 //  7-10 head parts
    {4, SRC_TS, 0, 0, 0, 0, CM_NONE, 1}, {4, SRC_TS, 0, 0, 1, 0, CM_OUT0, 1},
    {0, SRC_XM, 1, 1, 1, 1, CM_BOTH, 2},                                                      // 11 trunk0
    {4, SRC_TS, 1, 1, 0, 1, CM_NONE, 3}, {4, SRC_TS, 1, 0, 1, 0, CM_NONE, 3},                 // 12-19 trunk1 (H0, H1) x 4 parts
    {3, SRC_TS, 0, 1, 0, 0, CM_NONE, 3}, {3, SRC_TS, 0, 0, 1, 0, CM_NONE, 3},
    {3, SRC_TS, 0, 1, 0, 0, CM_NONE, 3}, {3, SRC_TS, 0, 0, 1, 0, CM_NONE, 3},
    {3, SRC_TS, 0, 1, 0, 0, CM_H0, 3}, {3, SRC_TS, 0, 0, 1, 0, CM_H1, 3},
    {4, SRC_TS, 1, 1, 1, 1, CM_NONE, 4}, {3, SRC_TS, 0, 1, 1, 0, CM_NONE, 4},                 // 20-23 diff hidden (N = 208) x 4 parts
    {3, SRC_TS, 0, 1, 1, 0, CM_NONE, 4}, {3, SRC_TS, 0, 1, 1, 0, CM_BOTH, 4},
    {4, SRC_TS, 1, 1, 0, 0, CM_NONE, 5}, {4, SRC_TS, 1, 0, 1, 1, CM_NONE, 5},                 // 24-31 log-var hidden (H1 first in part 0)
    {3, SRC_TS, 0, 1, 0, 0, CM_NONE, 5}, {3, SRC_TS, 0, 0, 1, 0, CM_NONE, 5},
    {3, SRC_TS, 0, 1, 0, 0, CM_NONE, 5}, {3, SRC_TS, 0, 0, 1, 0, CM_NONE, 5},
    {3, SRC_TS, 0, 1, 0, 0, CM_H0, 5}, {3, SRC_TS, 0, 0, 1, 0, CM_H1, 5},
    {4, SRC_TS, 1, 1, 0, 1, CM_NONE, 5}, {3, SRC_TS, 0, 0, 0, 0, CM_NONE, 5},                 // 32-35 diff head
    {3, SRC_TS, 0, 0, 0, 0, CM_NONE, 5}, {3, SRC_TS, 0, 0, 1, 0, CM_OUT1, 5},
    {4, SRC_TS, 1, 1, 0, 1, CM_NONE, 5}, {3, SRC_TS, 0, 0, 0, 0, CM_NONE, 5},                 // 36-39 log-var head
    {3, SRC_TS, 0, 0, 0, 0, CM_NONE, 5}, {3, SRC_TS, 0, 0, 1, 0, CM_OUT2, 5}};

// Upper word of an mbarrier object: bit 31 (bit 63 of the object) is the parity of the phase in progress, so "the phase with parity P
// has completed" <=> that bit != P.  A plain load: unlike try_wait / test_wait (whose predicate result stalls the in-order issue of
// the thread until the shared-memory unit answers, ~200 cycles while the tensor core streams operands), its result register is
// consumed a whole weight block later.  A negative answer falls back to the blocking wait, so a stale value is harmless.
__device__ __forceinline__ uint32_t mbar_peek_hi(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1+4];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// debug timing (kDebug build, dump_layer == 100): CTA 0 stamps clock() for its first 4 tiles into dump_out viewed as uint32
// [(tile_it * 64 + slot) * 8 + k].  slots 0..27 = weight blocks (issuer: k 0 ring stage full, 1 counter waits done, 2 token received,
// 3 block issued), 32 + 4*lh + j = hidden layer lh, group j (k 4 wait begin, 5 accumulator full, 6 drained, 7 activation published),
// 60 = output group (k 0 head full, 1 member input published, 2 next prologue done, 3 diff head full, 4 log-var head full, 5 stores done)
template <bool kDebug>
__device__ __forceinline__ void stamp(const StepParams& p, uint32_t tile_it, int slot, int k) {
  if (kDebug) {
    if (p.dump_layer == 100 && blockIdx.x == 0 && tile_it < 4)
      reinterpret_cast<uint32_t*>(p.dump_out)[(tile_it * 64 + slot) * 8 + k] = (uint32_t)clock();
  }
}

template <bool kDebug, bool kStdProg>
__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(NUM_THREADS, 1) rollout_step_fused_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const PlanDev& plan = p.plan;
  const int S = p.S, A = p.A, O = S + 1;
  // shared memory: [weight ring | xp | xm x2 | ones | raw states x2 | output rows | policy noise x2 | model noise x2 | control]
  uint8_t* ring = smem_raw;
  const uint32_t slot_bytes = plan.slot_bytes;
  uint8_t* xp = ring + (size_t)slot_bytes * p.stages;
  uint8_t* xm = xp + TILE_M * plan.K0p * 2;
  const uint32_t xm_bytes = TILE_M * plan.K0m * 2;                                    // xm is double-buffered: tile i+1's state part is staged during tile i
  uint8_t* ones = xm + 2 * xm_bytes;
  float* st_s = reinterpret_cast<float*>(ones + TILE_M * 16 * 2);                    // [2][128][SPs] raw states (fp32)
  float* st_o = st_s + 2 * TILE_M * p.SPs;                                           // [128][OPs] means, then [next state, reward] of the current tile
  float* st_np = st_o + TILE_M * p.OPs;                                              // [2][128][4]  policy noise
  float* st_nm = st_np + 2 * TILE_M * 4;                                             // [2][128][NM] model noise (NM = 0: drawn in place)
  SmemCtl* sm = reinterpret_cast<SmemCtl*>(st_nm + 2 * TILE_M * p.NM);
  int* err = p.err_flag;

  const int n = (int)min((int64_t)*p.n_dev, p.n_max);
  const int n_tiles = (n + TILE_M - 1) / TILE_M;
  // tile pairs are dealt to the clusters round-robin; both CTAs of a pair walk the same number of tiles in lock step (coupled by the
  // weight ring); an odd last tile leaves the second CTA a tile without rows
  const uint32_t crank = cluster_ctarank();
  const int n_clusters = (int)gridDim.x / CLUSTER, cid = (int)blockIdx.x / CLUSTER;
  const int n_pairs = (n_tiles + 1) / 2;
  const int my_tiles = cid < n_pairs ? (n_pairs - cid + n_clusters - 1) / n_clusters : 0;
  auto tile_of = [&](int it) { return 2 * (cid + it * n_clusters) + (int)crank; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], CLUSTER); }
    for (int h = 0; h < 2; ++h) { mbar_init(&sm->half_full[h][0], 1); mbar_init(&sm->half_full[h][1], 1); }
    for (int k = 0; k < 3; ++k) mbar_init(&sm->out_full[k], 1);
    fence_barrier_init();
  }
  {
    const int t0 = threadIdx.x;
    if (t0 < N_CNT) sm->cnt[t0] = make_uint4(0u, 0u, 0u, 0u);
    if (t0 < S) { sm->norm_mean[t0] = p.norm_mean[t0]; sm->norm_inv[t0] = 1.f / (p.norm_std[t0] + 1e-6f); }
    if (t0 <= S) {
      const float lo = p.min_lv[t0], hi = p.max_lv[t0];
      sm->lv_hi[t0] = hi; sm->lv_E[t0] = __expf(hi - lo); sm->lv_s0[t0] = __expf(0.5f * lo);
    }
    for (int i = t0; i < MAX_BLOCKS; i += NUM_THREADS) sm->blk[i] = plan.blk[i];
    for (int i = t0; i < N_HID * N_GROUPS; i += NUM_THREADS) sm->epi[i] = plan.epi[i];
    for (int i = t0; i < 2 * TILE_M * p.SPs; i += NUM_THREADS) st_s[i] = 0.f;      // rows past the end of the batch stay finite
    for (int i = t0; i < TILE_M * p.OPs; i += NUM_THREADS) st_o[i] = 0.f;
    for (int i = t0; i < TILE_M * 16; i += NUM_THREADS) {       // ones tile: element (row, k) at (k/8)*2048 + row*16 + (k%8)*2, k = 0, 1 -> 1.0
      const int k = (i >> 10) * 8 + (i & 7);
      reinterpret_cast<__nv_bfloat16*>(ones)[i] = __float2bfloat16_rn(k < 2 ? 1.f : 0.f);
    }
  }
  if (warp == ISSUER_WARP0) tmem_alloc(&sm->tmem_base, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp == PRODUCER_WARP) {
    // ===================== TMA producer: every weight block of every tile through the ring, half per CTA, multicast to the pair ======
    if (elect_one()) {
      uint32_t s = 0, ph = 0;
      for (int it = 0; it < my_tiles; ++it)
        for (int b = 0; b < plan.n_blocks; ++b) {
          mbar_wait(&sm->empty[s], ph ^ 1, err, 1);            // both CTAs' MMAs on the stage's previous contents are done
          const BlockRec blk = sm->blk[b];
          const uint8_t* src = (b < plan.n_pol_blocks ? p.policy_img : p.model_img) + blk.src_off;
          mbar_expect_tx(&sm->full[s], blk.bytes);
          const uint32_t half = blk.bytes >> 1;
          bulk_g2s_multicast(ring + (size_t)s * slot_bytes + crank * half, src + crank * half, half, &sm->full[s], (uint16_t)3);
          if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
        }
    }
  } else if (warp == ISSUER_WARP0) {
    // ===================== MMA issuer: one thread walks the static list of MMA groups ===================================================
    if (kStdProg) {
      if (elect_one()) {
        const uint32_t sm_a = smem_u32(sm), ring_a = (smem_u32(ring) >> 4) & 0x3FFFu, slot_a = slot_bytes >> 4, xm_step = xm_bytes >> 4;
        const uint32_t xp_a = (smem_u32(xp) >> 4) & 0x3FFFu, xm_a = (smem_u32(xm) >> 4) & 0x3FFFu, ones_a = (smem_u32(ones) >> 4) & 0x3FFFu;
        const uint32_t full0 = smem_u32(&sm->full[0]), empty0 = smem_u32(&sm->empty[0]);
        const uint32_t hf_a = smem_u32(&sm->half_full[0][0]), of_a = smem_u32(&sm->out_full[0]);
        int s = 0; uint32_t ring_par = 0, L0 = 0, stage_a = ring_a;
        for (uint32_t t = 0; t < (uint32_t)my_tiles; ++t, L0 += N_HID) {
          int blk_no = 0;
#pragma unroll
          for (int i = 0; i < N_STD_IG; ++i) {
            const StdIG g = kStd[i];
            const IG& r = plan.ig[i];                                // kernel parameters: constant bank, static offsets
            if (g.newblk) {
              mbar_wait_addr(full0 + 8u * (uint32_t)s, ring_par, err, 3);
              stamp<kDebug>(p, t, blk_no, 0);
            }
            if (g.waits) {
              for (uint32_t w = r.waits; w; w >>= 8) {
                const uint32_t id = (w >> 3) & 31u, c0 = w & 7u;
                cnt_wait(&sm->cnt[id], c0 + (id < CNT_TILE ? L0 : t), err, 20 + (int)id);
              }
            }
            if (g.newblk) stamp<kDebug>(p, t, blk_no, 1);
            if (g.newblk || g.waits) tc_fence_after();
            const uint32_t d = tmem + r.d_col, b_lo = ((r.b_off & 0x3FFFu) | (8u << 16)) + stage_a, b_hi = r.b_hi, idesc = r.idesc;
            const int nk = g.nk ? g.nk : (int)r.nk;
            if (g.src == SRC_TS) {
              const uint32_t a = tmem + r.a_src;
              if (g.first) mma_ts<false>(d, a, b_lo, b_hi, idesc); else mma_ts<true>(d, a, b_lo, b_hi, idesc);
#pragma unroll
              for (int k = 1; k < 4; ++k)
                if (k < nk) mma_ts<true>(d, a + 8u * k, b_lo + 16u * k, b_hi, idesc);
            } else {
              const uint32_t base = g.src == SRC_XP ? xp_a : (g.src == SRC_XM ? xm_a + (t & 1u) * xm_step : ones_a);
              const uint32_t a_lo = (base + r.a_src) | (128u << 16), a_hi = 8u | (1u << 14);            // LBO 2048 B, SBO 128 B
              if (g.first) mma_ss<false>(d, a_lo, a_hi, b_lo, b_hi, idesc); else mma_ss<true>(d, a_lo, a_hi, b_lo, b_hi, idesc);
#pragma unroll
              for (int k = 1; k < 4; ++k)
                if (k < nk) mma_ss<true>(d, a_lo + 256u * k, a_hi, b_lo + 16u * k, b_hi, idesc);
            }
            if (g.commit == CM_H0 || g.commit == CM_BOTH) tc_commit_addr(hf_a + 8u * (0u + (g.lh & 1)));
            if (g.commit == CM_H1 || g.commit == CM_BOTH) tc_commit_addr(hf_a + 8u * (2u + (g.lh & 1)));
            if (g.commit >= CM_OUT0) tc_commit_addr(of_a + 8u * (uint32_t)(g.commit - CM_OUT0));
            if (g.endblk) {
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                           ::"r"(empty0 + 8u * (uint32_t)s), "h"((uint16_t)3) : "memory");
              stamp<kDebug>(p, t, blk_no, 3);
              ++blk_no;
              stage_a += slot_a;
              if (++s == p.stages) { s = 0; ring_par ^= 1u; stage_a = ring_a; }
            }
          }
        }
      }
    } else {
    {   // resolve the plan's records against this CTA's TMEM base and shared-memory addresses (all 32 lanes, once)
      const uint32_t xp_a = smem_u32(xp) >> 4, xm_a = smem_u32(xm) >> 4, ones_a = smem_u32(ones) >> 4;
      const uint32_t hf = (uint32_t)((const uint8_t*)&sm->half_full[0][0] - (const uint8_t*)sm), of = (uint32_t)((const uint8_t*)&sm->out_full[0] - (const uint8_t*)sm);
      for (int i = lane; i < plan.n_ig; i += 32) {
        const IG g = plan.ig[i];
        const uint32_t src = (g.flags >> IGF_SRC_SHIFT) & 3u;
        const uint32_t ra_ = (g.flags & IGF_SS) ? ((((src == 0 ? xp_a : (src == 1 ? xm_a : ones_a)) + g.a_src) & 0x3FFFu) | (128u << 16)) : tmem + g.a_src;
        const uint32_t ctl_ = (uint32_t)g.nk | ((g.flags & IGF_FIRST) ? CTL_FIRST : 0u) | ((g.flags & IGF_SS) ? CTL_SS : 0u) | (((g.flags & IGF_SS) && src == 1) ? CTL_XM : 0u) |
                ((g.flags & IGF_NEWBLK) ? CTL_NEWBLK : 0u) | ((g.flags & IGF_ENDBLK) ? CTL_ENDBLK : 0u);
        uint32_t c0 = 0, c1 = 0;                                       // the barrier index of a hidden layer is static: (6 t + lh) & 1 = lh & 1
        if (g.commit & COMMIT_H0) c0 = hf + 8u * (0u + (g.pad & 1u));
        if (g.commit & COMMIT_H1) { const uint32_t x = hf + 8u * (2u + (g.pad & 1u)); if (c0) c1 = x; else c0 = x; }
        if (g.commit & (7u * COMMIT_OUT0)) c0 = of + 8u * (uint32_t)(__ffs(g.commit >> 2) - 1);
        sm->igd[i].lo = make_uint4(g.idesc, g.b_hi, (g.b_off & 0x3FFFu) | (8u << 16), tmem + g.d_col);
        sm->igd[i].hi = make_uint4(ra_, g.waits, ctl_, c0 | (c1 << 16));
      }
      __syncwarp();
    }
    if (elect_one()) {
      const int n_ig = plan.n_ig;
      // (descriptor address fields hold bits 4..17 of the CTA-local address; the 32-bit shared address of a CTA with cluster rank > 0
      // carries the rank in its upper bits, which must not leak into the descriptor's LBO field)
      const uint32_t sm_a = smem_u32(sm), ring_a = (smem_u32(ring) >> 4) & 0x3FFFu, slot_a = slot_bytes >> 4, xm_step = xm_bytes >> 4;
      const uint32_t full0 = smem_u32(&sm->full[0]), empty0 = smem_u32(&sm->empty[0]);
      int s = 0; uint32_t ring_par = 0, L0 = 0, stage_a = ring_a;  // L0 = hidden layers completed before this tile: 6 * t
      for (uint32_t t = 0; t < (uint32_t)my_tiles; ++t, L0 += N_HID) {
        uint4 ra = sm->igd[0].lo, rb = sm->igd[0].hi;
        int blk_no = 0;
        for (int i = 0; i < n_ig; ++i) {
          const int in = i + 1 < n_ig ? i + 1 : 0;
          const uint4 na = sm->igd[in].lo, nb = sm->igd[in].hi;  // next record: its shared-memory latency hides behind this group's issue
          const uint32_t ctl = rb.z;
          if (ctl & CTL_NEWBLK) { mbar_wait_addr(full0 + 8u * (uint32_t)s, ring_par, err, 3); stamp<kDebug>(p, t, blk_no, 0); }
          for (uint32_t w = rb.y; w; w >>= 8) {
            const uint32_t id = (w >> 3) & 31u, c0 = w & 7u;
            cnt_wait(&sm->cnt[id], c0 + (id < CNT_TILE ? L0 : t), err, 20 + (int)id);
          }
          if (ctl & CTL_NEWBLK) stamp<kDebug>(p, t, blk_no, 1);
          tc_fence_after();
          const uint32_t d = ra.w, b_lo = ra.z + stage_a, b_hi = ra.y, idesc = ra.x;
          const uint32_t first = ctl & CTL_FIRST;
          const int nk = (int)(ctl & CTL_NK);
          if (ctl & CTL_SS) {
            const uint32_t a_lo = rb.x + ((ctl & CTL_XM) ? (t & 1u) * xm_step : 0u), a_hi = 8u | (1u << 14);      // LBO 2048 B, SBO 128 B
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < nk) mma_ss_p(d, a_lo + 256u * k, a_hi, b_lo + 16u * k, b_hi, idesc, (k > 0 || !first) ? 1u : 0u);
          } else {
            const uint32_t a = rb.x;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < nk) mma_ts_p(d, a + 8u * k, b_lo + 16u * k, b_hi, idesc, (k > 0 || !first) ? 1u : 0u);
          }
          const uint32_t cm = rb.w;
          if (cm) {
            tc_commit_addr(sm_a + (cm & 0xFFFFu));
            if (cm >> 16) tc_commit_addr(sm_a + (cm >> 16));
          }
          if (ctl & CTL_ENDBLK) {
            // frees the stage in both CTAs of the pair when these MMAs retire
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(empty0 + 8u * (uint32_t)s), "h"((uint16_t)3) : "memory");
            stamp<kDebug>(p, t, blk_no, 3);
            ++blk_no;
            stage_a += slot_a;
            if (++s == p.stages) { s = 0; ring_par ^= 1u; stage_a = ring_a; }
          }
          ra = na; rb = nb;
        }
      }
    }
    }
  } else if (warp < EPI_WARPS) {
    // ===================== hidden-layer epilogue groups: group j drains output chunk j of every hidden layer =====================
    const int j = warp >> 2, q = warp & 3, h = j >> 1;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t L = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = tile_of(it);
      for (int lh = 0; lh < N_HID; ++lh, ++L) {
        const EpiRec e = sm->epi[lh * N_GROUPS + j];
        const bool lead = kDebug && q == 0 && lane == 0;
        if (lead) stamp<kDebug>(p, (uint32_t)it, 32 + 4 * lh + j, 4);
        mbar_wait(&sm->half_full[h][L & 1u], (L >> 1) & 1u, err, 4);
        tc_fence_after();
        if (lead) stamp<kDebug>(p, (uint32_t)it, 32 + 4 * lh + j, 5);
        if (kDebug && p.dump_layer == (int)e.layer) {          // debug hook: raw accumulator to global
          const int64_t row0 = (int64_t)tile * TILE_M; const int tr = q * 32 + lane;
          for (int c0 = 0; c0 < (int)e.nc; c0 += 16) {
            uint32_t r[16]; tmem_ld16(lane_base + e.acc_col + (uint32_t)c0, r); tmem_ld_wait();
            if (row0 + tr < n) for (int jj = 0; jj < 16; ++jj) if (e.n0 + c0 + jj < e.n_real) p.dump_out[(row0 + tr) * e.n_real + e.n0 + c0 + jj] = __uint_as_float(r[jj]);
          }
        }
        const uint32_t src = lane_base + e.acc_col, dst = lane_base + e.out_col;
        if (e.silu) hidden_chunk_any<true>(src, dst, e.nc, e.one_rel, &sm->cnt[CNT_DRAIN + j], q, L + 1u, lane);
        else hidden_chunk_any<false>(src, dst, e.nc, e.one_rel, &sm->cnt[CNT_DRAIN + j], q, L + 1u, lane);
        if (lead) stamp<kDebug>(p, (uint32_t)it, 32 + 4 * lh + j, 6);
        tmem_st_wait();
        tc_fence_before();
        if (lane == 0) cnt_publish(&sm->cnt[CNT_ACT + j], q, L + 1u);      // activation chunk j of this layer is visible to the MMAs
        if (lead) stamp<kDebug>(p, (uint32_t)it, 32 + 4 * lh + j, 7);
      }
    }
  } else if (warp >= OUT_WARP0 && warp < OUT_WARP0 + 4) {
    // ===================== output group: policy head, the two output heads, stores, next tile's prologue ==========================
    const int q = warp & 3;
    const int t = q * 32 + lane;                                          // trajectory row of the tile == TMEM lane
    const int ct = threadIdx.x - OUT_WARP0 * 32;                          // 0..127
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const int No = plan.No;

    auto prefetch = [&](int tile, int buf) {                              // cp.async of a tile's raw states
      const int64_t r0 = (int64_t)tile * TILE_M;
      const int rws = max(0, min(TILE_M, n - (int)r0));
      float* dst = st_s + buf * TILE_M * p.SPs;
      if (rws > 0 && p.ready_flags) wait_rows_ready(p.ready_flags, p.ready_shift, (int)r0, (int)r0 + rws - 1, err);
      {                                                                   // element i = ct + 128 k of the tile's rws x S block (no divisions in the loop)
        int r = ct / S, cc = ct - r * S;
        const int dr = GROUP_THREADS / S, dc = GROUP_THREADS - dr * S;
        for (int i = ct; i < rws * S; i += GROUP_THREADS) {
          cp_async4(dst + r * p.SPs + cc, p.cur + r0 * S + i);
          r += dr; cc += dc; if (cc >= S) { cc -= S; ++r; }
        }
      }
      cp_async_commit();
    };
    auto prologue = [&](int tile, int buf, uint32_t it) {
      // ---- policy input [s, 1] and the state part of the member input [(s - mean)/(std + 1e-6), ..] -> shared-memory A tiles ----
      cp_async_wait_all();
      named_bar_sync(1, GROUP_THREADS);                                   // the tile's states landed (all of the group's copies)
      const float* ps = st_s + buf * TILE_M * p.SPs + t * p.SPs;
      const int xp_one = plan.xp_one;
      write_input_octets(xp, t, 0, plan.K0p >> 3, [&](int k) { return k < S ? ps[k] : (k == xp_one ? 1.f : 0.f); });
      write_input_octets(xm + (it & 1u) * xm_bytes, t, 0, S >> 3, [&](int k) { return (ps[k] - sm->norm_mean[k]) * sm->norm_inv[k]; });   // src/dynamics.py:113
      fence_proxy_async();                                                // generic-proxy writes of xp -> visible to the tensor core
      if (lane == 0) cnt_publish(&sm->cnt[CNT_TILE], q, it + 1u);
      // torch.normal in policy.act, randn_like in ensemble.sample: keyed by the row's global trajectory id (independent of sharding
      // and compaction); generated long before the head epilogues need them
      const int prow = tile * TILE_M + t;
      const bool pvalid = prow < n;
      const int64_t pid = pvalid ? (int64_t)p.ids[prow] : 0;
      *reinterpret_cast<float4*>(st_np + (buf * TILE_M + t) * 4) = pvalid ? noise_get4(p.noise_p, pid, 0, A) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.NM > 0)
        for (int cg = 0; 4 * cg < O; ++cg)
          *reinterpret_cast<float4*>(st_nm + (buf * TILE_M + t) * p.NM + 4 * cg) = pvalid ? noise_get4(p.noise_m, pid, cg, O) : make_float4(0.f, 0.f, 0.f, 0.f);
    };

    if (my_tiles > 0) { prefetch(tile_of(0), 0); prologue(tile_of(0), 0, 0u); }
    if (my_tiles > 1) prefetch(tile_of(1), 1);
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = tile_of(it), buf = it & 1;
      const int64_t row0 = (int64_t)tile * TILE_M;
      const int rows = max(0, min(TILE_M, n - (int)row0));
      const bool valid = t < rows;
      const int64_t row = row0 + t;
      const float* my_s = st_s + buf * TILE_M * p.SPs + t * p.SPs;
      float* my_o = st_o + t * p.OPs;
      // ---- policy head: [mu, raw] -> a = tanh(mu + exp(-6 + 10 sigmoid(raw)) eps)      src/policy.py:89-97 ----
      mbar_wait(&sm->out_full[0], it & 1, err, 7);
      tc_fence_after();
      if (ct == 0) stamp<kDebug>(p, (uint32_t)it, 60, 0);
      {
        uint32_t r[16]; tmem_ld16(lane_base + plan.head_col, r); tmem_ld_wait();
        if (kDebug && p.dump_layer == 2 && valid) for (int jj = 0; jj < 2 * A; ++jj) p.dump_out[row * 2 * A + jj] = __uint_as_float(r[jj]);
        const float4 e4 = *reinterpret_cast<const float4*>(st_np + (buf * TILE_M + t) * 4);
        const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
        const float mu4[4] = {__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3])};
        float raw4[4] = {0.f, 0.f, 0.f, 0.f};                              // raw_j = out[A + j], statically indexed per A
        if (A == 1) { raw4[0] = __uint_as_float(r[1]); }
        else if (A == 2) { raw4[0] = __uint_as_float(r[2]); raw4[1] = __uint_as_float(r[3]); }
        else if (A == 3) { raw4[0] = __uint_as_float(r[3]); raw4[1] = __uint_as_float(r[4]); raw4[2] = __uint_as_float(r[5]); }
        else { raw4[0] = __uint_as_float(r[4]); raw4[1] = __uint_as_float(r[5]); raw4[2] = __uint_as_float(r[6]); raw4[3] = __uint_as_float(r[7]); }
        float act4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          if (jj < A) {
            const float sd = __expf(-6.f + __fdividef(10.f, 1.f + __expf(-raw4[jj])));
            act4[jj] = tanh_fast(fmaf(ev[jj], sd, mu4[jj]));
          }
        }
        // member input x0 = [(s - mean)/(std + 1e-6), a, 1]  (src/dynamics.py:113-114): the octets that hold only state elements were
        // staged by the prologue; the ones from the first action element on are written here
        const int xm_one = plan.xm_one;
        write_input_octets(xm + (it & 1) * xm_bytes, t, S >> 3, plan.K0m >> 3, [&](int k) {
          if (k < S) return (my_s[k] - sm->norm_mean[k]) * sm->norm_inv[k];
          const int ja = k - S;
          const float av = ja == 0 ? act4[0] : (ja == 1 ? act4[1] : (ja == 2 ? act4[2] : (ja == 3 ? act4[3] : 0.f)));     // 0 beyond the A real actions
          return k == xm_one ? 1.f : av;
        });
        fence_proxy_async();
        tc_fence_before();
        if (lane == 0) cnt_publish(&sm->cnt[CNT_XM], q, (uint32_t)it + 1u);   // also: the head accumulator has been read
        if (ct == 0) stamp<kDebug>(p, (uint32_t)it, 60, 1);
        if (valid) {                                                     // off the critical path: the actions go to global memory
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) if (jj < A) p.actions[row * A + jj] = act4[jj];
        }
      }
      // ---- off the critical path: the next tile's prologue (its states were prefetched a tile ago) ----
      if (it + 1 < my_tiles) prologue(tile_of(it + 1), buf ^ 1, (uint32_t)it + 1u);
      if (ct == 0) stamp<kDebug>(p, (uint32_t)it, 60, 2);
      // ---- diff head: means = diffs + [s, 0]  (kept in shared memory)                   src/dynamics.py:118 ----
      mbar_wait(&sm->out_full[1], it & 1, err, 7);
      tc_fence_after();
      if (ct == 0) stamp<kDebug>(p, (uint32_t)it, 60, 3);
      for (int c0 = 0; c0 < No; c0 += 16) {
        uint32_t r[16]; tmem_ld16(lane_base + plan.d1_col + (uint32_t)c0, r);
        float sv[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) sv[jj] = (c0 + jj < S) ? my_s[c0 + jj] : 0.f;
        tmem_ld_wait();
        if (kDebug && p.dump_layer == 7 && valid) for (int jj = 0; jj < 16; ++jj) if (c0 + jj < O) p.dump_out[row * O + c0 + jj] = __uint_as_float(r[jj]);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) if (c0 + jj < O) my_o[c0 + jj] = __uint_as_float(r[jj]) + sv[jj];
      }
      // ---- log-var head + Gaussian sample                                              src/dynamics.py:119-121,201-203 ----
      mbar_wait(&sm->out_full[2], it & 1, err, 7);
      tc_fence_after();
      if (ct == 0) stamp<kDebug>(p, (uint32_t)it, 60, 4);
      for (int c0 = 0; c0 < No; c0 += 16) {
        uint32_t r[16]; tmem_ld16(lane_base + plan.v1_col + (uint32_t)c0, r);
        float ev[16], res[16];
#pragma unroll
        for (int jg = 0; jg < 4; ++jg) {
          const int cg = (c0 >> 2) + jg;
          float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (4 * cg < O) {
            if (p.NM > 0) e4 = *reinterpret_cast<const float4*>(st_nm + (buf * TILE_M + t) * p.NM + 4 * cg);
            else if (valid) e4 = noise_get4(p.noise_m, (int64_t)p.ids[row], cg, O);      // wide states: drawn in place
          }
          ev[4 * jg] = e4.x; ev[4 * jg + 1] = e4.y; ev[4 * jg + 2] = e4.z; ev[4 * jg + 3] = e4.w;
        }
        tmem_ld_wait();
        if (kDebug && p.dump_layer == 8 && valid) for (int jj = 0; jj < 16; ++jj) if (c0 + jj < O) p.dump_out[row * O + c0 + jj] = __uint_as_float(r[jj]);
        if (c0 + 16 >= No) {                                              // both output accumulators are in registers: release their slots
          tc_fence_before();
          if (lane == 0) cnt_publish(&sm->cnt[CNT_OUT], q, (uint32_t)it + 1u);
        }
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const int cc = min(c0 + jj, O - 1);
          const float u = __expf(sm->lv_hi[cc] - __uint_as_float(r[jj]));
          res[jj] = fmaf(sm->lv_s0[cc] * sqrt_fast(1.f + __fdividef(sm->lv_E[cc], 1.f + u)), ev[jj], my_o[cc]);
        }
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) if (c0 + jj < O) my_o[c0 + jj] = res[jj];
      }
      if (valid) p.rewards[row] = my_o[S];
      named_bar_sync(2, GROUP_THREADS);                                  // every row of the tile is final in st_o
      const float* so = st_o;
      {                                                                   // coalesced store of the tile's next states (no divisions in the loop)
        int r = ct / S, cc = ct - r * S;
        const int dr = GROUP_THREADS / S, dc = GROUP_THREADS - dr * S;
        for (int i = ct; i < rows * S; i += GROUP_THREADS) {
          p.next_states[row0 * S + i] = so[r * p.OPs + cc];
          r += dr; cc += dc; if (cc >= S) { cc -= S; ++r; }
        }
      }
      named_bar_sync(2, GROUP_THREADS);                                  // st_o / st_s[buf] are rewritten from here on
      if (ct == 0) stamp<kDebug>(p, (uint32_t)it, 60, 5);
      if (it + 2 < my_tiles) prefetch(tile_of(it + 2), buf);              // this tile's buffers are free again
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // no CTA leaves while its peer may still multicast into it or signal its barriers
  if (warp == ISSUER_WARP0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// host side: the static plan (TMEM map, weight blocks, MMA groups with their waits, epilogue records)
// ---------------------------------------------------------------------------------------------------------------
struct Chunks { int n0[4], nc[4]; };
static inline Chunks split4(int np) {            // four chunks in units of 16 columns (208 -> 64,48,48,48; 256 -> 4 x 64)
  Chunks c; const int units = np / 16; int n0 = 0;
  for (int j = 0; j < 4; ++j) { const int u = units / 4 + (j < units % 4 ? 1 : 0); c.n0[j] = n0; c.nc[j] = 16 * u; n0 += 16 * u; }
  return c;
}
static inline uint32_t wait_code(int id, int c0) { return (uint32_t)((id << 3) | c0); }

static int build_plan(const drpo_rollout_args& a, Plan& P) {
  P = Plan();
  PlanDev& D = P.d;
  memset(&D, 0, sizeof(D));
  const int S = a.ensemble->state_dim, A = a.ensemble->action_dim, Hm = a.ensemble->hidden, Hp = a.actor->l0.out_dim, O = S + 1;
  if (a.actor->l1.out_dim != Hp || a.actor->l1.in_dim != Hp || a.actor->l2.out_dim != 2 * A || Hp > 256 || Hm > 208 || Hp < 16 || Hm < 16 ||
      S + A + 1 > 64 || A > 4 || O > 64) {
    set_error("bf16 rollout: dims outside the fused kernel's plan (S=%d A=%d actor hidden=%d model hidden=%d)", S, A, Hp, Hm);
    return DRPO_ERR_UNSUPPORTED;
  }
  // a layer whose K is not a multiple of 16 carries its bias in the K slot behind the last real input (the activation holds a
  // constant 1 there); otherwise the bias is one extra k-step against the constant ones tile
  auto slot_bias = [](int k_real) { return k_real % 16 != 0; };
  const int Hpp = round_up(Hp + (slot_bias(Hp) ? 1 : 0), 16), Hmp = round_up(Hm + (slot_bias(Hm) ? 1 : 0), 16);
  const int K0p = round_up(S + (slot_bias(S) ? 1 : 0), 16), K0m = round_up(S + A + (slot_bias(S + A) ? 1 : 0), 16), No = round_up(O, 16);
  const Chunks cp = split4(Hpp), cm = split4(Hmp);
  D.K0p = K0p; D.K0m = K0m; D.No = No;
  D.xp_one = slot_bias(S) ? S : -1; D.xm_one = slot_bias(S + A) ? S + A : -1;
  // ---- TMEM map (512 columns) -------------------------------------------------------------------------------------------------
  // policy phase : pa [0,128) | pb [128,256) | accumulators of L0 / L1 [256,512) (L1 reuses L0's once they are drained); head at 256
  // member phase : P = h1 / d1 [0,Hmp/2) | Q = h2 / l1 [Hmp/2,Hmp) | slots SA (chunks 0+1) SB SC (chunks 2+3, alternating layers)
  const int PA = 0, PB = 128, ACC = 256;
  const int w0 = cm.nc[0] + cm.nc[1], w1 = cm.nc[2] + cm.nc[3];
  const int Pm = 0, Qm = Hmp / 2, SA = Hmp, SB = SA + w0, SC = SB + w1;
  if (SC + w1 > 512 || Hpp > 256 || No > w1) {
    set_error("bf16 rollout: TMEM plan does not fit (actor hidden %d, model hidden %d, state dim %d)", Hp, Hm, S);
    return DRPO_ERR_UNSUPPORTED;
  }
  D.head_col = ACC; D.d1_col = SB; D.v1_col = SC;
  const int n_real[N_LAYERS] = {Hp, Hp, 2 * A, Hm, Hm, Hm, Hm, O, O};
  const int k_real[N_LAYERS] = {S, Hp, Hp, S + A, Hm, Hm, Hm, Hm, Hm};
  for (int l = 0; l < N_LAYERS; ++l) { P.n_real[l] = n_real[l]; P.k_real[l] = k_real[l]; }

  uint32_t img_off = 0; std::vector<PackSub>* pack = &P.pack_pol;
  int cur_first_ig = 0; uint32_t cur_block_bytes = 0, cur_block_off = 0;
  uint32_t max_block = 0;
  auto begin_block = [&]() { cur_first_ig = D.n_ig; cur_block_bytes = 0; cur_block_off = img_off; };
  auto end_block = [&]() -> int {
    if (D.n_blocks >= MAX_BLOCKS) return -1;
    BlockRec& b = D.blk[D.n_blocks++];
    b.src_off = cur_block_off; b.bytes = cur_block_bytes; b.first_ig = (uint16_t)cur_first_ig; b.n_ig = (uint16_t)(D.n_ig - cur_first_ig);
    if (b.n_ig == 0) return -1;
    D.ig[cur_first_ig].flags |= IGF_NEWBLK; D.ig[D.n_ig - 1].flags |= IGF_ENDBLK;     // the issuer waits for / releases the ring stage here
    max_block = std::max(max_block, cur_block_bytes);
    return 0;
  };
  // one sub-block [nc x kc] of layer l (rows n0.., K range k0..) appended to the current block; returns its offset inside the block
  auto add_sub = [&](int l, int n0, int nc, int k0, int kc, int mode) -> uint32_t {
    PackSub s; s.layer = l; s.n0 = n0; s.nc = nc; s.k0 = k0; s.kc = kc; s.mode = mode; s.slot = slot_bias(k_real[l]) ? 1 : 0; s.dst_off = img_off;
    pack->push_back(s);
    const uint32_t off = cur_block_bytes, bytes = (uint32_t)nc * kc * 2;
    cur_block_bytes += bytes; img_off += bytes;
    return off;
  };
  auto add_ig = [&](int N, int d_col, bool ss, int ss_src, uint32_t a_src, uint32_t sub_off, int kc, int ks0, int nk, bool first, uint32_t waits,
                    int commit, int lh) -> int {
    if (D.n_ig >= MAX_IG) return -1;
    IG& g = D.ig[D.n_ig++];
    g.idesc = make_idesc(N);
    g.b_hi = (((uint32_t)(kc >> 3) * 128u) >> 4) | (1u << 14);          // SBO = stride between 8-column groups = (kc/8) * 128 B
    g.b_off = (sub_off >> 4) + 16u * (uint32_t)ks0;
    g.d_col = (uint32_t)d_col; g.a_src = a_src; g.waits = waits; g.nk = (uint16_t)nk;
    g.flags = (uint8_t)((first ? IGF_FIRST : 0) | (ss ? (IGF_SS | (ss_src << IGF_SRC_SHIFT)) : 0));
    g.commit = (uint8_t)commit; g.pad = (uint32_t)lh;                    // pad = hidden-layer index of the commit (barrier parity)
    return 0;
  };
  auto W4 = [](uint32_t a, uint32_t b = 0, uint32_t c = 0, uint32_t d = 0) { return a | (b << 8) | (c << 16) | (d << 24); };
  int rc = 0;
  // ---- input layer (SS-mode A from the staged tile): all k-steps of [N x K0] in one block ---------------------------------------------
  auto input_layer = [&](int l, int lh, int K0, int ss_src, const Chunks& ch, int d0, int d1, bool split, uint32_t waits) {
    begin_block();
    const bool ones = !slot_bias(k_real[l]);
    const int np = ch.n0[3] + ch.nc[3];
    const int ngroups = split ? 2 : 1;
    uint32_t sub[2], subb[2] = {0, 0}; int gn0[2], gnc[2], gd[2];
    for (int g = 0; g < ngroups; ++g) {
      gn0[g] = split ? ch.n0[2 * g] : 0; gnc[g] = split ? ch.nc[2 * g] + ch.nc[2 * g + 1] : np; gd[g] = g == 0 ? d0 : d1;
      sub[g] = add_sub(l, gn0[g], gnc[g], 0, K0, 0);
      if (ones) subb[g] = add_sub(l, gn0[g], gnc[g], 0, 16, 1);
    }
    for (int g = 0; g < ngroups; ++g) {
      if (gnc[g] == 0) continue;
      const int nks = K0 / 16;
      for (int ks = 0; ks < nks; ks += 4) {
        const int nk = std::min(4, nks - ks);
        const bool last = ks + nk == nks && !ones;
        rc |= add_ig(gnc[g], gd[g], true, ss_src, (uint32_t)(ks * 256), sub[g], K0, ks, nk, ks == 0, ks == 0 && g == 0 ? waits : 0u,
                     last ? (split ? (g == 0 ? COMMIT_H0 : COMMIT_H1) : (COMMIT_H0 | COMMIT_H1)) : 0, lh);
      }
      if (ones) rc |= add_ig(gnc[g], gd[g], true, 2, 0u, subb[g], 16, 0, 1, false, 0u, split ? (g == 0 ? COMMIT_H0 : COMMIT_H1) : (COMMIT_H0 | COMMIT_H1), lh);
    }
    rc |= end_block();
  };
  // ---- hidden / output layer whose A operand is the previous hidden layer's activation in TMEM: one block per part (= producer chunk) ----
  // groups: up to 2 MMA groups (n0, nc, d_col); `prod` = chunks of the producer (k ranges of the parts), `a_col` its activation region;
  // act_c0 = counter value offset of the producer layer; extra0[g] = additional waits of group g's first MMA (slot reuse)
  auto tmem_layer = [&](int l, int lh, const Chunks& prod, int a_col, int ngroups, const int* gn0, const int* gnc, const int* gd, int act_c0,
                        const uint32_t* extra0, const int* commit, bool one_block, bool h1_first_in_part0, bool skip_act_waits) {
    const bool ones = !slot_bias(k_real[l]);
    if (one_block) begin_block();
    if (ones) {                                                            // bias k-step against the ones tile: no dependency on the producer
      if (!one_block) begin_block();
      for (int g = 0; g < ngroups; ++g) {
        if (gnc[g] == 0) continue;
        const uint32_t sb = add_sub(l, gn0[g], gnc[g], 0, 16, 1);
        rc |= add_ig(gnc[g], gd[g], true, 2, 0u, sb, 16, 0, 1, true, extra0[g], 0, lh);
      }
      if (!one_block) rc |= end_block();
    }
    int last_part = -1;
    for (int pa = 0; pa < 4; ++pa) if (prod.nc[pa] > 0) last_part = pa;
    for (int pa = 0; pa < 4; ++pa) {
      if (prod.nc[pa] == 0) continue;
      if (!one_block) begin_block();
      uint32_t sub[2];
      for (int g = 0; g < ngroups; ++g) sub[g] = gnc[g] ? add_sub(l, gn0[g], gnc[g], prod.n0[pa], prod.nc[pa], 0) : 0u;
      for (int gi = 0; gi < ngroups; ++gi) {
        const int g = (pa == 0 && h1_first_in_part0 && ngroups == 2) ? 1 - gi : gi;
        if (gnc[g] == 0) continue;
        const bool first = pa == 0 && !ones;
        uint32_t waits = 0;
        (void)first;
        // The four chunks of the producer finish their epilogues within ~200 cycles of each other (measured), so a K-sliced start
        // buys nothing, while every counter wait is a ~200-cycle shared-memory round trip of the issuing thread: the layer's first
        // MMA group waits for all four activation chunks (which implies that their accumulator columns are drained), later parts
        // wait for nothing.  extra0 = slot occupants that are NOT chunks of the producer.
        if (pa == 0 && gi == 0 && !skip_act_waits)
          waits = wait_code(CNT_ACT + 0, act_c0) | (wait_code(CNT_ACT + 1, act_c0) << 8) | (wait_code(CNT_ACT + 2, act_c0) << 16) | (wait_code(CNT_ACT + 3, act_c0) << 24);
        if (pa == 0 && extra0[g] && !ones) waits = waits ? waits : extra0[g];
        rc |= add_ig(gnc[g], gd[g], false, 0, (uint32_t)(a_col + prod.n0[pa] / 2), sub[g], prod.nc[pa], 0, prod.nc[pa] / 16, first, waits,
                     pa == last_part ? commit[g] : 0, lh);
      }
      if (!one_block) rc |= end_block();
    }
    if (one_block) rc |= end_block();
  };
  const int one_n0[1] = {0};
  // ================= policy =================
  {
    input_layer(0, 0, K0p, 0, cp, ACC, ACC, false, W4(wait_code(CNT_TILE, 1), wait_code(CNT_OUT, 0)));
    // L1 reuses L0's accumulator columns: its first MMA (bias k-step or part 0) waits until all four L0 chunks are drained
    const int gnc[1] = {Hpp}, gd[1] = {ACC}, cm1[1] = {COMMIT_H0 | COMMIT_H1};
    const uint32_t ex[1] = {W4(wait_code(CNT_DRAIN + 0, 1), wait_code(CNT_DRAIN + 1, 1), wait_code(CNT_DRAIN + 2, 1), wait_code(CNT_DRAIN + 3, 1))};
    tmem_layer(1, 1, cp, PA, 1, one_n0, gnc, gd, 1, ex, cm1, false, false, false);         // (ex is used by the bias k-step only)
    // head: N = 16 (2A padded), accumulator over L1's chunk 0 columns (drained before part 0 may be issued)
    const int hnc[1] = {16}, hd[1] = {ACC}, hc[1] = {COMMIT_OUT0};
    const uint32_t hex[1] = {W4(wait_code(CNT_DRAIN + 0, 2))};
    tmem_layer(2, 1, cp, PB, 1, one_n0, hnc, hd, 2, hex, hc, true, false, false);
    D.n_pol_blocks = D.n_blocks; P.pol_bytes = img_off;
  }
  // ================= ensemble member =================
  {
    img_off = 0; pack = &P.pack_mem;
    const int hn0[2] = {cm.n0[0], cm.n0[2]}, hnc[2] = {w0, w1};
    const int cmH[2] = {COMMIT_H0, COMMIT_H1};
    const bool sab = SB == SA + w0;                                   // SA and SB are adjacent: trunk0 and the diff hidden layer accumulate with ONE N = Hmp MMA per k-step
    input_layer(3, 2, K0m, 1, cm, SA, SB, !sab, W4(wait_code(CNT_XM, 1)));
    {   // trunk1: H0 -> SA (trunk0's chunks 0,1: chunk 0 is covered by the activation wait, chunk 1 needs its own), H1 -> SC
      const int gd[2] = {SA, SC}; const uint32_t ex[2] = {0u, 0u};
      tmem_layer(4, 3, cm, Pm, 2, hn0, hnc, gd, 3, ex, cmH, false, false, false);
    }
    {   // diff hidden: H0 -> SA, H1 -> SB (trunk0's H1 was drained before trunk1's parts 2,3 could be issued)
      const int gd[2] = {SA, SB}; const uint32_t ex[2] = {0u, 0u};
      const int gnc1[1] = {Hmp}, cm1[1] = {COMMIT_H0 | COMMIT_H1};
      if (sab) tmem_layer(5, 4, cm, Qm, 1, one_n0, gnc1, gd, 4, ex, cm1, false, false, false);
      else tmem_layer(5, 4, cm, Qm, 2, hn0, hnc, gd, 4, ex, cmH, false, false, false);
    }
    {   // log-var hidden: independent of the diff hidden layer's epilogue; H1 -> SC first, H0 -> SA once diff's chunks 0,1 are drained
      const int gd[2] = {SA, SC}; const uint32_t ex[2] = {W4(wait_code(CNT_DRAIN + 0, 5), wait_code(CNT_DRAIN + 1, 5)), 0u};
      tmem_layer(6, 5, cm, Qm, 2, hn0, hnc, gd, 4, ex, cmH, false, true, true);
    }
    {   // output heads: diff -> SB (diff hidden's H1 slot), log-var -> SC (log-var hidden's H1 slot)
      const int onc[1] = {No};
      const int dd[1] = {SB}, dc[1] = {COMMIT_OUT0 << 1}; const uint32_t dex[1] = {0u};
      tmem_layer(7, 5, cm, Pm, 1, one_n0, onc, dd, 5, dex, dc, true, false, false);
      const int vd[1] = {SC}, vc[1] = {COMMIT_OUT0 << 2}; const uint32_t vex[1] = {0u};
      tmem_layer(8, 5, cm, Qm, 1, one_n0, onc, vd, 6, vex, vc, true, false, false);
    }
    P.mem_bytes = img_off;
  }
  if (rc) { set_error("bf16 rollout: plan tables overflow (%d MMA groups, %d blocks)", D.n_ig, D.n_blocks); return DRPO_ERR_UNSUPPORTED; }
  // every block is fetched in two halves (one per CTA of the pair): 32-byte granularity
  for (int b = 0; b < D.n_blocks; ++b)
    if (D.blk[b].bytes % 32 != 0 || D.blk[b].src_off % 16 != 0) { set_error("bf16 rollout: internal block alignment (block %d)", b); return DRPO_ERR_UNSUPPORTED; }
  D.slot_bytes = (max_block + 1023u) & ~1023u;
  // ---- epilogue records: hidden layer lh (0 L0, 1 L1, 2 trunk0, 3 trunk1, 4 diff hidden, 5 log-var hidden), group j -------------------
  struct HL { int layer; const Chunks* ch; int acc0, acc1; int out; int silu; int width; };
  const HL hl[N_HID] = {{0, &cp, ACC, ACC + cp.n0[2], PA, 0, Hp}, {1, &cp, ACC, ACC + cp.n0[2], PB, 0, Hp},
                        {3, &cm, SA, SB, Pm, 1, Hm}, {4, &cm, SA, SC, Qm, 1, Hm}, {5, &cm, SA, SB, Pm, 1, Hm}, {6, &cm, SA, SC, Qm, 1, Hm}};
  for (int lh = 0; lh < N_HID; ++lh)
    for (int j = 0; j < N_GROUPS; ++j) {
      EpiRec& e = D.epi[lh * N_GROUPS + j];
      const Chunks& ch = *hl[lh].ch;
      e.n0 = (uint16_t)ch.n0[j]; e.nc = (uint16_t)ch.nc[j];
      e.acc_col = (uint16_t)((j < 2 ? hl[lh].acc0 : hl[lh].acc1) + ch.n0[j] - ch.n0[j < 2 ? 0 : 2]);
      e.out_col = (uint16_t)(hl[lh].out + ch.n0[j] / 2);
      e.silu = (uint8_t)hl[lh].silu; e.layer = (uint8_t)hl[lh].layer; e.n_real = (uint16_t)hl[lh].width;
      const int w = hl[lh].width;                                         // consumer's bias slot = feature index `width` when K % 16 != 0
      e.one_rel = (int16_t)((slot_bias(w) && w >= ch.n0[j] && w < ch.n0[j] + ch.nc[j]) ? w - ch.n0[j] : -1);
    }
  return DRPO_OK;
}

static const StdIG kStdHost[N_STD_IG] = {
    {0, SRC_XP, 1, 1, 1, 1, CM_BOTH, 0},                                                      //  0 actor L0
    {1, SRC_ONES, 1, 1, 1, 1, CM_NONE, 1},                                                    //  1 actor L1 bias
    {4, SRC_TS, 0, 1, 1, 1, CM_NONE, 1}, {4, SRC_TS, 0, 1, 1, 0, CM_NONE, 1},                 //  2-5 actor L1 parts
    {4, SRC_TS, 0, 1, 1, 0, CM_NONE, 1}, {4, SRC_TS, 0, 1, 1, 0, CM_BOTH, 1},
    {1, SRC_ONES, 1, 1, 0, 1, CM_NONE, 1},                                                    //  6 head bias
    {4, SRC_TS, 0, 0, 0, 1, CM_NONE, 1}, {4, SRC_TS, 0, 0, 0, 0, CM_NONE, 1},                 //  7-10 head parts
    {4, SRC_TS, 0, 0, 0, 0, CM_NONE, 1}, {4, SRC_TS, 0, 0, 1, 0, CM_OUT0, 1},
    {0, SRC_XM, 1, 1, 1, 1, CM_BOTH, 2},                                                      // 11 trunk0
    {4, SRC_TS, 1, 1, 0, 1, CM_NONE, 3}, {4, SRC_TS, 1, 0, 1, 0, CM_NONE, 3},                 // 12-19 trunk1 (H0, H1) x 4 parts
    {3, SRC_TS, 0, 1, 0, 0, CM_NONE, 3}, {3, SRC_TS, 0, 0, 1, 0, CM_NONE, 3},
    {3, SRC_TS, 0, 1, 0, 0, CM_NONE, 3}, {3, SRC_TS, 0, 0, 1, 0, CM_NONE, 3},
    {3, SRC_TS, 0, 1, 0, 0, CM_H0, 3}, {3, SRC_TS, 0, 0, 1, 0, CM_H1, 3},
    {4, SRC_TS, 1, 1, 1, 1, CM_NONE, 4}, {3, SRC_TS, 0, 1, 1, 0, CM_NONE, 4},                 // 20-23 diff hidden (N = 208) x 4 parts
    {3, SRC_TS, 0, 1, 1, 0, CM_NONE, 4}, {3, SRC_TS, 0, 1, 1, 0, CM_BOTH, 4},
    {4, SRC_TS, 1, 1, 0, 0, CM_NONE, 5}, {4, SRC_TS, 1, 0, 1, 1, CM_NONE, 5},                 // 24-31 log-var hidden (H1 first in part 0)
    {3, SRC_TS, 0, 1, 0, 0, CM_NONE, 5}, {3, SRC_TS, 0, 0, 1, 0, CM_NONE, 5},
    {3, SRC_TS, 0, 1, 0, 0, CM_NONE, 5}, {3, SRC_TS, 0, 0, 1, 0, CM_NONE, 5},
    {3, SRC_TS, 0, 1, 0, 0, CM_H0, 5}, {3, SRC_TS, 0, 0, 1, 0, CM_H1, 5},
    {4, SRC_TS, 1, 1, 0, 1, CM_NONE, 5}, {3, SRC_TS, 0, 0, 0, 0, CM_NONE, 5},                 // 32-35 diff head
    {3, SRC_TS, 0, 0, 0, 0, CM_NONE, 5}, {3, SRC_TS, 0, 0, 1, 0, CM_OUT1, 5},
    {4, SRC_TS, 1, 1, 0, 1, CM_NONE, 5}, {3, SRC_TS, 0, 0, 0, 0, CM_NONE, 5},                 // 36-39 log-var head
    {3, SRC_TS, 0, 0, 0, 0, CM_NONE, 5}, {3, SRC_TS, 0, 0, 1, 0, CM_OUT2, 5}};

// does the plan have the compile-time structure of kStd (same groups, k-steps, operand sources, block boundaries, waits, commits)?
static bool std_structure_matches(const Plan& P) {
  const StdIG* h = kStdHost;
  if (P.d.n_ig != N_STD_IG) return false;
  for (int i = 0; i < N_STD_IG; ++i) {
    const IG& g = P.d.ig[i]; const StdIG& e = h[i];
    const int src = (g.flags & IGF_SS) ? 1 + ((g.flags >> IGF_SRC_SHIFT) & 3) : SRC_TS;
    int cm = CM_NONE;
    if ((g.commit & 3) == 3) cm = CM_BOTH; else if (g.commit & COMMIT_H0) cm = CM_H0; else if (g.commit & COMMIT_H1) cm = CM_H1;
    else if (g.commit & (7 * COMMIT_OUT0)) cm = CM_OUT0 + (__builtin_ffs(g.commit >> 2) - 1);
    if ((e.nk && e.nk != (int)g.nk) || (int)g.nk < 1 || (int)g.nk > 4 || src != e.src || ((g.flags & IGF_FIRST) != 0) != (e.first != 0) ||
        ((g.flags & IGF_NEWBLK) != 0) != (e.newblk != 0) || ((g.flags & IGF_ENDBLK) != 0) != (e.endblk != 0) || (g.waits != 0) != (e.waits != 0) ||
        cm != e.commit || (cm >= CM_H0 && cm <= CM_BOTH && ((int)g.pad & 1) != (e.lh & 1)))
      return false;
  }
  return true;
}

static int smem_bytes_for(const Plan& P, int S, int A, int stages, int& SPs, int& OPs, int& NM) {
  SPs = S | 1;                                             // raw state row (odd stride: conflict-free row-per-thread access)
  OPs = (S + 1) | 1;                                       // [means], then [next state, reward] of the current tile
  NM = (S + 1) <= 16 ? round_up(S + 1, 4) + 4 : 0;         // staged model noise (+4: conflict-free float4 rows); wide states draw in place
  return (int)((size_t)P.d.slot_bytes * stages + (size_t)TILE_M * (P.d.K0p + 2 * P.d.K0m + 16) * 2 +
               (size_t)TILE_M * (2 * SPs + OPs + 2 * 4 + 2 * NM) * 4 + sizeof(SmemCtl) + 64);
}

// pack the actor (pack_pol) or one member (pack_mem) into its image
static void pack_jobs(const Plan& P, const std::vector<PackSub>& subs, const drpo_linear* lin /* indexed by layer - base */, int base, uint8_t* img,
                      std::vector<PackJob>& jobs) {
  for (const PackSub& s : subs) {
    const drpo_linear& L = lin[s.layer - base];
    PackJob j;
    j.W = L.w; j.b = L.b; j.n_real = P.n_real[s.layer]; j.k_real = P.k_real[s.layer];
    j.n0 = s.n0; j.nc = s.nc; j.k0 = s.k0; j.kc = s.kc; j.mode = s.mode; j.slot = s.slot;
    j.dst = reinterpret_cast<__nv_bfloat16*>(img + s.dst_off);
    jobs.push_back(j);
  }
}
static int pack_flush(std::vector<PackJob>& jobs, void* stream) {
  for (size_t i0 = 0; i0 < jobs.size(); i0 += PACK_JOBS) {
    PackTable t; memset(&t, 0, sizeof(t));
    const int nj = (int)std::min<size_t>(PACK_JOBS, jobs.size() - i0);
    for (int i = 0; i < nj; ++i) t.job[i] = jobs[i0 + i];
    dim3 grid(8, nj);
    DRPO_LAUNCH(pack_sub_kernel, grid, 256, 0, stream, t);
  }
  jobs.clear();
  return DRPO_OK;
}

}  // namespace r2
}  // namespace drpo
