// DRPO_PREC_BF16 rollout step, third generation: TWO 128-row tiles in flight per SM (src/smbpo.py:234-246, src/policy.py:89-97,
// src/dynamics.py:112-122,198-203).  One persistent, warp-specialised tcgen05 kernel per rollout step keeps both tiles on the SM for the
// whole   policy MLP -> squashed-Gaussian sample -> ensemble-member MLP (trunk + 2 heads) -> Gaussian next-state sample   chain.
//
// Why two tiles: the layers of one tile form a dependent chain  MMA -> epilogue -> MMA -> ...; with a single tile per SM the tensor pipe
// idles during every epilogue (generations 1 and 2 of this kernel: 23-25 % tensor-pipe activity, 30-40 k cycles per tile of which 7-10 k
// are MMAs).  Here the SM holds two tile SLOTS; the single MMA-issuing thread alternates  job(slot 0, j), job(slot 1, j), job(slot 0, j+1)..
// so the MMAs of one slot run while the other slot's epilogue drains its accumulator.
//
// What makes two tiles fit:
//   * TMEM (512 columns) = 2 slots x 256 columns: a hidden layer accumulates ALL its output columns at once (one N = 256 / 208 MMA per
//     k-step, the N/2-cycle floor of the tensor pipe), the three narrow heads reuse drained / spare columns of the slot's region;
//   * hidden activations live in SHARED memory, not in TMEM: the epilogue writes them as bf16 in the UMMA canonical K-major layout
//     ((k/8) * 2048 + row * 16 bytes: a warp's 16-byte stores are 512 contiguous bytes) into the slot's activation buffer X, the next
//     layer reads them as the A operand of SS-mode MMAs (full rate at N >= 208, tools/ubench_tcgen05.cu), and its own epilogue
//     overwrites X in place (all its MMAs have completed by then);
//   * the one activation that must coexist with its predecessor (the diff-hidden layer: the log-var hidden layer still needs the trunk
//     output) is packed in place over its own accumulator in TMEM and read by TS-mode MMAs.
// Per slot and tile the job list is  P0 P1 P2 | T0 T1 D0 D1 V0 V1  (actor L0, L1, head | trunk0, trunk1, diff hidden, diff head, log-var
// hidden, log-var head).  Weights: per job a sequence of k-step tiles [N x 16] (bias: K slot behind the last real input, or one extra
// k-step against a constant tile of ones with the bias split in (hi, lo) bf16 parts), streamed L2 -> shared-memory ring in blocks of <= 4
// k-steps by TMA bulk copies; CTA pairs (clusters of 2) fetch every block once and multicast it into both rings.
// Warp roles (704 threads): warps 0-11 hidden epilogue: three groups of four warps, group g drains column chunk g (96/80/80 or 80/64/64
// columns) of every hidden layer of BOTH slots in issue order (warp q of a group owns TMEM lanes 32q..32q+31, double-buffered
// tcgen05.ld); warps 12-15 / 16-19 output group of slot 0 / 1 (policy head -> action and member input, diff / log-var heads ->
// Gaussian sample, stores, the slot's next prologue, Philox draws); warp 20 TMA producer; warp 21 MMA issuer.  Epilogue -> issuer
// signalling: monotone counters in shared memory (release / acquire); MMA -> epilogue: mbarriers armed by tcgen05.commit.  Every wait
// is bounded (err_flag), never a hang.
#pragma once
#include <cuda_bf16.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "nets.cuh"
#include "tc05.cuh"

namespace drpo {
namespace r3 {
using namespace tc;

constexpr int TILE_M = 128, CLUSTER = 2, N_SLOTS = 2, N_JOBS = 9, N_HID = 6, N_GROUPS = 3, SLOT_COLS = 256;
constexpr int HID_WARP0 = 0, OUT_WARP0 = 4 * N_GROUPS, PRODUCER_WARP = OUT_WARP0 + 8, ISSUER_WARP = PRODUCER_WARP + 1;
constexpr int NUM_THREADS = (ISSUER_WARP + 1) * 32, GROUP_THREADS = 128;
constexpr int MAX_STAGES = 8;
// layer ids (debug dumps, weight packing): 0 actor L0, 1 actor L1, 2 actor head, 3 trunk0, 4 trunk1, 5 diff hidden, 6 log-var hidden,
// 7 diff head, 8 log-var head.  Job order of a tile: layers 0 1 2 3 4 5 7 6 8.
constexpr int N_LAYERS = 9;

// per-slot signals epilogue / output group -> issuer: mbarriers (one arrival per warp of the signalling group), every phase of which the
// issuer observes in order (phase k of a signal = the k-th event; a job waits for phase  c0 + inc * tile - 1)
enum { C_HID = 0, C_TILE, C_XM, C_D1R, C_OUT, N_CNT };
enum { A_XP = 0, A_X = 1, A_TMEM = 2 };
enum { BAR_HID = 0, BAR_OUT0 = 1 /* + k */ };

struct JobDev {
  uint32_t idesc, d_col, a_kind, a_off;          // a_off: TMEM column (A_TMEM) / 16-byte units inside the buffer (SS)
  uint32_t n_ks, ones_ks, kstep_units, kb, nfull;       // k-steps (incl. the bias k-step), index of the bias k-step or 0xFFFF, B tile size / 16, k-steps per block
  uint32_t img_off, wait0, wait1, bar;           // waits: id | c0 << 8 | per-tile increment << 16 ; 0 = none
};
// one column chunk [n0, n0 + nc) of a hidden layer: accumulator at column n0 of the slot region, packed in-place activation at out_col
struct EpiRec { uint16_t n0, nc, out_col, n_real; int16_t one; uint8_t silu, to_smem, layer, pad[3]; };

struct PlanDev {
  JobDev job[N_JOBS];
  EpiRec epi[N_HID * N_GROUPS];
  // The static schedule of one iteration: entries (slot, job, lag) in issue order; an entry with lag 1 works on the slot's tile of the
  // PREVIOUS iteration (slot 1 runs half a tile behind slot 0, so that one slot's head / output phases overlap the other slot's wide
  // layers).  hsched = the hidden-layer entries of sched (index into epi), in the same order.
  uint8_t sched[2 * N_JOBS][4], hsched[2 * N_HID][4];
  uint16_t ts_col[16];         // TMEM column (inside the slot region) of k-step ks of the in-place packed diff-hidden activation
  int K0p, K0m, No;            // padded K of the two input layers, padded N of the two output heads
  int xp_one, xm_one;          // position of the constant 1 (bias slot) in the staged inputs, -1: none
  int head_col, d1_col, v1_col, wide;       // TMEM columns (inside the slot region) of the three head accumulators; wide: O > 16
  uint32_t slot_bytes, x_bytes, xp_bytes;
};

struct Plan {
  PlanDev d;
  uint32_t pol_bytes, mem_bytes;
  int n_pad[N_LAYERS], n_real[N_LAYERS], k_real[N_LAYERS], n_ks[N_LAYERS], ones_ks[N_LAYERS], slot[N_LAYERS], job_of[N_LAYERS];
};

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------------------------
// weight images: per job a sequence of k-step tiles, each the canonical K-major tile  [n/8][2][8 rows][8 elems]  of W[n0.., 16 ks..]
// ---------------------------------------------------------------------------------------------------------------
struct PackJob { const float* W; const float* b; int n_real, k_real, n_pad, n_ks, ones_ks, slot; __nv_bfloat16* dst; };
constexpr int PACK_JOBS = 48;
struct PackTable { PackJob job[PACK_JOBS]; };
static __global__ void pack_steps_kernel(const __grid_constant__ PackTable t) {
  const PackJob& j = t.job[blockIdx.y];
  const int per_step = j.n_pad * 16, total = per_step * j.n_ks;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ks = i / per_step, r = i - ks * per_step;
    const int n = r >> 4, kk = r & 15;
    float v = 0.f;
    if (n < j.n_real) {
      if (ks == j.ones_ks) {                                     // bias k-step against the ones tile: k = 0 -> bf16(b), k = 1 -> bf16(b - bf16(b))
        const float bb = j.b[n];
        const float hi = __bfloat162float(__float2bfloat16_rn(bb));
        v = kk == 0 ? hi : (kk == 1 ? bb - hi : 0.f);
      } else {
        const int gk = 16 * ks + kk;
        v = gk < j.k_real ? j.W[(int64_t)n * j.k_real + gk] : (gk == j.k_real && j.slot ? j.b[n] : 0.f);
      }
    }
    j.dst[(int64_t)ks * per_step + ((n >> 3) * 2 + (kk >> 3)) * 64 + (n & 7) * 8 + (kk & 7)] = __float2bfloat16_rn(v);
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
  int S, A, stages;
  int* err_flag;
  int dump_layer; float* dump_out;       // debug: dump the fp32 accumulator of one layer (100: clock stamps)
  const int32_t* ready_flags; int ready_shift;     // step 0 with streamed start states: flag of every 2^shift-row block (else NULL)
};

struct SmemCtl {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES], hid_full[N_SLOTS], out_full[N_SLOTS][3], pad0;
  uint64_t sig[N_SLOTS][N_CNT];
  uint32_t tmem_base, pad[3];
  // per-dim constants of the member, staged once per CTA: normaliser, and the log-var soft clamp folded into
  //   std = exp(lv/2) = s0 * sqrt(1 + E / (1 + exp(hi - x)))   with s0 = exp(lo/2), E = exp(hi - lo)      (src/dynamics.py:120-121,201)
  float norm_mean[64], norm_inv[64], lv_hi[64], lv_E[64], lv_s0[64];
  EpiRec epi[N_HID * N_GROUPS];
};

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

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 16 accumulator columns (already in registers), columns n0 + c .. of the layer -> activation -> packed bf16 ->
//   kSmem : the slot's activation buffer (A operand of the next layer's SS-mode MMAs): octet o of this thread's row at  xrow + o * 2048
//   !kSmem: TMEM, packed in place over the chunk's own accumulator columns (tdst + c / 2): A operand of TS-mode MMAs
template <bool kSilu, bool kSmem>
__device__ __forceinline__ void act_store(const uint32_t (&r)[16], int c, int n0, int one, uint32_t xrow, uint32_t tdst) {
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (kSilu) pk[j] = silu_bf16x2(pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])));
    else pk[j] = pack_bf16_relu(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
  }
  const int o1 = one - n0 - c;
  if (o1 >= 0 && o1 < 16) {                                     // rare: the piece that holds the consumer's bias slot (constant 1)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (o1 == 2 * j) pk[j] = (pk[j] & 0xFFFF0000u) | 0x00003F80u;
      if (o1 == 2 * j + 1) pk[j] = (pk[j] & 0x0000FFFFu) | 0x3F800000u;
    }
  }
  if constexpr (kSmem) {
    const uint32_t a = xrow + (uint32_t)((n0 + c) >> 3) * 2048u;
    st_shared_v4(a, pk[0], pk[1], pk[2], pk[3]);
    st_shared_v4(a + 2048u, pk[4], pk[5], pk[6], pk[7]);
  } else {
    tmem_st8(tdst + (uint32_t)(c >> 1), pk);
  }
}
// One column chunk (nc columns, multiple of 16) of a hidden layer's epilogue for this thread's row: 32 columns per round (two
// tcgen05.ld in flight, one wait).  Deliberately a small rolled loop: the kernel's hot code has to stay inside the instruction
// cache (an earlier version with every chunk width and the whole issue program unrolled was ~170 KB of SASS and ran 2-3x slower
// in every role); the load latency of one warp is covered by the other two epilogue warps of its scheduler.
template <bool kSilu, bool kSmem>
__device__ __forceinline__ void hidden_chunk(uint32_t src, int nc, int n0, int one, uint32_t xrow, uint32_t tdst) {
#pragma unroll 1
  for (int c = 0; c < nc; c += 32) {
    uint32_t ra[16], rb[16];
    const bool two = c + 16 < nc;
    tmem_ld16(src + (uint32_t)c, ra);
    if (two) tmem_ld16(src + (uint32_t)(c + 16), rb);
    tmem_ld_wait();
    act_store<kSilu, kSmem>(ra, c, n0, one, xrow, tdst);
    if (two) act_store<kSilu, kSmem>(rb, c + 16, n0, one, xrow, tdst);
  }
}

// octets [o0, o1) of one row of a layer input -> bf16 K-major canonical tile in shared memory: element (row, k) at
// (k/8) * 2048 + row * 16 + (k%8) * 2  (A operand of an SS-mode MMA: LBO = 2048, SBO = 128); `val(k)` yields element k
template <typename F>
__device__ __forceinline__ void write_input_octets(uint32_t tile_row, int o0, int o1, F val) {
  for (int o = o0; o < o1; ++o) {
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = val(8 * o + e);
    st_shared_v4(tile_row + 2048u * o, pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

// four consecutive elements k0 .. k0+3 (k0 a multiple of 4) of one fp32 row of S elements: one 16-byte load when the row is 16-byte
// aligned (vec: S % 4 == 0) and the group lies inside it - each thread reads its OWN row, so a scalar load is 32 sectors per warp
// instruction and the instruction count is what the load/store unit pays for; zeros beyond the row or when the row is not valid
// (kLdg: read-only data of this launch through the non-coherent path; false for rows this very thread stored earlier in the kernel)
template <bool kLdg>
__device__ __forceinline__ void load_row4(const float* row, int k0, int S, bool valid, bool vec, float (&f)[4]) {
  if (!valid || k0 >= S) { f[0] = f[1] = f[2] = f[3] = 0.f; return; }
  if (vec && k0 + 4 <= S) {
    const float4 v = kLdg ? __ldg(reinterpret_cast<const float4*>(row + k0)) : *reinterpret_cast<const float4*>(row + k0);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) f[e] = k0 + e < S ? (kLdg ? __ldg(row + k0 + e) : row[k0 + e]) : 0.f;
  }
}
__device__ __forceinline__ void store_row4(float* row, int k0, int S, bool vec, const float (&f)[4]) {      // elements k0..k0+3 that lie inside the row
  if (k0 >= S) return;
  if (vec && k0 + 4 <= S) *reinterpret_cast<float4*>(row + k0) = make_float4(f[0], f[1], f[2], f[3]);
  else {
#pragma unroll
    for (int e = 0; e < 4; ++e) if (k0 + e < S) row[k0 + e] = f[e];
  }
}

// the same, four octets (32 elements) per round: the global-memory reads of the whole round (the thread's own state row `row`, S
// elements, zeros behind it) are issued before the first use, so a wide state costs two or three memory round trips instead of one
// per octet; `fin(k, x)` turns the loaded value into element k
// `keep` (a thread-local array of 64 floats, i.e. lane-interleaved local memory: every later access is one coalesced wavefront)
// receives the raw row, so that the tile's later passes over its state do not go back to global memory with 32-sector accesses
// `keepn` (32 words) receives nrm(k, x) of the same elements as packed bf16 pairs: the member input's state part, ready to be copied
// when the action arrives (that copy is on the tile's critical path, the prologue is not)
template <typename F, typename G>
__device__ __forceinline__ void write_input_octets_wide(uint32_t tile_row, int n_oct, const float* row, int S, bool valid, bool vec, F fin, float* keep,
                                                        uint32_t* keepn, G nrm) {
#pragma unroll 1
  for (int o0 = 0; o0 < n_oct; o0 += 4) {
    float f[32];
#pragma unroll
    for (int e4 = 0; e4 < 8; ++e4) {
      float v[4];
      load_row4<true>(row, 8 * o0 + 4 * e4, S, valid, vec, v);
      f[4 * e4] = v[0]; f[4 * e4 + 1] = v[1]; f[4 * e4 + 2] = v[2]; f[4 * e4 + 3] = v[3];
    }
    if (8 * o0 < 64) {
#pragma unroll
      for (int e = 0; e < 32; ++e) keep[8 * o0 + e] = f[e];
#pragma unroll
      for (int e = 0; e < 16; ++e) keepn[4 * o0 + e] = pack_bf16(nrm(8 * o0 + 2 * e, f[2 * e]), nrm(8 * o0 + 2 * e + 1, f[2 * e + 1]));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (o0 + q < n_oct) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pk[j] = pack_bf16(fin(8 * (o0 + q) + 2 * j, f[8 * q + 2 * j]), fin(8 * (o0 + q) + 2 * j + 1, f[8 * q + 2 * j + 1]));
        st_shared_v4(tile_row + 2048u * (o0 + q), pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
}

// the member input from the thread-local packed copy of the normalised state (`locn`, 32 words = 64 bf16, written by the prologue):
// octets that lie inside the state are copied, the others take `tail(k)` (actions, the bias slot, zeros) behind the state.  The copy
// lives in L2 (the output groups' local arrays exceed the small L1 left beside 227 KB of shared memory): all loads are issued first.
template <typename F>
__device__ __forceinline__ void write_input_octets_packed(uint32_t tile_row, int n_oct, const uint32_t* locn, int S, F tail) {
  uint32_t w[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) w[e] = locn[e];
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    if (o < n_oct) {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k0 = 8 * o + 2 * j;
        if (k0 + 1 < S) pk[j] = w[4 * o + j];
        else pk[j] = pack_bf16(k0 < S ? __uint_as_float(w[4 * o + j] << 16) : tail(k0), tail(k0 + 1));
      }
      st_shared_v4(tile_row + 2048u * o, pk[0], pk[1], pk[2], pk[3]);
    }
  }
  for (int o = 8; o < n_oct; ++o)                        // (K0m > 64 only when S + A + 1 > 64)
    st_shared_v4(tile_row + 2048u * o, pack_bf16(tail(8 * o), tail(8 * o + 1)), pack_bf16(tail(8 * o + 2), tail(8 * o + 3)),
                 pack_bf16(tail(8 * o + 4), tail(8 * o + 5)), pack_bf16(tail(8 * o + 6), tail(8 * o + 7)));
}

// debug timing (kDebug build, dump_layer == 100): CTA 0 stamps clock() for its first 4 iterations into dump_out viewed as uint32
// [(it * 64 + idx) * 8 + k].  idx = slot * 32 + e:  e 0..8 = job (issuer: k 0 counter waits done, 1 first block's weights there, 2 issued);
// e 10..15 = hidden epilogue i (k 4 wait begin, 5 accumulator full, 6 activation published); e 20 = output group (k 0 head full,
// 1 member input published, 2 next prologue done, 3 diff head full, 4 log-var head full, 5 stores done)
template <bool kDebug>
__device__ __forceinline__ void stamp(const StepParams& p, uint32_t it, int idx, int k) {
  if (kDebug) {
    if (p.dump_layer == 100 && blockIdx.x == 0 && it < 4)
      reinterpret_cast<uint32_t*>(p.dump_out)[(it * 64 + idx) * 8 + k] = (uint32_t)clock();
  }
}

// ---- CTA-pair (cta_group::2) forms: one tcgen05.mma issued by the leader CTA multiplies BOTH CTAs' 128-row tiles (M = 256) with a weight
// block of which each CTA holds half the rows in its own ring; completion is committed to the barrier at the same offset in both CTAs.
__device__ __forceinline__ void mma2_ss_p(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.u32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n"
               ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma2_ts_p(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 d;\n\tsetp.ne.u32 p, %5, 0;\n\tmov.b64 d, {%2, %3};\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], d, %4, p;\n\t}\n"
               ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit2_addr(uint32_t addr) {           // arrives on the barrier at this offset in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(addr), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t cols) {      // the same warp of both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// wait with cluster-scope acquire: the arrivals come from both CTAs of the pair
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int* err_flag, int code) { mbar_wait(bar, parity, err_flag, code); }
// signal to the leader CTA's issuer: an arrival on the barrier in the leader's shared memory (local for the leader's own warps)
// (default semantics - release at CTA scope - as CUTLASS's ClusterBarrier::arrive(cta_id): a cluster-scope release costs a ~1500-cycle
// L1 flush per arrival, and what is published here is this CTA's own shared memory / TMEM, consumed by this CTA's own tensor core)
__device__ __forceinline__ void arrive_leader(uint64_t* bar, uint32_t crank) {
  if (crank == 0) mbar_arrive(bar);
  else asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 0;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sig_arrive(uint64_t* bar, uint32_t crank) { arrive_leader(bar, crank); }

struct IssueCtx {
  uint32_t ring_a, slot_a, full0, empty0;
  uint32_t ring_par, stage_a; int s, stages;
  uint32_t look;                 // 1: the full barrier of the CURRENT stage was already seen complete (by the previous block's look-ahead)
};
__device__ __forceinline__ void ring_release_advance(IssueCtx& c) {
  // frees the stage in both CTAs of the pair when the MMAs issued so far retire
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(c.empty0 + 8u * (uint32_t)c.s), "h"((uint16_t)3) : "memory");
  c.stage_a += c.slot_a;
  if (++c.s == c.stages) { c.s = 0; c.ring_par ^= 1u; c.stage_a = c.ring_a; }
}
// the current stage's weights have landed (usually known from the look-ahead: no shared-memory round trip on the issue path)
__device__ __forceinline__ void ring_wait_full(IssueCtx& c, int* err) {
  if (!c.look) mbar_wait_addr(c.full0 + 8u * (uint32_t)c.s, c.ring_par, err, 3);      // (CTA-scope acquire + tcgen05 fence; the peer's half
                                                                                       // was written by ITS async proxy and is read by ITS tensor core)
  c.look = 0;
  tc_fence_after();
}
// KB MMAs of one weight block, issued back to back, with a non-blocking probe of the NEXT stage's full barrier started before them and
// read after them: a mbarrier probe costs the single issuing thread a 60-200-cycle shared-memory round trip (the tensor core is
// streaming operands out of shared memory), which this way overlaps the issue of the block instead of preceding it.
//   operands: d accumulator, A (SS: descriptor words a_lo / a_hi, +256 per k-step), B descriptor words (b_lo + k * kunits, b_hi), idesc,
//   acc0 = accumulate flag of the first MMA, (next_full, next_par) = barrier probed.  Returns 1 when the next stage is already full.
__device__ __forceinline__ uint32_t mma_block_ss2(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t kunits,
                                                  uint32_t idesc, uint32_t acc0, uint32_t next_full, uint32_t next_par) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q, [%9], %10;\n\t"
      "setp.ne.u32 p, %8, 0;\n\t"
      "mov.b64 da, {%2, %3};\n\tmov.b64 db, {%4, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %7, p;\n\t"
      "setp.eq.u32 p, 1, 1;\n\t"
      "add.u32 ta, %2, 256;\n\tmad.lo.u32 tb, %6, 1, %4;\n\tmov.b64 da, {ta, %3};\n\tmov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %7, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}\n"
      : "=r"(ok)
      : "r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(kunits), "r"(idesc), "r"(acc0), "r"(next_full), "r"(next_par)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t mma_block_ss3(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t kunits,
                                                  uint32_t idesc, uint32_t acc0, uint32_t next_full, uint32_t next_par) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q, [%9], %10;\n\t"
      "setp.ne.u32 p, %8, 0;\n\t"
      "mov.b64 da, {%2, %3};\n\tmov.b64 db, {%4, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %7, p;\n\t"
      "setp.eq.u32 p, 1, 1;\n\t"
      "add.u32 ta, %2, 256;\n\tmad.lo.u32 tb, %6, 1, %4;\n\tmov.b64 da, {ta, %3};\n\tmov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %7, p;\n\t"
      "add.u32 ta, %2, 512;\n\tmad.lo.u32 tb, %6, 2, %4;\n\tmov.b64 da, {ta, %3};\n\tmov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %7, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}\n"
      : "=r"(ok)
      : "r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(kunits), "r"(idesc), "r"(acc0), "r"(next_full), "r"(next_par)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t mma_block_ss4(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t kunits,
                                                  uint32_t idesc, uint32_t acc0, uint32_t next_full, uint32_t next_par) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q, [%9], %10;\n\t"
      "setp.ne.u32 p, %8, 0;\n\t"
      "mov.b64 da, {%2, %3};\n\tmov.b64 db, {%4, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %7, p;\n\t"
      "setp.eq.u32 p, 1, 1;\n\t"
      "add.u32 ta, %2, 256;\n\tmad.lo.u32 tb, %6, 1, %4;\n\tmov.b64 da, {ta, %3};\n\tmov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %7, p;\n\t"
      "add.u32 ta, %2, 512;\n\tmad.lo.u32 tb, %6, 2, %4;\n\tmov.b64 da, {ta, %3};\n\tmov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %7, p;\n\t"
      "add.u32 ta, %2, 768;\n\tmad.lo.u32 tb, %6, 3, %4;\n\tmov.b64 da, {ta, %3};\n\tmov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %7, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}\n"
      : "=r"(ok)
      : "r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(kunits), "r"(idesc), "r"(acc0), "r"(next_full), "r"(next_par)
      : "memory");
  return ok;
}
// TS-mode variant (A = packed activation in TMEM at columns a0..a3)
__device__ __forceinline__ uint32_t mma_block_ts4(uint32_t d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t kunits, uint32_t idesc, uint32_t acc0, uint32_t next_full, uint32_t next_par) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q, [%9], %10;\n\t"
      "setp.ne.u32 p, %8, 0;\n\t"
      "mov.b64 db, {%4, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], [%2], db, %7, p;\n\t"
      "setp.eq.u32 p, 1, 1;\n\t"
      "mad.lo.u32 tb, %6, 1, %4;\n\tmov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], [%11], db, %7, p;\n\t"
      "mad.lo.u32 tb, %6, 2, %4;\n\tmov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], [%12], db, %7, p;\n\t"
      "mad.lo.u32 tb, %6, 3, %4;\n\tmov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], [%13], db, %7, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}\n"
      : "=r"(ok)
      : "r"(d), "r"(a0), "r"(0u), "r"(b_lo), "r"(b_hi), "r"(kunits), "r"(idesc), "r"(acc0), "r"(next_full), "r"(next_par), "r"(a1), "r"(a2), "r"(a3)
      : "memory");
  return ok;
}
// `nfull` weight blocks of KB k-steps each, starting at k-step ks
template <int KB, bool kTS>
__device__ __forceinline__ void issue_blocks(const StepParams& p, IssueCtx& c, uint32_t nfull, uint32_t& ks, uint32_t d, uint32_t a0, uint32_t kunits,
                                             uint32_t idesc) {
  constexpr uint32_t a_hi = 8u | (1u << 14), b_hi = 16u | (1u << 14);     // A: LBO 2048 B, SBO 128 B ; B: LBO 128 B, SBO 256 B
#pragma unroll 1
  for (uint32_t b = 0; b < nfull; ++b, ks += KB) {
    ring_wait_full(c, p.err_flag);
    const uint32_t b_lo = c.stage_a | (8u << 16);
    const int sn = c.s + 1 == c.stages ? 0 : c.s + 1;
    const uint32_t next_full = c.full0 + 8u * (uint32_t)sn, next_par = sn == 0 ? c.ring_par ^ 1u : c.ring_par;
    const uint32_t acc0 = ks != 0 ? 1u : 0u;
    uint32_t ok;
    if constexpr (kTS) ok = mma_block_ts4(d, a0 + p.plan.ts_col[ks], a0 + p.plan.ts_col[ks + 1], a0 + p.plan.ts_col[ks + 2], a0 + p.plan.ts_col[ks + 3], b_lo, b_hi,
                                          kunits, idesc, acc0, next_full, next_par);
    else if constexpr (KB == 2) ok = mma_block_ss2(d, a0 + 256u * ks, a_hi, b_lo, b_hi, kunits, idesc, acc0, next_full, next_par);
    else if constexpr (KB == 3) ok = mma_block_ss3(d, a0 + 256u * ks, a_hi, b_lo, b_hi, kunits, idesc, acc0, next_full, next_par);
    else ok = mma_block_ss4(d, a0 + 256u * ks, a_hi, b_lo, b_hi, kunits, idesc, acc0, next_full, next_par);
    ring_release_advance(c);
    c.look = ok;
  }
}

// kWide: states wider than 15 (O > 16).  A separate instantiation: the wide output path (global / local-memory parking of the means and
// draws) would otherwise cost the narrow variant instruction-cache space and a larger local frame (measured: 7 % of the step).
template <bool kDebug, bool kWide>
__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(NUM_THREADS, 1) rollout_step_pipe_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const PlanDev& plan = p.plan;
  const int S = p.S, A = p.A, O = S + 1;
  // shared memory: [weight ring | X slot 0 | X slot 1 | xp slot 0 | xp slot 1 | ones | control]
  uint8_t* ring = smem_raw;
  const uint32_t slot_bytes = plan.slot_bytes;
  uint8_t* Xb = ring + (size_t)slot_bytes * p.stages;
  uint8_t* xpb = Xb + (size_t)N_SLOTS * plan.x_bytes;
  uint8_t* ones = xpb + (size_t)N_SLOTS * plan.xp_bytes;
  SmemCtl* sm = reinterpret_cast<SmemCtl*>(ones + TILE_M * 16 * 2);
  int* err = p.err_flag;

  const int n = (int)min((int64_t)*p.n_dev, p.n_max);
  const int n_tiles = (n + TILE_M - 1) / TILE_M;
  // tile quads (2 CTAs x 2 slots) are dealt to the clusters round-robin; both CTAs of a pair walk the same number of iterations in
  // lock step (coupled by the weight ring); tiles past the end of the batch run without rows
  const uint32_t crank = cluster_ctarank();
  const int n_clusters = (int)gridDim.x / CLUSTER, cid = (int)blockIdx.x / CLUSTER;
  const int n_quads = (n_tiles + 3) / 4;
  const int my_iters = cid < n_quads ? (n_quads - cid + n_clusters - 1) / n_clusters : 0;
  auto tile_of = [&](int it, int slot) { return 4 * (cid + it * n_clusters) + 2 * (int)crank + slot; };

  if (threadIdx.x == 0) {
    // full: the leader's barrier also collects the peer's "my half has landed" relay; empty: one multicast commit of the leader
    for (int s = 0; s < p.stages; ++s) { mbar_init(&sm->full[s], crank == 0 ? 2 : 1); mbar_init(&sm->empty[s], 1); }
    for (int s = 0; s < N_SLOTS; ++s) {
      mbar_init(&sm->hid_full[s], 1);
      for (int k = 0; k < 3; ++k) mbar_init(&sm->out_full[s][k], 1);
      for (int k = 0; k < N_CNT; ++k) mbar_init(&sm->sig[s][k], CLUSTER * (k == C_HID ? 4 * N_GROUPS : 4));     // (used in the leader only)
    }
    fence_barrier_init();
  }
  {
    const int t0 = threadIdx.x;
    if (t0 < S) { sm->norm_mean[t0] = p.norm_mean[t0]; sm->norm_inv[t0] = 1.f / (p.norm_std[t0] + 1e-6f); }
    if (t0 <= S) {
      const float lo = p.min_lv[t0], hi = p.max_lv[t0];
      sm->lv_hi[t0] = hi; sm->lv_E[t0] = __expf(hi - lo); sm->lv_s0[t0] = __expf(0.5f * lo);
    }
    if (t0 < N_HID * N_GROUPS) sm->epi[t0] = plan.epi[t0];
    for (int i = t0; i < TILE_M * 16; i += NUM_THREADS) {       // ones tile: element (row, k) at (k/8)*2048 + row*16 + (k%8)*2, k = 0, 1 -> 1.0
      const int k = (i >> 10) * 8 + (i & 7);
      reinterpret_cast<__nv_bfloat16*>(ones)[i] = __float2bfloat16_rn(k < 2 ? 1.f : 0.f);
    }
  }
  if (warp == ISSUER_WARP) tmem_alloc2(&sm->tmem_base, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp == PRODUCER_WARP) {
    // ===================== TMA producer: every weight block of every job through the ring, half per CTA, multicast to the pair =========
    if (elect_one()) {
      uint32_t s = 0, ph = 0;
      for (int it = 0; it <= my_iters; ++it)
#pragma unroll 1
        for (int e = 0; e < 2 * N_JOBS; ++e) {
          const int j = plan.sched[e][1], tit = it - (int)plan.sched[e][2];
          if (tit < 0 || tit >= my_iters) continue;
          const JobDev& jb = plan.job[j];
          const uint8_t* img = (j < 3 ? p.policy_img : p.model_img) + jb.img_off;
          const uint32_t kbytes = jb.kstep_units << 4;
          const uint32_t n_reg = jb.n_ks - (jb.ones_ks != 0xFFFFu ? 1u : 0u);
          for (uint32_t ks = 0; ks < jb.n_ks;) {                     // blocks of kb regular k-steps, then the bias k-step on its own
            const uint32_t nk = ks < n_reg ? min(jb.kb, n_reg - ks) : 1u, hb = kbytes >> 1;
            mbar_wait(&sm->empty[s], ph ^ 1, err, 1);              // the pair's MMAs on the stage's previous contents are done
            mbar_expect_tx(&sm->full[s], nk * hb);
            // this CTA's half of the rows of each k-step tile (k-step tiles are [N x 16]: the halves are contiguous)
            for (uint32_t k = 0; k < nk; ++k)
              bulk_g2s(ring + (size_t)s * slot_bytes + k * hb, img + (size_t)(ks + k) * kbytes + crank * hb, hb, &sm->full[s]);
            if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
            ks += nk;
          }
        }
    }
  } else if (warp == ISSUER_WARP) {
    // ===================== MMA issuer: one thread alternates the two slots job by job ===================================================
    // A compact program: the job and slot loops are rolled (the whole issuer is a few hundred instructions), only the MMAs of one weight
    // block are issued from straight-line code.  TMEM base = 0 (the CTA owns all 512 columns), checked here.
    if (crank != 0) {
      // ---- peer CTA: its MMAs are issued by the leader; relay "my half of the block has landed" to the leader's full barrier ----
      if (elect_one()) {
        if (tmem != 0u && atomicCAS(err, 0, 88) == 0) printf("drpo_b200: unexpected TMEM base %u\n", tmem);
        uint32_t s = 0, ph = 0;
        for (int it = 0; it <= my_iters; ++it)
#pragma unroll 1
          for (int e = 0; e < 2 * N_JOBS; ++e) {
            const int j = plan.sched[e][1], tit = it - (int)plan.sched[e][2];
            if (tit < 0 || tit >= my_iters) continue;
            const JobDev& jb = plan.job[j];
            const uint32_t n_reg = jb.n_ks - (jb.ones_ks != 0xFFFFu ? 1u : 0u);
            for (uint32_t ks = 0; ks < jb.n_ks;) {
              const uint32_t nk = ks < n_reg ? min(jb.kb, n_reg - ks) : 1u;
              mbar_wait(&sm->full[s], ph, err, 2);
              arrive_leader(&sm->full[s], 1u);
              if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
              ks += nk;
            }
          }
      }
    } else if (elect_one()) {
      if (tmem != 0u && atomicCAS(err, 0, 88) == 0) printf("drpo_b200: unexpected TMEM base %u\n", tmem);
      IssueCtx c;
      c.ring_a = (smem_u32(ring) >> 4) & 0x3FFFu; c.slot_a = slot_bytes >> 4;
      c.full0 = smem_u32(&sm->full[0]); c.empty0 = smem_u32(&sm->empty[0]);
      c.s = 0; c.ring_par = 0; c.stage_a = c.ring_a; c.stages = p.stages; c.look = 0;
      const uint32_t x_a = (smem_u32(Xb) >> 4) & 0x3FFFu, x_step = plan.x_bytes >> 4;
      const uint32_t xp_a = (smem_u32(xpb) >> 4) & 0x3FFFu, xp_step = plan.xp_bytes >> 4;
      const uint32_t ones_a = ((smem_u32(ones) >> 4) & 0x3FFFu) | (128u << 16);
      const uint32_t hid0 = smem_u32(&sm->hid_full[0]), out0 = smem_u32(&sm->out_full[0][0]);
      constexpr uint32_t a_hi = 8u | (1u << 14), b_hi = 16u | (1u << 14);
      for (int it = 0; it <= my_iters; ++it) {
#pragma unroll 1
        for (int e = 0; e < 2 * N_JOBS; ++e) {
          const uint32_t slot = plan.sched[e][0];
          const int j = plan.sched[e][1], tit = it - (int)plan.sched[e][2];
          if (tit < 0 || tit >= my_iters) continue;
          const uint32_t t = (uint32_t)tit;
          const JobDev& jb = plan.job[j];
          const uint32_t idesc = jb.idesc, kunits = jb.kstep_units >> 1 /* this CTA's half of a k-step tile */, kb = jb.kb, wait0 = jb.wait0, wait1 = jb.wait1;
          const uint32_t n_reg = jb.n_ks - (jb.ones_ks != 0xFFFFu ? 1u : 0u);
          const uint32_t nfull = jb.nfull;
          const bool ts = jb.a_kind == A_TMEM;
          {
            {
              const uint32_t want0 = ((wait0 >> 8) & 0xFFu) + (wait0 >> 16) * t, want1 = ((wait1 >> 8) & 0xFFu) + (wait1 >> 16) * t;
              if (wait0 && want0) mbar_wait_cluster(&sm->sig[slot][wait0 & 0xFFu], (want0 - 1u) & 1u, err, 20 + (int)(wait0 & 0xFFu));
              if (wait1 && want1) mbar_wait_cluster(&sm->sig[slot][wait1 & 0xFFu], (want1 - 1u) & 1u, err, 20 + (int)(wait1 & 0xFFu));
            }
            stamp<kDebug>(p, t, (int)slot * 32 + j, 0);
            const uint32_t d = slot * SLOT_COLS + jb.d_col;
            const uint32_t a0 = ts ? slot * SLOT_COLS + jb.a_off
                                   : (((jb.a_kind == A_XP ? xp_a + slot * xp_step : x_a + slot * x_step) + jb.a_off) | (128u << 16));
            uint32_t ks = 0;
            if (ts) issue_blocks<4, true>(p, c, kb == 4 ? nfull : 0u, ks, d, a0, kunits, idesc);
            else if (kb == 2) issue_blocks<2, false>(p, c, nfull, ks, d, a0, kunits, idesc);
            else if (kb == 3) issue_blocks<3, false>(p, c, nfull, ks, d, a0, kunits, idesc);
            else if (kb == 4) issue_blocks<4, false>(p, c, nfull, ks, d, a0, kunits, idesc);
            // what is left, in the producer's block sizes ([kb] * nfull + [rem]): the remainder block, or every block when the block
            // size has no static variant above
            if (ks < n_reg) {
              uint32_t left = n_reg - ks;
              while (left) {
                const uint32_t nk = min(kb, left);
                ring_wait_full(c, err);
                const uint32_t b_lo = c.stage_a | (8u << 16);
                for (uint32_t k = 0; k < nk; ++k) {
                  const uint32_t acc = (ks + k) != 0 ? 1u : 0u;
                  if (ts) mma2_ts_p(d, a0 + plan.ts_col[ks + k], b_lo + k * kunits, b_hi, idesc, acc);
                  else mma2_ss_p(d, a0 + 256u * (ks + k), a_hi, b_lo + k * kunits, b_hi, idesc, acc);
                }
                ring_release_advance(c);
                ks += nk; left -= nk;
              }
            }
            if (jb.ones_ks != 0xFFFFu) {                                     // bias k-step: constant tile of ones x (hi, lo) bias block
              ring_wait_full(c, err);
              mma2_ss_p(d, ones_a, a_hi, c.stage_a | (8u << 16), b_hi, idesc, n_reg != 0 ? 1u : 0u);
              ring_release_advance(c);
            }
            tc_commit2_addr(jb.bar == BAR_HID ? hid0 + 8u * slot : out0 + 8u * (3u * slot + (jb.bar - BAR_OUT0)));
            stamp<kDebug>(p, t, (int)slot * 32 + j, 2);
          }
        }
      }
    }
  } else if (warp < OUT_WARP0) {
    // ===================== hidden-layer epilogues: group g = warp / 4 drains column chunk g of every hidden layer of BOTH slots, in the
    // issuer's order (layer i of slot 0, layer i of slot 1, layer i+1 of slot 0, ..): in the steady state the two slots' epilogues
    // alternate in time, so each one gets all twelve warps =====================
    const int g = warp >> 2, q = warp & 3;
    for (int itx = 0; itx <= my_iters; ++itx) {
#pragma unroll 1
      for (int he = 0; he < 2 * N_HID; ++he) {
        {
          const int slot = plan.hsched[he][0], i = plan.hsched[he][1], it = itx - (int)plan.hsched[he][2];
          if (it < 0 || it >= my_iters) continue;
          const uint32_t L = (uint32_t)(N_HID * it + i);
          const EpiRec e = sm->epi[i * N_GROUPS + g];
          const uint32_t region = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * SLOT_COLS;
          const uint32_t xrow = smem_u32(Xb + (size_t)slot * plan.x_bytes) + (uint32_t)(q * 32 + lane) * 16u;
          const bool lead = kDebug && warp == 0 && lane == 0;
          if (lead) stamp<kDebug>(p, (uint32_t)it, slot * 32 + 10 + i, 4);
          mbar_wait(&sm->hid_full[slot], L & 1u, err, 4);
          tc_fence_after();
          if (lead) stamp<kDebug>(p, (uint32_t)it, slot * 32 + 10 + i, 5);
          if (kDebug && p.dump_layer == (int)e.layer) {          // debug hook: raw accumulator to global
            const int64_t row0 = (int64_t)tile_of(it, slot) * TILE_M; const int tr = q * 32 + lane;
            for (int c0 = 0; c0 < (int)e.nc; c0 += 16) {
              uint32_t r[16]; tmem_ld16(region + e.n0 + (uint32_t)c0, r); tmem_ld_wait();
              if (row0 + tr < n) for (int jj = 0; jj < 16; ++jj) if (e.n0 + c0 + jj < e.n_real) p.dump_out[(row0 + tr) * e.n_real + e.n0 + c0 + jj] = __uint_as_float(r[jj]);
            }
          }
          const uint32_t src = region + e.n0, tdst = region + e.out_col;
          if (e.to_smem) {
            if (e.silu) hidden_chunk<true, true>(src, e.nc, e.n0, e.one, xrow, tdst);
            else hidden_chunk<false, true>(src, e.nc, e.n0, e.one, xrow, tdst);
            if (lead) stamp<kDebug>(p, (uint32_t)it, slot * 32 + 10 + i, 7);
            fence_proxy_async();                                 // generic-proxy writes of X -> visible to the tensor core
          } else {
            hidden_chunk<true, false>(src, e.nc, e.n0, e.one, xrow, tdst);
            tmem_st_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) sig_arrive(&sm->sig[slot][C_HID], crank);                  // activation visible to the MMAs, accumulator drained
          if (lead) stamp<kDebug>(p, (uint32_t)it, slot * 32 + 10 + i, 6);
        }
      }
    }
  } else if (warp < PRODUCER_WARP) {
    // ===================== output groups: policy head, the two output heads, stores, the slot's next prologue ==========================
    const int slot = (warp - OUT_WARP0) >> 2, q = warp & 3;
    const int t = q * 32 + lane;                                          // trajectory row of the tile == TMEM lane
    const uint32_t region = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * SLOT_COLS;
    const uint32_t xrow = smem_u32(Xb + (size_t)slot * plan.x_bytes) + (uint32_t)t * 16u;
    const uint32_t xprow = smem_u32(xpb + (size_t)slot * plan.xp_bytes) + (uint32_t)t * 16u;
    const int No = plan.No;
    const int sidx = slot * 32 + 20;
    float4 np4 = make_float4(0.f, 0.f, 0.f, 0.f);                          // policy draws of the staged tile
    uint32_t xn[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};                     // narrow states: the staged tile's normalised state, packed bf16 pairs
    constexpr bool narrow = !kWide;                                        // S <= 15
    // wide states: the raw state row of the tile in flight and of the staged next tile (filled by the prologue, read again by the
    // member-input and diff-head passes) in thread-local memory
    float srow[kWide ? 2 : 1][kWide ? 64 : 1];
    uint32_t nrow[kWide ? 2 : 1][kWide ? 32 : 1];                        // the same rows normalised, packed bf16 pairs (the member input's state part)

    auto prologue = [&](int tile, uint32_t it) {
      // ---- policy input [s, 1] -> the slot's xp tile; torch.normal of policy.act keyed by the row's global trajectory id ----
      const int64_t row = (int64_t)tile * TILE_M + t;
      const bool valid = row < n;
      if (valid && p.ready_flags) wait_rows_ready(p.ready_flags, p.ready_shift, (int)row, (int)row, err);
      const float* ps = p.cur + row * S;
      const int xp_one = plan.xp_one;
      if constexpr (narrow) {                                             // one pass over the row: raw -> xp, normalised -> registers (the
        float sv[16];                                                     // member input is assembled from them when the action arrives)
#pragma unroll
        for (int k = 0; k < 16; ++k) sv[k] = (k < S && valid) ? __ldg(ps + k) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k0 = 2 * j, k1 = 2 * j + 1;
          xn[j] = pack_bf16(k0 < S ? (sv[k0] - sm->norm_mean[k0]) * sm->norm_inv[k0] : 0.f, k1 < S ? (sv[k1] - sm->norm_mean[k1]) * sm->norm_inv[k1] : 0.f);
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(2 * j < S ? sv[2 * j] : (2 * j == xp_one ? 1.f : 0.f), 2 * j + 1 < S ? sv[2 * j + 1] : (2 * j + 1 == xp_one ? 1.f : 0.f));
        st_shared_v4(xprow, pk[0], pk[1], pk[2], pk[3]);
        st_shared_v4(xprow + 2048u, pk[4], pk[5], pk[6], pk[7]);          // K0p = 16
      } else {
        write_input_octets_wide(xprow, plan.K0p >> 3, ps, S, valid, (S & 3) == 0,
                                [&](int k, float x) { return k < S ? x : (k == xp_one ? 1.f : 0.f); }, srow[it & 1u], nrow[it & 1u],
                                [&](int k, float x) { return (k < S && valid) ? (x - sm->norm_mean[k]) * sm->norm_inv[k] : 0.f; });
      }
      fence_proxy_async();                                                // generic-proxy writes of xp -> visible to the tensor core
      __syncwarp();
      if (lane == 0) sig_arrive(&sm->sig[slot][C_TILE], crank);
      np4 = valid ? noise_get4(p.noise_p, (int64_t)p.ids[row], 0, A) : make_float4(0.f, 0.f, 0.f, 0.f);
    };

    if (my_iters > 0) prologue(tile_of(0, slot), 0u);
    for (int it = 0; it < my_iters; ++it) {
      const int tile = tile_of(it, slot);
      const int64_t row = (int64_t)tile * TILE_M + t;
      const bool valid = row < n;
      const float* my_s = p.cur + row * S;
      const uint32_t par = (uint32_t)it & 1u;
      // ---- policy head: [mu, raw] -> a = tanh(mu + exp(-6 + 10 sigmoid(raw)) eps)      src/policy.py:89-97 ----
      mbar_wait(&sm->out_full[slot][0], par, err, 7);
      tc_fence_after();
      if (t == 0) stamp<kDebug>(p, (uint32_t)it, sidx, 0);
      {
        uint32_t r[16]; tmem_ld16(region + plan.head_col, r); tmem_ld_wait();
        if (kDebug && p.dump_layer == 2 && valid) for (int jj = 0; jj < 2 * A; ++jj) p.dump_out[row * 2 * A + jj] = __uint_as_float(r[jj]);
        const float ev[4] = {np4.x, np4.y, np4.z, np4.w};
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
        // member input x0 = [(s - mean)/(std + 1e-6), a, 1]  (src/dynamics.py:113-114) -> the first octets of the slot's X (the policy's
        // activations in it are dead: the head's MMAs have completed)
        const int xm_one = plan.xm_one;
        auto tail = [&](int k) {                                           // elements behind the state: actions, the bias slot, zeros
          const int ja = k - S;
          const float av = ja == 0 ? act4[0] : (ja == 1 ? act4[1] : (ja == 2 ? act4[2] : (ja == 3 ? act4[3] : 0.f)));     // 0 beyond the A real actions
          return k == xm_one ? 1.f : av;
        };
        if constexpr (narrow) {                                            // the state part comes from registers (staged by the prologue)
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k0 = 2 * j, k1 = 2 * j + 1;
            const float lo = k0 < S ? __uint_as_float(xn[j] << 16) : tail(k0), hi = k1 < S ? __uint_as_float(xn[j] & 0xFFFF0000u) : tail(k1);
            pk[j] = pack_bf16(lo, hi);
          }
          st_shared_v4(xrow, pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(xrow + 2048u, pk[4], pk[5], pk[6], pk[7]);
          write_input_octets(xrow, 2, plan.K0m >> 3, tail);                // (K0m = 32 when S + A + 1 > 16)
        } else {
          write_input_octets_packed(xrow, plan.K0m >> 3, nrow[it & 1], S, tail);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) sig_arrive(&sm->sig[slot][C_XM], crank);                    // also: the head accumulator has been read
        if (t == 0) stamp<kDebug>(p, (uint32_t)it, sidx, 1);
        if (valid) {                                                     // off the critical path: the actions go to global memory
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) if (jj < A) p.actions[row * A + jj] = act4[jj];
        }
      }
      // ---- off the critical path: the slot's next prologue ----
      if (it + 1 < my_iters) prologue(tile_of(it + 1, slot), (uint32_t)it + 1u);
      if (t == 0) stamp<kDebug>(p, (uint32_t)it, sidx, 2);
      if constexpr (!kWide) {
        // ---- diff head: means = diffs + [s, 0]                                           src/dynamics.py:118 ----
        float mean[16], ev[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) mean[jj] = (jj < S && valid) ? __ldg(my_s + jj) : 0.f;      // in flight while the diff head is computed
        mbar_wait(&sm->out_full[slot][1], par, err, 7);
        tc_fence_after();
        if (t == 0) stamp<kDebug>(p, (uint32_t)it, sidx, 3);
        {
          uint32_t r[16]; tmem_ld16(region + plan.d1_col, r);
          tmem_ld_wait();
          if (kDebug && p.dump_layer == 7 && valid) for (int jj = 0; jj < 16; ++jj) if (jj < O) p.dump_out[row * O + jj] = __uint_as_float(r[jj]);
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) mean[jj] += __uint_as_float(r[jj]);
        }
        // randn_like of ensemble.sample (src/dynamics.py:202), keyed by the row's global trajectory id
        {
          const int64_t pid = valid ? (int64_t)p.ids[row] : 0;
#pragma unroll
          for (int cg = 0; cg < 4; ++cg) {
            float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * cg < O && valid) e4 = noise_get4(p.noise_m, pid, cg, O);
            ev[4 * cg] = e4.x; ev[4 * cg + 1] = e4.y; ev[4 * cg + 2] = e4.z; ev[4 * cg + 3] = e4.w;
          }
        }
        // ---- log-var head + Gaussian sample                                              src/dynamics.py:119-121,201-203 ----
        mbar_wait(&sm->out_full[slot][2], par, err, 7);
        tc_fence_after();
        if (t == 0) stamp<kDebug>(p, (uint32_t)it, sidx, 4);
        {
          uint32_t r[16]; tmem_ld16(region + plan.v1_col, r); tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) sig_arrive(&sm->sig[slot][C_OUT], crank);                   // both head accumulators are in registers
          if (kDebug && p.dump_layer == 8 && valid) for (int jj = 0; jj < 16; ++jj) if (jj < O) p.dump_out[row * O + jj] = __uint_as_float(r[jj]);
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const int cc = min(jj, O - 1);
            const float u = __expf(sm->lv_hi[cc] - __uint_as_float(r[jj]));
            mean[jj] = fmaf(sm->lv_s0[cc] * sqrt_fast(1.f + __fdividef(sm->lv_E[cc], 1.f + u)), ev[jj], mean[jj]);
          }
        }
        if (valid) {
          float* dst = p.next_states + row * S;
          if ((S & 3) == 0) {
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) if (4 * c4 < S) reinterpret_cast<float4*>(dst)[c4] = make_float4(mean[4 * c4], mean[4 * c4 + 1], mean[4 * c4 + 2], mean[4 * c4 + 3]);
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) if (jj < S) dst[jj] = mean[jj];
          }
          float rw = mean[15];                                             // mean[S] without a run-time register index
#pragma unroll
          for (int jj = 14; jj >= 0; --jj) rw = jj >= S ? mean[jj] : rw;
          p.rewards[row] = rw;
        }
      } else {
        // wide states (O > 16): the diff head's accumulator sits in columns the log-var hidden layer overwrites, so the means are
        // parked until the log-var head arrives; the Gaussian draws are generated while the log-var layers run.  Both are parked in
        // thread-local memory (lane-interleaved: every access is one coalesced wavefront; a thread's own 240-byte row in global
        // memory costs 32 sectors per warp instruction) and the state row comes from the copy the prologue kept: the only
        // global-memory traffic of the tile's output is ONE pass of 16-byte stores.
        float evw[64], mvw[64];
        const bool vec = (S & 3) == 0;                                     // 16-byte aligned rows: one vector store per four columns
        float* my_ns = p.next_states + row * S;
        const float* my_row = srow[it & 1];
        mbar_wait(&sm->out_full[slot][1], par, err, 7);
        tc_fence_after();
        if (t == 0) stamp<kDebug>(p, (uint32_t)it, sidx, 3);
        for (int c0 = 0; c0 < No; c0 += 16) {
          uint32_t r[16]; tmem_ld16(region + plan.d1_col + (uint32_t)c0, r);
          tmem_ld_wait();
          if (kDebug && p.dump_layer == 7 && valid) for (int jj = 0; jj < 16; ++jj) if (c0 + jj < O) p.dump_out[row * O + c0 + jj] = __uint_as_float(r[jj]);
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) mvw[c0 + jj] = __uint_as_float(r[jj]);       // (the state is added in the last pass: this one is on the
        }                                                                            // critical path of the log-var hidden layer)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) sig_arrive(&sm->sig[slot][C_D1R], crank);                     // the log-var hidden layer may overwrite the diff head
        {
          const int64_t pid = valid ? (int64_t)p.ids[row] : 0;
          for (int cg = 0; 4 * cg < O; ++cg) {
            const float4 e4 = valid ? noise_get4(p.noise_m, pid, cg, O) : make_float4(0.f, 0.f, 0.f, 0.f);
            evw[4 * cg] = e4.x; evw[4 * cg + 1] = e4.y; evw[4 * cg + 2] = e4.z; evw[4 * cg + 3] = e4.w;
          }
        }
        mbar_wait(&sm->out_full[slot][2], par, err, 7);
        tc_fence_after();
        if (t == 0) stamp<kDebug>(p, (uint32_t)it, sidx, 4);
        // first pass: drain the log-var head's accumulator (std * eps replaces eps in the local array) so that the slot's next tile
        // may start (C_OUT) before the sums and the global stores of the second pass
        for (int c0 = 0; c0 < No; c0 += 16) {
          uint32_t r[16]; tmem_ld16(region + plan.v1_col + (uint32_t)c0, r);
          float ev[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) ev[jj] = c0 + jj < O ? evw[c0 + jj] : 0.f;
          tmem_ld_wait();
          if (c0 + 16 >= No) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) sig_arrive(&sm->sig[slot][C_OUT], crank);
          }
          if (kDebug && p.dump_layer == 8 && valid) for (int jj = 0; jj < 16; ++jj) if (c0 + jj < O) p.dump_out[row * O + c0 + jj] = __uint_as_float(r[jj]);
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const int c = min(c0 + jj, O - 1);
            const float u = __expf(sm->lv_hi[c] - __uint_as_float(r[jj]));
            evw[c0 + jj] = sm->lv_s0[c] * sqrt_fast(1.f + __fdividef(sm->lv_E[c], 1.f + u)) * ev[jj];
          }
        }
        if (valid) {
          for (int c0 = 0; c0 < No; c0 += 16) {
            float o16[16];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)                               // means = diffs + [s, 0] (src/dynamics.py:118), + std * eps (:201-203)
              o16[jj] = (mvw[c0 + jj] + (c0 + jj < S ? my_row[c0 + jj] : 0.f)) + evw[c0 + jj];
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const float v[4] = {o16[4 * g4], o16[4 * g4 + 1], o16[4 * g4 + 2], o16[4 * g4 + 3]};
#pragma unroll
              for (int e = 0; e < 4; ++e) if (c0 + 4 * g4 + e == S) p.rewards[row] = v[e];
              store_row4(my_ns, c0 + 4 * g4, S, vec, v);
            }
          }
        }
      }
      if (t == 0) stamp<kDebug>(p, (uint32_t)it, sidx, 5);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // no CTA leaves while its peer may still multicast into it or signal its barriers
  if (warp == ISSUER_WARP) tmem_dealloc2(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// host side: the static plan (TMEM map, jobs with their waits, epilogue records)
// ---------------------------------------------------------------------------------------------------------------
static inline uint32_t wait_code(int id, int c0, int inc) { return (uint32_t)(id | (c0 << 8) | (inc << 16)); }

static int build_plan(const drpo_rollout_args& a, Plan& P, int max_smem, int& stages_out, int& smem_out) {
  memset(&P, 0, sizeof(P));
  PlanDev& D = P.d;
  const int S = a.ensemble->state_dim, A = a.ensemble->action_dim, Hm = a.ensemble->hidden, Hp = a.actor->l0.out_dim, O = S + 1;
  if (a.actor->l1.out_dim != Hp || a.actor->l1.in_dim != Hp || a.actor->l2.out_dim != 2 * A || Hp < 16 || Hm < 16 || A > 4 || A < 1 || O > 64 ||
      S + A + 1 > 64) {
    set_error("bf16 rollout: dims outside the fused kernel's plan (S=%d A=%d actor hidden=%d model hidden=%d)", S, A, Hp, Hm);
    return DRPO_ERR_UNSUPPORTED;
  }
  // a layer whose K is not a multiple of 16 carries its bias in the K slot behind the last real input (the activation holds a
  // constant 1 there); otherwise the bias is one extra k-step against the constant ones tile
  auto slot_bias = [](int k_real) { return k_real % 16 != 0; };
  const int Hpp = round_up(Hp + (slot_bias(Hp) ? 1 : 0), 16), Hmp = round_up(Hm + (slot_bias(Hm) ? 1 : 0), 16);
  const int K0p = round_up(S + (slot_bias(S) ? 1 : 0), 16), K0m = round_up(S + A + (slot_bias(S + A) ? 1 : 0), 16), No = round_up(O, 16);
  if (Hpp > SLOT_COLS || Hmp > SLOT_COLS) {
    set_error("bf16 rollout: hidden widths beyond one TMEM slot (actor %d, model %d)", Hp, Hm);
    return DRPO_ERR_UNSUPPORTED;
  }
  D.K0p = K0p; D.K0m = K0m; D.No = No;
  D.xp_one = slot_bias(S) ? S : -1; D.xm_one = slot_bias(S + A) ? S + A : -1;
  // ---- TMEM map of a slot (256 columns) ------------------------------------------------------------------------------------------
  // hidden accumulators at column 0; policy head over the drained L1 accumulator; diff hidden activation packed in place at the start
  // of each epilogue group's column chunk.
  // narrow outputs: diff head at Hmp, log-var head at Hmp + No (both survive until the output group reads them);
  // wide outputs  : diff head behind the last packed chunk (overwritten by the log-var hidden layer, which therefore waits for the
  //                 output group), log-var head at 0
  const int Nd1 = std::max(No, 32);                       // the diff head is a TS-mode pair MMA: N must be a multiple of 32
  D.wide = (No > 16 || Hmp + Nd1 + No > SLOT_COLS) ? 1 : 0;
  D.head_col = 0;
  if (!D.wide) { D.d1_col = Hmp; D.v1_col = Hmp + Nd1; }
  else {
    {   // behind the packed activation of the last column chunk (chunk starts as in the epilogue records below)
      const int units = Hmp / 16; int n0 = 0, last_n0 = 0, last_nc = 0;
      for (int g = 0; g < N_GROUPS; ++g) { const int u = units / N_GROUPS + (g < units % N_GROUPS ? 1 : 0); if (u) { last_n0 = n0; last_nc = 16 * u; } n0 += 16 * u; }
      D.d1_col = round_up(last_n0 + last_nc / 2, 16);
    }
    D.v1_col = 0;
    if (D.d1_col + Nd1 > SLOT_COLS) { set_error("bf16 rollout: TMEM plan does not fit (model hidden %d, state dim %d)", Hm, S); return DRPO_ERR_UNSUPPORTED; }
  }
  const int n_real[N_LAYERS] = {Hp, Hp, 2 * A, Hm, Hm, Hm, Hm, O, O};
  const int k_real[N_LAYERS] = {S, Hp, Hp, S + A, Hm, Hm, Hm, Hm, Hm};
  const int n_pad[N_LAYERS] = {Hpp, Hpp, 16, Hmp, Hmp, Hmp, Hmp, Nd1, No};
  const int k_pad[N_LAYERS] = {K0p, Hpp, Hpp, K0m, Hmp, Hmp, Hmp, Hmp, Hmp};
  const int job_layer[N_JOBS] = {0, 1, 2, 3, 4, 5, 7, 6, 8};
  D.x_bytes = (uint32_t)(TILE_M * std::max(std::max(Hpp, Hmp), K0m) * 2);
  D.xp_bytes = (uint32_t)(TILE_M * K0p * 2);
  // ---- ring geometry: blocks of kb k-steps, kb per job = what fits a ring slot (<= 4) ----
  const int fixed = (int)(N_SLOTS * (D.x_bytes + D.xp_bytes) + TILE_M * 16 * 2 + sizeof(SmemCtl) + 64);
  max_smem -= 1024;                                                        // the kernel's static shared memory (alignment of the dynamic block)
  uint32_t slot_bytes = 0; int stages = 0;
  for (uint32_t cand : {16384u, 12288u, 8192u}) {
    const int st = std::min<int>(MAX_STAGES, (max_smem - fixed) / (int)cand);
    slot_bytes = cand; stages = st;
    if (st >= 4) break;
  }
  if (stages < 2 || slot_bytes < (uint32_t)std::max(Hpp, Hmp) * 16u) {
    set_error("bf16 rollout: needs more shared memory than the device offers (%d B fixed, %d B available)", fixed, max_smem);
    return DRPO_ERR_UNSUPPORTED;
  }
  D.slot_bytes = slot_bytes; stages_out = stages; smem_out = fixed + stages * (int)slot_bytes;
  uint32_t off_pol = 0, off_mem = 0;
  for (int j = 0; j < N_JOBS; ++j) {
    const int l = job_layer[j];
    JobDev& J = D.job[j];
    const bool ones = !slot_bias(k_real[l]);
    const int nks = k_pad[l] / 16 + (ones ? 1 : 0);
    P.n_pad[l] = n_pad[l]; P.n_real[l] = n_real[l]; P.k_real[l] = k_real[l]; P.n_ks[l] = nks; P.ones_ks[l] = ones ? nks - 1 : 0xFFFF;
    P.slot[l] = ones ? 0 : 1; P.job_of[l] = j;
    J.idesc = (make_idesc(n_pad[l]) & ~(0x1Fu << 24)) | ((256u >> 4) << 24);        // M = 256: both CTAs' 128-row tiles
    J.n_ks = (uint32_t)nks; J.ones_ks = ones ? (uint32_t)(nks - 1) : 0xFFFFu;
    J.kstep_units = (uint32_t)(n_pad[l] * 32) >> 4;                                  // whole k-step tile (both halves) / 16
    J.kb = std::max(1u, std::min(4u, slot_bytes / (uint32_t)(n_pad[l] * 16)));       // each CTA holds half the rows of a block
    J.nfull = (uint32_t)(nks - (ones ? 1 : 0)) / J.kb;
    uint32_t& off = j < 3 ? off_pol : off_mem;
    J.img_off = off; off += (uint32_t)nks * (uint32_t)n_pad[l] * 32u;
  }
  P.pol_bytes = off_pol; P.mem_bytes = off_mem;
  // ---- per job: accumulator, A operand, waits, completion barrier -----------------------------------------------------------------
  auto set = [&](int j, int d_col, int a_kind, int a_off, uint32_t w0, uint32_t w1, int bar) {
    JobDev& J = D.job[j]; J.d_col = (uint32_t)d_col; J.a_kind = (uint32_t)a_kind; J.a_off = (uint32_t)a_off; J.wait0 = w0; J.wait1 = w1; J.bar = (uint32_t)bar;
  };
  set(0, 0, A_XP, 0, wait_code(C_TILE, 1, 1), wait_code(C_OUT, 0, 1), BAR_HID);         // P0: the slot's input staged, previous tile's heads read
  set(1, 0, A_X, 0, wait_code(C_HID, 1, N_HID), 0, BAR_HID);                            // P1
  set(2, D.head_col, A_X, 0, wait_code(C_HID, 2, N_HID), 0, BAR_OUT0 + 0);              // P2 (head)
  set(3, 0, A_X, 0, wait_code(C_XM, 1, 1), 0, BAR_HID);                                 // T0: member input written (head accumulator read)
  set(4, 0, A_X, 0, wait_code(C_HID, 3, N_HID), 0, BAR_HID);                            // T1
  set(5, 0, A_X, 0, wait_code(C_HID, 4, N_HID), 0, BAR_HID);                            // D0
  set(6, D.d1_col, A_TMEM, 0, wait_code(C_HID, 5, N_HID), 0, BAR_OUT0 + 1);             // D1: TS-mode from the in-place packed activation
  set(7, 0, A_X, 0, D.wide ? wait_code(C_D1R, 1, 1) : 0u, 0, BAR_HID);                  // V0: X still holds the trunk output
  set(8, D.v1_col, A_X, 0, wait_code(C_HID, 6, N_HID), 0, BAR_OUT0 + 2);                // V1
  // ---- the static schedule: the two slots alternate job by job, except that a slot's diff head is directly followed by its log-var
  // hidden layer (which needs no epilogue in between), so the epilogue warps are not left idle while two narrow heads are issued back
  // to back.  (A half-tile offset between the slots - lag 1 entries - was measured slower: with a static order a job that waits for
  // its slot's epilogue blocks the other slot's ready job behind it.)
  {
    const int order[2 * N_JOBS][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0}, {0, 2, 0}, {1, 2, 0}, {0, 3, 0}, {1, 3, 0}, {0, 4, 0},
                                      {1, 4, 0}, {0, 5, 0}, {1, 5, 0}, {0, 6, 0}, {0, 7, 0}, {1, 6, 0}, {1, 7, 0}, {0, 8, 0}, {1, 8, 0}};
    const int hid_of_job[N_JOBS] = {0, 1, -1, 2, 3, 4, -1, 5, -1};
    int nh = 0;
    for (int e = 0; e < 2 * N_JOBS; ++e) {
      D.sched[e][0] = (uint8_t)order[e][0]; D.sched[e][1] = (uint8_t)order[e][1]; D.sched[e][2] = (uint8_t)order[e][2]; D.sched[e][3] = 0;
      const int h = hid_of_job[order[e][1]];
      if (h >= 0) { D.hsched[nh][0] = (uint8_t)order[e][0]; D.hsched[nh][1] = (uint8_t)h; D.hsched[nh][2] = (uint8_t)order[e][2]; D.hsched[nh][3] = 0; ++nh; }
    }
  }
  // ---- epilogue records of the six hidden layers (job order) x column groups: chunks in units of 16 columns (256 -> 96,80,80; 208 -> 80,64,64)
  const int hid_layer[N_HID] = {0, 1, 3, 4, 5, 6};
  for (int i = 0; i < N_HID; ++i) {
    const int l = hid_layer[i];
    const int units = n_pad[l] / 16; int n0 = 0;
    for (int g = 0; g < N_GROUPS; ++g) {
      const int u = units / N_GROUPS + (g < units % N_GROUPS ? 1 : 0);
      EpiRec& e = D.epi[i * N_GROUPS + g];
      e.n0 = (uint16_t)n0; e.nc = (uint16_t)(16 * u); e.out_col = (uint16_t)n0; e.n_real = (uint16_t)n_real[l];
      e.one = (int16_t)(slot_bias(n_real[l]) ? n_real[l] : -1);          // the consumer's bias slot = feature index `width` when K % 16 != 0
      e.silu = l >= 3; e.to_smem = l != 5; e.layer = (uint8_t)l;
      if (l == 5)                                                        // diff hidden: packed in place at the start of each chunk
        for (int c = 0; c < 16 * u; c += 16) D.ts_col[(n0 + c) / 16] = (uint16_t)(n0 + c / 2);
      n0 += 16 * u;
    }
  }
  return DRPO_OK;
}

// pack the actor (layers 0-2) or one member (layers 3-8) into its image
static void pack_jobs(const Plan& P, const drpo_linear* lin /* indexed by layer - base */, int base, int count, uint8_t* img, std::vector<PackJob>& jobs) {
  for (int l = base; l < base + count; ++l) {
    const drpo_linear& L = lin[l - base];
    PackJob j;
    j.W = L.w; j.b = L.b; j.n_real = P.n_real[l]; j.k_real = P.k_real[l]; j.n_pad = P.n_pad[l]; j.n_ks = P.n_ks[l]; j.ones_ks = P.ones_ks[l]; j.slot = P.slot[l];
    j.dst = reinterpret_cast<__nv_bfloat16*>(img + P.d.job[P.job_of[l]].img_off);
    jobs.push_back(j);
  }
}
static int pack_flush(std::vector<PackJob>& jobs, void* stream) {
  for (size_t i0 = 0; i0 < jobs.size(); i0 += PACK_JOBS) {
    PackTable t; memset(&t, 0, sizeof(t));
    const int nj = (int)std::min<size_t>(PACK_JOBS, jobs.size() - i0);
    for (int i = 0; i < nj; ++i) t.job[i] = jobs[i0 + i];
    dim3 grid(16, nj);
    DRPO_LAUNCH(pack_steps_kernel, grid, 256, 0, stream, t);
  }
  jobs.clear();
  return DRPO_OK;
}

}  // namespace r3
}  // namespace drpo
