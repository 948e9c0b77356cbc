// Shared machinery of the fused tcgen05 SSAC update kernels (critic_umma.cu: critic step; solver_umma.cu: multiplier and
// actor steps): op tables, weight-image packing, the generic forward / backward / head epilogues of a 128-row tile, the weight
// producer and the MMA issuer, the split-K dW kernel over the octet layout and the gradient-assembly kernel.
#pragma once
#include <cuda_bf16.h>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "critic.cuh"
#include "tc05.cuh"

namespace drpo {
namespace cu {

using namespace tc;

constexpr int HID = 256;                   // hidden width of every SSAC net (src/ssac.py:18-21)
constexpr int TILE = 128;                  // batch rows per tile = TMEM lanes
constexpr int NGROUPS = 4;                 // epilogue groups = 64-column slabs = accumulators
constexpr int EPI_THREADS = NGROUPS * 128;
constexpr int F_THREADS = EPI_THREADS + 64;   // + warp 16 TMA producer, warp 17 MMA issuer (owns the TMEM allocation)
constexpr int OCT_ROWS = 64;                // rows per slab of a saved octet array = batch rows (K) per stage of the dW kernel
// Saved [rows, F] bf16 matrix, slab-octet layout: element (r, f) at (((r / 64) * (F / 8) + f / 8) * 64 + r % 64) * 8 + f % 8, i.e. per
// 64-row slab one [64 rows][8] panel per 8 features, the slab's F/8 panels contiguous.  A warp of the fused kernel (32 rows, one
// 16-byte vector each) writes 512 contiguous bytes per panel; a dW stage (one slab of dH and of H) is ONE contiguous block per operand
// - one bulk copy instead of F/8 copies of 1 KB, whose per-request cost held the dW kernel at 3.5 TB/s - and already in the order
// of the MN-major UMMA core matrices ([panel][64 rows][8]).
__host__ __device__ inline int64_t oct_index(int64_t row, int oct, int n_oct) { return ((((row >> 6) * n_oct + oct) << 6) + (row & 63)) * 8; }
constexpr int MAX_OPS = 24;
constexpr int MAXO = 4;                    // widest CUDA-core head: 2*action_dim <= 4, con_dim <= 4
constexpr int KBIAS = 16;                  // extra K block of every forward weight chunk: bias (hi, lo) against a constant-ones A tile
constexpr uint32_t CHUNK_BYTES = 64 * (HID + KBIAS) * 2;   // one ring stage: 64 output columns x (K=256 + bias block) bf16
constexpr uint32_t TM_ACC = 0, TM_R0 = 256, TM_R1 = 384;

enum ASrc { A_XS0 = 0, A_XS1 = 1, A_R0 = 2, A_R1 = 3 };

// bias: the chunks carry a bias K block.  early: the op reads only shared-memory inputs that were complete before the previous op
// (its A operand is [s,a] staged at the tile start), so its MMAs need not wait for the previous op's epilogues and post step
// nchunks: 64-column output chunks the op computes (0 = all four); a narrow output layer (ensemble heads) sets 1 and the issuer
// completes the other accumulators' barriers without MMAs
struct FOp { uint32_t w_off[2]; uint16_t kp; uint8_t a_src[2]; uint8_t parts; uint8_t bias; uint8_t early; uint8_t nchunks; uint8_t pad[2]; };
// epilogue side of an op
enum Post { POST_NONE = 0, POST_POLICY0, POST_POLICY1, POST_QT0, POST_QT1, POST_KEEP, POST_QCT, POST_Q0, POST_Q1, POST_QC };
struct EOp {
  int hw_off, hb_off;                      // offsets into the constant table: head weights [no][256], head bias [no]
  uint8_t backward;                        // 0 forward epilogue (bias + ReLU), 1 backward epilogue (mask + column sums)
  uint8_t no;                              // head outputs evaluated on the CUDA cores from this layer's activation (0 = none)
  uint8_t out_region;                      // 0 none, 1 R0, 2 R1: TMEM region the packed result is stored to
  uint8_t wait_all;                        // out_region is one of the op's own A operands: wait for all of its MMAs first
  uint8_t save, hsave;                     // saved-array ids (0 = none): result to save / forward activation giving the ReLU mask
  uint8_t bias_slot;                       // backward: column-sum slot of the bias gradient
  uint8_t post;                            // Post
  uint8_t act;                             // 0 ReLU, 1 tanh (the multiplier net, src/ssac.py:105): activation of a forward op / derivative mask of a backward op
};
constexpr uint8_t NO_SLOT = 255;           // EOp::bias_slot of a backward op through a frozen net: no column sums

// Order in which a CTA pair walks the op tables.  split = 0: every tile pair runs order[0] (all ops).  split = 1: the update has two
// chains of ops that do not depend on each other (critic step: actor -> target Q's -> Q losses | safe actor -> target Qc -> Qc
// loss); job j of the grid is chain (j & 1) of tile pair (j >> 1), so small batches spread over twice as many SMs and a large
// batch is dealt in half-size pieces (shorter tail).
struct JobSched { uint8_t order[2][32]; uint8_t n[2]; uint8_t split; uint8_t pad; };
// The two chains are not equally long (critic: 79 k vs 89 k cycles), and with an even number of CTA pairs job parity = CTA-pair parity:
// the chain of a tile pair's two jobs is swapped every `n_clusters / 2` tile pairs, so a CTA pair alternates chains round by round
// instead of running the long one every time.
__device__ __forceinline__ int sched_chain(const JobSched& s, int job, int n_clusters) {
  if (!s.split) return 0;
  const int half = n_clusters > 1 ? (n_clusters >> 1) : 1;
  return (job & 1) ^ (((job >> 1) / half) & 1);
}
__device__ __forceinline__ int sched_pair(const JobSched& s, int job) { return s.split ? (job >> 1) : job; }
__device__ __forceinline__ int sched_my_jobs(const JobSched& s, int n_tiles, int cid, int n_clusters) {
  const int jobs = (n_tiles / 2) * (s.split ? 2 : 1);
  return (jobs - cid + n_clusters - 1) / n_clusters;
}

struct FusedParams {
  FOp op[MAX_OPS];
  EOp eop[MAX_OPS];
  int n_ops;
  JobSched sch;
  const uint8_t* wimg;
  const float* ctab; int ctab_floats;
  int hw_q[2], hw_cm, hw_cl;               // head weights needed again by the backward steps (offsets into ctab)
  // batch
  const float *obs, *act, *next_obs, *rew, *cv; const uint8_t* done;
  NoiseView n_actor, n_safe, n_qc;
  const float* log_alpha;
  float gamma, one_minus_gamma, td_bound, inv_bg, inv_bgc;
  int64_t B, Bpad; int S, A, C, D, Kx, stages, n_tiles;
  // saved activations (octet layout, Bpad rows)
  __nv_bfloat16* x_sa;
  __nv_bfloat16* sv[13];                   // [0] unused; ids below

  float* gacc_out; int nv;                 // [grid][nv*256] per-CTA column sums
  double* loss_part;                       // [grid][2]
  int* err_flag;
  float* dbg;                              // optional [B,16] per-row intermediates (tests)
  long long* prof;                         // optional [MAX_OPS][4] clock stamps of block 0's second tile (tools/prof_critic_ops.py)
};

enum SaveId { SV_Q_H1 = 1, SV_Q_DH2 = 3, SV_Q_DH1 = 5, SV_C_T1 = 7, SV_C_T2, SV_C_DM1, SV_C_DL1, SV_C_DT2, SV_C_DT1 };   // Q ids: + net index

// slots of the per-CTA column-sum accumulators (each 256 floats)
__host__ __device__ inline int slot_q_w2(int i) { return 3 * i; }
__host__ __device__ inline int slot_q_b1(int i) { return 3 * i + 1; }
__host__ __device__ inline int slot_q_b0(int i) { return 3 * i + 2; }
constexpr int SLOT_C_BM0 = 6, SLOT_C_BL0 = 7, SLOT_C_BT1 = 8, SLOT_C_BT0 = 9, SLOT_SCAL = 10, SLOT_C_WM = 11;
__host__ __device__ inline int slot_c_wm(int c) { return SLOT_C_WM + c; }
__host__ __device__ inline int slot_c_wl(int c, int C) { return SLOT_C_WM + C + c; }
__host__ __device__ inline int n_slots(int C) { return SLOT_C_WM + 2 * C; }
// scalars inside SLOT_SCAL: [0,1] dL/d q_i bias, [2..2+C) mean-head bias, [2+C..2+2C) log-std-head bias

constexpr int CLUSTER = 2;                 // CTA pair: every weight chunk is fetched from L2 once and multicast to both
struct FusedSmem {
  uint64_t full[6], empty[6], acc_full[NGROUPS], acc_free[NGROUPS], act_ready[2];
  uint32_t tmem_base, pad[3];
};

// ---------------------------------------------------------------------------------------------------------------
// weight images: 256 x kp bf16 per image, four 64-column chunks, each in the canonical K-major no-swizzle layout
//   [n/8][k/8][8 rows][8 elems];  transposed images hold W^T (the backward op's B operand)
// ---------------------------------------------------------------------------------------------------------------
struct PackEntry { const float* W; const float* bias; int n_real, k_real, kp, transposed; int64_t dst; };
struct PackTable { PackEntry e[48]; int n; };
// Forward images (bias != NULL) carry a 16-wide extra K block per chunk: k = kp holds bf16(b), k = kp+1 holds bf16(b - bf16(b));
// the issuer multiplies it with a constant tile of ones, so the accumulator already contains the fp32-accurate bias.
static __global__ void pack_images_kernel(PackTable t, __nv_bfloat16* __restrict__ img) {
  const PackEntry e = t.e[blockIdx.y];
  const int kt = e.kp + (e.bias ? KBIAS : 0);
  const int total = HID * kt;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / kt, k = i - n * kt;
    float v = 0.f;
    if (n < e.n_real) {
      if (k < e.k_real) v = e.transposed ? e.W[(int64_t)k * e.n_real + n] : e.W[(int64_t)n * e.k_real + k];
      else if (k == e.kp) v = e.bias[n];
      else if (k == e.kp + 1) { const float b = e.bias[n]; v = b - __bfloat162float(__float2bfloat16_rn(b)); }
    }
    const int c = n >> 6, nin = n & 63;
    const int64_t idx = (int64_t)c * 64 * kt + ((int64_t)(nin >> 3) * (kt >> 3) + (k >> 3)) * 64 + (nin & 7) * 8 + (k & 7);
    img[e.dst + idx] = __float2bfloat16_rn(v);
  }
}
// constant table: gathers biases / head weights (fp32) into one contiguous block
struct CopyEntry { const float* src; int n; int dst; int stride; };   // src == NULL: zeros
struct CopyTable { CopyEntry e[48]; int n; };

// both preparations of an update step in ONE launch (blockIdx.y < images: weight images; the rest: constant-table entries): at the
// 8-GPU shard size every launch is ~2 % of the step
static __global__ void pack_gather_kernel(const __grid_constant__ PackTable pt, __nv_bfloat16* __restrict__ img, const __grid_constant__ CopyTable ct,
                                          float* __restrict__ out) {
  if ((int)blockIdx.y < pt.n) {
    const PackEntry& e = pt.e[blockIdx.y];
    const int kt = e.kp + (e.bias ? KBIAS : 0);
    const int total = HID * kt;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int n = i / kt, k = i - n * kt;
      float v = 0.f;
      if (n < e.n_real) {
        if (k < e.k_real) v = e.transposed ? e.W[(int64_t)k * e.n_real + n] : e.W[(int64_t)n * e.k_real + k];
        else if (k == e.kp) v = e.bias[n];
        else if (k == e.kp + 1) { const float b = e.bias[n]; v = b - __bfloat162float(__float2bfloat16_rn(b)); }
      }
      const int c = n >> 6, nin = n & 63;
      const int64_t idx = (int64_t)c * 64 * kt + ((int64_t)(nin >> 3) * (kt >> 3) + (k >> 3)) * 64 + (nin & 7) * 8 + (k & 7);
      img[e.dst + idx] = __float2bfloat16_rn(v);
    }
  } else if (blockIdx.x == 0) {
    const CopyEntry& e = ct.e[blockIdx.y - pt.n];
    for (int i = threadIdx.x; i < e.n; i += blockDim.x) out[e.dst + i] = e.src ? e.src[(int64_t)i * e.stride] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf_hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }

// column sums over the 32 rows of a warp: f(j) = this row's value of column j (j is a compile-time constant at every call site,
// so the 32 values need not exist at the same time as an array); returns the sum of column `lane`
template <class F>
__device__ __forceinline__ float colsum32_f(int lane, F f) {
  float w[16];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float lo = f(i), hi = f(i + 16);
    const float keep = b4 ? hi : lo, send = b4 ? lo : hi;
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = b3 ? w[i + 8] : w[i], send = b3 ? w[i] : w[i + 8];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = b2 ? w[i + 4] : w[i], send = b2 ? w[i] : w[i + 4];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = b1 ? w[i + 2] : w[i], send = b1 ? w[i] : w[i + 2];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const float keep = b0 ? w[1] : w[0], send = b0 ? w[0] : w[1];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}
__device__ __forceinline__ float colsum32(const float (&v)[32], int lane) {
  return colsum32_f(lane, [&](int j) { return v[j]; });
}

// per-thread view of the fused kernel's epilogue state
struct Epi {
  FusedSmem* sm;
  const float* ctab;        // shared-memory copy of the constant table
  float* gacc;              // shared-memory column-sum accumulators [nv][256]
  float4* hp;               // head partials [4 groups][128 rows] (float4 = up to MAXO outputs)
  uint32_t tm;              // TMEM base + this warp's lane offset
  int g, row, lane;         // epilogue group (64-column slab), row within the tile, lane
  int64_t grow, Bpad;       // global row, padded row count
  bool valid;
  uint32_t it;              // global index of the op whose epilogue runs next (parity of the per-op barriers)
  uint32_t arr;             // act_ready arrivals made so far
  int* err;
  long long* prof;          // non-null only in the one thread that records clock stamps
};

__device__ __forceinline__ void epi_wait_acc(Epi& e) {
  mbar_wait(&e.sm->acc_full[e.g], e.it & 1, e.err, 100 + e.g);
  tc_fence_after();
  if (e.prof && e.it >= MAX_OPS && e.it < 2 * MAX_OPS) e.prof[(e.it - MAX_OPS) * 32 + 16 + 4 * e.g] = clock64();
}
// (one arrival per WARP on the epilogue -> issuer barriers: every lane fences, the warp converges, lane 0 arrives - 16 + 16 arrivals
// per op instead of 512 + 512 serialised shared-memory atomics on the hand-off path)
__device__ __forceinline__ void epi_free_acc(Epi& e) {
  tc_fence_before();
  __syncwarp();
  if (e.lane == 0) mbar_arrive(&e.sm->acc_free[e.g]);
}
// all of this op's MMAs (every chunk) have completed: the op's A regions may be overwritten
__device__ __forceinline__ void epi_wait_all_mma(Epi& e) {
  mbar_wait(&e.sm->acc_full[NGROUPS - 1], e.it & 1, e.err, 110);
  tc_fence_after();
}
__device__ __forceinline__ void epi_op_done(Epi& e) {      // this thread's TMEM / shared writes for the next op are complete
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (e.lane == 0) mbar_arrive(&e.sm->act_ready[e.arr & 1]);      // arrival number k enables op k; even / odd ops use separate barriers so
  ++e.arr;                                                        // that an issuer that observes op k's phase late can never be two phases behind
}
// 16-byte panels of 32 packed columns -> global octet layout
__device__ __forceinline__ void save_octets(const Epi& e, __nv_bfloat16* base, int half, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int oct = e.g * 8 + half * 4 + o;
    uint4 v = make_uint4(pk[4 * o], pk[4 * o + 1], pk[4 * o + 2], pk[4 * o + 3]);
    *reinterpret_cast<uint4*>(base + oct_index(e.grow, oct, HID / 8)) = v;
  }
}
__device__ __forceinline__ void add_colsum(const Epi& e, int slot, int half, const float (&v)[32]) {
  const float s = colsum32(v, e.lane);
  atomicAdd(&e.gacc[slot * HID + e.g * 64 + half * 32 + e.lane], s);
}
template <class F>
__device__ __forceinline__ void add_colsum_f(const Epi& e, int slot, int half, F f) {
  const float s = colsum32_f(e.lane, f);
  atomicAdd(&e.gacc[slot * HID + e.g * 64 + half * 32 + e.lane], s);
}

// forward epilogue of one 64-column slab: h = relu(acc + bias).
//   out_region != 0 : store packed bf16 to that TMEM region (next op's A operand / stash)
//   save != nullptr : store to the global octet array
//   no > 0          : accumulate the head dot products hpart[o] += h . hw[o][cols]
//   wait_all        : wait for every MMA of the op before touching out_region (it is one of the op's own A regions)
// One copy of this code serves all 20 forward ops (the per-op parameters come from the EOp table): the first version inlined
// one specialised copy per op, 240 KB of SASS, and paid an instruction-cache miss chain at every site of every tile.
__device__ __forceinline__ void epi_forward(Epi& e, uint32_t out_region, __nv_bfloat16* save, int hw_off, int no,
                                            float (&hpart)[MAXO], bool wait_all, int act = 0) {
  epi_wait_acc(e);
  uint32_t raw[2][32];
  tmem_ld32(e.tm + TM_ACC + e.g * 64, raw[0]);
  tmem_ld32(e.tm + TM_ACC + e.g * 64 + 32, raw[1]);
  tmem_ld_wait();
  epi_free_acc(e);
  if (e.prof && e.it >= MAX_OPS && e.it < 2 * MAX_OPS) e.prof[(e.it - MAX_OPS) * 32 + 17 + 4 * e.g] = clock64();
  if (act) {                               // tanh (1) / SiLU (2) layer: applied in place (tanh.approx, error far below the bf16 rounding that follows)
#pragma unroll
    for (int half = 0; half < 2; ++half)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = __uint_as_float(raw[half][j]);
        const float hx = 0.5f * x;
        raw[half][j] = __float_as_uint(act == 1 ? tanh_fast(x) : fmaf(hx, tanh_fast(hx), hx));      // x*sigmoid(x) = h + h*tanh(h), h = x/2
      }
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    // the accumulator already holds x W^T + b (bias block of the weight chunk): the epilogue is ReLU + pack
#pragma unroll
    for (int o = 0; o < MAXO; ++o) {
      if (o < no) {
        const float4* w4 = reinterpret_cast<const float4*>(e.ctab + hw_off + o * HID + e.g * 64 + half * 32);
        float acc = hpart[o];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 w = w4[j];
          const float h0 = __uint_as_float(raw[half][4 * j]), h1 = __uint_as_float(raw[half][4 * j + 1]);
          const float h2 = __uint_as_float(raw[half][4 * j + 2]), h3 = __uint_as_float(raw[half][4 * j + 3]);
          acc = fmaf(act ? h0 : fmaxf(h0, 0.f), w.x, acc);
          acc = fmaf(act ? h1 : fmaxf(h1, 0.f), w.y, acc);
          acc = fmaf(act ? h2 : fmaxf(h2, 0.f), w.z, acc);
          acc = fmaf(act ? h3 : fmaxf(h3, 0.f), w.w, acc);
        }
        hpart[o] = acc;
      }
    }
    if (out_region || save) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        pk[j] = act ? pack_bf16(__uint_as_float(raw[half][2 * j]), __uint_as_float(raw[half][2 * j + 1]))
                    : pack_bf16_relu(__uint_as_float(raw[half][2 * j]), __uint_as_float(raw[half][2 * j + 1]));
      if (save) save_octets(e, save, half, pk);
      if (out_region) {
        if (wait_all && half == 0) epi_wait_all_mma(e);
        tmem_st16(e.tm + out_region + e.g * 32 + half * 16, pk);
      }
    }
  }
  if (e.prof && e.it >= MAX_OPS && e.it < 2 * MAX_OPS) e.prof[(e.it - MAX_OPS) * 32 + 18 + 4 * e.g] = clock64();
  ++e.it;
}

// The same epilogue, 32 accumulator columns at a time: half the live registers.  With 18 warps per CTA the register file gives 96
// registers per thread; epi_forward's 64-register accumulator block plus an activation's temporaries exceed that, and ptxas moves the
// whole block through local memory (ncu source page of the SiLU ensemble kernel: 58 STL.64 + as many LDL.64 per thread and layer,
// 0.8 G L2 sectors per launch - the epilogue was waiting on L2 round trips of its own spills).  Costs a second tcgen05.ld round trip
// and frees the accumulator ~200 cycles later.
// n_cols: columns of the layer the NEXT op reads (its K); 32-column pieces beyond it are neither loaded nor stored (the ensemble's
// hidden width 200 -> K = 208: the last group, whose accumulator completes last, skips its second half)
__device__ __forceinline__ void epi_forward_halfwise(Epi& e, uint32_t out_region, __nv_bfloat16* save, int hw_off, int no,
                                                     float (&hpart)[MAXO], bool wait_all, int act = 0, int n_cols = HID) {
  epi_wait_acc(e);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if (e.g * 64 + half * 32 >= n_cols) {                    // (warp-uniform)
      if (half == 1) epi_free_acc(e);
      continue;
    }
    uint32_t raw[32];
    tmem_ld32(e.tm + TM_ACC + e.g * 64 + half * 32, raw);
    tmem_ld_wait();
    if (half == 1) epi_free_acc(e);
    if (half == 0 && e.prof && e.it >= MAX_OPS && e.it < 2 * MAX_OPS) e.prof[(e.it - MAX_OPS) * 32 + 17 + 4 * e.g] = clock64();
    if (act) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = __uint_as_float(raw[j]);
        const float hx = 0.5f * x;
        raw[j] = __float_as_uint(act == 1 ? tanh_fast(x) : fmaf(hx, tanh_fast(hx), hx));
      }
    }
#pragma unroll
    for (int o = 0; o < MAXO; ++o) {
      if (o < no) {
        const float4* w4 = reinterpret_cast<const float4*>(e.ctab + hw_off + o * HID + e.g * 64 + half * 32);
        float acc = hpart[o];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 w = w4[j];
          const float h0 = __uint_as_float(raw[4 * j]), h1 = __uint_as_float(raw[4 * j + 1]);
          const float h2 = __uint_as_float(raw[4 * j + 2]), h3 = __uint_as_float(raw[4 * j + 3]);
          acc = fmaf(act ? h0 : fmaxf(h0, 0.f), w.x, acc);
          acc = fmaf(act ? h1 : fmaxf(h1, 0.f), w.y, acc);
          acc = fmaf(act ? h2 : fmaxf(h2, 0.f), w.z, acc);
          acc = fmaf(act ? h3 : fmaxf(h3, 0.f), w.w, acc);
        }
        hpart[o] = acc;
      }
    }
    if (out_region || save) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        pk[j] = act ? pack_bf16(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]))
                    : pack_bf16_relu(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
      if (save) save_octets(e, save, half, pk);
      if (out_region) {
        if (wait_all && half == 0) epi_wait_all_mma(e);
        tmem_st16(e.tm + out_region + e.g * 32 + half * 16, pk);
      }
    }
  }
  if (e.prof && e.it >= MAX_OPS && e.it < 2 * MAX_OPS) e.prof[(e.it - MAX_OPS) * 32 + 18 + 4 * e.g] = clock64();
  ++e.it;
}

// backward epilogue: dh = (h > 0) ? acc : 0 with h re-read from the saved forward activation (bf16 octets, written by this
// very thread a few ops earlier, L2-resident); column sums -> bias gradient; save; optional TMEM store for the next backward op
// act = 1: the layer was a tanh layer, dh = acc * (1 - h^2).  bias_slot == NO_SLOT / save == NULL: frozen net, nothing kept.
// no > 0: dot products of the masked gradient with `no` fp32 rows of the constant table (the action columns of a frozen first
// layer: d loss / d action without an N = 256 MMA for two useful columns)
__device__ __forceinline__ void epi_backward(Epi& e, const __nv_bfloat16* hsave, int bias_slot, __nv_bfloat16* save, uint32_t out_region,
                                          bool wait_all, int act = 0, int hw_off = 0, int no = 0, float* hpart = nullptr) {
  epi_wait_acc(e);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint4 hv[4];
#pragma unroll
    for (int o = 0; o < 4; ++o)
      hv[o] = *reinterpret_cast<const uint4*>(hsave + oct_index(e.grow, e.g * 8 + half * 4 + o, HID / 8));
    uint32_t raw[32];
    tmem_ld32(e.tm + TM_ACC + e.g * 64 + half * 32, raw);
    tmem_ld_wait();
    if (half == 1) epi_free_acc(e);
    float v[32];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const uint32_t w[4] = {hv[o].x, hv[o].y, hv[o].z, hv[o].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (act) {
          const float hl = bf_lo(w[q]), hh = bf_hi(w[q]);
          v[8 * o + 2 * q] = __uint_as_float(raw[8 * o + 2 * q]) * (1.f - hl * hl);
          v[8 * o + 2 * q + 1] = __uint_as_float(raw[8 * o + 2 * q + 1]) * (1.f - hh * hh);
        } else {
          v[8 * o + 2 * q] = (w[q] & 0xFFFFu) ? __uint_as_float(raw[8 * o + 2 * q]) : 0.f;
          v[8 * o + 2 * q + 1] = (w[q] >> 16) ? __uint_as_float(raw[8 * o + 2 * q + 1]) : 0.f;
        }
      }
    }
    if (bias_slot != NO_SLOT) add_colsum(e, bias_slot, half, v);
#pragma unroll
    for (int o = 0; o < MAXO; ++o) {
      if (o < no) {
        const float4* w4 = reinterpret_cast<const float4*>(e.ctab + hw_off + o * HID + e.g * 64 + half * 32);
        float acc = hpart[o];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 w = w4[j];
          acc = fmaf(v[4 * j], w.x, acc); acc = fmaf(v[4 * j + 1], w.y, acc);
          acc = fmaf(v[4 * j + 2], w.z, acc); acc = fmaf(v[4 * j + 3], w.w, acc);
        }
        hpart[o] = acc;
      }
    }
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
    if (save) save_octets(e, save, half, pk);
    if (out_region) {
      if (wait_all && half == 0) epi_wait_all_mma(e);
      tmem_st16(e.tm + out_region + e.g * 32 + half * 16, pk);
    }
  }
  if (e.prof && e.it >= MAX_OPS && e.it < 2 * MAX_OPS) e.prof[(e.it - MAX_OPS) * 32 + 18 + 4 * e.g] = clock64();
  ++e.it;
}

// After the row's output gradients d[o] are known: read the stashed activation h (packed bf16 in `region`), accumulate the
// head-weight gradient column sums d[o]*h, form dh = (h > 0) * sum_o d[o]*hw[o][col], its column sums (bias gradient of the
// layer that produced h), store dh over h and save it.
// hglobal != NULL: h is re-read from that saved octet array instead of `region` (the activation left TMEM many ops ago), dh is
// still stored to `region`.  act = 1: tanh layer (1 - h^2 instead of the ReLU mask).  w_slot0 < 0 / bias_slot == NO_SLOT /
// save == NULL: the net is frozen, only dh is wanted.
__device__ __forceinline__ void epi_head_backward(Epi& e, uint32_t region, const float (&d)[MAXO], int no, int hw_off, int w_slot0,
                                                  int bias_slot, __nv_bfloat16* save, int act = 0, const __nv_bfloat16* hglobal = nullptr) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t hp16[16];
    if (hglobal) {
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const uint4 hv = *reinterpret_cast<const uint4*>(hglobal + oct_index(e.grow, e.g * 8 + half * 4 + o, HID / 8));
        hp16[4 * o] = hv.x; hp16[4 * o + 1] = hv.y; hp16[4 * o + 2] = hv.z; hp16[4 * o + 3] = hv.w;
      }
    } else {
      tmem_ld16(e.tm + region + e.g * 32 + half * 16, hp16);
      tmem_ld_wait();
    }
    // (h stays packed: 16 registers instead of 32, unpacked where it is used; the products d[o] * h feeding the weight-gradient
    // column sums are formed inside the butterfly - the register file gives 96 registers per thread and spills go to L2)
    auto hval = [&](int j) { return (j & 1) ? bf_hi(hp16[j >> 1]) : bf_lo(hp16[j >> 1]); };
    float dh[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) dh[j] = 0.f;
#pragma unroll
    for (int o = 0; o < MAXO; ++o) {
      if (o < no) {
        if (w_slot0 >= 0) {
          const float dd = d[o];
          add_colsum_f(e, w_slot0 + o, half, [&](int j) { return dd * hval(j); });
        }
        const float4* w4 = reinterpret_cast<const float4*>(e.ctab + hw_off + o * HID + e.g * 64 + half * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 w = w4[j];
          dh[4 * j] = fmaf(d[o], w.x, dh[4 * j]); dh[4 * j + 1] = fmaf(d[o], w.y, dh[4 * j + 1]);
          dh[4 * j + 2] = fmaf(d[o], w.z, dh[4 * j + 2]); dh[4 * j + 3] = fmaf(d[o], w.w, dh[4 * j + 3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) { const float hj = hval(j); dh[j] = act ? dh[j] * (1.f - hj * hj) : (hj > 0.f ? dh[j] : 0.f); }
    if (bias_slot != NO_SLOT) add_colsum(e, bias_slot, half, dh);
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(dh[2 * j], dh[2 * j + 1]);
    if (save) save_octets(e, save, half, pk);
    tmem_st16(e.tm + region + e.g * 32 + half * 16, pk);
  }
}

// head outputs of the row: sum of the four groups' partials + bias (all MAXO lanes; unused ones carry zeros + padding)
__device__ __forceinline__ void head_combine(Epi& e, const float (&hpart)[MAXO], int hb_off, float (&out)[MAXO]) {
  e.hp[e.g * TILE + e.row] = make_float4(hpart[0], hpart[1], hpart[2], hpart[3]);
  named_bar_sync(1, EPI_THREADS);
  float4 s = e.hp[e.row];
#pragma unroll
  for (int gg = 1; gg < NGROUPS; ++gg) {
    const float4 p = e.hp[gg * TILE + e.row];
    s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
  }
  const float4 b = *reinterpret_cast<const float4*>(e.ctab + hb_off);
  out[0] = s.x + b.x; out[1] = s.y + b.y; out[2] = s.z + b.z; out[3] = s.w + b.w;
}

// ---------------------------------------------------------------------------------------------------------------
// the two single-warp roles every fused SSAC kernel shares (P = its parameter block: op[], n_ops, wimg, stages)
// ---------------------------------------------------------------------------------------------------------------
template <class P>
__device__ __forceinline__ void fused_producer(const P& p, FusedSmem* sm, uint8_t* ring, int my_jobs, int cid, int n_clusters, uint32_t crank,
                                               int* err) {
  const int stages = p.stages;
  // ---- TMA producer: (tile, op, chunk, part) weight blocks through the ring ---------------------------------------------
  if (elect_one()) {
    uint32_t s = 0, ph = 0;
    for (int t = 0; t < my_jobs; ++t) {
      const int chain = sched_chain(p.sch, cid + t * n_clusters, n_clusters);
      for (int oi = 0; oi < p.sch.n[chain]; ++oi) {
        const FOp op = p.op[p.sch.order[chain][oi]];
        const uint32_t bytes = 64u * (op.kp + (op.bias ? KBIAS : 0)) * 2u;
        const int nch = op.nchunks ? op.nchunks : NGROUPS;
        for (int c = 0; c < nch; ++c)
          for (int part = 0; part < op.parts; ++part) {
            mbar_wait(&sm->empty[s], ph ^ 1, err, 1);   // both CTAs' MMAs on the stage's previous contents are done (multicast commits)
            mbar_expect_tx(&sm->full[s], bytes);
            const uint32_t half = bytes >> 1;
            bulk_g2s_multicast(ring + (size_t)s * CHUNK_BYTES + crank * half,
                               p.wimg + op.w_off[part] + (size_t)c * bytes + crank * half, half, &sm->full[s], (uint16_t)3);
            if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
          }
      }
    }
  }
}

template <class P>
__device__ __forceinline__ void fused_issuer(const P& p, FusedSmem* sm, uint8_t* ring, uint8_t* xs0, uint8_t* xs1, uint8_t* ones, uint32_t tmem,
                                             int my_jobs, int cid, int n_clusters, int lane, int* err, long long* prof) {
  const int stages = p.stages;
  // ---- MMA issuer ---------------------------------------------------------------------------------------------------------
  // The tensor pipe accepts only ~4 MMAs ahead of execution, so every barrier wait between two chunks is a bubble in the
  // pipe: the op's accumulators and up to `round` ring stages are awaited first, then that many chunks are issued back to back.
  uint32_t it = 0, stage = 0, phase = 0;
  const uint32_t ring_addr = smem_u32(ring), xs_addr[2] = {smem_u32(xs0), smem_u32(xs1)};
  const uint64_t ones_d = make_desc(smem_u32(ones), 2048, 128);
  const uint32_t ones_lo = (uint32_t)ones_d, ones_hi = (uint32_t)(ones_d >> 32);
  const int round = stages >= 4 ? 4 : 2;
  const uint32_t idesc = make_idesc(64);
  for (int t = 0; t < my_jobs; ++t) {
    const int chain = sched_chain(p.sch, cid + t * n_clusters, n_clusters);
    for (int oi = 0; oi < p.sch.n[chain]; ++oi, ++it) {
      const int o = p.sch.order[chain][oi];
      const FOp op = p.op[o];
      const int nk = op.kp >> 4;
      const int pshift = op.parts - 1;                              // parts is 1 or 2
      const bool stamp = prof && blockIdx.x == 0 && lane == 0 && t == 1;
      const uint32_t sbo = (uint32_t)(op.kp + (op.bias ? KBIAS : 0)) * 16;
      const int total = (op.nchunks ? op.nchunks : NGROUPS) << pshift;
      for (int i0 = 0; i0 < total; i0 += round) {
        const int nr = min(round, total - i0);
        // (the first round's weights and the accumulators become available while the previous op is still in its epilogue:
        // these waits are off the critical path; the last one - the previous op's activations - is the real dependency)
        uint32_t sw = stage, pw = phase;
        for (int r = 0; r < nr; ++r) {
          mbar_wait(&sm->full[sw], pw, err, 4);
          if (++sw == (uint32_t)stages) { sw = 0; pw ^= 1; }
        }
        if (i0 == 0) {
          if (stamp) prof[o * 32 + 8] = clock64();
          for (int c = 0; c < NGROUPS; ++c) mbar_wait(&sm->acc_free[c], (it & 1) ^ 1, err, 3);
          if (stamp) prof[o * 32 + 9] = clock64();
          // the previous op's activations / post step; an `early` op does not read them: its MMAs go first and the phase is
          // observed right after (every phase is observed, in order, before the next one of the same barrier can complete)
          if (!op.early || oi == 0) mbar_wait(&sm->act_ready[it & 1], (it >> 1) & 1, err, 2);      // (a job's first op always waits for the staging)
          if (stamp) prof[o * 32] = clock64();
        }
        tc_fence_after();
        if (stamp) prof[o * 32 + 4 + (i0 >> pshift)] = clock64();
        if (elect_one()) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            if (r < nr) {
              const int i = i0 + r, c = i >> pshift, part = i & pshift;
              uint32_t s = stage + r; if (s >= (uint32_t)stages) s -= stages;
              // B: chunk image [64 cols][kp] K-major: LBO = 128 B (next K octet), SBO = kp*16 B (next 8 columns)
              const uint64_t bd = make_desc(ring_addr + s * CHUNK_BYTES, 128, sbo);
              const uint32_t b_lo = (uint32_t)bd, b_hi = (uint32_t)(bd >> 32);
              const uint32_t d_tmem = tmem + TM_ACC + c * 64;
              const int src = op.a_src[part];
              if (src <= A_XS1) {
                // A: [128 rows][Kx] K-major in shared memory: LBO = 2048 B (next K octet), SBO = 128 B (next 8 rows)
                const uint64_t ad = make_desc(xs_addr[src], 2048, 128);
                const uint32_t a_lo = (uint32_t)ad, a_hi = (uint32_t)(ad >> 32);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (k < nk) mma_ss_p(d_tmem, a_lo + k * 256, a_hi, b_lo + k * 16, b_hi, idesc, (part | k) != 0);
              } else {
                const uint32_t a_tmem = tmem + (src == A_R0 ? TM_R0 : TM_R1);
#pragma unroll
                for (int k = 0; k < 16; ++k)                       // (nk < 16: a hidden width below 256, e.g. the ensemble's 200 -> 13 k-steps)
                  if (k < nk) mma_ts_p(d_tmem, a_tmem + k * 8, b_lo + k * 16, b_hi, idesc, (part | k) != 0);
              }
              if (op.bias) mma_ss_p(d_tmem, ones_lo, ones_hi, b_lo + nk * 16, b_hi, idesc, 1u);      // + bias (ones x bias block)
              tc_commit_multicast(&sm->empty[s], (uint16_t)3);      // frees the stage in both CTAs of the pair
              if (part == pshift) tc_commit(&sm->acc_full[c]);
            }
          }
          // a narrow op: the accumulators it does not use complete their phase too (their epilogue groups only pass the barriers)
          if (op.nchunks && i0 + nr >= total)
            for (int c = op.nchunks; c < NGROUPS; ++c) tc_commit(&sm->acc_full[c]);
        }
        __syncwarp();
        stage = sw; phase = pw;
      }
      if (op.early && oi != 0) mbar_wait(&sm->act_ready[it & 1], (it >> 1) & 1, err, 6);
      if (stamp) prof[o * 32 + 1] = clock64();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dW kernel
// ---------------------------------------------------------------------------------------------------------------
constexpr int DW_ROWS = 64;                                   // batch rows (K) per stage
constexpr int DW_STAGES = 3;
constexpr uint32_t DW_PANEL = DW_ROWS * 16;                   // one 8-feature panel of a stage: 1 KB
constexpr uint32_t DW_STAGE_BYTES = 64 * DW_PANEL;            // 32 A panels + up to 32 B panels
constexpr int DW_THREADS = 192;
constexpr int MAX_DW_JOBS = 8;
struct DwJob { const __nv_bfloat16* a; const __nv_bfloat16* b; int b_octets; int cta0, ksplit; float* partial; };
struct DwParams { DwJob job[MAX_DW_JOBS]; int n_jobs; int64_t Bpad; int n_slabs; int* err_flag; uint32_t lbo, sbo; };
struct DwSmem { uint64_t full[DW_STAGES], empty[DW_STAGES], done; uint32_t tmem_base, pad[3]; };

static __global__ void __launch_bounds__(DW_THREADS, 1) critic_dw_kernel(const __grid_constant__ DwParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  DwSmem* sm = reinterpret_cast<DwSmem*>(smem + DW_STAGES * DW_STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* err = p.err_flag;
  int ji = 0;
  while (ji + 1 < p.n_jobs && (int)blockIdx.x >= p.job[ji + 1].cta0) ++ji;
  const DwJob job = p.job[ji];
  const int split = (int)blockIdx.x - job.cta0;
  const int slab0 = (int)((int64_t)p.n_slabs * split / job.ksplit), slab1 = (int)((int64_t)p.n_slabs * (split + 1) / job.ksplit);
  const int N = job.b_octets * 8;

  if (threadIdx.x == 0) {
    for (int s = 0; s < DW_STAGES; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], 1); }
    mbar_init(&sm->done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&sm->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp == 0) {
    // producer: a stage = one 64-row slab of dH (32 panels, 32 KB contiguous) and of H (b_octets panels, contiguous): two bulk copies
    const uint32_t bytes = (32 + job.b_octets) * DW_PANEL;
    if (lane == 0) {
      for (int sl = slab0, n = 0; sl < slab1; ++sl, ++n) {
        const uint32_t s = n % DW_STAGES, ph = (n / DW_STAGES) & 1;
        mbar_wait(&sm->empty[s], ph ^ 1, err, 21);
        mbar_expect_tx(&sm->full[s], bytes);
        uint8_t* dst = smem + s * DW_STAGE_BYTES;
        bulk_g2s(dst, job.a + (int64_t)sl * 32 * (OCT_ROWS * 8), 32 * DW_PANEL, &sm->full[s]);
        bulk_g2s(dst + 32 * DW_PANEL, job.b + (int64_t)sl * job.b_octets * (OCT_ROWS * 8), job.b_octets * DW_PANEL, &sm->full[s]);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(N, 1, 1);
    const uint32_t base = smem_u32(smem);
    for (int sl = slab0, n = 0; sl < slab1; ++sl, ++n) {
      const uint32_t s = n % DW_STAGES, ph = (n / DW_STAGES) & 1;
      mbar_wait(&sm->full[s], ph, err, 22);
      tc_fence_after();
      // MN-major operands: LBO = 128 B (next 8 batch rows), SBO = 1 KB (next 8-feature panel); 16 rows per MMA = 256 B
      const uint64_t ad = make_desc(base + s * DW_STAGE_BYTES, p.lbo, p.sbo);
      const uint64_t bd = make_desc(base + s * DW_STAGE_BYTES + 32 * DW_PANEL, p.lbo, p.sbo);
      const uint32_t a_lo = (uint32_t)ad, a_hi = (uint32_t)(ad >> 32), b_lo = (uint32_t)bd, b_hi = (uint32_t)(bd >> 32);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < DW_ROWS / 16; ++kk) {
          const uint32_t accum = (n | kk) != 0;
          mma_ss_p(tmem, a_lo + kk * 16, a_hi, b_lo + kk * 16, b_hi, idesc, accum);                                 // out features 0..127
          mma_ss_p(tmem + 256, a_lo + (16 * DW_PANEL >> 4) + kk * 16, a_hi, b_lo + kk * 16, b_hi, idesc, accum);    // 128..255
        }
        tc_commit(&sm->empty[s]);
        if (sl == slab1 - 1) tc_commit(&sm->done);
      }
      __syncwarp();
    }
  } else {
    // epilogue warps 2..5: TMEM lanes 32*(warp%4)..  = output feature within the half
    mbar_wait(&sm->done, 0, err, 23);
    tc_fence_after();
    const int q = warp & 3;
    float* out = job.partial + (int64_t)split * HID * N;
    for (int half = 0; half < 2; ++half) {
      float* orow = out + (int64_t)(half * 128 + q * 32 + lane) * N;
      for (int cb = 0; cb < N; cb += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + half * 256 + cb, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(orow + cb + 4 * j) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// gradient assembly: grads[dst + r*cols + c] = sum_s src[s*stride + r*ld + c]
// ---------------------------------------------------------------------------------------------------------------
struct ReduceEntry { int64_t dst; const float* src; int n_src; int64_t stride; int rows, cols, ld; };
struct ReduceTable { ReduceEntry e[40]; int n; };
static __global__ void __launch_bounds__(256) critic_grad_reduce_kernel(ReduceTable t, float* __restrict__ grads) {
  const ReduceEntry e = t.e[blockIdx.y];
  const int total = e.rows * e.cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / e.cols, c = i - r * e.cols;
    const float* s = e.src + (int64_t)r * e.ld + c;
    float acc = 0.f;
    for (int k = 0; k < e.n_src; ++k) acc += s[k * e.stride];
    grads[e.dst + i] = acc;
  }
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace cu
}  // namespace drpo
