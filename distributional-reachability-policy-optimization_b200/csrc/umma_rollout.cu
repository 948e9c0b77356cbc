// DRPO_PREC_BF16 rollout: one persistent, warp-specialised tcgen05 kernel per rollout step that runs the whole
//   policy MLP -> squashed-Gaussian sample -> ensemble-member MLP (trunk + 2 heads) -> Gaussian next-state sample
//   -> env hooks
// chain for a 128-row tile without ever leaving the SM:
//   * every dense layer is a tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) with M = 128 rows of trajectories,
//   * the accumulator AND the activations live in TMEM (the A operand of layer l+1 is read from TMEM, where the epilogue
//     of layer l stored it as packed bf16), so activations never touch shared or global memory,
//   * weights are pre-packed once per rollout into the UMMA canonical K-major (no-swizzle) layout and streamed from L2
//     into a shared-memory ring by TMA bulk copies (cp.async.bulk, mbarrier complete_tx),
//   * bias rides in the GEMM: every activation tile carries a constant-1 column and the packed weights carry the bias in
//     the matching K slot, so the epilogue is activation + bf16 pack only.
// Warp roles: warps 0-3 = epilogue (thread t owns TMEM lane t = trajectory row t), warp 4 = TMA producer,
// warp 5 = MMA issuer (one elected thread) + TMEM allocator.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "hooks.cuh"
#include "nets.cuh"
#include "rollout.cuh"
#include "umma_api.h"

namespace drpo {
namespace umma {

constexpr int TILE_M = 128;
constexpr int KCHUNK = 64;                 // K elements per weight chunk (4 MMA k-steps)
constexpr int MAX_CHUNKS = 48;
constexpr int MAX_LAYERS = 9;
constexpr int NUM_THREADS = 192;
constexpr uint32_t TM_ACC = 0, TM_ACTA = 256, TM_ACTB = 392, TM_COLS = 512;
constexpr int ACTA_COLS = 136, ACTB_COLS = 120;

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
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (fails the launch) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag, int code) {
  for (uint32_t it = 0; it < (1u << 22); ++it)
    if (mbar_try_wait(bar, parity)) return;
  if (err_flag) atomicExch(err_flag, code);
  __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T      (A: 128 lanes x K bf16 packed two per column; B: K-major canonical layout)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {       // element 2j in the low half, 2j+1 in the high half
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float silu_fast(float x) { const float h = 0.5f * x; return fmaf(h, tanh_fast(h), h); }   // x*sigmoid(x)

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
// weight image: per layer, per 64-wide K chunk, a contiguous block in canonical layout  [n/8][k/8][8 rows][8 elems]
// ---------------------------------------------------------------------------------------------------------------
struct LayerSpec {
  int n_real, k_real;      // nn.Linear out / in
  int np, kp;              // padded MMA N (x16) and K (x16, includes the bias slot at k_real)
  int act;                 // ACT_RELU / ACT_SILU / ACT_NONE (final layers)
  int a_col, out_col;      // TMEM column of the A operand / of the activation written by the epilogue
  int first_chunk, n_chunks;
};
struct ChunkSpec { uint32_t offset, bytes; uint16_t n, kc; };     // byte offset inside the net image
struct NetPlan {
  LayerSpec layer[MAX_LAYERS];
  ChunkSpec chunk[MAX_CHUNKS];
  int n_layers, n_chunks;
  uint32_t policy_bytes, model_bytes;     // image sizes; policy chunks index into the policy image, model chunks into the member image
  int n_policy_chunks;
  uint32_t max_chunk_bytes;
};

__global__ void pack_layer_kernel(const float* __restrict__ W, const float* __restrict__ b, int n_real, int k_real, int np, int kp,
                                  __nv_bfloat16* __restrict__ dst /* start of this layer inside the image */) {
  // one thread per padded element (n, k)
  const int total = np * kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / kp, k = i % kp;
    float v = 0.f;
    if (n < n_real) v = k < k_real ? W[(int64_t)n * k_real + k] : (k == k_real ? b[n] : 0.f);
    const int kc0 = (k / KCHUNK) * KCHUNK, kc = min(KCHUNK, kp - kc0), kk = k - kc0;
    const int64_t chunk_off = (int64_t)np * kc0;                                   // elements before this chunk
    const int64_t idx = chunk_off + ((int64_t)(n >> 3) * (kc >> 3) + (kk >> 3)) * 64 + (n & 7) * 8 + (kk & 7);
    dst[idx] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the fused step kernel
// ---------------------------------------------------------------------------------------------------------------
struct StepParams {
  NetPlan plan;
  const uint8_t* policy_img; const uint8_t* model_img;
  // data
  const float* cur; const int32_t* ids; const int* n_dev; int64_t n_max;
  float *actions, *next_states, *rewards, *cv; uint8_t *done, *viol;
  const float *norm_mean, *norm_std, *min_lv, *max_lv;
  NoiseView noise_p, noise_m;
  drpo_env_params env;
  int S, A, C, SP, OP, stages;
  int* err_flag;
  // debug / self-test: dump the fp32 accumulator of layer `dump_layer` (n_real columns) into dump_out[row, col]
  int dump_layer; float* dump_out;
};

struct SmemLayout {
  uint64_t full[8], empty[8], in_ready, acc_ready;
  uint32_t tmem_base, pad;
};

__device__ __forceinline__ void act_store_const_tail(uint32_t lane_base, uint32_t out_col, int from_elem, int kp, int one_at) {
  // activation elements [from_elem, kp) are constants: 1.0 at `one_at`, 0 elsewhere (from_elem, kp multiples of 16)
  for (int e0 = from_elem; e0 < kp; e0 += 16) {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pk[j] = pack_bf16((e0 + 2 * j) == one_at ? 1.f : 0.f, (e0 + 2 * j + 1) == one_at ? 1.f : 0.f);
    tmem_st8(lane_base + out_col + (uint32_t)(e0 >> 1), pk);
  }
}

// epilogue of a hidden layer: ACC[0,np) -> act -> bf16 pairs -> TMEM activation buffer with the constant-1 bias column
__device__ __forceinline__ void hidden_epilogue(uint32_t lane_base, const LayerSpec& L, int next_kp) {
  for (int c0 = 0; c0 < L.np; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(lane_base + TM_ACC + (uint32_t)c0, r);
    tmem_ld_wait();
    float v[16];
    if (L.act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(__uint_as_float(r[j]), 0.f);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = silu_fast(__uint_as_float(r[j]));
    }
    if (L.n_real >= c0 && L.n_real < c0 + 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) if (c0 + j == L.n_real) v[j] = 1.f;           // bias slot of the next layer
    }
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
    tmem_st8(lane_base + (uint32_t)L.out_col + (uint32_t)(c0 >> 1), pk);
  }
  if (next_kp > L.np) act_store_const_tail(lane_base, (uint32_t)L.out_col, L.np, next_kp, L.n_real);
}

// write one input row (k_real values from smem scratch + constant 1) as packed bf16 into a TMEM activation buffer
__device__ __forceinline__ void write_input_row(uint32_t lane_base, uint32_t col, const float* row, int k_real, int kp) {
  for (int e0 = 0; e0 < kp; e0 += 16) {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int a = e0 + 2 * j, b = a + 1;
      const float lo = a < k_real ? row[a] : (a == k_real ? 1.f : 0.f);
      const float hi = b < k_real ? row[b] : (b == k_real ? 1.f : 0.f);
      pk[j] = pack_bf16(lo, hi);
    }
    tmem_st8(lane_base + col + (uint32_t)(e0 >> 1), pk);
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1) rollout_step_umma_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const NetPlan& plan = p.plan;
  // carve shared memory: [weight ring | state tile | out tile | x tile | barriers]
  uint8_t* ring = smem_raw;
  const uint32_t slot_bytes = (plan.max_chunk_bytes + 1023u) & ~1023u;
  float* st_s = reinterpret_cast<float*>(ring + (size_t)slot_bytes * p.stages);      // [128][SP] current states (fp32)
  float* st_o = st_s + TILE_M * p.SP;                                                // [128][OP] diff-head output -> next state
  SmemLayout* sl = reinterpret_cast<SmemLayout*>(st_o + TILE_M * p.OP);

  int n = min((int64_t)*p.n_dev, p.n_max);
  const int n_tiles = (n + TILE_M - 1) / TILE_M;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&sl->full[s], 1); mbar_init(&sl->empty[s], 1); }
    mbar_init(&sl->in_ready, TILE_M);
    mbar_init(&sl->acc_ready, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(&sl->tmem_base, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sl->tmem_base;

  if (warp == 4) {
    // ===================== TMA producer: stream every weight chunk of every tile through the ring =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int c = 0; c < plan.n_chunks; ++c, ++it) {
          const int s = it % p.stages; const uint32_t round = it / p.stages;
          if (round > 0) mbar_wait(&sl->empty[s], (round - 1) & 1, p.err_flag, 1);
          const ChunkSpec& ch = plan.chunk[c];
          const uint8_t* src = (c < plan.n_policy_chunks ? p.policy_img : p.model_img) + ch.offset;
          mbar_expect_tx(&sl->full[s], ch.bytes);
          bulk_g2s(ring + (size_t)s * slot_bytes, src, ch.bytes, &sl->full[s]);
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0, lphase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < plan.n_layers; ++l, ++lphase) {
          const LayerSpec& L = plan.layer[l];
          mbar_wait(&sl->in_ready, lphase & 1, p.err_flag, 2);           // A operand written, accumulator free
          tc_fence_after();
          const uint32_t idesc = make_idesc(L.np);
          int kdone = 0;
          for (int c = 0; c < L.n_chunks; ++c, ++it) {
            const int s = it % p.stages; const uint32_t round = it / p.stages;
            const ChunkSpec& ch = plan.chunk[L.first_chunk + c];
            mbar_wait(&sl->full[s], round & 1, p.err_flag, 3);
            tc_fence_after();
            const uint32_t b_base = smem_u32(ring + (size_t)s * slot_bytes);
            const uint32_t sbo = (uint32_t)(ch.kc >> 3) * 128u;
            for (int ks = 0; ks < ch.kc; ks += 16, kdone += 16) {
              const uint64_t bdesc = make_b_desc(b_base + (uint32_t)(ks >> 3) * 128u, 128u, sbo);
              mma_ts(tmem + TM_ACC, tmem + (uint32_t)L.a_col + (uint32_t)(kdone >> 1), bdesc, idesc, kdone > 0 ? 1u : 0u);
            }
            tc_commit(&sl->empty[s]);                                    // frees the ring slot when these MMAs retire
          }
          tc_commit(&sl->acc_ready);                                     // accumulator complete -> epilogue
        }
      }
    }
  } else {
    // ===================== epilogue warps: thread t <-> TMEM lane t <-> trajectory row t of the tile =====================
    const int t = threadIdx.x;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    const int S = p.S, A = p.A, O = S + 1;
    uint32_t aphase = 0;
    float* my_s = st_s + t * p.SP;
    float* my_o = st_o + t * p.OP;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t row0 = (int64_t)tile * TILE_M;
      const int rows = min(TILE_M, n - (int)row0);
      const bool valid = t < rows;
      const int64_t row = row0 + t;
      // ---- E0: stage the tile's states (coalesced), write the policy input [s, 1] into TMEM ----
      asm volatile("bar.sync 1, 128;" ::: "memory");                     // previous tile's readers of st_s/st_o are done
      for (int i = t; i < TILE_M * S; i += TILE_M) {
        const int r = i / S, c = i - r * S;
        st_s[r * p.SP + c] = r < rows ? p.cur[row0 * S + i] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      int l = 0;
      write_input_row(lane_base, (uint32_t)plan.layer[0].a_col, my_s, S, plan.layer[0].kp);
      tmem_st_wait(); tc_fence_before(); mbar_arrive(&sl->in_ready);

      float act_v[4];          // sampled action (A <= 4)
      for (; l < plan.n_layers; ++l, ++aphase) {
        const LayerSpec& L = plan.layer[l];
        mbar_wait(&sl->acc_ready, aphase & 1, p.err_flag, 4);
        tc_fence_after();
        if (p.dump_layer == l) {                                         // self-test hook: raw accumulator to global
          for (int c0 = 0; c0 < L.np; c0 += 16) {
            uint32_t r[16]; tmem_ld16(lane_base + TM_ACC + (uint32_t)c0, r); tmem_ld_wait();
            if (valid) for (int j = 0; j < 16; ++j) if (c0 + j < L.n_real) p.dump_out[row * L.n_real + c0 + j] = __uint_as_float(r[j]);
          }
        }
        if (L.act != ACT_NONE) {
          hidden_epilogue(lane_base, L, plan.layer[l + 1].kp);
        } else if (l == 2) {
          // ---- policy head: [mu, raw] -> a = tanh(mu + exp(-6 + 10 sigmoid(raw)) eps)      src/policy.py:89-97 ----
          uint32_t r[16]; tmem_ld16(lane_base + TM_ACC, r); tmem_ld_wait();
          const int64_t id = valid ? (int64_t)p.ids[row] : 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < A) {
              float mu = 0.f, raw = 0.f;
#pragma unroll
              for (int q = 0; q < 16; ++q) { if (q == j) mu = __uint_as_float(r[q]); if (q == A + j) raw = __uint_as_float(r[q]); }
              const float sd = __expf(-6.f + 10.f / (1.f + __expf(-raw)));
              const float x = valid ? fmaf(p.noise_p.get(id, j), sd, mu) : 0.f;
              act_v[j] = tanhf(x);
              if (valid) p.actions[row * A + j] = act_v[j];
            }
          }
          // model input x0 = [(s - mean)/(std + 1e-6), a, 1]                               src/dynamics.py:113-114
          for (int c = 0; c < S; ++c) my_o[c] = (my_s[c] - p.norm_mean[c]) / (p.norm_std[c] + 1e-6f);
#pragma unroll
          for (int j = 0; j < 4; ++j) if (j < A) my_o[S + j] = act_v[j];
          write_input_row(lane_base, (uint32_t)plan.layer[3].a_col, my_o, S + A, plan.layer[3].kp);
        } else if (l == 6) {
          // ---- diff head: means = diffs + [s, 0]  -> shared-memory row                  src/dynamics.py:118 ----
          for (int c0 = 0; c0 < L.np; c0 += 16) {
            uint32_t r[16]; tmem_ld16(lane_base + TM_ACC + (uint32_t)c0, r); tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) if (c0 + j < O) my_o[c0 + j] = __uint_as_float(r[j]) + (c0 + j < S ? my_s[c0 + j] : 0.f);
          }
        } else {
          // ---- log-var head + sampling + hooks                                         src/dynamics.py:119-121,201-203 ----
          const int64_t id = valid ? (int64_t)p.ids[row] : 0;
          float reward = 0.f;
          for (int c0 = 0; c0 < L.np; c0 += 16) {
            uint32_t r[16]; tmem_ld16(lane_base + TM_ACC + (uint32_t)c0, r); tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = c0 + j;
              if (c < O) {
                const float lv = soft_clamp(__uint_as_float(r[j]), p.min_lv[c], p.max_lv[c]);
                const float sd = sqrtf(__expf(lv));
                const float y = valid ? fmaf(sd, p.noise_m.get(id, c), my_o[c]) : 0.f;
                if (c < S) my_o[c] = y; else reward = y;
              }
            }
          }
          if (valid) {
            HookOut ho;
            eval_hooks(p.env, [my_o](int d) { return my_o[d]; }, ho);
            p.rewards[row] = reward; p.done[row] = ho.done; p.viol[row] = ho.viol;
            for (int c = 0; c < p.C; ++c) p.cv[row * p.C + c] = ho.cv[c];
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int i = t; i < rows * S; i += TILE_M) {                   // coalesced store of the tile's next states
            const int r = i / S, c = i - r * S;
            p.next_states[row0 * S + i] = st_o[r * p.OP + c];
          }
        }
        if (l + 1 < plan.n_layers) {                                     // hand the accumulator (and new activations) back
          tmem_st_wait(); tc_fence_before(); mbar_arrive(&sl->in_ready);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, TM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static void add_layer(NetPlan& P, int n_real, int k_real, int act, int a_col, int out_col, uint32_t& img_off) {
  LayerSpec& L = P.layer[P.n_layers++];
  L.n_real = n_real; L.k_real = k_real; L.np = round_up(n_real, 16); L.kp = round_up(k_real + 1, 16);
  L.act = act; L.a_col = a_col; L.out_col = out_col; L.first_chunk = P.n_chunks; L.n_chunks = 0;
  for (int k0 = 0; k0 < L.kp; k0 += KCHUNK) {
    ChunkSpec& c = P.chunk[P.n_chunks++];
    c.kc = (uint16_t)std::min(KCHUNK, L.kp - k0); c.n = (uint16_t)L.np;
    c.offset = img_off + (uint32_t)L.np * k0 * 2; c.bytes = (uint32_t)L.np * c.kc * 2;
    P.max_chunk_bytes = std::max(P.max_chunk_bytes, c.bytes);
    ++L.n_chunks;
  }
  img_off += (uint32_t)L.np * L.kp * 2;
}

static int build_plan(const drpo_rollout_args& a, NetPlan& P) {
  memset(&P, 0, sizeof(P));
  const int S = a.ensemble->state_dim, A = a.ensemble->action_dim, Hm = a.ensemble->hidden;
  const int Hp = a.actor->l0.out_dim;
  if (a.actor->l1.out_dim != Hp || Hp > 256 || round_up(Hp + 1, 16) / 2 > ACTA_COLS || round_up(Hm + 1, 16) / 2 > ACTB_COLS ||
      round_up(S + A + 1, 16) / 2 > 32 || 2 * A > 16 || A > 4 || round_up(S + 1, 16) > 256 || Hm > 256) {
    set_error("bf16 rollout: dims outside the TMEM plan (S=%d A=%d Hp=%d Hm=%d)", S, A, Hp, Hm);
    return DRPO_ERR_UNSUPPORTED;
  }
  uint32_t off = 0;
  add_layer(P, Hp, S, ACT_RELU, TM_ACTB, TM_ACTA, off);          // P1
  add_layer(P, Hp, Hp, ACT_RELU, TM_ACTA, TM_ACTA, off);         // P2 (in place: all MMAs retired before the epilogue)
  add_layer(P, 2 * A, Hp, ACT_NONE, TM_ACTA, 0, off);            // P3 -> policy head
  P.policy_bytes = off; P.n_policy_chunks = P.n_chunks;
  off = 0;
  add_layer(P, Hm, S + A, ACT_SILU, TM_ACTB, TM_ACTA, off);      // M1
  add_layer(P, Hm, Hm, ACT_SILU, TM_ACTA, TM_ACTB, off);         // M2 -> h2 kept in ACTB for both heads
  add_layer(P, Hm, Hm, ACT_SILU, TM_ACTB, TM_ACTA, off);         // M3d
  add_layer(P, S + 1, Hm, ACT_NONE, TM_ACTA, 0, off);            // M4d -> diffs
  add_layer(P, Hm, Hm, ACT_SILU, TM_ACTB, TM_ACTA, off);         // M3l
  add_layer(P, S + 1, Hm, ACT_NONE, TM_ACTA, 0, off);            // M4l -> log-vars
  P.model_bytes = off;
  return DRPO_OK;
}

static int pack_net(const drpo_linear* lin, const LayerSpec* L, int count, uint8_t* img, void* stream) {
  uint32_t off = 0;
  for (int i = 0; i < count; ++i) {
    DRPO_LAUNCH(pack_layer_kernel, grid_for((int64_t)L[i].np * L[i].kp), 256, 0, stream, lin[i].w, lin[i].b, L[i].n_real, L[i].k_real,
                L[i].np, L[i].kp, reinterpret_cast<__nv_bfloat16*>(img + off));
    off += (uint32_t)L[i].np * L[i].kp * 2;
  }
  return DRPO_OK;
}

static int smem_bytes_for(const NetPlan& P, int S, int stages, int& SP, int& OP) {
  SP = S | 1; OP = std::max(S + 1, S + 4 + 1) | 1;           // st_o also holds the model input row [norm s, a]
  const uint32_t slot = (P.max_chunk_bytes + 1023u) & ~1023u;
  return (int)(slot * stages + (size_t)TILE_M * (SP + OP) * 4 + sizeof(SmemLayout) + 64);
}

}  // namespace umma

using namespace umma;

int64_t umma_rollout_ws_bytes(const drpo_rollout_args& a) {
  NetPlan P;
  if (build_plan(a, P) != DRPO_OK) return 0;
  return rollout_ws_bytes_fp32(a) + (int64_t)align_up(P.policy_bytes, 1024) + (int64_t)a.ensemble->ensemble_size * align_up(P.model_bytes, 1024) + 4096;
}

int umma_rollout_impl(const drpo_rollout_args& a, int dump_layer, float* dump_out) {
  NetPlan P; int rc;
  if ((rc = build_plan(a, P))) return rc;
  const int64_t B = a.batch; const int S = a.ensemble->state_dim, A = a.ensemble->action_dim, C = a.env->con_dim;
  const int H = dump_layer >= 0 ? 1 : a.horizon; void* stream = a.stream;
  if (dump_layer >= P.n_layers) { set_error("debug dump: layer %d out of range", dump_layer); return DRPO_ERR_ARG; }
  Arena ar(a.workspace, a.workspace_bytes);
  RolloutScratch w;
  w.curA = ar.take<float>(B * S); w.curB = ar.take<float>(B * S); w.actions = ar.take<float>(B * A);
  w.next_states = ar.take<float>(B * S); w.rewards = ar.take<float>(B); w.cv = ar.take<float>(B * C);
  w.done = ar.take<uint8_t>(B); w.viol = ar.take<uint8_t>(B);
  w.idsA = ar.take<int32_t>(B); w.idsB = ar.take<int32_t>(B); w.n_alive = ar.take<int32_t>(a.horizon + 2);
  const int nblocks = (int)((B + CBLK - 1) / CBLK);
  w.block_counts = ar.take<int32_t>(nblocks + 1);
  w.st = ar.take<RolloutState>(1);
  int* err_flag = ar.take<int>(4);
  uint8_t* pol_img = ar.take<uint8_t>(align_up(P.policy_bytes, 1024));
  const int E = a.ensemble->ensemble_size;
  uint8_t* mem_img = ar.take<uint8_t>((int64_t)E * align_up(P.model_bytes, 1024));
  if (!ar.ok()) { set_error("drpo_rollout(bf16): workspace too small (%lld needed, %lld given)", (long long)ar.off, (long long)a.workspace_bytes); return DRPO_ERR_WORKSPACE; }

  // ---- pack the actor and every member this rollout uses into the UMMA layout (bf16, bias folded in) ----
  {
    drpo_linear pl[3] = {a.actor->l0, a.actor->l1, a.actor->l2};
    if ((rc = pack_net(pl, &P.layer[0], 3, pol_img, stream))) return rc;
    bool used[64] = {false};
    for (int t = 0; t < H; ++t) used[a.member_idx_host[t]] = true;
    for (int m = 0; m < E; ++m) {
      if (!used[m]) continue;
      MemberNet mn = member_of(*a.ensemble, m);
      drpo_linear ml[6] = {mn.t0, mn.t1, mn.d0, mn.d1, mn.l0, mn.l1};
      if ((rc = pack_net(ml, &P.layer[3], 6, mem_img + (int64_t)m * align_up(P.model_bytes, 1024), stream))) return rc;
    }
  }
  int dev = 0, sms = 148, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  int SP, OP, stages = 6;
  while (stages > 2 && smem_bytes_for(P, S, stages, SP, OP) > max_smem) --stages;
  const int smem = smem_bytes_for(P, S, stages, SP, OP);
  DRPO_CUDA_OK(cudaFuncSetAttribute(rollout_step_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));

  DRPO_CUDA_OK(cudaMemsetAsync(err_flag, 0, 16, (cudaStream_t)stream));
  DRPO_CUDA_OK(cudaMemcpyAsync(w.curA, a.initial_states, sizeof(float) * B * S, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  DRPO_LAUNCH(rollout_init_kernel, grid_for(B), 256, 0, stream, w.idsA, B, a.traj_id_offset, w.n_alive, w.st, a.virt.pointer);
  float* cur = w.curA; float* nxt = w.curB; int32_t* ids = w.idsA; int32_t* ids_n = w.idsB;
  const int grid = (int)std::min<int64_t>(sms, (B + TILE_M - 1) / TILE_M);
  for (int t = 0; t < H; ++t) {
    const int* n_dev = w.n_alive + t;
    StepParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.plan = P; sp.policy_img = pol_img; sp.model_img = mem_img + (int64_t)a.member_idx_host[t] * align_up(P.model_bytes, 1024);
    sp.cur = cur; sp.ids = ids; sp.n_dev = n_dev; sp.n_max = B;
    sp.actions = w.actions; sp.next_states = w.next_states; sp.rewards = w.rewards; sp.cv = w.cv; sp.done = w.done; sp.viol = w.viol;
    sp.norm_mean = a.ensemble->norm_mean; sp.norm_std = a.ensemble->norm_std; sp.min_lv = a.ensemble->min_log_var; sp.max_lv = a.ensemble->max_log_var;
    sp.noise_p = make_noise(a.eps_policy ? a.eps_policy + (int64_t)t * a.eps_batch_stride * A : nullptr, A, a.seed, TAG_ROLLOUT_POLICY, (uint32_t)t);
    sp.noise_m = make_noise(a.eps_model ? a.eps_model + (int64_t)t * a.eps_batch_stride * (S + 1) : nullptr, S + 1, a.seed, TAG_ROLLOUT_MODEL, (uint32_t)t);
    sp.env = *a.env; sp.S = S; sp.A = A; sp.C = C; sp.SP = SP; sp.OP = OP; sp.stages = stages; sp.err_flag = err_flag;
    sp.dump_layer = dump_layer; sp.dump_out = dump_out;
    DRPO_LAUNCH(rollout_step_umma_kernel, grid, NUM_THREADS, smem, stream, sp);
    DRPO_LAUNCH(rollout_store_kernel, grid_for(B * S), 256, 0, stream, a.virt, w.st, n_dev, cur, w.actions, w.next_states,
                w.rewards, w.done, w.viol, w.cv);
    DRPO_LAUNCH(compact_count_kernel, nblocks, CBLK, 0, stream, w.done, n_dev, w.block_counts);
    DRPO_LAUNCH(compact_scan_kernel, 1, CBLK, 0, stream, w.block_counts, nblocks, w.n_alive, t, w.st, a.step_counts);
    DRPO_LAUNCH(compact_scatter_kernel, nblocks, CBLK, 0, stream, w.done, n_dev, w.block_counts, w.next_states, ids, nxt, ids_n, S);
    float* tf = cur; cur = nxt; nxt = tf;
    int32_t* ti = ids; ids = ids_n; ids_n = ti;
  }
  DRPO_LAUNCH(rollout_finish_kernel, 1, 1, 0, stream, a.virt.pointer, w.st, a.step_counts, H);
  return DRPO_OK;
}

int umma_rollout(const drpo_rollout_args& a) { return umma_rollout_impl(a, -1, nullptr); }
int umma_debug_layer(const drpo_rollout_args& a, int layer, float* out) { return umma_rollout_impl(a, layer, out); }

}  // namespace drpo
