// DRPO_PREC_BF16 ensemble forward OUTSIDE the rollout (BatchedGaussianEnsemble._forward1 / _forward_all / means / sample /
// elite_samples, src/dynamics.py:112-134, 198-234) on the fused tcgen05 op tables of umma_ops.cuh.  A job is (128-row tile pair,
// member): the member's six dense layers run back to back on one CTA pair -
//     x0 = [(s - mean)/(std + 1e-6), a] -> trunk0 -> SiLU -> trunk1 -> SiLU -> { diff0 -> SiLU -> diff1 ; lvar0 -> SiLU -> lvar1 }
// - as tcgen05.mma bf16 x bf16 -> fp32 (hidden width 200 zero-padded to 256, biases folded into the MMA), activations kept in TMEM
// as the next layer's A operand; the two narrow output layers compute ONE 64-column chunk; their epilogue adds the state (means),
// soft-clamps the log-variance and - in sample mode - draws next_state / reward, all from the fp32 accumulator.  Nothing but the
// inputs and the [B, S+1] outputs touches HBM.  The rollout itself uses the dedicated kernel of rollout_pipe.cuh.
#include "critic_umma_api.h"
#include "umma_ops.cuh"

namespace drpo {
namespace cu {

constexpr int ENS_MAX_MEMBERS = 8, ENS_OPS = 6;
struct EnsSched { uint8_t order[ENS_MAX_MEMBERS][8]; uint8_t n[ENS_MAX_MEMBERS]; int members; };
__device__ __forceinline__ int sched_chain(const EnsSched& s, int job, int /*n_clusters*/) { return job % s.members; }
__device__ __forceinline__ int sched_pair(const EnsSched& s, int job) { return job / s.members; }
__device__ __forceinline__ int sched_my_jobs(const EnsSched& s, int n_tiles, int cid, int n_clusters) {
  return ((n_tiles / 2) * s.members - cid + n_clusters - 1) / n_clusters;
}

struct EnsParams {
  FOp op[ENS_MAX_MEMBERS * ENS_OPS];
  int n_ops;
  EnsSched sch;
  const uint8_t* wimg;
  int stages;
  const float *states, *actions; int64_t slot_stride_s, slot_stride_a;      // per member slot (0: every member reads the same rows)
  const float *norm_mean, *norm_std, *min_lv, *max_lv;
  float *means, *log_vars; int64_t slot_stride_out;                          // forward mode
  int sample; NoiseView noise; float *next_states, *rewards;                 // sample mode (one member)
  int64_t B, Bpad; int S, A, Kx, Kh, n_tiles;               // Kh: padded hidden width = K of the hidden-input layers
  int* err_flag;
};

__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(F_THREADS, 1) ens_fused_kernel(const __grid_constant__ EnsParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int stages = p.stages;
  uint8_t* ring = smem;
  uint8_t* xs0 = ring + (size_t)stages * CHUNK_BYTES;
  uint8_t* ones = xs0 + TILE * p.Kx * 2;
  float* obuf = reinterpret_cast<float*>(ones + TILE * KBIAS * 2);          // [128 rows][S + 1]: the tile's output rows, staged for coalesced stores
  FusedSmem* sm = reinterpret_cast<FusedSmem*>(obuf + TILE * 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int PRODUCER = EPI_THREADS / 32, ISSUER = PRODUCER + 1;
  int* err = p.err_flag;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], CLUSTER); }
    for (int g = 0; g < NGROUPS; ++g) { mbar_init(&sm->acc_full[g], 1); mbar_init(&sm->acc_free[g], 4); }
    mbar_init(&sm->act_ready[0], EPI_THREADS / 32); mbar_init(&sm->act_ready[1], EPI_THREADS / 32);
    fence_barrier_init();
  }
  if (warp == ISSUER) tmem_alloc(&sm->tmem_base, 512);
  for (int i = threadIdx.x; i < TILE * KBIAS; i += F_THREADS) {
    const int k = (i >> 10) * 8 + (i & 7);
    reinterpret_cast<__nv_bfloat16*>(ones)[i] = __float2bfloat16_rn(k < 2 ? 1.f : 0.f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  const uint32_t crank = cluster_ctarank();
  const int n_clusters = (int)gridDim.x / CLUSTER, cid = (int)blockIdx.x / CLUSTER;
  const int my_jobs = sched_my_jobs(p.sch, p.n_tiles, cid, n_clusters);

  if (warp == PRODUCER) {
    fused_producer(p, sm, ring, my_jobs, cid, n_clusters, crank, err);
  } else if (warp == ISSUER) {
    fused_issuer(p, sm, ring, xs0, xs0, ones, tmem, my_jobs, cid, n_clusters, lane, err, (long long*)nullptr);
  } else {
    Epi e;
    e.sm = sm; e.ctab = nullptr; e.gacc = nullptr; e.hp = nullptr; e.g = warp >> 2; e.lane = lane; e.row = (warp & 3) * 32 + lane;
    e.tm = tmem + ((uint32_t)((warp & 3) * 32) << 16); e.it = 0; e.arr = 0; e.err = err; e.Bpad = p.Bpad; e.prof = nullptr;
    const int S = p.S, A = p.A, D = S + A, O = S + 1;
    float hpart[MAXO] = {0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < my_jobs; ++t) {
      const int job = cid + t * n_clusters, slot = sched_chain(p.sch, job, n_clusters);
      const int tile = CLUSTER * sched_pair(p.sch, job) + (int)crank;
      e.grow = (int64_t)tile * TILE + e.row;
      e.valid = e.grow < p.B;
      const int64_t gr = e.valid ? e.grow : 0;
      const float* st = p.states + slot * p.slot_stride_s + gr * S;
      const float* ac = p.actions + slot * p.slot_stride_a + gr * A;
      // ---- x0 = [(s - mean) / (std + 1e-6), a]                                   src/normalization.py:23-24, src/dynamics.py:113-114
      for (int j = e.g; j < (p.Kx >> 3); j += NGROUPS) {
        uint32_t w0[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int k = j * 8 + q * 2 + u;
            v[u] = !e.valid ? 0.f : (k < S ? (st[k] - p.norm_mean[k]) / (p.norm_std[k] + 1e-6f) : (k < D ? ac[k - S] : 0.f));
          }
          w0[q] = pack_bf16(v[0], v[1]);
        }
        *reinterpret_cast<uint4*>(xs0 + j * 2048 + e.row * 16) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
      }
      fence_proxy_async();
      epi_op_done(e);
      // ---- trunk0 -> R0, trunk1 -> R1, diff0 -> R0, diff1 (narrow), lvar0 -> R0, lvar1 (narrow) ---------------------------------
#pragma unroll 1
      for (int oi = 0; oi < ENS_OPS; ++oi) {
        if (oi != 3 && oi != 5) {
          epi_forward_halfwise(e, oi == 1 ? TM_R1 : TM_R0, nullptr, 0, 0, hpart, false, 2, p.Kh);
        } else {
          // output layer: columns 0..S of accumulator 0 hold [next_state - state, reward] (oi == 3) or the raw log-variance (oi == 5)
          epi_wait_acc(e);
          float* orow = obuf + e.row * O;
          // The thread's row goes to shared memory first and the 128 rows leave as ONE contiguous block ([B, S+1] row-major: a tile's
          // rows are adjacent) instead of 4-byte pieces at a 4(S+1)-byte stride per thread.  32 columns at a time (register budget).
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t raw[32];
            if (e.g == 0 && 32 * half < O) {
              tmem_ld32(e.tm + TM_ACC + 32 * half, raw);
              tmem_ld_wait();
            }
            if (half == 1) { epi_free_acc(e); ++e.it; }
            if (e.g == 0 && 32 * half < O) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int c = 32 * half + j;
                if (c < O) {
                  const float v = __uint_as_float(raw[j]);
                  if (oi == 3) orow[c] = __fadd_rn(v, (c < S && e.valid) ? st[c] : 0.f);                  // means = diffs + [s, 0]   src/dynamics.py:118
                  else if (!p.sample) orow[c] = soft_clamp(v, p.min_lv[c], p.max_lv[c]);                // soft-clamped log-variance     :119-121
                  else orow[c] = fmaf(sqrtf(expf(soft_clamp(v, p.min_lv[c], p.max_lv[c]))), e.valid ? p.noise.get(e.grow, c) : 0.f, orow[c]);   // :201-203
                }
              }
            }
          }
          if (e.g == 0) {
            named_bar_sync(2, TILE);
            const int64_t row0 = (int64_t)tile * TILE;
            const int rows = (int)max((int64_t)0, min((int64_t)TILE, p.B - row0));
            if (!p.sample) {
              float* dst = (oi == 3 ? p.means : p.log_vars) + slot * p.slot_stride_out + row0 * O;
              for (int i = e.row; i < rows * O; i += TILE) dst[i] = obuf[i];
            } else if (oi == 5) {
              float* dst = p.next_states + row0 * S;
              for (int i = e.row; i < rows * S; i += TILE) { const int r = i / S; dst[i] = obuf[r * O + (i - r * S)]; }
              if (e.valid) p.rewards[e.grow] = orow[S];
            }
            if (!p.sample || oi == 5) named_bar_sync(2, TILE);                                         // (sample mode keeps the means until the log-var head)
          }
        }
        if (oi != ENS_OPS - 1) epi_op_done(e);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == ISSUER) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int64_t ens_img_bytes(int Kx, int members) {
  return (int64_t)members * ((int64_t)HID * (Kx + KBIAS) * 2 + (int64_t)5 * HID * (HID + KBIAS) * 2);
}
bool ens_bf16_supported(const drpo_ensemble& e) { return e.hidden <= HID && e.state_dim + e.action_dim <= 64 && e.state_dim + 1 <= 64; }
int64_t ens_bf16_ws_bytes(const drpo_ensemble& e) {
  return align_up(ens_img_bytes(round_up(e.state_dim + e.action_dim, 16), ENS_MAX_MEMBERS), 256) + 8192;
}

struct EnsLayers { drpo_linear l[ENS_OPS]; };
static EnsLayers ens_member(const drpo_ensemble& e, int m) {
  const int S = e.state_dim, A = e.action_dim, H = e.hidden, D = S + A, O = S + 1;
  EnsLayers n;
  n.l[0] = {e.trunk0_w + (int64_t)m * H * D, e.trunk0_b + (int64_t)m * H, D, H};
  n.l[1] = {e.trunk1_w + (int64_t)m * H * H, e.trunk1_b + (int64_t)m * H, H, H};
  n.l[2] = {e.diff0_w + (int64_t)m * H * H, e.diff0_b + (int64_t)m * H, H, H};
  n.l[3] = {e.diff1_w + (int64_t)m * O * H, e.diff1_b + (int64_t)m * O, H, O};
  n.l[4] = {e.lvar0_w + (int64_t)m * H * H, e.lvar0_b + (int64_t)m * H, H, H};
  n.l[5] = {e.lvar1_w + (int64_t)m * O * H, e.lvar1_b + (int64_t)m * O, H, O};
  return n;
}

// members m0..m1-1 on `batch` rows; forward mode (means / log_vars, slot = member - m0) or sample mode (one member)
int ens_bf16_run(const drpo_ensemble& e, int m0, int m1, int per_member_inputs, const float* states, const float* actions, int64_t batch,
                 float* means, float* log_vars, const drpo_noise* noise, float* next_states, float* rewards, void* workspace,
                 int64_t workspace_bytes, void* stream, int* err_flag) {
  cudaStream_t st = (cudaStream_t)stream;
  const int S = e.state_dim, A = e.action_dim, O = S + 1, Kx = round_up(S + A, 16);
  const int Kh = round_up(e.hidden, 16);                     // K of the hidden-input layers: 208 for the reference's 200 (13 k-steps, not 16)
  DRPO_CHECK_ARG(ens_bf16_supported(e), "ensemble forward (bf16): needs hidden <= 256 and state_dim + action_dim <= 64");
  EnsParams fp; memset(&fp, 0, sizeof(fp));
  const int64_t Bpad = (batch + CLUSTER * TILE - 1) / (CLUSTER * TILE) * (CLUSTER * TILE);
  fp.B = batch; fp.Bpad = Bpad; fp.S = S; fp.A = A; fp.Kx = Kx; fp.Kh = Kh; fp.n_tiles = (int)(Bpad / TILE);
  const size_t fixed = (size_t)TILE * Kx * 2 + TILE * KBIAS * 2 + (size_t)TILE * 64 * sizeof(float) + sizeof(FusedSmem);
  int stages = (int)((232448 - 1024 - fixed) / CHUNK_BYTES);
  if (stages > 6) stages = 6;
  fp.stages = stages;
  const size_t smem = fixed + (size_t)stages * CHUNK_BYTES;
  Arena ar(workspace, workspace_bytes);
  for (int g0 = m0; g0 < m1; g0 += ENS_MAX_MEMBERS) {           // (ensembles of more than 8 members: several launches)
    const int g1 = std::min(m1, g0 + ENS_MAX_MEMBERS), nm = g1 - g0;
    uint8_t* img = g0 == m0 ? ar.take<uint8_t>(ens_img_bytes(Kx, ENS_MAX_MEMBERS)) : const_cast<uint8_t*>(fp.wimg);
    if (!ar.ok()) { set_error("ensemble forward (bf16): workspace too small"); return DRPO_ERR_WORKSPACE; }
    PackTable pt; pt.n = 0;
    int64_t img_off = 0;
    for (int s = 0; s < nm; ++s) {
      const EnsLayers n = ens_member(e, g0 + s);
      for (int l = 0; l < ENS_OPS; ++l) {
        PackEntry& pe = pt.e[pt.n++];
        pe.W = n.l[l].w; pe.bias = n.l[l].b; pe.transposed = 0; pe.n_real = n.l[l].out_dim; pe.k_real = n.l[l].in_dim;
        pe.kp = l == 0 ? Kx : Kh; pe.dst = img_off / 2;
        FOp& op = fp.op[s * ENS_OPS + l];
        op.w_off[0] = (uint32_t)img_off; op.kp = (uint16_t)pe.kp; op.parts = 1; op.bias = 1; op.early = 0;
        op.a_src[0] = (uint8_t)(l == 0 ? A_XS0 : (l == 1 || l == 3 || l == 5 ? A_R0 : A_R1));
        op.nchunks = (l == 3 || l == 5) ? 1 : 0;
        img_off += (int64_t)HID * (pe.kp + KBIAS) * 2;
        fp.sch.order[s][l] = (uint8_t)(s * ENS_OPS + l);
      }
      fp.sch.n[s] = ENS_OPS;
    }
    fp.sch.members = nm; fp.n_ops = nm * ENS_OPS; fp.wimg = img;
    {
      dim3 grid(32, pt.n);
      DRPO_LAUNCH(pack_images_kernel, grid, 256, 0, st, pt, reinterpret_cast<__nv_bfloat16*>(img));
    }
    const int64_t slot0 = g0 - m0;
    fp.states = states + (per_member_inputs ? slot0 * batch * S : 0); fp.actions = actions + (per_member_inputs ? slot0 * batch * A : 0);
    fp.slot_stride_s = per_member_inputs ? batch * S : 0; fp.slot_stride_a = per_member_inputs ? batch * A : 0;
    fp.norm_mean = e.norm_mean; fp.norm_std = e.norm_std; fp.min_lv = e.min_log_var; fp.max_lv = e.max_log_var;
    fp.sample = noise ? 1 : 0;
    if (noise) {
      fp.noise = make_noise(noise->eps, noise->row_stride, noise->seed, noise->stream_tag, noise->step);
      fp.next_states = next_states; fp.rewards = rewards;
    } else {
      fp.means = means + slot0 * batch * O; fp.log_vars = log_vars + slot0 * batch * O; fp.slot_stride_out = batch * O;
    }
    fp.err_flag = err_flag;
    static int max_clusters = 0;
    if (!max_clusters) {
      DRPO_CUDA_OK(cudaFuncSetAttribute(ens_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
      cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(148); cfg.blockDim = dim3(F_THREADS); cfg.dynamicSmemBytes = 232448 - 1024;
      cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = CLUSTER; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, ens_fused_kernel, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = 148 / CLUSTER; }
      max_clusters = n;
    }
    const int64_t jobs = (int64_t)(fp.n_tiles / CLUSTER) * nm;
    const int grid = (int)std::min<int64_t>(jobs, max_clusters) * CLUSTER;
    DRPO_LAUNCH(ens_fused_kernel, grid, F_THREADS, smem, st, fp);
  }
  return DRPO_OK;
}

}  // namespace cu
}  // namespace drpo
