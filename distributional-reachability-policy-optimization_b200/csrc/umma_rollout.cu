// DRPO_PREC_BF16 rollout (src/smbpo.py:229-249): host side of the fused tcgen05 step kernel (rollout_pipe.cuh: policy MLP ->
// squashed-Gaussian sample -> ensemble-member MLP -> Gaussian next-state sample, two 128-row tiles in flight per SM, CTA-pair MMAs)
// and the two HBM-bound satellites of a step: env hooks fused with the replay-ring store, and the order-preserving compaction of the
// surviving trajectories.  One launch of each per rollout step; the per-step row counts stay on the device.
#include <cuda_bf16.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "common.cuh"
#include "hooks.cuh"
#include "nets.cuh"
#include "rollout.cuh"
#include "umma_api.h"
#include "critic_umma_api.h"
#include "rollout_pipe.cuh"

namespace drpo {

// check_done / check_violation / get_constraint_values of the step's next states (src/smbpo.py:238-240) fused with
// buffer.extend: the step's rows go into the ring at (base + r) % capacity (src/sampling.py:128-145, src/smbpo.py:241-242,248).
__global__ void __launch_bounds__(256) hooks_store_kernel(drpo_env_params env, drpo_buffer buf, const RolloutState* st, const int32_t* n_dev,
                                                          const float* __restrict__ s, const float* __restrict__ a, const float* __restrict__ ns,
                                                          const float* __restrict__ rew, uint8_t* __restrict__ done_out,
                                                          int32_t* __restrict__ chunk_counts /* survivors per 256 rows */) {
  // 32-bit index arithmetic and one 64-bit modulo per thread: the step's rows are contiguous in the ring, slot(r) = first + r
  // with at most one wrap (the first version spent its time in 64-bit divisions, not in memory traffic)
  const int n = *n_dev;
  const int64_t cap = buf.capacity, first = st->base % cap;
  const int S = buf.state_dim, A = buf.action_dim, C = buf.con_dim;
  const unsigned stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  auto slot_of = [&](unsigned r) { int64_t x = first + r; return x >= cap ? x - cap : x; };
  const unsigned nS = (unsigned)n * (unsigned)S, nA = (unsigned)n * (unsigned)A;
  if ((S & 3) == 0) {                                                  // 16-byte rows: float4 copies (all row bases are 16-byte aligned)
    const unsigned S4 = (unsigned)S >> 2, nS4 = (unsigned)n * S4;
    const float4* s4 = reinterpret_cast<const float4*>(s); const float4* ns4 = reinterpret_cast<const float4*>(ns);
    float4* d4 = reinterpret_cast<float4*>(buf.states); float4* dn4 = reinterpret_cast<float4*>(buf.next_states);
    unsigned i = t0;
    for (; i + stride < nS4; i += 2 * stride) {                       // two rows of loads in flight per thread
      const unsigned j = i + stride;
      const float4 a0 = s4[i], b0 = ns4[i], a1 = s4[j], b1 = ns4[j];
      const unsigned r0 = i / S4, r1 = j / S4;
      const int64_t o0 = slot_of(r0) * S4 + (i - r0 * S4), o1 = slot_of(r1) * S4 + (j - r1 * S4);
      d4[o0] = a0; dn4[o0] = b0; d4[o1] = a1; dn4[o1] = b1;
    }
    for (; i < nS4; i += stride) {
      const unsigned r = i / S4, c = i - r * S4; const int64_t o = slot_of(r) * S4 + c;
      d4[o] = s4[i];
      dn4[o] = ns4[i];
    }
  } else {
    for (unsigned i = t0; i < nS; i += stride) {
      const unsigned r = i / (unsigned)S, c = i - r * (unsigned)S; const int64_t o = slot_of(r) * S + c;
      buf.states[o] = s[i];
      buf.next_states[o] = ns[i];
    }
  }
  for (unsigned i = t0; i < nA; i += stride) {
    const unsigned r = i / (unsigned)A, c = i - r * (unsigned)A;
    buf.actions[slot_of(r) * A + c] = a[i];
  }
  // env hooks + per-row record fields, walked in chunks of 256 consecutive rows dealt to ALL blocks (an even tail), so that the
  // kernel also produces the compaction's survivor counts per chunk (this used to be a separate pass over `done`)
  __shared__ int kept[8];
  for (unsigned rb = blockIdx.x; rb * 256u < (unsigned)n; rb += gridDim.x) {
    const unsigned r = rb * 256u + threadIdx.x;
    bool keep = false;
    if (r < (unsigned)n) {
      const float* row = ns + (int64_t)r * S;
      HookOut ho;
      eval_hooks(env, [row](int d) { return row[d]; }, ho);
      const int64_t slot = slot_of(r);
      buf.rewards[slot] = rew[r]; buf.dones[slot] = ho.done; buf.violations[slot] = ho.viol;
      done_out[r] = ho.done;
      keep = !ho.done;
#pragma unroll
      for (int cc = 0; cc < DRPO_MAX_CON; ++cc) if (cc < C) buf.constraint_values[slot * C + cc] = ho.cv[cc];
    }
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0) kept[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) chunk_counts[rb] = kept[0] + kept[1] + kept[2] + kept[3] + kept[4] + kept[5] + kept[6] + kept[7];
    __syncthreads();
  }
}

// order-preserving compaction of the survivors (states = next_states[~dones], src/smbpo.py:243-246) with the scan folded in:
// every block sums the counts of the blocks before it (<= 1024 ints out of L2), the last active block publishes the next
// step's row count and advances the ring base
__global__ void __launch_bounds__(CBLK) compact_scatter_scan_kernel(const uint8_t* __restrict__ done, const int32_t* n_dev,
                                                                    const int32_t* __restrict__ chunk_counts, const float* __restrict__ ns,
                                                                    const int32_t* __restrict__ ids, float* __restrict__ cur_next,
                                                                    int32_t* __restrict__ ids_next, int S, int32_t* n_alive, int t,
                                                                    RolloutState* st, int32_t* step_counts) {
  __shared__ int warp_off[CBLK / 32];
  __shared__ int dst_row[CBLK];                 // destination row of every kept row of this block, -1 for dropped rows
  __shared__ int base_sh;
  const int n = *n_dev;
  if (n == 0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { step_counts[t] = 0; n_alive[t + 1] = 0; }
    return;
  }
  const int64_t r0 = (int64_t)blockIdx.x * CBLK;
  if (r0 >= n) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int rows = (int)min((int64_t)CBLK, (int64_t)n - r0);
  const int64_t r = r0 + threadIdx.x;
  // issue every independent load first (this row's flag and id, the block's state rows for the 16-byte fast path): their latency
  // overlaps the prefix over the chunk counts instead of following it
  const bool in = r < n;
  const uint8_t dflag = in ? done[r] : (uint8_t)1;
  const int32_t my_id = in ? ids[r] : 0;
  const bool vec = (S & 3) == 0 && S <= 16;     // <= 4 float4 per thread prefetched in registers
  const int S4 = S >> 2;
  float4 pre[4];
  if (vec) {
    const float4* src = reinterpret_cast<const float4*>(ns) + r0 * S4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * CBLK;
      if (q < S4 && i < rows * S4) pre[q] = src[i];
    }
  }
  // exclusive prefix of this block: survivors of all 256-row chunks before it
  int pref = 0;
  for (int i = threadIdx.x; i < 4 * (int)blockIdx.x; i += CBLK) pref += chunk_counts[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pref += __shfl_xor_sync(0xffffffffu, pref, o);
  if (lane == 0) warp_off[w] = pref;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = warp_off[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) base_sh = v;
  }
  __syncthreads();
  const int base = base_sh;
  const bool keep = in && !dflag;
  const unsigned b = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_off[w] = __popc(b);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = warp_off[threadIdx.x], incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int x = __shfl_up_sync(0xffffffffu, incl, o); if (threadIdx.x >= o) incl += x; }
    warp_off[threadIdx.x] = incl - v;
    if (threadIdx.x == 31 && r0 + CBLK >= n) {          // last active block: totals
      step_counts[t] = n;
      st->base += n;
      n_alive[t + 1] = base + incl;
    }
  }
  __syncthreads();
  const int pos = keep ? base + warp_off[w] + __popc(b & ((1u << lane) - 1u)) : -1;
  dst_row[threadIdx.x] = pos;
  if (keep) ids_next[pos] = my_id;
  __syncthreads();
  if (vec) {
    float4* dst = reinterpret_cast<float4*>(cur_next);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * CBLK;
      if (q < S4 && i < rows * S4) {
        const int rr = i / S4, c = i - rr * S4;
        const int d = dst_row[rr];
        if (d >= 0) dst[(int64_t)d * S4 + c] = pre[q];
      }
    }
    return;
  }
  if ((S & 3) == 0) {                                     // 16-byte rows: float4 copies
    const float4* src = reinterpret_cast<const float4*>(ns) + r0 * S4;
    float4* dst = reinterpret_cast<float4*>(cur_next);
    for (int i = threadIdx.x; i < rows * S4; i += CBLK) {
      const int rr = i / S4, c = i - rr * S4;
      const int d = dst_row[rr];
      if (d >= 0) dst[(int64_t)d * S4 + c] = src[i];
    }
    return;
  }
  for (int i = threadIdx.x; i < rows * S; i += CBLK) {
    const int rr = i / S, c = i - rr * S;
    const int d = dst_row[rr];
    if (d >= 0) cur_next[(int64_t)d * S + c] = ns[r0 * S + i];
  }
}

static thread_local int g_timing_on = 0;
static thread_local std::vector<std::pair<cudaEvent_t, cudaEvent_t>>* g_timing_events = nullptr;
static thread_local std::vector<cudaEvent_t>* g_timing_tail = nullptr;     // end of each step's HBM-bound satellites (hooks/store + compaction)

void umma_timing_enable(int on) {
  g_timing_on = on;
  if (!g_timing_events) { g_timing_events = new std::vector<std::pair<cudaEvent_t, cudaEvent_t>>(); g_timing_tail = new std::vector<cudaEvent_t>(); }
  for (auto& e : *g_timing_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  for (auto& e : *g_timing_tail) cudaEventDestroy(e);
  g_timing_events->clear(); g_timing_tail->clear();
}
int umma_timing_read(double* total_ms, int64_t* launches, double* tail_ms) {
  double t = 0, tt = 0; int64_t n = 0;
  if (g_timing_events)
    for (size_t i = 0; i < g_timing_events->size(); ++i) {
      auto& e = (*g_timing_events)[i];
      float ms = 0.f, ms2 = 0.f;
      if (cudaEventSynchronize(e.second) != cudaSuccess || cudaEventElapsedTime(&ms, e.first, e.second) != cudaSuccess) { set_error("drpo_timing_read: event query failed"); return DRPO_ERR_CUDA; }
      if (i < g_timing_tail->size() && cudaEventSynchronize((*g_timing_tail)[i]) == cudaSuccess) cudaEventElapsedTime(&ms2, e.second, (*g_timing_tail)[i]);
      t += ms; tt += ms2; ++n;
    }
  *total_ms = t; *launches = n; if (tail_ms) *tail_ms = tt;
  return DRPO_OK;
}

static int device_max_smem() { int dev = 0, m = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&m, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); return m; }

static int64_t rollout_ws_bytes_v3(const drpo_rollout_args& a) {
  r3::Plan P; int stages, smem;
  if (r3::build_plan(a, P, device_max_smem(), stages, smem) != DRPO_OK) return 0;
  return rollout_ws_bytes_fp32(a) + (int64_t)align_up(P.pol_bytes, 1024) +
         (int64_t)a.ensemble->ensemble_size * align_up(P.mem_bytes, 1024) + 4096 + ((a.batch + CBLK - 1) / CBLK) * 16 + 64;
}

// third generation (rollout_pipe.cuh): two tiles in flight per SM
static int rollout_impl_v3(const drpo_rollout_args& a, int dump_layer, float* dump_out) {
  r3::Plan P; int rc, stages = 0, smem = 0;
  if ((rc = r3::build_plan(a, P, device_max_smem(), stages, smem))) return rc;
  const int64_t B = a.batch; const int S = a.ensemble->state_dim, A = a.ensemble->action_dim;
  const int H = dump_layer >= 0 ? 1 : a.horizon; void* stream = a.stream;
  if (dump_layer >= r3::N_LAYERS && dump_layer != 100) { set_error("debug dump: layer %d out of range", dump_layer); return DRPO_ERR_ARG; }
  if (a.env->con_dim > DRPO_MAX_CON) { set_error("con_dim %d > %d", a.env->con_dim, DRPO_MAX_CON); return DRPO_ERR_ARG; }
  const int E = a.ensemble->ensemble_size;
  if (E > 64) { set_error("bf16 rollout: ensemble_size %d > 64", E); return DRPO_ERR_ARG; }

  Arena ar(a.workspace, a.workspace_bytes);
  RolloutScratch w;
  w.curA = ar.take<float>(B * S); w.curB = ar.take<float>(B * S); w.actions = ar.take<float>(B * A);
  w.next_states = ar.take<float>(B * S); w.rewards = ar.take<float>(B);
  w.done = ar.take<uint8_t>(B);
  w.idsA = ar.take<int32_t>(B); w.idsB = ar.take<int32_t>(B); w.n_alive = ar.take<int32_t>(a.horizon + 2);
  const int nblocks = (int)((B + CBLK - 1) / CBLK);
  w.block_counts = ar.take<int32_t>(4 * nblocks + 4);       // survivors per 256-row chunk
  w.st = ar.take<RolloutState>(1);
  int* err_flag = ar.take<int>(4);
  uint8_t* pol_img = ar.take<uint8_t>(align_up(P.pol_bytes, 1024));
  const int64_t mem_stride = align_up(P.mem_bytes, 1024);
  uint8_t* mem_img = ar.take<uint8_t>((int64_t)E * mem_stride);
  if (!ar.ok()) { set_error("drpo_rollout(bf16): workspace too small (%lld needed, %lld given)", (long long)ar.off, (long long)a.workspace_bytes); return DRPO_ERR_WORKSPACE; }

  // ---- pack the actor and every member this rollout uses into the k-step tile images (bf16, canonical K-major) ----
  {
    drpo_linear pl[3] = {a.actor->l0, a.actor->l1, a.actor->l2};
    std::vector<r3::PackJob> jobs;
    r3::pack_jobs(P, pl, 0, 3, pol_img, jobs);
    bool used[64] = {false};
    for (int t = 0; t < H; ++t) used[a.member_idx_host[t]] = true;
    for (int m = 0; m < E; ++m) {
      if (!used[m]) continue;
      MemberNet mn = member_of(*a.ensemble, m);
      drpo_linear ml[6] = {mn.t0, mn.t1, mn.d0, mn.l0, mn.d1, mn.l1};       // layer ids 3 trunk0, 4 trunk1, 5 diff0, 6 lvar0, 7 diff1, 8 lvar1
      r3::pack_jobs(P, ml, 3, 6, mem_img + (int64_t)m * mem_stride, jobs);
    }
    if ((rc = r3::pack_flush(jobs, stream))) return rc;
  }
  static int max_clusters = 0;
  {
    static int attr_smem = 0;
    if (smem > attr_smem) {
      DRPO_CUDA_OK(cudaFuncSetAttribute(r3::rollout_step_pipe_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      DRPO_CUDA_OK(cudaFuncSetAttribute(r3::rollout_step_pipe_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      DRPO_CUDA_OK(cudaFuncSetAttribute(r3::rollout_step_pipe_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      DRPO_CUDA_OK(cudaFuncSetAttribute(r3::rollout_step_pipe_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_smem = smem; max_clusters = 0;
    }
    if (!max_clusters) {
      cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(148); cfg.blockDim = dim3(r3::NUM_THREADS); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = r3::CLUSTER; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, r3::rollout_step_pipe_kernel<false, false>, &cfg) != cudaSuccess || nc < 1) { cudaGetLastError(); nc = 148 / r3::CLUSTER; }
      max_clusters = nc;
    }
  }
  DRPO_CUDA_OK(cudaMemsetAsync(err_flag, 0, 16, (cudaStream_t)stream));
  const bool streamed = a.init_ready_flags != nullptr && dump_layer < 0;
  if (!streamed) DRPO_CUDA_OK(cudaMemcpyAsync(w.curA, a.initial_states, sizeof(float) * B * S, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  DRPO_LAUNCH(rollout_init_kernel, grid_for(B), 256, 0, stream, w.idsA, B, a.traj_id_offset, w.n_alive, w.st, a.virt.pointer);
  float* cur = streamed ? const_cast<float*>(a.initial_states) : w.curA;
  float* nxt = streamed ? w.curA : w.curB; int32_t* ids = w.idsA; int32_t* ids_n = w.idsB;
  const int64_t n_quads = ((B + r3::TILE_M - 1) / r3::TILE_M + 3) / 4;
  const int grid = r3::CLUSTER * (int)std::max<int64_t>(1, std::min<int64_t>(max_clusters, n_quads));
  for (int t = 0; t < H; ++t) {
    const int* n_dev = w.n_alive + t;
    NoiseView np_ = make_noise(a.eps_policy ? a.eps_policy + (int64_t)t * a.eps_batch_stride * A : nullptr, A, a.seed, TAG_ROLLOUT_POLICY, (uint32_t)t);
    NoiseView nm_ = make_noise(a.eps_model ? a.eps_model + (int64_t)t * a.eps_batch_stride * (S + 1) : nullptr, S + 1, a.seed, TAG_ROLLOUT_MODEL, (uint32_t)t);
    r3::StepParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.plan = P.d; sp.policy_img = pol_img; sp.model_img = mem_img + (int64_t)a.member_idx_host[t] * mem_stride;
    sp.cur = cur; sp.n_dev = n_dev; sp.n_max = B; sp.ids = ids; sp.noise_p = np_; sp.noise_m = nm_;
    sp.actions = w.actions; sp.next_states = w.next_states; sp.rewards = w.rewards;
    sp.norm_mean = a.ensemble->norm_mean; sp.norm_std = a.ensemble->norm_std; sp.min_lv = a.ensemble->min_log_var; sp.max_lv = a.ensemble->max_log_var;
    sp.S = S; sp.A = A; sp.stages = stages; sp.err_flag = err_flag;
    sp.dump_layer = dump_layer; sp.dump_out = dump_out;
    sp.ready_flags = (streamed && t == 0) ? a.init_ready_flags : nullptr; sp.ready_shift = a.init_rows_per_flag_log2;
    std::pair<cudaEvent_t, cudaEvent_t> ev{};
    if (g_timing_on) { cudaEventCreate(&ev.first); cudaEventCreate(&ev.second); cudaEventRecord(ev.first, (cudaStream_t)stream); }
    if (B > 0) {
      if (P.d.wide) {
        if (dump_layer >= 0) { DRPO_LAUNCH((r3::rollout_step_pipe_kernel<true, true>), grid, r3::NUM_THREADS, smem, stream, sp); }
        else { DRPO_LAUNCH((r3::rollout_step_pipe_kernel<false, true>), grid, r3::NUM_THREADS, smem, stream, sp); }
      } else {
        if (dump_layer >= 0) { DRPO_LAUNCH((r3::rollout_step_pipe_kernel<true, false>), grid, r3::NUM_THREADS, smem, stream, sp); }
        else { DRPO_LAUNCH((r3::rollout_step_pipe_kernel<false, false>), grid, r3::NUM_THREADS, smem, stream, sp); }
      }
    }
    if (g_timing_on) { cudaEventRecord(ev.second, (cudaStream_t)stream); g_timing_events->push_back(ev); }
    DRPO_LAUNCH(hooks_store_kernel, grid_for(B * S), 256, 0, stream, *a.env, a.virt, w.st, n_dev, cur, w.actions, w.next_states, w.rewards, w.done,
                w.block_counts);
    DRPO_LAUNCH(compact_scatter_scan_kernel, nblocks, CBLK, 0, stream, w.done, n_dev, w.block_counts, w.next_states, ids, nxt, ids_n, S,
                w.n_alive, t, w.st, a.step_counts);
    if (g_timing_on) { cudaEvent_t e3; cudaEventCreate(&e3); cudaEventRecord(e3, (cudaStream_t)stream); g_timing_tail->push_back(e3); }
    if (streamed && t == 0) { cur = w.curA; nxt = w.curB; }         // never write into the caller's start states
    else { float* tf = cur; cur = nxt; nxt = tf; }
    int32_t* ti = ids; ids = ids_n; ids_n = ti;
  }
  DRPO_LAUNCH(rollout_finish_kernel, 1, 1, 0, stream, a.virt.pointer, w.st, a.step_counts, H, (const int*)err_flag);
  return publish_status(err_flag, 0, nullptr, stream);
}

int64_t umma_rollout_ws_bytes(const drpo_rollout_args& a) { return rollout_ws_bytes_v3(a); }
int umma_rollout(const drpo_rollout_args& a) { return rollout_impl_v3(a, -1, nullptr); }
int umma_debug_layer(const drpo_rollout_args& a, int layer, float* out) { return rollout_impl_v3(a, layer, out); }

}  // namespace drpo
