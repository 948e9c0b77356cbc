// SMBPO.rollout (src/smbpo.py:229-249) on the device, with no host synchronisation:
//   per step  policy sample -> member sample -> hooks -> ring-buffer store -> order-preserving compaction.
// The number of alive rows lives in device memory (n_alive[t]); every kernel of step t is launched for the upper
// bound B0 and trims itself.  Compaction is a 3-kernel order-preserving stream compaction (block counts, one-block
// scan, scatter) so transitions land step-major in survivor order exactly like the reference's boolean-mask
// indexing (src/smbpo.py:243-246).
#pragma once
#include "nets.cuh"

namespace drpo {

struct RolloutState {         // device-resident control block
  int64_t base;               // ring position of the current step's first row (monotone, like SampleBuffer._pointer)
  int32_t pad[2];
};

constexpr int CBLK = 1024;

static __global__ void rollout_init_kernel(int32_t* ids, int64_t n, int64_t id_offset, int32_t* n_alive, RolloutState* st,
                                    const int64_t* pointer) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    ids[i] = (int32_t)(id_offset + i);
  if (blockIdx.x == 0 && threadIdx.x == 0) { n_alive[0] = (int32_t)n; st->base = *pointer; }
}

// write the step's rows into the ring at (base + r) % capacity  (SampleBuffer.extend with wrap, src/sampling.py:128-145)
static __global__ void rollout_store_kernel(drpo_buffer buf, const RolloutState* st, const int32_t* n_dev,
                                     const float* __restrict__ s, const float* __restrict__ a,
                                     const float* __restrict__ ns, const float* __restrict__ rew,
                                     const uint8_t* __restrict__ done, const uint8_t* __restrict__ viol,
                                     const float* __restrict__ cv) {
  const int64_t n = *n_dev, base = st->base, cap = buf.capacity;
  const int S = buf.state_dim, A = buf.action_dim, C = buf.con_dim;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = t0; i < n * S; i += stride) {
    const int64_t r = i / S; const int c = (int)(i % S); const int64_t slot = (base + r) % cap;
    buf.states[slot * S + c] = s[i];
    buf.next_states[slot * S + c] = ns[i];
  }
  for (int64_t i = t0; i < n * A; i += stride) {
    const int64_t r = i / A; const int c = (int)(i % A);
    buf.actions[((base + r) % cap) * A + c] = a[i];
  }
  for (int64_t i = t0; i < n * C; i += stride) {
    const int64_t r = i / C; const int c = (int)(i % C);
    buf.constraint_values[((base + r) % cap) * C + c] = cv[i];
  }
  for (int64_t r = t0; r < n; r += stride) {
    const int64_t slot = (base + r) % cap;
    buf.rewards[slot] = rew[r]; buf.dones[slot] = done[r]; buf.violations[slot] = viol[r];
  }
}

static __global__ void __launch_bounds__(CBLK) compact_count_kernel(const uint8_t* __restrict__ done, const int32_t* n_dev,
                                                             int32_t* __restrict__ block_counts) {
  __shared__ int warp_cnt[CBLK / 32];
  const int n = *n_dev;
  const int64_t r = (int64_t)blockIdx.x * CBLK + threadIdx.x;
  const bool keep = r < n && !done[r];
  const unsigned b = __ballot_sync(0xffffffffu, keep);
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = warp_cnt[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = v;
  }
}

// one block: exclusive scan of the block counts; publishes next step's row count and advances the ring base
static __global__ void __launch_bounds__(CBLK) compact_scan_kernel(int32_t* __restrict__ block_counts, int nblocks, int32_t* n_alive,
                                                            int t, RolloutState* st, int32_t* step_counts) {
  __shared__ int tmp[CBLK];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nblocks; b0 += CBLK) {
    const int i = b0 + threadIdx.x;
    const int v = i < nblocks ? block_counts[i] : 0;
    tmp[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < CBLK; o <<= 1) {            // Hillis-Steele inclusive scan
      int add = threadIdx.x >= o ? tmp[threadIdx.x - o] : 0;
      __syncthreads();
      tmp[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < nblocks) block_counts[i] = carry + tmp[threadIdx.x] - v;   // exclusive
    __syncthreads();
    if (threadIdx.x == 0) carry += tmp[CBLK - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int n_t = n_alive[t];
    step_counts[t] = n_t;
    st->base += n_t;
    n_alive[t + 1] = carry;
  }
}

static __global__ void __launch_bounds__(CBLK) compact_scatter_kernel(const uint8_t* __restrict__ done, const int32_t* n_dev,
                                                               const int32_t* __restrict__ block_offsets,
                                                               const float* __restrict__ ns, const int32_t* __restrict__ ids,
                                                               float* __restrict__ cur_next, int32_t* __restrict__ ids_next, int S) {
  __shared__ int warp_off[CBLK / 32];
  __shared__ int dst_row[CBLK];                 // destination row of every kept row of this block, -1 for dropped rows
  const int n = *n_dev;
  const int64_t r0 = (int64_t)blockIdx.x * CBLK;
  if (r0 >= n) return;
  const int64_t r = r0 + threadIdx.x;
  const bool keep = r < n && !done[r];
  const unsigned b = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) warp_off[w] = __popc(b);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = warp_off[threadIdx.x], incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int x = __shfl_up_sync(0xffffffffu, incl, o); if (threadIdx.x >= o) incl += x; }
    warp_off[threadIdx.x] = incl - v;
  }
  __syncthreads();
  const int pos = keep ? block_offsets[blockIdx.x] + warp_off[w] + __popc(b & ((1u << lane) - 1u)) : -1;
  dst_row[threadIdx.x] = pos;
  if (keep) ids_next[pos] = ids[r];
  __syncthreads();
  // element-parallel copy: consecutive threads read consecutive floats of the block's rows (coalesced) and write the kept
  // rows' floats, which are contiguous in the destination as well (order-preserving compaction)
  const int rows = (int)min((int64_t)CBLK, (int64_t)n - r0);
  if ((S & 3) == 0) {                                     // 16-byte rows: float4 copies
    const int S4 = S >> 2;
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

// err_flag (bf16 path): a non-zero watchdog flag means the fused kernel reported a pipeline time-out - the rows it wrote are garbage,
// so the ring pointer does NOT advance and the rollout reports zero transitions (the host raises at its next status check)
static __global__ void rollout_finish_kernel(int64_t* pointer, const RolloutState* st, int32_t* step_counts, int horizon,
                                             const int* err_flag = nullptr) {
  if (err_flag && *err_flag != 0) { step_counts[horizon] = 0; return; }
  int total = 0;
  for (int t = 0; t < horizon; ++t) total += step_counts[t];
  step_counts[horizon] = total;
  *pointer = st->base;
}

struct RolloutScratch {
  float *curA, *curB, *actions, *next_states, *rewards, *cv;
  uint8_t *done, *viol;
  int32_t *idsA, *idsB, *n_alive, *block_counts;
  RolloutState* st;
  EnsScratch ens; PolScratch pol;
};

static inline int64_t rollout_ws_bytes_fp32(const drpo_rollout_args& a) {
  const int64_t B = a.batch; const int S = a.ensemble->state_dim, A = a.ensemble->action_dim, C = a.env->con_dim;
  int64_t floats = 3 * B * S + B * A + B + B * C + ens_scratch_floats(*a.ensemble, B) + pol_scratch_floats(*a.actor, B);
  int64_t bytes = floats * 4 + 2 * B + 2 * B * 4 + (a.horizon + 2) * 4 + ((B + CBLK - 1) / CBLK + 1) * 4 + sizeof(RolloutState);
  return bytes + 32 * 256;
}

static inline int rollout_fp32(const drpo_rollout_args& a) {
  const int64_t B = a.batch; const int S = a.ensemble->state_dim, A = a.ensemble->action_dim, C = a.env->con_dim;
  const int H = a.horizon; void* stream = a.stream;
  Arena ar(a.workspace, a.workspace_bytes);
  RolloutScratch w;
  w.curA = ar.take<float>(B * S); w.curB = ar.take<float>(B * S); w.actions = ar.take<float>(B * A);
  w.next_states = ar.take<float>(B * S); w.rewards = ar.take<float>(B); w.cv = ar.take<float>(B * C);
  w.done = ar.take<uint8_t>(B); w.viol = ar.take<uint8_t>(B);
  w.idsA = ar.take<int32_t>(B); w.idsB = ar.take<int32_t>(B); w.n_alive = ar.take<int32_t>(H + 2);
  const int nblocks = (int)((B + CBLK - 1) / CBLK);
  w.block_counts = ar.take<int32_t>(nblocks + 1);
  w.st = ar.take<RolloutState>(1);
  w.ens = ens_scratch(ar, *a.ensemble, B); w.pol = pol_scratch(ar, *a.actor, B);
  if (!ar.ok()) { set_error("drpo_rollout: workspace too small (%lld needed, %lld given)", (long long)ar.off, (long long)a.workspace_bytes); return DRPO_ERR_WORKSPACE; }

  DRPO_CUDA_OK(cudaMemcpyAsync(w.curA, a.initial_states, sizeof(float) * B * S, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  DRPO_LAUNCH(rollout_init_kernel, grid_for(B), 256, 0, stream, w.idsA, B, a.traj_id_offset, w.n_alive, w.st, a.virt.pointer);
  float* cur = w.curA; float* nxt = w.curB; int32_t* ids = w.idsA; int32_t* ids_n = w.idsB;
  int rc;
  for (int t = 0; t < H; ++t) {
    const int* n_dev = w.n_alive + t;
    // policy.act(states, eval=False)                                   src/smbpo.py:236
    if ((rc = mlp3_fwd(*a.actor, cur, S, (int)B, ACT_RELU, w.pol.hA, w.pol.hB, w.pol.out, n_dev, stream))) return rc;
    NoiseView np_ = make_noise(a.eps_policy ? a.eps_policy + (int64_t)t * a.eps_batch_stride * A : nullptr, A, a.seed,
                               TAG_ROLLOUT_POLICY, (uint32_t)t);
    DRPO_LAUNCH(policy_head_kernel, grid_for(B), 256, 0, stream, w.pol.out, np_, ids, 0, w.actions, (float*)nullptr, B, A, n_dev);
    // model_ensemble.sample(states, actions)                           src/smbpo.py:237
    if ((rc = ens_member_raw(*a.ensemble, a.member_idx_host[t], cur, w.actions, (int)B, w.ens, n_dev, stream))) return rc;
    NoiseView nm = make_noise(a.eps_model ? a.eps_model + (int64_t)t * a.eps_batch_stride * (S + 1) : nullptr, S + 1, a.seed,
                              TAG_ROLLOUT_MODEL, (uint32_t)t);
    DRPO_LAUNCH(ens_sample_kernel, grid_for(B * (S + 1)), 256, 0, stream, w.ens.dd, w.ens.lr, cur, a.ensemble->min_log_var,
                a.ensemble->max_log_var, nm, ids, w.next_states, w.rewards, B, S, n_dev);
    // check_done / check_violation / get_constraint_value               src/smbpo.py:238-240
    DRPO_LAUNCH(hooks_kernel, grid_for(B), 256, 0, stream, *a.env, w.next_states, B, w.done, w.viol, w.cv, n_dev);
    // buffer.extend(...)                                                 src/smbpo.py:241-242,248
    DRPO_LAUNCH(rollout_store_kernel, grid_for(B * S), 256, 0, stream, a.virt, w.st, n_dev, cur, w.actions, w.next_states,
                w.rewards, w.done, w.viol, w.cv);
    // states = next_states[~dones]                                       src/smbpo.py:243-246
    DRPO_LAUNCH(compact_count_kernel, nblocks, CBLK, 0, stream, w.done, n_dev, w.block_counts);
    DRPO_LAUNCH(compact_scan_kernel, 1, CBLK, 0, stream, w.block_counts, nblocks, w.n_alive, t, w.st, a.step_counts);
    DRPO_LAUNCH(compact_scatter_kernel, nblocks, CBLK, 0, stream, w.done, n_dev, w.block_counts, w.next_states, ids, nxt, ids_n, S);
    float* tf = cur; cur = nxt; nxt = tf;
    int32_t* ti = ids; ids = ids_n; ids_n = ti;
  }
  DRPO_LAUNCH(rollout_finish_kernel, 1, 1, 0, stream, a.virt.pointer, w.st, a.step_counts, H);
  return DRPO_OK;
}

// SampleBuffer.sample + SMBPO.update_solver assembly (src/sampling.py:147-151, src/smbpo.py:253-270)
static __global__ void buffer_gather_kernel(drpo_buffer real, drpo_buffer virt, const int64_t* __restrict__ idx, int64_t n_real,
                                     int64_t n, float reward_scale, float alive_bonus, float c_scale, float c_off, drpo_batch out) {
  const int S = real.state_dim, A = real.action_dim, C = real.con_dim;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = t0; i < n * S; i += stride) {
    const int64_t r = i / S; const int c = (int)(i % S);
    const drpo_buffer& b = r < n_real ? real : virt;
    out.obs[i] = b.states[idx[r] * S + c];
    out.next_obs[i] = b.next_states[idx[r] * S + c];
  }
  for (int64_t i = t0; i < n * A; i += stride) {
    const int64_t r = i / A; const int c = (int)(i % A);
    const drpo_buffer& b = r < n_real ? real : virt;
    out.act[i] = b.actions[idx[r] * A + c];
  }
  for (int64_t i = t0; i < n * C; i += stride) {
    const int64_t r = i / C; const int c = (int)(i % C);
    const drpo_buffer& b = r < n_real ? real : virt;
    float v = __fmul_rn(b.constraint_values[idx[r] * C + c], c_scale);
    v = __fadd_rn(v, (v > 0.f ? 1.f : 0.f) * c_off);
    out.cv[i] = v;
  }
  for (int64_t r = t0; r < n; r += stride) {
    const drpo_buffer& b = r < n_real ? real : virt;
    float rw = b.rewards[idx[r]];
    if (reward_scale != 0.f) rw = __fmul_rn(rw, reward_scale);
    if (alive_bonus != 0.f) rw = __fadd_rn(rw, alive_bonus);
    out.rew[r] = rw; out.done[r] = b.dones[idx[r]]; out.viol[r] = b.violations[idx[r]];
  }
}

}  // namespace drpo
