// Safety shield: action selection of SMBPO.step_generator (src/smbpo.py:124-136) and of
// sample_episodes_batched (src/sampling.py:420-439).  Latency-bound (1..n_envs rows): the candidates of the "linear" shield
// are independent rows, so the 11 constraint-critic evaluations of the reference's loop run as ONE Qc forward over 11*n rows.
#pragma once
#include "common.cuh"

namespace drpo {

constexpr int SHIELD_MIX = 11;                 // src/sampling.py:433: for i in range(11): ratio = (10 - i) / 10
struct ShieldRatios { float r[SHIELD_MIX], one_minus_r[SHIELD_MIX]; };
inline ShieldRatios shield_ratios() {
  ShieldRatios s;
  for (int i = 0; i < SHIELD_MIX; ++i) {        // Python evaluates both factors in double; torch multiplies by float(scalar)
    const double ratio = (10 - i) / 10.0;
    s.r[i] = (float)ratio; s.one_minus_r[i] = (float)(1.0 - ratio);
  }
  return s;
}

// candidate-major inputs of the batched Qc forward: row (i*n + r) = [s_r, a_safe_r*ratio_i + a_perf_r*(1-ratio_i)]  (:434-435).
// n_mix == 1 -> the single candidate is the performance action itself ("safe" shield / training-step shield).
static __global__ void shield_candidates_kernel(const float* __restrict__ s, const float* __restrict__ a_perf, const float* __restrict__ a_safe,
                                                ShieldRatios ratios, int n_mix, int64_t n, int S, int A, float* __restrict__ cand_s,
                                                float* __restrict__ cand_a) {
  const int64_t total = (int64_t)n_mix * n * (S + A);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / (S + A); const int c = (int)(i - row * (S + A));
    const int64_t m = row / n, r = row - m * n;
    if (c < S) cand_s[row * S + c] = s[r * S + c];
    else {
      const int j = c - S;
      const float ap = a_perf[r * A + j];
      // two rounded products and a rounded sum, as the three separate ATen ops do (no FMA contraction)
      cand_a[row * A + j] = n_mix == 1 ? ap : __fadd_rn(__fmul_rn(a_safe[r * A + j], ratios.r[m]), __fmul_rn(ap, ratios.one_minus_r[m]));
    }
  }
}

// _get_qc (src/ssac.py:588-600: torch.max over the constraint dims, NaN-propagating) + the selection rule.
//   n_mix == 1 : actions = where(qc > thr, a_safe, a_perf)                                    (:430 ; src/smbpo.py:134-135)
//   n_mix == 11: actions = a_safe; for i: actions = where(qc_i <= thr, mix_i, actions)        (:432-438)
static __global__ void shield_select_kernel(const float* __restrict__ q /* [n_mix*n, C] */, const float* __restrict__ cand_a,
                                            const float* __restrict__ a_safe, float thr, int n_mix, int64_t n, int A, int C,
                                            float* __restrict__ actions, float* __restrict__ qc_perf, int32_t* __restrict__ choice) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    int pick = -1; float q_last = 0.f;
    for (int m = 0; m < n_mix; ++m) {
      const float* qr = q + ((int64_t)m * n + r) * C;
      float v = qr[0];
      for (int c = 1; c < C; ++c) { const float x = qr[c]; if (x > v || x != x) v = x; }
      q_last = v;
      if (n_mix == 1) pick = (v > thr) ? -1 : 0;
      else if (v <= thr) pick = m;
    }
    const float* src = pick < 0 ? a_safe + r * A : cand_a + ((int64_t)pick * n + r) * A;
    for (int j = 0; j < A; ++j) actions[r * A + j] = src[j];
    if (qc_perf) qc_perf[r] = q_last;                      // the last candidate is the performance action (ratio 0)
    if (choice) choice[r] = n_mix == 1 ? (pick < 0 ? 1 : 0) : pick;
  }
}


// ---- latency path: two launches for up to SHIELD_FUSED_MAX_ROWS candidate rows ----------------------------------------------
// The shield runs on 1 state per training step and on the 10 evaluation envs: the work is ~1 MFLOP per row and the cost of the
// batched path above is its 16-18 launches.  Here one CTA walks a whole network for one row (thread j = output neuron j, the
// row's activations in shared memory, weights straight from L2), so a call is: one launch for both actors, one launch for the
// constraint critic of every candidate row, whose last CTA per state also applies the selection rule.
constexpr int SHIELD_FUSED_THREADS = 256;
constexpr int SHIELD_FUSED_MAX_ROWS = 256;

template <bool kRelu>
__device__ __forceinline__ void shield_dense(const drpo_linear& L, const float* __restrict__ x /* shared */, float* __restrict__ y /* shared */) {
  const int j = threadIdx.x;
  if (j < L.out_dim) {
    const float* w = L.w + (int64_t)j * L.in_dim;
    float acc = __ldg(L.b + j);
    if ((L.in_dim & 3) == 0 && (reinterpret_cast<uintptr_t>(L.w) & 15) == 0) {
      const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll 8
      for (int k = 0; k < (L.in_dim >> 2); ++k) {
        const float4 v = __ldg(w4 + k);
        acc = fmaf(v.x, x[4 * k], acc); acc = fmaf(v.y, x[4 * k + 1], acc); acc = fmaf(v.z, x[4 * k + 2], acc); acc = fmaf(v.w, x[4 * k + 3], acc);
      }
    } else {
#pragma unroll 4
      for (int k = 0; k < L.in_dim; ++k) acc = fmaf(__ldg(w + k), x[k], acc);
    }
    y[j] = kRelu ? (acc < 0.f ? 0.f : acc) : acc;
  }
  __syncthreads();
}

struct ShieldFusedArgs {
  drpo_mlp3 actor, actor_safe; drpo_qc qc;
  const float* states; int n, S, A, C, n_mix, eval_perf, uncertainty;
  float std_ratio, threshold; NoiseView noise; ShieldRatios ratios;
  float *a_perf, *a_safe, *qrow; int32_t* done_cnt;
  float* actions; float* qc_perf; int32_t* choice;
};

// blockIdx.x = state row, blockIdx.y = 0 performance actor (eval or sampled), 1 safe actor (eval)       src/policy.py:77-97
static __global__ void __launch_bounds__(SHIELD_FUSED_THREADS) shield_policy_kernel(const __grid_constant__ ShieldFusedArgs a) {
  __shared__ float bufA[SHIELD_FUSED_THREADS], bufB[SHIELD_FUSED_THREADS];
  const int r = blockIdx.x, which = blockIdx.y, tid = threadIdx.x;
  const drpo_mlp3& net = which ? a.actor_safe : a.actor;
  if (tid < a.S) bufA[tid] = a.states[(int64_t)r * a.S + tid];
  if (which == 0 && tid == 0 && a.done_cnt) a.done_cnt[r] = 0;
  __syncthreads();
  shield_dense<true>(net.l0, bufA, bufB);
  shield_dense<true>(net.l1, bufB, bufA);
  shield_dense<false>(net.l2, bufA, bufB);
  if (tid < a.A) {
    const float mu = bufB[tid], raw = bufB[a.A + tid];
    const float sd = expf(__fadd_rn(-6.f, __fmul_rn(10.f, sigmoid_f(raw))));
    const bool eval = which || a.eval_perf;
    const float x = eval ? mu : __fadd_rn(__fmul_rn(a.noise.get(r, tid), sd), mu);
    float* dst = which ? a.a_safe : (a.n_mix == 0 ? a.actions : a.a_perf);      // n_mix == 0: no shield, the action is the result
    dst[(int64_t)r * a.A + tid] = tanhf(x);
  }
}

// blockIdx.x = candidate row (m * n + r): Qc(s_r, mix_m) -> _get_qc; the last CTA of state r selects its action
static __global__ void __launch_bounds__(SHIELD_FUSED_THREADS) shield_qc_select_kernel(const __grid_constant__ ShieldFusedArgs a) {
  __shared__ float bufA[SHIELD_FUSED_THREADS], bufB[SHIELD_FUSED_THREADS], bufC[SHIELD_FUSED_THREADS], head[2 * DRPO_MAX_CON];
  __shared__ int is_last;
  const int m = blockIdx.x / a.n, r = blockIdx.x - m * a.n, tid = threadIdx.x;
  auto mix = [&](int mm, int j) {
    const float ap = a.a_perf[(int64_t)r * a.A + j];
    return a.n_mix == 1 ? ap : __fadd_rn(__fmul_rn(a.a_safe[(int64_t)r * a.A + j], a.ratios.r[mm]), __fmul_rn(ap, a.ratios.one_minus_r[mm]));
  };
  if (tid < a.S) bufA[tid] = a.states[(int64_t)r * a.S + tid];
  else if (tid < a.S + a.A) bufA[tid] = mix(m, tid - a.S);
  __syncthreads();
  shield_dense<true>(a.qc.trunk0, bufA, bufB);
  shield_dense<true>(a.qc.trunk1, bufB, bufC);              // shared hidden (both heads read it)
  shield_dense<true>(a.qc.mean0, bufC, bufA);
  shield_dense<false>(a.qc.mean1, bufA, head);
  if (a.uncertainty) {
    shield_dense<true>(a.qc.lstd0, bufC, bufA);
    shield_dense<false>(a.qc.lstd1, bufA, head + DRPO_MAX_CON);
  }
  if (tid == 0) {
    float v = 0.f;
    for (int c = 0; c < a.C; ++c) {
      float q = head[c];
      if (a.uncertainty) q = __fadd_rn(q, __fmul_rn(a.std_ratio, expf(soft_clamp(head[DRPO_MAX_CON + c], -4.f, 4.f))));      // src/ssac.py:74-85
      if (c == 0 || q > v || q != q) v = q;                   // torch.max over the constraint dims (NaN propagates)
    }
    a.qrow[blockIdx.x] = v;
    __threadfence();
    is_last = atomicAdd(a.done_cnt + r, 1) == a.n_mix - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (tid == 0) {
    int pick = -1; float q_last = 0.f;
    for (int mm = 0; mm < a.n_mix; ++mm) {
      const float v = __ldcg(a.qrow + (int64_t)mm * a.n + r);
      q_last = v;
      if (a.n_mix == 1) pick = (v > a.threshold) ? -1 : 0;
      else if (v <= a.threshold) pick = mm;
    }
    for (int j = 0; j < a.A; ++j) a.actions[(int64_t)r * a.A + j] = pick < 0 ? a.a_safe[(int64_t)r * a.A + j] : mix(pick, j);
    if (a.qc_perf) a.qc_perf[r] = q_last;
    if (a.choice) a.choice[r] = a.n_mix == 1 ? (pick < 0 ? 1 : 0) : pick;
  }
}

}  // namespace drpo
