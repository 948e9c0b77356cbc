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

}  // namespace drpo
