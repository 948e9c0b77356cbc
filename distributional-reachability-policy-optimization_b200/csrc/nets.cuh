// fp32 (DRPO_PREC_FP32) forward passes of the hot path's networks: ensemble member, squashed-Gaussian actor,
// twin Q, distributional constraint critic.  Dense layers go through gemm_simt.cuh; everything between them is
// fused into one elementwise kernel per stage.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#include "hooks.cuh"

namespace drpo {

// ---------------------------------------------------------------------------------------------------------------
// elementwise stages
// ---------------------------------------------------------------------------------------------------------------

// x0 = [(s - mean)/(std + 1e-6), a]      src/normalization.py:23-24 + src/dynamics.py:113-114
static __global__ void ens_pack_kernel(const float* __restrict__ s, const float* __restrict__ a, const float* __restrict__ mean,
                                const float* __restrict__ stdv, float* __restrict__ x0, int64_t n, int S, int A,
                                const int* n_dev) {
  if (n_dev) n = min(n, (int64_t)*n_dev);
  const int D = S + A;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * D; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D; const int c = (int)(i % D);
    x0[i] = c < S ? __fdiv_rn(__fsub_rn(s[r * S + c], mean[c]), __fadd_rn(stdv[c], 1e-6f)) : a[r * A + (c - S)];
  }
}

// sa = [s, a]
static __global__ void cat2_kernel(const float* __restrict__ s, const float* __restrict__ a, float* __restrict__ out, int64_t n,
                            int S, int A) {
  const int D = S + A;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * D; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D; const int c = (int)(i % D);
    out[i] = c < S ? s[r * S + c] : a[r * A + (c - S)];
  }
}

// means = diffs + [s,0]; log_vars soft-clamped      src/dynamics.py:118-121
static __global__ void ens_head_kernel(const float* __restrict__ dd, const float* __restrict__ lr, const float* __restrict__ s,
                                const float* __restrict__ min_lv, const float* __restrict__ max_lv,
                                float* __restrict__ means, float* __restrict__ log_vars, int64_t n, int S) {
  const int O = S + 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * O; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / O; const int c = (int)(i % O);
    means[i] = __fadd_rn(dd[i], c < S ? s[r * S + c] : 0.f);
    log_vars[i] = soft_clamp(lr[i], min_lv[c], max_lv[c]);
  }
}

// samples = means + sqrt(exp(log_vars)) * eps ; split into next_states / rewards     src/dynamics.py:201-203
static __global__ void ens_sample_kernel(const float* __restrict__ dd, const float* __restrict__ lr, const float* __restrict__ s,
                                  const float* __restrict__ min_lv, const float* __restrict__ max_lv, NoiseView noise,
                                  const int32_t* __restrict__ row_ids, float* __restrict__ next_states,
                                  float* __restrict__ rewards, int64_t n, int S, const int* n_dev) {
  if (n_dev) n = min(n, (int64_t)*n_dev);
  const int O = S + 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * O; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / O; const int c = (int)(i % O);
    const float mean = __fadd_rn(dd[i], c < S ? s[r * S + c] : 0.f);
    const float lv = soft_clamp(lr[i], min_lv[c], max_lv[c]);
    const float sd = sqrtf(expf(lv));
    const int64_t id = row_ids ? row_ids[r] : r;
    const float y = __fadd_rn(mean, __fmul_rn(sd, noise.get(id, c)));
    if (c < S) next_states[r * S + c] = y; else rewards[r] = y;
  }
}

// squashed-Gaussian head: out[n,2A] -> action (+ log-prob)       src/policy.py:89-97, src/ssac.py:286-288
static __global__ void policy_head_kernel(const float* __restrict__ out, NoiseView noise, const int32_t* __restrict__ row_ids,
                                   int eval_mode, float* __restrict__ actions, float* __restrict__ log_prob, int64_t n,
                                   int A, const int* n_dev) {
  if (n_dev) n = min(n, (int64_t)*n_dev);
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    float lp = 0.f;
    const int64_t id = row_ids ? row_ids[r] : r;
    for (int j = 0; j < A; ++j) {
      const float mu = out[r * 2 * A + j], raw = out[r * 2 * A + A + j];
      const float log_std = __fadd_rn(-6.f, __fmul_rn(10.f, sigmoid_f(raw)));
      const float sd = expf(log_std);
      const float x = eval_mode ? mu : __fadd_rn(__fmul_rn(noise.get(id, j), sd), mu);
      actions[r * A + j] = tanhf(x);
      if (log_prob) {
        const float ladj = 2.f * (0.69314718055994531f - x - softplus_f(-2.f * x));
        const float d = x - mu;
        const float base = -(d * d) / (2.f * (sd * sd)) - logf(sd) - 0.91893853320467267f;
        lp += (0.f - ladj) + base;
      }
    }
    if (log_prob) log_prob[r] = lp;
  }
}

// constraint-critic head: raw mean [n,C], raw log-std [n,C] -> mean / std / shifted or sampled value
// mode 0: mean ; 1: mean + std_ratio*std (src/ssac.py:85) ; 2: mean, std, mean + clamp(eps,-2,2)*std (:88-90)
static __global__ void qc_head_kernel(const float* __restrict__ mean_raw, const float* __restrict__ ls_raw, int mode,
                               float std_ratio, NoiseView noise, int64_t row_off, float* __restrict__ out_mean,
                               float* __restrict__ out_std, float* __restrict__ out_sample, int64_t n, int C) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * C; i += (int64_t)gridDim.x * blockDim.x) {
    const float mu = mean_raw[i];
    if (out_mean) out_mean[i] = mu;
    if (mode == 0) continue;
    const float sd = expf(soft_clamp(ls_raw[i], -4.f, 4.f));
    if (out_std) out_std[i] = sd;
    if (mode == 1) { out_sample[i] = __fadd_rn(mu, __fmul_rn(std_ratio, sd)); continue; }
    const float e = fminf(fmaxf(noise.get(row_off + i / C, (int)(i % C)), -2.f), 2.f);
    out_sample[i] = __fadd_rn(mu, __fmul_rn(e, sd));
  }
}

// hooks over a batch of rows
static __global__ void hooks_kernel(drpo_env_params p, const float* __restrict__ states, int64_t n, uint8_t* __restrict__ done,
                             uint8_t* __restrict__ viol, float* __restrict__ cv, const int* n_dev) {
  if (n_dev) n = min(n, (int64_t)*n_dev);
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const float* row = states + r * p.state_dim;
    HookOut o;
    eval_hooks(p, [row](int d) { return row[d]; }, o);
    if (done) done[r] = o.done;
    if (viol) viol[r] = o.viol;
    if (cv) for (int c = 0; c < p.con_dim; ++c) cv[r * p.con_dim + c] = o.cv[c];
  }
}

static inline unsigned grid_for(int64_t work, int block = 256) {
  int64_t g = (work + block - 1) / block;
  if (g < 1) g = 1;
  if (g > 148 * 16) g = 148 * 16;
  return (unsigned)g;
}

// ---------------------------------------------------------------------------------------------------------------
// member views of the ensemble weights
// ---------------------------------------------------------------------------------------------------------------
struct MemberNet { drpo_linear t0, t1, d0, d1, l0, l1; };
static inline MemberNet member_of(const drpo_ensemble& e, int m) {
  const int S = e.state_dim, A = e.action_dim, H = e.hidden, D = S + A, O = S + 1;
  MemberNet n;
  n.t0 = {e.trunk0_w + (int64_t)m * H * D, e.trunk0_b + (int64_t)m * H, D, H};
  n.t1 = {e.trunk1_w + (int64_t)m * H * H, e.trunk1_b + (int64_t)m * H, H, H};
  n.d0 = {e.diff0_w + (int64_t)m * H * H, e.diff0_b + (int64_t)m * H, H, H};
  n.d1 = {e.diff1_w + (int64_t)m * O * H, e.diff1_b + (int64_t)m * O, H, O};
  n.l0 = {e.lvar0_w + (int64_t)m * H * H, e.lvar0_b + (int64_t)m * H, H, H};
  n.l1 = {e.lvar1_w + (int64_t)m * O * H, e.lvar1_b + (int64_t)m * O, H, O};
  return n;
}

struct EnsScratch { float *x0, *hA, *hB, *hC, *dd, *lr; };
static inline int64_t ens_scratch_floats(const drpo_ensemble& e, int64_t B) {
  return B * (e.state_dim + e.action_dim) + 3 * B * e.hidden + 2 * B * (e.state_dim + 1) + 6 * 64;
}
static inline EnsScratch ens_scratch(Arena& ar, const drpo_ensemble& e, int64_t B) {
  EnsScratch s;
  s.x0 = ar.take<float>(B * (e.state_dim + e.action_dim));
  s.hA = ar.take<float>(B * e.hidden); s.hB = ar.take<float>(B * e.hidden); s.hC = ar.take<float>(B * e.hidden);
  s.dd = ar.take<float>(B * (e.state_dim + 1)); s.lr = ar.take<float>(B * (e.state_dim + 1));
  return s;
}

// trunk + both heads of one member up to the raw head outputs dd / lr   (src/dynamics.py:112-119)
static inline int ens_member_raw(const drpo_ensemble& e, int member, const float* states, const float* actions, int B,
                                 const EnsScratch& w, const int* n_dev, void* stream) {
  const int S = e.state_dim, A = e.action_dim, H = e.hidden;
  MemberNet n = member_of(e, member);
  DRPO_LAUNCH(ens_pack_kernel, grid_for((int64_t)B * (S + A)), 256, 0, stream, states, actions, e.norm_mean, e.norm_std,
              w.x0, (int64_t)B, S, A, n_dev);
  int rc;
  if ((rc = linear_fwd(w.x0, S + A, n.t0, w.hA, H, B, ACT_SILU, n_dev, stream))) return rc;
  if ((rc = linear_fwd(w.hA, H, n.t1, w.hB, H, B, ACT_SILU, n_dev, stream))) return rc;
  if ((rc = linear_fwd(w.hB, H, n.d0, w.hC, H, B, ACT_SILU, n_dev, stream))) return rc;
  if ((rc = linear_fwd(w.hC, H, n.d1, w.dd, S + 1, B, ACT_NONE, n_dev, stream))) return rc;
  if ((rc = linear_fwd(w.hB, H, n.l0, w.hC, H, B, ACT_SILU, n_dev, stream))) return rc;
  if ((rc = linear_fwd(w.hC, H, n.l1, w.lr, S + 1, B, ACT_NONE, n_dev, stream))) return rc;
  return DRPO_OK;
}

// actor MLP -> raw head output [B,2A]
struct PolScratch { float *hA, *hB, *out; };
static inline int64_t pol_scratch_floats(const drpo_mlp3& a, int64_t B) {
  return 2 * B * a.l0.out_dim + B * a.l2.out_dim + 3 * 64;
}
static inline PolScratch pol_scratch(Arena& ar, const drpo_mlp3& a, int64_t B) {
  PolScratch s;
  s.hA = ar.take<float>(B * a.l0.out_dim); s.hB = ar.take<float>(B * a.l1.out_dim); s.out = ar.take<float>(B * a.l2.out_dim);
  return s;
}
static inline int mlp3_fwd(const drpo_mlp3& net, const float* x, int64_t ldx, int B, int act, float* hA, float* hB, float* out,
                           const int* n_dev, void* stream) {
  int rc;
  if ((rc = linear_fwd(x, ldx, net.l0, hA, net.l0.out_dim, B, act, n_dev, stream))) return rc;
  if ((rc = linear_fwd(hA, net.l0.out_dim, net.l1, hB, net.l1.out_dim, B, act, n_dev, stream))) return rc;
  return linear_fwd(hB, net.l1.out_dim, net.l2, out, net.l2.out_dim, B, ACT_NONE, n_dev, stream);
}

// constraint critic forward keeping every activation (needed by the backward pass)
struct QcActs { float *t1, *t2, *m1, *l1, *mean_raw, *ls_raw; };
static inline int qc_fwd(const drpo_qc& q, const float* sa, int ldsa, int B, const QcActs& a, bool need_std, void* stream) {
  int rc; const int H = q.trunk0.out_dim;
  if ((rc = linear_fwd(sa, ldsa, q.trunk0, a.t1, H, B, ACT_RELU, nullptr, stream))) return rc;
  if ((rc = linear_fwd(a.t1, H, q.trunk1, a.t2, H, B, ACT_RELU, nullptr, stream))) return rc;
  if ((rc = linear_fwd(a.t2, H, q.mean0, a.m1, H, B, ACT_RELU, nullptr, stream))) return rc;
  if ((rc = linear_fwd(a.m1, H, q.mean1, a.mean_raw, q.mean1.out_dim, B, ACT_NONE, nullptr, stream))) return rc;
  if (!need_std) return DRPO_OK;
  if ((rc = linear_fwd(a.t2, H, q.lstd0, a.l1, H, B, ACT_RELU, nullptr, stream))) return rc;
  return linear_fwd(a.l1, H, q.lstd1, a.ls_raw, q.lstd1.out_dim, B, ACT_NONE, nullptr, stream);
}

}  // namespace drpo
