// SSAC.update_actor_and_alpha (src/ssac.py:458-527) in DRPO mode (reachability, distributional_qc, mlp_multiplier,
// autotune_alpha): performance actor (alpha*log_prob - Q_k + lambda*Qc_ub), temperature alpha, safe actor (Qc_ub).
// Forward of every pass, one fused per-row loss / output-gradient kernel, hand-written backward: dX chains through the
// FROZEN Q_k and Qc nets down to the action columns, the squashed-Gaussian rsample / log-prob backward, then the actor's
// own dX chain + split-K dW (mlp3_bwd); two grad-norm clips, three Adam steps (the alpha one on a device scalar).
// Dense layers go through gemm_simt.cuh (fp32 FFMA, or TF32 tensor-op GEMMs when the tensor mode is selected).
#pragma once
#include "critic.cuh"

namespace drpo {

// xaug = [obs, max_c(mean + ratio*std)]  (the multiplier's input, src/ssac.py:107-108 with _get_qc :588-600)
static __global__ void actor_xaug_kernel(const float* __restrict__ obs, const float* __restrict__ mean_raw, const float* __restrict__ ls_raw,
                                         float ratio, float* __restrict__ xaug, int64_t B, int S, int C) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x) {
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, __fadd_rn(mean_raw[r * C + c], __fmul_rn(ratio, expf(soft_clamp(ls_raw[r * C + c], -4.f, 4.f)))));
    for (int c = 0; c < S; ++c) xaug[r * (S + 1) + c] = obs[r * S + c];
    xaug[r * (S + 1) + S] = m;
  }
}

struct ActorLossArgs {
  const float *logp, *q;                         // actor log-prob, Q_k(obs, a)
  const float *mean1, *ls1;                      // Qc heads at (obs, a)        [B,C]
  const float *lam_raw;                          // multiplier net output       [B]
  const float *mean2, *ls2;                      // Qc heads at (obs, a_safe')  [B,C]
  const float* log_alpha;
  float ratio, ub, inv_bg, target_entropy;
  float *dq, *dmean1, *dls1, *dmean2, *dls2;     // output gradients
  double* partials;                              // [grid,4]: actor loss sum, safe loss sum, sum(log_prob + target_entropy), -
  int64_t B; int C;
};
// upper bound mean + ratio*exp(soft_clamp(ls)) of every constraint, its max (first index on ties, like torch.max) and the
// gradient of the max w.r.t. the raw head outputs scaled by `w`
__device__ __forceinline__ float qc_ub_max_grad(const float* mean, const float* ls, int C, float ratio, float w, float* dmean, float* dls) {
  float best = -INFINITY; int bi = 0; float bsd = 0.f, bx = 0.f;
  for (int c = 0; c < C; ++c) {
    const float x = ls[c];
    const float sd = expf(soft_clamp(x, -4.f, 4.f));
    const float v = __fadd_rn(mean[c], __fmul_rn(ratio, sd));
    if (v > best) { best = v; bi = c; bsd = sd; bx = x; }
    dmean[c] = 0.f; dls[c] = 0.f;
  }
  const float y1 = 4.f - softplus_f(4.f - bx);
  dmean[bi] = w;
  dls[bi] = w * ratio * bsd * dsoftplus(y1 + 4.f) * dsoftplus(4.f - bx);
  return best;
}
static __global__ void __launch_bounds__(256) actor_loss_kernel(ActorLossArgs a) {
  double la = 0.0, ls = 0.0, lm = 0.0;
  const float alpha = expf(*a.log_alpha);
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.B; r += (int64_t)gridDim.x * blockDim.x) {
    const float th = tanhf(a.lam_raw[r] / a.ub * 2.f);
    const float lam = a.ub / 2.f * (1.f + th);                                   // src/ssac.py:109-110 (no gradient: detached)
    const float qc1 = qc_ub_max_grad(a.mean1 + r * a.C, a.ls1 + r * a.C, a.C, a.ratio, lam * a.inv_bg, a.dmean1 + r * a.C, a.dls1 + r * a.C);
    const float qc2 = qc_ub_max_grad(a.mean2 + r * a.C, a.ls2 + r * a.C, a.C, a.ratio, a.inv_bg, a.dmean2 + r * a.C, a.dls2 + r * a.C);
    a.dq[r] = -a.inv_bg;
    la += (double)(alpha * a.logp[r] - a.q[r]) + (double)(lam * qc1);
    ls += (double)qc2;
    lm += (double)(a.logp[r] + a.target_entropy);
  }
  la = warp_sum_d(la); ls = warp_sum_d(ls); lm = warp_sum_d(lm);
  __shared__ double s0[8], s1[8], s2[8];
  if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = la; s1[threadIdx.x >> 5] = ls; s2[threadIdx.x >> 5] = lm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0, t2 = 0;
    for (int w = 0; w < 8; ++w) { t0 += s0[w]; t1 += s1[w]; t2 += s2[w]; }
    a.partials[4 * blockIdx.x] = t0; a.partials[4 * blockIdx.x + 1] = t1; a.partials[4 * blockIdx.x + 2] = t2; a.partials[4 * blockIdx.x + 3] = 0;
  }
}
// losses[0] = actor loss, [1] = alpha loss = -alpha*M, [2] = safe-actor loss, [5] = d(alpha loss)/d(log_alpha) = -alpha*M
// (M = mean(log_prob + target_entropy); use_log_alpha_loss = False: the coefficient is alpha = exp(log_alpha))
static __global__ void actor_loss_finalize_kernel(const double* partials, int nblocks, double inv_bg, const float* log_alpha, float* losses) {
  double s[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double v = 0;
    for (int b = threadIdx.x; b < nblocks; b += 32) v += partials[4 * b + k];
    s[k] = warp_sum_d(v);
  }
  if (threadIdx.x == 0) {
    const float alpha = expf(*log_alpha);
    const float M = (float)(s[2] * inv_bg);
    losses[0] = (float)(s[0] * inv_bg); losses[1] = -alpha * M; losses[2] = (float)(s[1] * inv_bg);
    losses[5] = -alpha * M;
  }
}

// Backward of the squashed-Gaussian rsample (+ log-prob) w.r.t. the policy net output [mu | raw log-std]:
//   x = mu + sd*eps, a = tanh(x), sd = exp(-6 + 10*sigmoid(raw));  log_prob = sum_j(-eps^2/2 - log sd - c - 2(log 2 - x - softplus(-2x)))
//   => d log_prob/dx = 2 tanh(x), d log_prob/d log sd = -1 (+ through x)
// da = dL/da (from the critics), w_logp = dL/d log_prob (alpha/B for the performance actor, 0 for the safe actor)
static __global__ void policy_rsample_bwd_kernel(const float* __restrict__ pout, NoiseView noise, const float* __restrict__ dsa, int ldsa, int S,
                                                 const float* log_alpha, float w_logp_scale, float* __restrict__ dpout, int64_t B, int A) {
  const float wl = w_logp_scale != 0.f ? expf(*log_alpha) * w_logp_scale : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * A; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / A; const int j = (int)(i - r * A);
    const float mu = pout[r * 2 * A + j], raw = pout[r * 2 * A + A + j];
    const float sg = sigmoid_f(raw);
    const float sd = expf(__fadd_rn(-6.f, __fmul_rn(10.f, sg)));
    const float eps = noise.get(r, j);
    const float x = __fadd_rn(__fmul_rn(eps, sd), mu);
    const float t = tanhf(x);
    const float dx = dsa[r * ldsa + S + j] * (1.f - t * t) + wl * 2.f * t;
    dpout[r * 2 * A + j] = dx;
    dpout[r * 2 * A + A + j] = (dx * sd * eps - wl) * 10.f * sg * (1.f - sg);
  }
}

// Adam on the device scalar log_alpha (torch.optim.Adam, no weight decay), gradient in losses[5]
static __global__ void alpha_adam_kernel(float* log_alpha, const float* grad, float* m, float* v, AdamScalars s) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float g = *grad;
    const float mm = *m + s.one_minus_b1 * (g - *m);                   // lerp
    const float vv = s.b2 * *v + s.one_minus_b2 * g * g;
    *m = mm; *v = vv;
    *log_alpha = *log_alpha + s.neg_step_size * (mm / (sqrtf(vv) / s.bc2_sqrt + s.eps));
  }
}

// dX chain through the frozen constraint critic: (dmean, dls) [B,C] -> d[s,a] accumulated into dsa (beta as given)
static inline int qc_bwd_input(const drpo_qc& q, const QcActs& acts, const float* dmean, const float* dls, int B, int C, int H, int D,
                               float* dhA, float* dt2, float* dsa, float beta_dsa, void* stream) {
  int rc;
  if ((rc = linear_bwd_data(dmean, C, q.mean1, dhA, H, B, acts.m1, H, 1, 0.f, stream))) return rc;
  if ((rc = linear_bwd_data(dhA, H, q.mean0, dt2, H, B, acts.t2, H, 1, 0.f, stream))) return rc;
  if ((rc = linear_bwd_data(dls, C, q.lstd1, dhA, H, B, acts.l1, H, 1, 0.f, stream))) return rc;
  if ((rc = linear_bwd_data(dhA, H, q.lstd0, dt2, H, B, acts.t2, H, 1, 1.f, stream))) return rc;
  if ((rc = linear_bwd_data(dt2, H, q.trunk1, dhA, H, B, acts.t1, H, 1, 0.f, stream))) return rc;
  return linear_bwd_data(dhA, H, q.trunk0, dsa, D, B, nullptr, 0, 0, beta_dsa, stream);
}

static inline int64_t actor_ws_bytes(int64_t B, int S, int A, int C, int H) {
  const int64_t D = S + A;
  int64_t f = 3 * B * D + 3 * B * A + B                 // sa1, sa_e, sa2, a1, a_e, a2, logp
            + 4 * B * H + 2 * B * 2 * A                 // actor / safe hidden + outputs
            + 2 * B * H + B                             // Q hidden + value
            + 3 * (4 * B * H + 2 * B * C)               // three Qc passes
            + B * (S + 1) + 2 * B * H + B               // multiplier input, hidden, raw
            + B + 4 * B * C                             // dq, dmean1, dls1, dmean2, dls2
            + 3 * B * H + B * D + 2 * B * 2 * A         // dhA, dhB, dt2, dsa, dpout x2
            + PARTIAL_FLOATS;
  return f * 4 + (LOSS_BLOCKS * 4 + 2 * 1184 + 8) * 8 + 64 * 256 + LT_WORKSPACE_BYTES;
}

static inline int actor_step_fp32(const drpo_actor_args& a) {
  const int64_t B = a.batch_size; const int S = a.state_dim, A = a.action_dim, C = a.con_dim, D = S + A;
  const int H = a.actor.l0.out_dim; void* stream = a.stream; int rc;
  Arena ar(a.workspace, a.workspace_bytes);
  float* sa1 = ar.take<float>(B * D); float* sae = ar.take<float>(B * D); float* sa2 = ar.take<float>(B * D);
  float* a1 = ar.take<float>(B * A); float* ae = ar.take<float>(B * A); float* a2 = ar.take<float>(B * A); float* logp = ar.take<float>(B);
  float* ahA = ar.take<float>(B * H); float* ahB = ar.take<float>(B * H); float* apout = ar.take<float>(B * 2 * A);
  float* shA = ar.take<float>(B * H); float* shB = ar.take<float>(B * H); float* spout = ar.take<float>(B * 2 * A);
  float* qh1 = ar.take<float>(B * H); float* qh2 = ar.take<float>(B * H); float* qv = ar.take<float>(B);
  QcActs c1, ce, c2;
  for (QcActs* c : {&c1, &ce, &c2}) {
    c->t1 = ar.take<float>(B * H); c->t2 = ar.take<float>(B * H); c->m1 = ar.take<float>(B * H); c->l1 = ar.take<float>(B * H);
    c->mean_raw = ar.take<float>(B * C); c->ls_raw = ar.take<float>(B * C);
  }
  float* xaug = ar.take<float>(B * (S + 1)); float* lh1 = ar.take<float>(B * H); float* lh2 = ar.take<float>(B * H); float* lraw = ar.take<float>(B);
  float* dq = ar.take<float>(B); float* dmean1 = ar.take<float>(B * C); float* dls1 = ar.take<float>(B * C);
  float* dmean2 = ar.take<float>(B * C); float* dls2 = ar.take<float>(B * C);
  float* dhA = ar.take<float>(B * H); float* dhB = ar.take<float>(B * H); float* dt2 = ar.take<float>(B * H);
  float* dsa = ar.take<float>(B * D); float* dp1 = ar.take<float>(B * 2 * A); float* dp2 = ar.take<float>(B * 2 * A);
  float* partial = ar.take<float>(PARTIAL_FLOATS);
  double* loss_part = ar.take<double>(LOSS_BLOCKS * 4); double* nrm_part = ar.take<double>(2 * 1184);
  float* coef = ar.take<float>(4);
  g_lt_workspace = ar.take<char>(LT_WORKSPACE_BYTES); g_lt_workspace_bytes = (size_t)LT_WORKSPACE_BYTES;
  if (!ar.ok()) { set_error("drpo_actor_step: workspace too small (%lld needed, %lld given)", (long long)ar.off, (long long)a.workspace_bytes); return DRPO_ERR_WORKSPACE; }
  NoiseView none = make_noise(nullptr, 0, 0, 0, 0);
  NoiseView n1 = make_noise(a.eps_actor, A, a.seed, TAG_ACTOR_ACTOR, a.noise_step, a.row_id_offset);
  NoiseView n2 = make_noise(a.eps_safe, A, a.seed, TAG_ACTOR_SAFE, a.noise_step, a.row_id_offset);
  const float inv_bg = (float)(1.0 / (double)a.global_batch_size);

  if (a.phases & 1) {
    DRPO_CUDA_OK(cudaMemsetAsync(a.losses + DRPO_LOSS_ERR_SLOT, 0, sizeof(float), (cudaStream_t)stream));   // watchdog slot: no fused kernel on this path
    // ---- forward ------------------------------------------------------------------------------------------------------
    // action = actor.distr(obs).rsample(), log_prob                                  src/ssac.py:459-461
    if ((rc = mlp3_fwd(a.actor, a.obs, S, (int)B, ACT_RELU, ahA, ahB, apout, nullptr, stream))) return rc;
    DRPO_LAUNCH(policy_head_kernel, grid_for(B), 256, 0, stream, apout, n1, (const int32_t*)nullptr, 0, a1, logp, B, A, (const int*)nullptr);
    DRPO_LAUNCH(cat2_kernel, grid_for(B * D), 256, 0, stream, a.obs, a1, sa1, B, S, A);
    // actor_Q = critic.random_choice(obs, action)                                    src/ssac.py:462 ; :41-43
    if ((rc = mlp3_fwd(*a.q, sa1, D, (int)B, ACT_RELU, qh1, qh2, qv, nullptr, stream))) return rc;
    // actor_Qc = max_c constraint_critic(obs, action, uncertainty=True)              src/ssac.py:468-469 ; :85
    if ((rc = qc_fwd(*a.qc, sa1, D, (int)B, c1, true, stream))) return rc;
    // no grad: action_safe = actor_safe.act(obs, eval=True); safe_Qc; lams = multiplier(obs, safe_Qc)     src/ssac.py:473-478
    if ((rc = mlp3_fwd(a.actor_safe, a.obs, S, (int)B, ACT_RELU, shA, shB, spout, nullptr, stream))) return rc;
    DRPO_LAUNCH(policy_head_kernel, grid_for(B), 256, 0, stream, spout, none, (const int32_t*)nullptr, 1, ae, (float*)nullptr, B, A, (const int*)nullptr);
    DRPO_LAUNCH(cat2_kernel, grid_for(B * D), 256, 0, stream, a.obs, ae, sae, B, S, A);
    if ((rc = qc_fwd(*a.qc, sae, D, (int)B, ce, true, stream))) return rc;
    DRPO_LAUNCH(actor_xaug_kernel, grid_for(B), 256, 0, stream, a.obs, ce.mean_raw, ce.ls_raw, (float)a.std_ratio, xaug, B, S, C);
    if ((rc = mlp3_fwd(*a.lam, xaug, S + 1, (int)B, ACT_TANH, lh1, lh2, lraw, nullptr, stream))) return rc;
    // safe actor: action_safe' = actor_safe.distr(obs).rsample() (same net output as the eval action)     src/ssac.py:488-492
    DRPO_LAUNCH(policy_head_kernel, grid_for(B), 256, 0, stream, spout, n2, (const int32_t*)nullptr, 0, a2, (float*)nullptr, B, A, (const int*)nullptr);
    DRPO_LAUNCH(cat2_kernel, grid_for(B * D), 256, 0, stream, a.obs, a2, sa2, B, S, A);
    if ((rc = qc_fwd(*a.qc, sa2, D, (int)B, c2, true, stream))) return rc;
    // ---- losses and output gradients --------------------------------------------------------------------------------------
    ActorLossArgs L;
    L.logp = logp; L.q = qv; L.mean1 = c1.mean_raw; L.ls1 = c1.ls_raw; L.lam_raw = lraw; L.mean2 = c2.mean_raw; L.ls2 = c2.ls_raw;
    L.log_alpha = a.log_alpha; L.ratio = (float)a.std_ratio; L.ub = (float)a.multiplier_ub; L.inv_bg = inv_bg;
    L.target_entropy = (float)a.target_entropy;
    L.dq = dq; L.dmean1 = dmean1; L.dls1 = dls1; L.dmean2 = dmean2; L.dls2 = dls2; L.partials = loss_part; L.B = B; L.C = C;
    DRPO_LAUNCH(actor_loss_kernel, LOSS_BLOCKS, 256, 0, stream, L);
    DRPO_LAUNCH(actor_loss_finalize_kernel, 1, 32, 0, stream, loss_part, LOSS_BLOCKS, 1.0 / (double)a.global_batch_size, a.log_alpha, a.losses);
    // ---- backward: performance actor ----------------------------------------------------------------------------------------
    //   d[s,a] through Q_k ...
    if ((rc = linear_bwd_data(dq, 1, a.q->l2, dhA, H, (int)B, qh2, H, 1, 0.f, stream))) return rc;
    if ((rc = linear_bwd_data(dhA, H, a.q->l1, dhB, H, (int)B, qh1, H, 1, 0.f, stream))) return rc;
    if ((rc = linear_bwd_data(dhB, H, a.q->l0, dsa, D, (int)B, nullptr, 0, 0, 0.f, stream))) return rc;
    //   ... plus through Qc (both heads, trunk)
    if ((rc = qc_bwd_input(*a.qc, c1, dmean1, dls1, (int)B, C, H, D, dhA, dt2, dsa, 1.f, stream))) return rc;
    //   tanh / rsample / log-prob, then the actor's own layers
    DRPO_LAUNCH(policy_rsample_bwd_kernel, grid_for(B * A), 256, 0, stream, apout, n1, dsa, D, S, a.log_alpha, inv_bg, dp1, B, A);
    {
      Mlp3Grads g = mlp3_grads(a.actor, a.params_actor, a.grads_actor);
      if ((rc = mlp3_bwd(a.actor, g, a.obs, S, ahA, ahB, dp1, (int)B, 1, dhA, dhB, partial, PARTIAL_FLOATS, stream))) return rc;
    }
    // ---- backward: safe actor ---------------------------------------------------------------------------------------------
    if ((rc = qc_bwd_input(*a.qc, c2, dmean2, dls2, (int)B, C, H, D, dhA, dt2, dsa, 0.f, stream))) return rc;
    DRPO_LAUNCH(policy_rsample_bwd_kernel, grid_for(B * A), 256, 0, stream, spout, n2, dsa, D, S, a.log_alpha, 0.f, dp2, B, A);
    {
      Mlp3Grads g = mlp3_grads(a.actor_safe, a.params_safe, a.grads_safe);
      if ((rc = mlp3_bwd(a.actor_safe, g, a.obs, S, shA, shB, dp2, (int)B, 1, dhA, dhB, partial, PARTIAL_FLOATS, stream))) return rc;
    }
  }
  if (a.phases & 2) {
    // clip_grad_norm_ (actor, actor_safe), three Adam steps                           src/ssac.py:516-527
    const int nb = 592;
    DRPO_LAUNCH(sumsq_kernel, nb, 256, 0, stream, a.grads_actor, a.n_actor, (int64_t)0, nrm_part);
    DRPO_LAUNCH(clip_coef_kernel, 1, 32, 0, stream, nrm_part, nb, (float)a.grad_norm, a.losses + 3, coef);
    AdamScalars s = adam_scalars(a.adam_actor, 0.0);
    DRPO_LAUNCH(adam_ema_kernel, grid_for(a.n_actor), 256, 0, stream, a.params_actor, a.grads_actor, a.m_actor, a.v_actor, (float*)nullptr,
                a.n_actor, a.n_actor, coef, s, (const float*)(a.losses + DRPO_LOSS_ERR_SLOT));
    DRPO_LAUNCH(alpha_adam_kernel, 1, 32, 0, stream, a.log_alpha, a.losses + 5, a.alpha_m, a.alpha_v, adam_scalars(a.adam_alpha, 0.0));
    DRPO_LAUNCH(sumsq_kernel, nb, 256, 0, stream, a.grads_safe, a.n_safe, (int64_t)0, nrm_part);
    DRPO_LAUNCH(clip_coef_kernel, 1, 32, 0, stream, nrm_part, nb, (float)a.grad_norm, a.losses + 6, coef + 2);
    AdamScalars s2 = adam_scalars(a.adam_safe, 0.0);
    DRPO_LAUNCH(adam_ema_kernel, grid_for(a.n_safe), 256, 0, stream, a.params_safe, a.grads_safe, a.m_safe, a.v_safe, (float*)nullptr,
                a.n_safe, a.n_safe, coef + 2, s2, (const float*)(a.losses + DRPO_LOSS_ERR_SLOT));
  }
  return DRPO_OK;
}

}  // namespace drpo
