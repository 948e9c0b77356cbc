// SSAC.update_critic (src/ssac.py:437-456) and SSAC.update_multiplier (:570-578), fp32 path:
// forward of every pass, fused target/loss/output-gradient kernel, hand-written backward (dX chain + split-K dW),
// grad-norm clips with warp-shuffle reductions, and one fused clip+L2+Adam+EMA kernel over the flat arenas.
#pragma once
#include "nets.cuh"

namespace drpo {

// ---------------------------------------------------------------------------------------------------------------
// per-row target / loss / dLoss kernel
// ---------------------------------------------------------------------------------------------------------------
struct CriticLossArgs {
  const float *rew, *cv; const uint8_t* done;                    // batch
  const float *logp, *q1t, *q2t;                                 // actor(next_obs) log-prob, target Q's
  const float *nqc;                                              // target Qc sample  mu' + clamp(eps)*sigma'   [B,C]
  const float *q1, *q2;                                          // online Q's
  const float *mean_raw, *ls_raw;                                // online Qc heads [B,C]
  const float* log_alpha;
  float gamma, one_minus_gamma, td_bound, inv_bg, inv_bgc;
  float *dq1, *dq2, *dmean, *dls;                                // output gradients
  double* partials;                                              // [grid,2] loss partial sums
  int64_t B; int C;
};

__device__ __forceinline__ float dsoftplus(float x) { return x > 20.f ? 1.f : sigmoid_f(x); }

static __global__ void __launch_bounds__(256) critic_loss_kernel(CriticLossArgs a) {
  double lq = 0.0, lc = 0.0;
  const float alpha = expf(*a.log_alpha);
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.B; r += (int64_t)gridDim.x * blockDim.x) {
    const float d = a.done[r] ? 1.f : 0.f;
    // compute_target                                                       src/ssac.py:284-294
    const float nv = fminf(a.q1t[r], a.q2t[r]) - alpha * a.logp[r];
    const float q = a.rew[r] + a.gamma * (1.f - d) * nv;
    const float e1 = a.q1[r] - q, e2 = a.q2[r] - q;
    a.dq1[r] = e1 * a.inv_bg; a.dq2[r] = e2 * a.inv_bg;
    lq += 0.5 * ((double)e1 * e1 + (double)e2 * e2);
    for (int c = 0; c < a.C; ++c) {
      const int64_t i = r * a.C + c;
      // compute_cons_target                                                src/ssac.py:345-354
      const float h = a.cv[i], mu = a.mean_raw[i];
      const float nonterm = a.one_minus_gamma * h + a.gamma * fmaxf(h, a.nqc[i]);
      const float tu = nonterm * (1.f - d) + h * d;
      const float tb = fminf(fmaxf(tu - mu, -a.td_bound), a.td_bound) + mu;
      // cons_critic_loss_given_target                                      src/ssac.py:416-423
      const float x = a.ls_raw[i];
      const float y1 = 4.f - softplus_f(4.f - x);
      const float ls = -4.f + softplus_f(y1 + 4.f);
      const float sd = expf(ls), var = sd * sd;
      const float du = mu - tu, db = mu - tb;
      lc += (double)(du * du / (2.f * var) + db * db / (2.f * var) + logf(sd));
      a.dmean[i] = du / var * a.inv_bgc;
      a.dls[i] = (1.f - db * db / var) * a.inv_bgc * dsoftplus(y1 + 4.f) * dsoftplus(4.f - x);
    }
  }
  lq = warp_sum_d(lq); lc = warp_sum_d(lc);
  __shared__ double sq[8], sc[8];
  if ((threadIdx.x & 31) == 0) { sq[threadIdx.x >> 5] = lq; sc[threadIdx.x >> 5] = lc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tq = 0, tc = 0;
    for (int w = 0; w < 8; ++w) { tq += sq[w]; tc += sc[w]; }
    a.partials[2 * blockIdx.x] = tq; a.partials[2 * blockIdx.x + 1] = tc;
  }
}

static __global__ void loss_finalize_kernel(const double* partials, int nblocks, int ncols, const double* scale, float* out) {
  if (threadIdx.x < ncols) {
    double s = 0;
    for (int b = 0; b < nblocks; ++b) s += partials[ncols * b + threadIdx.x];
    out[threadIdx.x] = (float)(s * scale[threadIdx.x]);
  }
}
struct Scale2 { double v[2]; };
static __global__ void loss_finalize2_kernel(const double* partials, int nblocks, Scale2 sc, float* out) {   // one warp
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    double s = 0;
    for (int b = threadIdx.x; b < nblocks; b += 32) s += partials[2 * b + k];
    s = warp_sum_d(s);
    if (threadIdx.x == 0) out[k] = (float)(s * sc.v[k]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// optimiser: grad norms (two clip groups), fused clip + coupled-L2 + Adam + EMA
// ---------------------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n0, int64_t n1, double* partials) {
  // partials[2*block + k] = sum of squares of group k handled by this block (group 0 = [0,n0), group 1 = [n0,n0+n1))
  double s0 = 0, s1 = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = g[i];
    if (i < n0) s0 += v * v; else s1 += v * v;
  }
  s0 = warp_sum_d(s0); s1 = warp_sum_d(s1);
  __shared__ double a0[8], a1[8];
  if ((threadIdx.x & 31) == 0) { a0[threadIdx.x >> 5] = s0; a1[threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0;
    for (int w = 0; w < 8; ++w) { t0 += a0[w]; t1 += a1[w]; }
    partials[2 * blockIdx.x] = t0; partials[2 * blockIdx.x + 1] = t1;
  }
}
// norms[k] = sqrt(sum), coef[k] = min(1, max_norm/(norm+1e-6))        torch.nn.utils.clip_grad_norm_
static __global__ void clip_coef_kernel(const double* partials, int nblocks, float max_norm, float* norms_out, float* coef) {   // one warp
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    double s = 0;
    for (int b = threadIdx.x; b < nblocks; b += 32) s += partials[2 * b + k];
    s = warp_sum_d(s);
    if (threadIdx.x == 0) {
      const float nrm = (float)sqrt(s);
      if (norms_out) norms_out[k] = nrm;
      coef[k] = fminf(max_norm / (nrm + 1e-6f), 1.f);
    }
  }
}

struct AdamScalars { float wd, one_minus_b1, b2, one_minus_b2, bc2_sqrt, eps, neg_step_size, tau, one_minus_tau; };
static inline AdamScalars adam_scalars(const drpo_adam& a, double tau) {
  AdamScalars s;
  const double bc1 = 1.0 - pow(a.beta1, a.step), bc2 = 1.0 - pow(a.beta2, a.step);
  s.wd = (float)a.weight_decay; s.one_minus_b1 = (float)(1.0 - a.beta1); s.b2 = (float)a.beta2;
  s.one_minus_b2 = (float)(1.0 - a.beta2); s.bc2_sqrt = (float)sqrt(bc2); s.eps = (float)a.eps;
  s.neg_step_size = (float)(-(a.lr / bc1)); s.tau = (float)tau; s.one_minus_tau = (float)(1.0 - tau);
  return s;
}
// torch.optim.Adam (coupled L2, src/ssac.py:199-203) + update_ema (src/torch_util.py:223-226) in one pass
static __global__ void __launch_bounds__(256) adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, float* __restrict__ tgt, int64_t n0, int64_t n,
                                                       const float* __restrict__ coef, AdamScalars s, const float* __restrict__ skip = nullptr) {
  // `skip` = the update's watchdog slot (losses[DRPO_LOSS_ERR_SLOT], summed over ranks by the gradient all-reduce): non-zero means a
  // fused kernel of phase 1 reported a pipeline time-out on some rank - the gradients are garbage and NO rank may apply them
  if (skip && *skip != 0.f) return;
  const float c0 = coef[0], c1 = coef[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = __fmul_rn(g[i], i < n0 ? c0 : c1);
    gi = __fadd_rn(gi, __fmul_rn(s.wd, pi));                                  // grad.add(param, alpha=wd)
    float mi = m[i]; mi = __fadd_rn(mi, __fmul_rn(s.one_minus_b1, __fsub_rn(gi, mi)));   // lerp_
    float vi = __fadd_rn(__fmul_rn(v[i], s.b2), __fmul_rn(s.one_minus_b2, __fmul_rn(gi, gi)));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), s.bc2_sqrt), s.eps);
    pi = __fadd_rn(pi, __fmul_rn(s.neg_step_size, __fdiv_rn(mi, denom)));    // addcdiv_
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (tgt) tgt[i] = __fadd_rn(__fmul_rn(s.tau, pi), __fmul_rn(s.one_minus_tau, tgt[i]));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward of a 3-layer MLP  x -> act -> act -> out    (Q nets: relu; multiplier: tanh)
// ---------------------------------------------------------------------------------------------------------------
struct Mlp3Grads { float *w0, *b0, *w1, *b1, *w2, *b2; };
static inline Mlp3Grads mlp3_grads(const drpo_mlp3& net, const float* params, float* grads) {
  Mlp3Grads g;
  g.w0 = grads + (net.l0.w - params); g.b0 = grads + (net.l0.b - params);
  g.w1 = grads + (net.l1.w - params); g.b1 = grads + (net.l1.b - params);
  g.w2 = grads + (net.l2.w - params); g.b2 = grads + (net.l2.b - params);
  return g;
}
static inline int mlp3_bwd(const drpo_mlp3& net, const Mlp3Grads& g, const float* x, int ldx, const float* h1, const float* h2,
                           const float* dout, int B, int mask_mode, float* dhA, float* dhB, float* partial, int64_t partial_floats,
                           void* stream) {
  const int H1 = net.l0.out_dim, H2 = net.l1.out_dim, O = net.l2.out_dim; int rc;
  if ((rc = linear_bwd_weight(dout, O, h2, H2, B, O, H2, g.w2, g.b2, partial, partial_floats, stream))) return rc;
  if ((rc = linear_bwd_data(dout, O, net.l2, dhA, H2, B, h2, H2, mask_mode, 0.f, stream))) return rc;
  if ((rc = linear_bwd_weight(dhA, H2, h1, H1, B, H2, H1, g.w1, g.b1, partial, partial_floats, stream))) return rc;
  if ((rc = linear_bwd_data(dhA, H2, net.l1, dhB, H1, B, h1, H1, mask_mode, 0.f, stream))) return rc;
  return linear_bwd_weight(dhB, H1, x, ldx, B, H1, net.l0.in_dim, g.w0, g.b0, partial, partial_floats, stream);
}

constexpr int64_t PARTIAL_FLOATS = 16 * 256 * 320;
constexpr int64_t LT_WORKSPACE_BYTES = 32ll << 20;      // scratch lent to cuBLASLt in tensor-core mode
constexpr int LOSS_BLOCKS = 296;

static inline int64_t critic_ws_bytes(int64_t B, int S, int A, int C, int H) {
  const int64_t D = S + A;
  int64_t f = 3 * B * D + 2 * B * A + B            // sa, nsa1, nsa2, a1, a2, logp
            + 2 * B * H + B * 2 * A               // policy scratch
            + 2 * B                               // q1t, q2t
            + 4 * B * H + 3 * B * C               // target-Qc scratch acts + mean/ls raw + nqc
            + 4 * B * H + 2 * B                   // Q1,Q2 saved h1,h2 + q
            + 4 * B * H + 2 * B * C               // Qc saved
            + 2 * B + 2 * B * C                   // dq1,dq2,dmean,dls
            + 3 * B * H                           // dhA, dhB, dt2
            + PARTIAL_FLOATS;
  return f * 4 + (LOSS_BLOCKS * 2 + 2 * 1184 + 8) * 8 + 64 * 256 + LT_WORKSPACE_BYTES;
}

static inline drpo_mlp3 const_view(const drpo_mlp3& n) { return n; }

static inline int critic_step_fp32(const drpo_critic_args& a) {
  const int64_t B = a.batch_size; const int S = a.state_dim, A = a.action_dim, C = a.con_dim, D = S + A;
  const int H = a.q[0].l0.out_dim; void* stream = a.stream; int rc;
  Arena ar(a.workspace, a.workspace_bytes);
  float* sa = ar.take<float>(B * D); float* nsa1 = ar.take<float>(B * D); float* nsa2 = ar.take<float>(B * D);
  float* a1 = ar.take<float>(B * A); float* a2 = ar.take<float>(B * A); float* logp = ar.take<float>(B);
  float* phA = ar.take<float>(B * H); float* phB = ar.take<float>(B * H); float* pout = ar.take<float>(B * 2 * A);
  float* q1t = ar.take<float>(B); float* q2t = ar.take<float>(B);
  QcActs ta; ta.t1 = ar.take<float>(B * H); ta.t2 = ar.take<float>(B * H); ta.m1 = ar.take<float>(B * H); ta.l1 = ar.take<float>(B * H);
  ta.mean_raw = ar.take<float>(B * C); ta.ls_raw = ar.take<float>(B * C); float* nqc = ar.take<float>(B * C);
  float* qh1[2]; float* qh2[2]; float* qv[2];
  for (int i = 0; i < 2; ++i) { qh1[i] = ar.take<float>(B * H); qh2[i] = ar.take<float>(B * H); qv[i] = ar.take<float>(B); }
  QcActs oa; oa.t1 = ar.take<float>(B * H); oa.t2 = ar.take<float>(B * H); oa.m1 = ar.take<float>(B * H); oa.l1 = ar.take<float>(B * H);
  oa.mean_raw = ar.take<float>(B * C); oa.ls_raw = ar.take<float>(B * C);
  float* dq[2] = {ar.take<float>(B), ar.take<float>(B)}; float* dmean = ar.take<float>(B * C); float* dls = ar.take<float>(B * C);
  float* dhA = ar.take<float>(B * H); float* dhB = ar.take<float>(B * H); float* dt2 = ar.take<float>(B * H);
  float* partial = ar.take<float>(PARTIAL_FLOATS);
  double* loss_part = ar.take<double>(LOSS_BLOCKS * 2); double* nrm_part = ar.take<double>(2 * 1184);
  float* coef = ar.take<float>(4);
  g_lt_workspace = ar.take<char>(LT_WORKSPACE_BYTES); g_lt_workspace_bytes = (size_t)LT_WORKSPACE_BYTES;
  if (!ar.ok()) { set_error("drpo_critic_step: workspace too small (%lld needed, %lld given)", (long long)ar.off, (long long)a.workspace_bytes); return DRPO_ERR_WORKSPACE; }
  const int64_t n_all = a.n_params_q + a.n_params_qc;

  if (a.phases & 1) {
    DRPO_CUDA_OK(cudaMemsetAsync(a.losses + DRPO_LOSS_ERR_SLOT, 0, sizeof(float), (cudaStream_t)stream));   // watchdog slot: no fused kernel on this path
    const drpo_batch& b = a.batch;
    // ---- no-grad passes -------------------------------------------------------------------------------------
    // actor.distr(next_obs).sample(), log_prob                                  src/ssac.py:286-288
    if ((rc = mlp3_fwd(*a.actor, b.next_obs, S, (int)B, ACT_RELU, phA, phB, pout, nullptr, stream))) return rc;
    NoiseView n1 = make_noise(a.eps_actor, A, a.seed, TAG_CRITIC_ACTOR, a.noise_step, a.row_id_offset);
    DRPO_LAUNCH(policy_head_kernel, grid_for(B), 256, 0, stream, pout, n1, (const int32_t*)nullptr, 0, a1, logp, B, A, (const int*)nullptr);
    // actor_safe.distr(next_obs).sample()                                       src/ssac.py:340-341
    if ((rc = mlp3_fwd(*a.actor_safe, b.next_obs, S, (int)B, ACT_RELU, phA, phB, pout, nullptr, stream))) return rc;
    NoiseView n2 = make_noise(a.eps_safe, A, a.seed, TAG_CRITIC_SAFE, a.noise_step, a.row_id_offset);
    DRPO_LAUNCH(policy_head_kernel, grid_for(B), 256, 0, stream, pout, n2, (const int32_t*)nullptr, 0, a2, (float*)nullptr, B, A, (const int*)nullptr);
    DRPO_LAUNCH(cat2_kernel, grid_for(B * D), 256, 0, stream, b.obs, b.act, sa, B, S, A);
    DRPO_LAUNCH(cat2_kernel, grid_for(B * D), 256, 0, stream, b.next_obs, a1, nsa1, B, S, A);
    DRPO_LAUNCH(cat2_kernel, grid_for(B * D), 256, 0, stream, b.next_obs, a2, nsa2, B, S, A);
    // critic_target.min(next_obs, next_action)                                  src/ssac.py:289
    if ((rc = mlp3_fwd(a.q_target[0], nsa1, D, (int)B, ACT_RELU, phA, phB, q1t, nullptr, stream))) return rc;
    if ((rc = mlp3_fwd(a.q_target[1], nsa1, D, (int)B, ACT_RELU, phA, phB, q2t, nullptr, stream))) return rc;
    // constraint_critic_target(next_obs, next_action, sample=True)              src/ssac.py:342-344
    if ((rc = qc_fwd(a.qc_target, nsa2, D, (int)B, ta, true, stream))) return rc;
    NoiseView n3 = make_noise(a.eps_qc, C, a.seed, TAG_CRITIC_QC, a.noise_step, a.row_id_offset);
    DRPO_LAUNCH(qc_head_kernel, grid_for(B * C), 256, 0, stream, ta.mean_raw, ta.ls_raw, 2, 0.f, n3, (int64_t)0, (float*)nullptr,
                (float*)nullptr, nqc, B, C);
    // ---- passes with gradient -----------------------------------------------------------------------------
    for (int i = 0; i < 2; ++i)
      if ((rc = mlp3_fwd(a.q[i], sa, D, (int)B, ACT_RELU, qh1[i], qh2[i], qv[i], nullptr, stream))) return rc;
    if ((rc = qc_fwd(a.qc, sa, D, (int)B, oa, true, stream))) return rc;
    // ---- targets, losses, output gradients --------------------------------------------------------------------
    CriticLossArgs L;
    L.rew = b.rew; L.cv = b.cv; L.done = b.done; L.logp = logp; L.q1t = q1t; L.q2t = q2t; L.nqc = nqc; L.q1 = qv[0]; L.q2 = qv[1];
    L.mean_raw = oa.mean_raw; L.ls_raw = oa.ls_raw; L.log_alpha = a.log_alpha;
    L.gamma = (float)a.discount; L.one_minus_gamma = (float)(1.0 - a.discount); L.td_bound = (float)a.qc_td_bound;
    L.inv_bg = (float)(1.0 / (double)a.global_batch_size); L.inv_bgc = (float)(1.0 / ((double)a.global_batch_size * C));
    L.dq1 = dq[0]; L.dq2 = dq[1]; L.dmean = dmean; L.dls = dls; L.partials = loss_part; L.B = B; L.C = C;
    DRPO_LAUNCH(critic_loss_kernel, LOSS_BLOCKS, 256, 0, stream, L);
    Scale2 sc; sc.v[0] = 1.0 / (double)a.global_batch_size; sc.v[1] = 1.0 / ((double)a.global_batch_size * C);
    DRPO_LAUNCH(loss_finalize2_kernel, 1, 32, 0, stream, loss_part, LOSS_BLOCKS, sc, a.losses);
    // ---- backward ---------------------------------------------------------------------------------------------
    for (int i = 0; i < 2; ++i) {
      Mlp3Grads g = mlp3_grads(a.q[i], a.params, a.grads);
      if ((rc = mlp3_bwd(a.q[i], g, sa, D, qh1[i], qh2[i], dq[i], (int)B, 1, dhA, dhB, partial, PARTIAL_FLOATS, stream))) return rc;
    }
    {
      const drpo_qc& q = a.qc; float* G = a.grads; const float* P = a.params;
      auto gp = [&](const float* p) { return G + (p - P); };
      // mean head
      if ((rc = linear_bwd_weight(dmean, C, oa.m1, H, (int)B, C, H, gp(q.mean1.w), gp(q.mean1.b), partial, PARTIAL_FLOATS, stream))) return rc;
      if ((rc = linear_bwd_data(dmean, C, q.mean1, dhA, H, (int)B, oa.m1, H, 1, 0.f, stream))) return rc;
      if ((rc = linear_bwd_weight(dhA, H, oa.t2, H, (int)B, H, H, gp(q.mean0.w), gp(q.mean0.b), partial, PARTIAL_FLOATS, stream))) return rc;
      if ((rc = linear_bwd_data(dhA, H, q.mean0, dt2, H, (int)B, oa.t2, H, 1, 0.f, stream))) return rc;
      // log-std head
      if ((rc = linear_bwd_weight(dls, C, oa.l1, H, (int)B, C, H, gp(q.lstd1.w), gp(q.lstd1.b), partial, PARTIAL_FLOATS, stream))) return rc;
      if ((rc = linear_bwd_data(dls, C, q.lstd1, dhA, H, (int)B, oa.l1, H, 1, 0.f, stream))) return rc;
      if ((rc = linear_bwd_weight(dhA, H, oa.t2, H, (int)B, H, H, gp(q.lstd0.w), gp(q.lstd0.b), partial, PARTIAL_FLOATS, stream))) return rc;
      if ((rc = linear_bwd_data(dhA, H, q.lstd0, dt2, H, (int)B, oa.t2, H, 1, 1.f, stream))) return rc;   // dt2 += ...
      // trunk
      if ((rc = linear_bwd_weight(dt2, H, oa.t1, H, (int)B, H, H, gp(q.trunk1.w), gp(q.trunk1.b), partial, PARTIAL_FLOATS, stream))) return rc;
      if ((rc = linear_bwd_data(dt2, H, q.trunk1, dhA, H, (int)B, oa.t1, H, 1, 0.f, stream))) return rc;
      if ((rc = linear_bwd_weight(dhA, H, sa, D, (int)B, H, D, gp(q.trunk0.w), gp(q.trunk0.b), partial, PARTIAL_FLOATS, stream))) return rc;
    }
  }
  if (a.phases & 2) {
    // clip_grad_norm_ x2 (src/ssac.py:449-450), Adam, CosineAnnealingLR is host-side, EMA x2 (:452-455)
    const int nb = 592;
    DRPO_LAUNCH(sumsq_kernel, nb, 256, 0, stream, a.grads, a.n_params_q, a.n_params_qc, nrm_part);
    DRPO_LAUNCH(clip_coef_kernel, 1, 32, 0, stream, nrm_part, nb, (float)a.grad_norm, a.losses + 2, coef);
    AdamScalars s = adam_scalars(a.adam, a.tau);
    DRPO_LAUNCH(adam_ema_kernel, grid_for(n_all), 256, 0, stream, a.params, a.grads, a.adam_m, a.adam_v, a.target_params,
                a.n_params_q, n_all, coef, s, (const float*)(a.losses + DRPO_LOSS_ERR_SLOT));
  }
  return DRPO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// multiplier step
// ---------------------------------------------------------------------------------------------------------------
// x_aug = [obs, safe_qc], with safe_qc = max_c(mean + ratio*std) of the safe action; also penalty from the actor action
static __global__ void mult_prep_kernel(const float* __restrict__ obs, const float* __restrict__ qc_a, const float* __restrict__ qc_s,
                                 float* __restrict__ xaug, float* __restrict__ penalty, float* __restrict__ safe_qc, int64_t B, int S,
                                 int C, float thr, float lb, float ub) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x) {
    float ma = qc_a[r * C], ms = qc_s[r * C];
    for (int c = 1; c < C; ++c) { ma = fmaxf(ma, qc_a[r * C + c]); ms = fmaxf(ms, qc_s[r * C + c]); }   // _get_qc
    penalty[r] = fminf(fmaxf(ma - thr, lb), ub);
    safe_qc[r] = ms;
    for (int c = 0; c < S; ++c) xaug[r * (S + 1) + c] = obs[r * S + c];
    xaug[r * (S + 1) + S] = ms;
  }
}
static __global__ void __launch_bounds__(256) mult_loss_kernel(const float* __restrict__ raw, const float* __restrict__ penalty,
                                                        const float* __restrict__ safe_qc, float ub, float lam_eps, float inv_bg,
                                                        float* __restrict__ draw, double* partials, int64_t B) {
  double l0 = 0, l1 = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x) {
    const float th = tanhf(raw[r] / ub * 2.f);
    const float lam = ub / 2.f * (1.f + th);                                 // src/ssac.py:109-110
    const bool unsafe = safe_qc[r] > 0.f;
    const float ls = unsafe ? 0.f : lam, lu = unsafe ? lam : 0.f;
    const float tgt = unsafe ? (ub - lam_eps) : 0.f;
    l0 += (double)(ls * penalty[r]);
    l1 += (double)((lu - tgt) * (lu - tgt));
    const float dlam = unsafe ? 2.f * (lu - tgt) * inv_bg : -0.5f * penalty[r] * inv_bg;
    draw[r] = dlam * (1.f - th * th);                                        // d lam / d raw = (ub/2)(1-th^2)(2/ub)
  }
  l0 = warp_sum_d(l0); l1 = warp_sum_d(l1);
  __shared__ double s0[8], s1[8];
  if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = l0; s1[threadIdx.x >> 5] = l1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0;
    for (int w = 0; w < 8; ++w) { t0 += s0[w]; t1 += s1[w]; }
    partials[2 * blockIdx.x] = -0.5 * t0 + t1; partials[2 * blockIdx.x + 1] = 0.0;
  }
}

static inline int64_t mult_ws_bytes(int64_t B, int S, int A, int C, int H) {
  int64_t f = B * (S + A) * 2 + 2 * B * A + 2 * B * H + B * 2 * A + 4 * B * H + 2 * B * C + 2 * B * C + B * (S + 1) + 3 * B
            + 2 * B * H + B + 2 * B * H + PARTIAL_FLOATS;
  return f * 4 + (LOSS_BLOCKS * 2 + 2 * 1184 + 8) * 8 + 64 * 256 + LT_WORKSPACE_BYTES;
}

static inline int multiplier_step_fp32(const drpo_multiplier_args& a) {
  const int64_t B = a.batch_size; const int S = a.state_dim, A = a.action_dim, C = a.con_dim, D = S + A;
  const int H = a.lam.l0.out_dim; void* stream = a.stream; int rc;
  Arena ar(a.workspace, a.workspace_bytes);
  float* sa1 = ar.take<float>(B * D); float* sa2 = ar.take<float>(B * D);
  float* a1 = ar.take<float>(B * A); float* a2 = ar.take<float>(B * A);
  float* phA = ar.take<float>(B * H); float* phB = ar.take<float>(B * H); float* pout = ar.take<float>(B * 2 * A);
  QcActs qa; qa.t1 = ar.take<float>(B * H); qa.t2 = ar.take<float>(B * H); qa.m1 = ar.take<float>(B * H); qa.l1 = ar.take<float>(B * H);
  qa.mean_raw = ar.take<float>(B * C); qa.ls_raw = ar.take<float>(B * C);
  float* qc_a = ar.take<float>(B * C); float* qc_s = ar.take<float>(B * C);
  float* xaug = ar.take<float>(B * (S + 1)); float* penalty = ar.take<float>(B); float* safe_qc = ar.take<float>(B); float* raw = ar.take<float>(B);
  float* h1 = ar.take<float>(B * H); float* h2 = ar.take<float>(B * H); float* draw = ar.take<float>(B);
  float* dhA = ar.take<float>(B * H); float* dhB = ar.take<float>(B * H);
  float* partial = ar.take<float>(PARTIAL_FLOATS);
  double* loss_part = ar.take<double>(LOSS_BLOCKS * 2); double* nrm_part = ar.take<double>(2 * 1184);
  float* coef = ar.take<float>(4);
  g_lt_workspace = ar.take<char>(LT_WORKSPACE_BYTES); g_lt_workspace_bytes = (size_t)LT_WORKSPACE_BYTES;
  if (!ar.ok()) { set_error("drpo_multiplier_step: workspace too small (%lld needed, %lld given)", (long long)ar.off, (long long)a.workspace_bytes); return DRPO_ERR_WORKSPACE; }
  NoiseView none = make_noise(nullptr, 0, 0, 0, 0);
  if (a.phases & 1) {
    DRPO_CUDA_OK(cudaMemsetAsync(a.losses + DRPO_LOSS_ERR_SLOT, 0, sizeof(float), (cudaStream_t)stream));   // watchdog slot: no fused kernel on this path
    // action = actor.distr(obs).rsample()                                       src/ssac.py:530-531
    if ((rc = mlp3_fwd(*a.actor, a.obs, S, (int)B, ACT_RELU, phA, phB, pout, nullptr, stream))) return rc;
    NoiseView n1 = make_noise(a.eps_actor, A, a.seed, TAG_MULT_ACTOR, a.noise_step, a.row_id_offset);
    DRPO_LAUNCH(policy_head_kernel, grid_for(B), 256, 0, stream, pout, n1, (const int32_t*)nullptr, 0, a1, (float*)nullptr, B, A, (const int*)nullptr);
    // action_safe = actor_safe.act(obs, eval=True)                              src/ssac.py:546
    if ((rc = mlp3_fwd(*a.actor_safe, a.obs, S, (int)B, ACT_RELU, phA, phB, pout, nullptr, stream))) return rc;
    DRPO_LAUNCH(policy_head_kernel, grid_for(B), 256, 0, stream, pout, none, (const int32_t*)nullptr, 1, a2, (float*)nullptr, B, A, (const int*)nullptr);
    DRPO_LAUNCH(cat2_kernel, grid_for(B * D), 256, 0, stream, a.obs, a1, sa1, B, S, A);
    DRPO_LAUNCH(cat2_kernel, grid_for(B * D), 256, 0, stream, a.obs, a2, sa2, B, S, A);
    // constraint_critic(obs, action, uncertainty=True)  = mean + std_ratio*std   src/ssac.py:534,548 ; :85
    if ((rc = qc_fwd(*a.qc, sa1, D, (int)B, qa, true, stream))) return rc;
    DRPO_LAUNCH(qc_head_kernel, grid_for(B * C), 256, 0, stream, qa.mean_raw, qa.ls_raw, 1, (float)a.std_ratio, none, (int64_t)0,
                (float*)nullptr, (float*)nullptr, qc_a, B, C);
    if ((rc = qc_fwd(*a.qc, sa2, D, (int)B, qa, true, stream))) return rc;
    DRPO_LAUNCH(qc_head_kernel, grid_for(B * C), 256, 0, stream, qa.mean_raw, qa.ls_raw, 1, (float)a.std_ratio, none, (int64_t)0,
                (float*)nullptr, (float*)nullptr, qc_s, B, C);
    DRPO_LAUNCH(mult_prep_kernel, grid_for(B), 256, 0, stream, a.obs, qc_a, qc_s, xaug, penalty, safe_qc, B, S, C,
                (float)a.constraint_threshold, (float)a.penalty_lb, (float)a.penalty_ub);
    // lams = multiplier(obs, safe_Qc)                                           src/ssac.py:549 ; :107-111
    if ((rc = mlp3_fwd(a.lam, xaug, S + 1, (int)B, ACT_TANH, h1, h2, raw, nullptr, stream))) return rc;
    DRPO_LAUNCH(mult_loss_kernel, LOSS_BLOCKS, 256, 0, stream, raw, penalty, safe_qc, (float)a.upper_bound, (float)a.lam_epsilon,
                (float)(1.0 / (double)a.global_batch_size), draw, loss_part, B);
    Scale2 sc; sc.v[0] = 1.0 / (double)a.global_batch_size; sc.v[1] = 0.0;
    DRPO_LAUNCH(loss_finalize2_kernel, 1, 32, 0, stream, loss_part, LOSS_BLOCKS, sc, a.losses);
    Mlp3Grads g = mlp3_grads(a.lam, a.params, a.grads);
    if ((rc = mlp3_bwd(a.lam, g, xaug, S + 1, h1, h2, draw, (int)B, 2, dhA, dhB, partial, PARTIAL_FLOATS, stream))) return rc;
  }
  if (a.phases & 2) {
    const int nb = 592;
    DRPO_LAUNCH(sumsq_kernel, nb, 256, 0, stream, a.grads, a.n_params, (int64_t)0, nrm_part);
    DRPO_LAUNCH(clip_coef_kernel, 1, 32, 0, stream, nrm_part, nb, (float)a.grad_norm, a.losses + 1, coef);
    AdamScalars s = adam_scalars(a.adam, 0.0);
    DRPO_LAUNCH(adam_ema_kernel, grid_for(a.n_params), 256, 0, stream, a.params, a.grads, a.adam_m, a.adam_v, (float*)nullptr,
                a.n_params, a.n_params, coef, s, (const float*)(a.losses + DRPO_LOSS_ERR_SLOT));
  }
  return DRPO_OK;
}

}  // namespace drpo
