// One iteration of BatchedGaussianEnsemble.fit's loop (src/dynamics.py:143-170): compute_loss (Gaussian NLL of every member on
// its own contiguous block of the batch + the log-var bound regulariser), hand-written backward, Adam (coupled L2 on every
// trainable tensor incl. the bounds, :93-101); and the forward-only holdout scoring at the end of fit (:172-186).
// Members are independent networks: each one runs forward / backward on its rows with the dense layers of gemm_simt.cuh
// (fp32 FFMA, or TF32 tensor-op GEMMs in the tensor mode); SiLU and its derivative act on the saved pre-activations.
#pragma once
#include "critic.cuh"

namespace drpo {

static __global__ void silu_fwd_kernel(const float* __restrict__ z, float* __restrict__ h, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) h[i] = silu_f(z[i]);
}
// dz = dh * silu'(z),  silu'(z) = s + z*s*(1-s), s = sigmoid(z)        (in place on dh)
static __global__ void silu_bwd_kernel(float* __restrict__ dh, const float* __restrict__ z, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = sigmoid_f(z[i]);
    dh[i] *= s + z[i] * s * (1.f - s);
  }
}

// per-member NLL (src/dynamics.py:236-243) and its gradients w.r.t. the raw head outputs and the log-var bounds
//   mean = dd + [s,0]; lv = lo + softplus(y - lo), y = hi - softplus(hi - x);  loss = mean_{b,c}((t-mean)^2 e^{-lv}) + mean_{b,c}(lv)
struct EnsNllArgs {
  const float *dd, *lr, *s, *t, *min_lv, *max_lv;
  float *g_dd, *g_lr, *g_hi, *g_lo;         // [B,O] each (g_hi / g_lo: per-row contributions to d loss / d max_log_var, min_log_var)
  double* partials;                          // [gridDim.y members][gridDim.x] loss partial sums
  int64_t B; int S; float inv_n;             // rows per member; inv_n = 1 / (B * (S+1))
  int64_t row_stride;                        // rows between consecutive members' blocks of s / t (0: all members score the same rows)
};
static __global__ void __launch_bounds__(256) ens_nll_kernel(EnsNllArgs a) {
  const int O = a.S + 1;
  double acc = 0.0;
  // member blockIdx.y: its head outputs / gradients are the blockIdx.y-th [B,O] block, its rows start at blockIdx.y*row_stride
  const int64_t mo = (int64_t)blockIdx.y * a.B * O;
  a.dd += mo; a.lr += mo;
  if (a.g_dd) { a.g_dd += mo; a.g_lr += mo; a.g_hi += mo; a.g_lo += mo; }
  a.s += (int64_t)blockIdx.y * a.row_stride * a.S; a.t += (int64_t)blockIdx.y * a.row_stride * O;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.B * O; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / O; const int c = (int)(i - r * O);
    const float mean = __fadd_rn(a.dd[i], c < a.S ? a.s[r * a.S + c] : 0.f);
    const float x = a.lr[i], hi = a.max_lv[c], lo = a.min_lv[c];
    const float y = hi - softplus_f(hi - x);
    const float lv = lo + softplus_f(y - lo);
    const float iv = expf(-lv);
    const float err = a.t[i] - mean;
    acc += (double)(err * err * iv) + (double)lv;
    if (a.g_dd) {
      const float glv = (1.f - err * err * iv) * a.inv_n;
      const float s1 = dsoftplus(y - lo), s2 = dsoftplus(hi - x);       // d lv / d y, d softplus(hi - x) / d (hi - x)
      a.g_dd[i] = -2.f * err * iv * a.inv_n;
      a.g_lr[i] = glv * s1 * s2;
      a.g_hi[i] = glv * s1 * (1.f - s2);
      a.g_lo[i] = glv * (1.f - s1);
    }
  }
  acc = warp_sum_d(acc);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    a.partials[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}
// column sums of g_hi / g_lo over the member's rows, accumulated into the bound gradients (one block per output column;
// members are processed one after the other on the stream, so the accumulation order is fixed)
static __global__ void __launch_bounds__(256) ens_bounds_grad_kernel(const float* __restrict__ g_hi, const float* __restrict__ g_lo, int64_t B, int O,
                                                                     float* __restrict__ grad_max, float* __restrict__ grad_min) {
  const int c = blockIdx.x;
  double a = 0, b = 0;
  for (int64_t r = threadIdx.x; r < B; r += blockDim.x) { a += g_hi[r * O + c]; b += g_lo[r * O + c]; }
  a = warp_sum_d(a); b = warp_sum_d(b);
  __shared__ double sa[8], sb[8];
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tb = 0;
    for (int w = 0; w < 8; ++w) { ta += sa[w]; tb += sb[w]; }
    grad_max[c] += (float)ta; grad_min[c] += (float)tb;
  }
}
static __global__ void ens_bounds_init_kernel(float* grad_max, float* grad_min, int O, float w, float* coef) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < O) { grad_max[i] = w; grad_min[i] = -w; }
  if (i < 2) coef[i] = 1.f;                    // the shared Adam kernel multiplies gradients by a clip coefficient: none here
}
// losses[1+e] = member e's NLL; losses[0] = their sum + w * (sum max_lv - sum min_lv)
static __global__ void ens_loss_finalize_kernel(const double* partials, int nblocks, int E, double inv_n, const float* max_lv, const float* min_lv,
                                                int O, float w, float* losses) {
  __shared__ double tot;
  if (threadIdx.x == 0) tot = 0;
  __syncthreads();
  for (int e = 0; e < E; ++e) {
    double v = 0;
    for (int b = threadIdx.x; b < nblocks; b += 32) v += partials[(int64_t)e * nblocks + b];
    v = warp_sum_d(v);
    if (threadIdx.x == 0) { losses[1 + e] = (float)(v * inv_n); tot += (double)(float)(v * inv_n); }
  }
  double reg = 0;
  for (int c = threadIdx.x; c < O; c += 32) reg += (double)max_lv[c] - (double)min_lv[c];
  reg = warp_sum_d(reg);
  if (threadIdx.x == 0) losses[0] = (float)(tot + (double)w * reg);
}

constexpr int ENS_LOSS_BLOCKS = 64;

static inline int64_t ens_train_ws_bytes(int64_t rows_per_member, int S, int A, int H, int E) {
  const int64_t B = rows_per_member, D = S + A, O = S + 1;
  int64_t f = (int64_t)E * (B * D + 10 * B * H + 6 * B * O + 3 * B * H) + PARTIAL_FLOATS + 64;      // [E][B][*] scratch of the batched path
  return f * 4 + (int64_t)E * ENS_LOSS_BLOCKS * 8 + 64 * 256 + LT_WORKSPACE_BYTES;
}

static inline int ensemble_train_step(const drpo_ensemble_train_args& a) {
  const drpo_ensemble& e = a.ens;
  const int S = e.state_dim, A = e.action_dim, H = e.hidden, E = e.ensemble_size, D = S + A, O = S + 1;
  const int64_t Bm = a.shared_rows ? a.n_rows : a.n_rows / E;            // rows per member (a remainder is dropped, src/dynamics.py:146-150)
  void* stream = a.stream; int rc;
  if (Bm <= 0) { set_error("drpo_ensemble_train_step: fewer rows than members"); return DRPO_ERR_ARG; }
  Arena ar(a.workspace, a.workspace_bytes);
  float* x0 = ar.take<float>(Bm * D);
  float *z0 = ar.take<float>(Bm * H), *h0 = ar.take<float>(Bm * H), *z1 = ar.take<float>(Bm * H), *h1 = ar.take<float>(Bm * H);
  float *zd = ar.take<float>(Bm * H), *hd = ar.take<float>(Bm * H), *zl = ar.take<float>(Bm * H), *hl = ar.take<float>(Bm * H);
  float *dd = ar.take<float>(Bm * O), *lr = ar.take<float>(Bm * O);
  float *g_dd = ar.take<float>(Bm * O), *g_lr = ar.take<float>(Bm * O), *g_hi = ar.take<float>(Bm * O), *g_lo = ar.take<float>(Bm * O);
  float *dhA = ar.take<float>(Bm * H), *dhB = ar.take<float>(Bm * H), *dh1 = ar.take<float>(Bm * H);
  float* partial = ar.take<float>(PARTIAL_FLOATS);
  float* coef = ar.take<float>(4);
  double* loss_part = ar.take<double>((int64_t)E * ENS_LOSS_BLOCKS);
  g_lt_workspace = ar.take<char>(LT_WORKSPACE_BYTES); g_lt_workspace_bytes = (size_t)LT_WORKSPACE_BYTES;
  if (!ar.ok()) { set_error("drpo_ensemble_train_step: workspace too small (%lld needed, %lld given)", (long long)ar.off, (long long)a.workspace_bytes); return DRPO_ERR_WORKSPACE; }
  const bool train = (a.phases & 1) != 0;
  auto gp = [&](const float* p) { return a.grads + (p - a.params); };
  const unsigned gH = grid_for(Bm * H);

  // ---- all members in one launch per layer (fp32 FFMA kernel, batched over blockIdx.z): the per-member loop below needs ~35
  //      small launches per member and is launch-bound at the reference's 256 rows per member ------------------------------
  // (used in the tensor mode as well: at <= 512 rows per member the layers are too small for the tensor-op GEMMs to pay)
  if ((a.phases & 5) && Bm <= 512) {
    const int saved_mode = g_gemm_mode;
    g_gemm_mode = 0;
    struct Restore { int v; ~Restore() { g_gemm_mode = v; } } restore{saved_mode};
    Arena br(a.workspace, a.workspace_bytes);
    const int64_t R = (int64_t)E * Bm;                                  // rows over all members
    float* X0 = br.take<float>(R * D);
    float *Z0 = br.take<float>(R * H), *H0 = br.take<float>(R * H), *Z1 = br.take<float>(R * H), *H1 = br.take<float>(R * H);
    float *ZD = br.take<float>(R * H), *HD = br.take<float>(R * H), *ZL = br.take<float>(R * H), *HL = br.take<float>(R * H);
    float *DD = br.take<float>(R * O), *LR = br.take<float>(R * O);
    float *GDD = br.take<float>(R * O), *GLR = br.take<float>(R * O), *GHI = br.take<float>(R * O), *GLO = br.take<float>(R * O);
    float *DA = br.take<float>(R * H), *DB = br.take<float>(R * H), *DH1 = br.take<float>(R * H);
    float* bcoef = br.take<float>(4);
    double* bloss = br.take<double>((int64_t)E * ENS_LOSS_BLOCKS);
    if (!br.ok()) { set_error("drpo_ensemble_train_step: workspace too small (%lld needed, %lld given)", (long long)br.off, (long long)a.workspace_bytes); return DRPO_ERR_WORKSPACE; }
    coef = bcoef;
    const int64_t sH = Bm * H, sO = Bm * O;
    const int64_t xrows = a.shared_rows ? Bm : R;                       // shared rows: one packed input block read by every member
    const int64_t sX = a.shared_rows ? 0 : Bm * D;
    MemberNet n = member_of(e, 0);
    const unsigned gA = grid_for(R * H);
    DRPO_LAUNCH(ens_pack_kernel, grid_for(xrows * D), 256, 0, stream, a.states, a.actions, e.norm_mean, e.norm_std, X0, xrows, S, A, (const int*)nullptr);
    if ((rc = linear_fwd_batched(E, X0, D, sX, n.t0, Z0, H, sH, (int)Bm, ACT_NONE, stream))) return rc;
    DRPO_LAUNCH(silu_fwd_kernel, gA, 256, 0, stream, Z0, H0, R * H);
    if ((rc = linear_fwd_batched(E, H0, H, sH, n.t1, Z1, H, sH, (int)Bm, ACT_NONE, stream))) return rc;
    DRPO_LAUNCH(silu_fwd_kernel, gA, 256, 0, stream, Z1, H1, R * H);
    if ((rc = linear_fwd_batched(E, H1, H, sH, n.d0, ZD, H, sH, (int)Bm, ACT_NONE, stream))) return rc;
    DRPO_LAUNCH(silu_fwd_kernel, gA, 256, 0, stream, ZD, HD, R * H);
    if ((rc = linear_fwd_batched(E, HD, H, sH, n.d1, DD, O, sO, (int)Bm, ACT_NONE, stream))) return rc;
    if ((rc = linear_fwd_batched(E, H1, H, sH, n.l0, ZL, H, sH, (int)Bm, ACT_NONE, stream))) return rc;
    DRPO_LAUNCH(silu_fwd_kernel, gA, 256, 0, stream, ZL, HL, R * H);
    if ((rc = linear_fwd_batched(E, HL, H, sH, n.l1, LR, O, sO, (int)Bm, ACT_NONE, stream))) return rc;
    EnsNllArgs L;
    L.dd = DD; L.lr = LR; L.s = a.states; L.t = a.targets; L.min_lv = e.min_log_var; L.max_lv = e.max_log_var;
    L.g_dd = train ? GDD : nullptr; L.g_lr = GLR; L.g_hi = GHI; L.g_lo = GLO;
    L.partials = bloss; L.B = Bm; L.S = S; L.inv_n = (float)(1.0 / ((double)Bm * O)); L.row_stride = a.shared_rows ? 0 : Bm;
    DRPO_LAUNCH(ens_nll_kernel, dim3(ENS_LOSS_BLOCKS, E), 256, 0, stream, L);
    if (train) {
      DRPO_LAUNCH(ens_bounds_init_kernel, 1, 64, 0, stream, gp(e.max_log_var), gp(e.min_log_var), O, (float)a.log_var_bound_weight, coef);
      DRPO_LAUNCH(ens_bounds_grad_kernel, O, 256, 0, stream, GHI, GLO, R, O, gp(e.max_log_var), gp(e.min_log_var));
      // log-var head
      if ((rc = linear_bwd_weight_batched(E, GLR, O, sO, HL, H, sH, (int)Bm, O, H, gp(n.l1.w), gp(n.l1.b), stream))) return rc;
      if ((rc = linear_bwd_data_batched(E, GLR, O, sO, n.l1, DA, H, sH, (int)Bm, 0.f, stream))) return rc;
      DRPO_LAUNCH(silu_bwd_kernel, gA, 256, 0, stream, DA, ZL, R * H);
      if ((rc = linear_bwd_weight_batched(E, DA, H, sH, H1, H, sH, (int)Bm, H, H, gp(n.l0.w), gp(n.l0.b), stream))) return rc;
      if ((rc = linear_bwd_data_batched(E, DA, H, sH, n.l0, DH1, H, sH, (int)Bm, 0.f, stream))) return rc;
      // diff head
      if ((rc = linear_bwd_weight_batched(E, GDD, O, sO, HD, H, sH, (int)Bm, O, H, gp(n.d1.w), gp(n.d1.b), stream))) return rc;
      if ((rc = linear_bwd_data_batched(E, GDD, O, sO, n.d1, DB, H, sH, (int)Bm, 0.f, stream))) return rc;
      DRPO_LAUNCH(silu_bwd_kernel, gA, 256, 0, stream, DB, ZD, R * H);
      if ((rc = linear_bwd_weight_batched(E, DB, H, sH, H1, H, sH, (int)Bm, H, H, gp(n.d0.w), gp(n.d0.b), stream))) return rc;
      if ((rc = linear_bwd_data_batched(E, DB, H, sH, n.d0, DH1, H, sH, (int)Bm, 1.f, stream))) return rc;
      // trunk
      DRPO_LAUNCH(silu_bwd_kernel, gA, 256, 0, stream, DH1, Z1, R * H);
      if ((rc = linear_bwd_weight_batched(E, DH1, H, sH, H0, H, sH, (int)Bm, H, H, gp(n.t1.w), gp(n.t1.b), stream))) return rc;
      if ((rc = linear_bwd_data_batched(E, DH1, H, sH, n.t1, DA, H, sH, (int)Bm, 0.f, stream))) return rc;
      DRPO_LAUNCH(silu_bwd_kernel, gA, 256, 0, stream, DA, Z0, R * H);
      if ((rc = linear_bwd_weight_batched(E, DA, H, sH, X0, D, sX, (int)Bm, H, D, gp(n.t0.w), gp(n.t0.b), stream))) return rc;
    }
    DRPO_LAUNCH(ens_loss_finalize_kernel, 1, 32, 0, stream, bloss, ENS_LOSS_BLOCKS, E, 1.0 / ((double)Bm * O), e.max_log_var, e.min_log_var, O,
                (float)a.log_var_bound_weight, a.losses);
  } else if (a.phases & 5) {
    if (train) DRPO_LAUNCH(ens_bounds_init_kernel, 1, 64, 0, stream, gp(e.max_log_var), gp(e.min_log_var), O, (float)a.log_var_bound_weight, coef);
    for (int m = 0; m < E; ++m) {
      const int64_t r0 = a.shared_rows ? 0 : (int64_t)m * Bm;
      const float* s = a.states + r0 * S; const float* ac = a.actions + r0 * A; const float* t = a.targets + r0 * O;
      MemberNet n = member_of(e, m);
      // ---- forward, keeping the pre-activations (src/dynamics.py:124-134) ------------------------------------------------------
      DRPO_LAUNCH(ens_pack_kernel, grid_for(Bm * D), 256, 0, stream, s, ac, e.norm_mean, e.norm_std, x0, Bm, S, A, (const int*)nullptr);
      if ((rc = linear_fwd(x0, D, n.t0, z0, H, (int)Bm, ACT_NONE, nullptr, stream))) return rc;
      DRPO_LAUNCH(silu_fwd_kernel, gH, 256, 0, stream, z0, h0, Bm * H);
      if ((rc = linear_fwd(h0, H, n.t1, z1, H, (int)Bm, ACT_NONE, nullptr, stream))) return rc;
      DRPO_LAUNCH(silu_fwd_kernel, gH, 256, 0, stream, z1, h1, Bm * H);
      if ((rc = linear_fwd(h1, H, n.d0, zd, H, (int)Bm, ACT_NONE, nullptr, stream))) return rc;
      DRPO_LAUNCH(silu_fwd_kernel, gH, 256, 0, stream, zd, hd, Bm * H);
      if ((rc = linear_fwd(hd, H, n.d1, dd, O, (int)Bm, ACT_NONE, nullptr, stream))) return rc;
      if ((rc = linear_fwd(h1, H, n.l0, zl, H, (int)Bm, ACT_NONE, nullptr, stream))) return rc;
      DRPO_LAUNCH(silu_fwd_kernel, gH, 256, 0, stream, zl, hl, Bm * H);
      if ((rc = linear_fwd(hl, H, n.l1, lr, O, (int)Bm, ACT_NONE, nullptr, stream))) return rc;
      // ---- loss and output gradients ------------------------------------------------------------------------------------------
      EnsNllArgs L;
      L.dd = dd; L.lr = lr; L.s = s; L.t = t; L.min_lv = e.min_log_var; L.max_lv = e.max_log_var;
      L.g_dd = train ? g_dd : nullptr; L.g_lr = g_lr; L.g_hi = g_hi; L.g_lo = g_lo;
      L.partials = loss_part + (int64_t)m * ENS_LOSS_BLOCKS; L.B = Bm; L.S = S; L.inv_n = (float)(1.0 / ((double)Bm * O)); L.row_stride = 0;
      DRPO_LAUNCH(ens_nll_kernel, ENS_LOSS_BLOCKS, 256, 0, stream, L);
      if (!train) continue;
      DRPO_LAUNCH(ens_bounds_grad_kernel, O, 256, 0, stream, g_hi, g_lo, Bm, O, gp(e.max_log_var), gp(e.min_log_var));
      // ---- backward ------------------------------------------------------------------------------------------------------------
      // log-var head
      if ((rc = linear_bwd_weight(g_lr, O, hl, H, (int)Bm, O, H, gp(n.l1.w), gp(n.l1.b), partial, PARTIAL_FLOATS, stream))) return rc;
      if ((rc = linear_bwd_data(g_lr, O, n.l1, dhA, H, (int)Bm, nullptr, 0, 0, 0.f, stream))) return rc;
      DRPO_LAUNCH(silu_bwd_kernel, gH, 256, 0, stream, dhA, zl, Bm * H);
      if ((rc = linear_bwd_weight(dhA, H, h1, H, (int)Bm, H, H, gp(n.l0.w), gp(n.l0.b), partial, PARTIAL_FLOATS, stream))) return rc;
      if ((rc = linear_bwd_data(dhA, H, n.l0, dh1, H, (int)Bm, nullptr, 0, 0, 0.f, stream))) return rc;
      // diff head
      if ((rc = linear_bwd_weight(g_dd, O, hd, H, (int)Bm, O, H, gp(n.d1.w), gp(n.d1.b), partial, PARTIAL_FLOATS, stream))) return rc;
      if ((rc = linear_bwd_data(g_dd, O, n.d1, dhB, H, (int)Bm, nullptr, 0, 0, 0.f, stream))) return rc;
      DRPO_LAUNCH(silu_bwd_kernel, gH, 256, 0, stream, dhB, zd, Bm * H);
      if ((rc = linear_bwd_weight(dhB, H, h1, H, (int)Bm, H, H, gp(n.d0.w), gp(n.d0.b), partial, PARTIAL_FLOATS, stream))) return rc;
      if ((rc = linear_bwd_data(dhB, H, n.d0, dh1, H, (int)Bm, nullptr, 0, 0, 1.f, stream))) return rc;      // dh1 += ...
      // trunk
      DRPO_LAUNCH(silu_bwd_kernel, gH, 256, 0, stream, dh1, z1, Bm * H);
      if ((rc = linear_bwd_weight(dh1, H, h0, H, (int)Bm, H, H, gp(n.t1.w), gp(n.t1.b), partial, PARTIAL_FLOATS, stream))) return rc;
      if ((rc = linear_bwd_data(dh1, H, n.t1, dhA, H, (int)Bm, nullptr, 0, 0, 0.f, stream))) return rc;
      DRPO_LAUNCH(silu_bwd_kernel, gH, 256, 0, stream, dhA, z0, Bm * H);
      if ((rc = linear_bwd_weight(dhA, H, x0, D, (int)Bm, H, D, gp(n.t0.w), gp(n.t0.b), partial, PARTIAL_FLOATS, stream))) return rc;
    }
    DRPO_LAUNCH(ens_loss_finalize_kernel, 1, 32, 0, stream, loss_part, ENS_LOSS_BLOCKS, E, 1.0 / ((double)Bm * O), e.max_log_var, e.min_log_var, O,
                (float)a.log_var_bound_weight, a.losses);
  }
  if (a.phases & 2) {
    if (!train) DRPO_LAUNCH(ens_bounds_init_kernel, 1, 64, 0, stream, coef + 2, coef + 2, 0, 0.f, coef);   // (coef = 1 only)
    AdamScalars sc = adam_scalars(a.adam, 0.0);
    DRPO_LAUNCH(adam_ema_kernel, grid_for(a.n_params), 256, 0, stream, a.params, a.grads, a.adam_m, a.adam_v, (float*)nullptr, a.n_params,
                a.n_params, coef, sc);
  }
  return DRPO_OK;
}

}  // namespace drpo
