// extern "C" entry points of libdrpo_sm100.so (see include/drpo_b200.h for the contract of each).
#include <stdarg.h>

#include "common.cuh"
#include "critic.cuh"
#include "actor.cuh"
#include "ensemble_train.cuh"
#include "nets.cuh"
#include "shield.cuh"
#include "rollout.cuh"
#include "umma_api.h"
#include "critic_umma_api.h"

namespace drpo {
static thread_local char g_err[1024] = "";
int64_t g_launch_count = 0;
thread_local int g_gemm_mode = 0;
thread_local void* g_lt_workspace = nullptr; thread_local size_t g_lt_workspace_bytes = 0;
cublasHandle_t gemm_cublas_handle() {
  static cublasHandle_t h = nullptr;
  if (!h) {
    if (cublasCreate(&h) != CUBLAS_STATUS_SUCCESS) { h = nullptr; return nullptr; }
    cublasSetMathMode(h, CUBLAS_TF32_TENSOR_OP_MATH);
  }
  return h;
}
void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}

// ---- sticky watchdog status (see common.cuh) ----------------------------------------------------------------------------------
static int* g_status_host = nullptr;
static int* g_status_dev = nullptr;
int* status_words_host() {
  if (!g_status_host) {
    if (cudaHostAlloc((void**)&g_status_host, 2 * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); g_status_host = nullptr; return nullptr; }
    g_status_host[0] = g_status_host[1] = 0;
    if (cudaHostGetDevicePointer((void**)&g_status_dev, g_status_host, 0) != cudaSuccess) { cudaGetLastError(); g_status_dev = g_status_host; }
  }
  return g_status_host;
}
__global__ void publish_status_kernel(const int* __restrict__ err_flag, int* host_word, float* loss_flag) {
  const int code = *err_flag;
  if (code != 0) { *(volatile int*)host_word = code; __threadfence_system(); }
  if (loss_flag) *loss_flag = code != 0 ? 1.f : 0.f;
}
int publish_status(const int* err_flag, int which, float* loss_flag, void* stream) {
  if (!status_words_host()) { set_error("cannot allocate the pinned status words"); return DRPO_ERR_CUDA; }
  DRPO_LAUNCH(publish_status_kernel, 1, 1, 0, stream, err_flag, g_status_dev + which, loss_flag);
  return DRPO_OK;
}
static int status_report() {
  if (!g_status_host) return 0;
  const int r = ((volatile int*)g_status_host)[0], c = ((volatile int*)g_status_host)[1];
  const int code = r ? r : c;
  if (code) set_error("bf16 %s kernel: an in-kernel wait timed out (code %d): pipeline protocol bug, the results of that call are invalid "
                      "(the rollout did not advance the ring pointer / the update step applied no parameter update)", r ? "rollout" : "update-step", code);
  return code;
}
int kernel_status_peek() { return status_report(); }
int kernel_status_sync() {
  if (!g_status_host) return 0;
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("drpo_kernel_status: %s", cudaGetErrorString(cudaGetLastError())); return DRPO_ERR_CUDA; }
  return status_report();
}

__global__ void philox_fill_kernel(float* out, int64_t n, int cols, const int32_t* row_ids, uint64_t seed, uint32_t tag, uint32_t step) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * cols; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols; const int c = (int)(i % cols);
    out[i] = philox_normal1(seed, (uint32_t)(row_ids ? row_ids[r] : r), (uint32_t)c, tag, step);
  }
}
}  // namespace drpo

using namespace drpo;

extern "C" {

const char* drpo_last_error(void) { return g_err; }
int drpo_abi_version(void) { return DRPO_ABI_VERSION; }
int64_t drpo_launch_count(void) { return g_launch_count; }
int drpo_kernel_status(void) { return kernel_status_sync(); }
int drpo_kernel_status_peek(void) { return kernel_status_peek(); }
void drpo_timing_enable(int32_t on) { umma_timing_enable(on); }
int drpo_timing_read(double* total_ms_host, int64_t* launches_host, double* satellites_ms_host) {
  DRPO_CHECK_ARG(total_ms_host && launches_host, "drpo_timing_read: NULL output");
  return umma_timing_read(total_ms_host, launches_host, satellites_ms_host);
}

int drpo_philox_normal(float* out, int64_t n, int32_t cols, const int32_t* row_ids, uint64_t seed, uint32_t stream_tag,
                       uint32_t step, void* stream) {
  DRPO_CHECK_ARG(out && n >= 0 && cols > 0, "drpo_philox_normal: bad arguments");
  if (n == 0) return DRPO_OK;
  DRPO_LAUNCH(philox_fill_kernel, grid_for(n * cols), 256, 0, stream, out, n, cols, row_ids, seed, stream_tag, step);
  return DRPO_OK;
}

static int check_env(const drpo_env_params* e) {
  DRPO_CHECK_ARG(e, "env params are NULL");
  DRPO_CHECK_ARG(e->kind >= DRPO_ENV_POINT_ROBOT && e->kind <= DRPO_ENV_TRACKING, "unknown env kind %d", e->kind);
  DRPO_CHECK_ARG(e->state_dim > 0 && e->con_dim > 0 && e->con_dim <= DRPO_MAX_CON, "bad env dims S=%d C=%d", e->state_dim, e->con_dim);
  if (e->kind == DRPO_ENV_POINT_ROBOT)
    DRPO_CHECK_ARG(e->state_dim >= 2 && e->con_dim == 1 && e->n_hazards >= 1 && e->n_hazards <= DRPO_MAX_HAZARDS, "bad point-robot params");
  if (e->kind == DRPO_ENV_BOUNDED) {
    DRPO_CHECK_ARG(e->n_active >= 1 && e->n_active <= DRPO_MAX_ACTIVE && e->con_dim == 2 * e->n_active, "bounded env: con_dim must be 2*n_active");
    for (int i = 0; i < e->n_active; ++i) DRPO_CHECK_ARG(e->active_dims[i] >= 0 && e->active_dims[i] < e->state_dim, "active dim out of range");
    DRPO_CHECK_ARG(e->n_done_dims >= 0 && e->n_done_dims <= DRPO_MAX_DONE_DIMS, "bad n_done_dims");
    for (int i = 0; i < e->n_done_dims; ++i) DRPO_CHECK_ARG(e->done_dims[i] >= 0 && e->done_dims[i] < e->state_dim, "done dim out of range");
  }
  if (e->kind == DRPO_ENV_TRACKING)
    DRPO_CHECK_ARG(e->con_dim == 1 && e->surr_veh_num >= 1 && e->surr_start >= 7 && e->surr_start + 4 * e->surr_veh_num <= e->state_dim,
                   "bad tracking params");
  return DRPO_OK;
}

int drpo_hooks_eval(const drpo_env_params* env, const float* states, int64_t n, uint8_t* done, uint8_t* violation,
                    float* constraint_values, void* stream) {
  int rc = check_env(env); if (rc) return rc;
  DRPO_CHECK_ARG(n >= 0 && (n == 0 || states), "drpo_hooks_eval: bad arguments");
  if (n == 0) return DRPO_OK;
  DRPO_LAUNCH(hooks_kernel, grid_for(n), 256, 0, stream, *env, states, n, done, violation, constraint_values, (const int*)nullptr);
  return DRPO_OK;
}

static int check_ens(const drpo_ensemble* e) {
  DRPO_CHECK_ARG(e && e->state_dim > 0 && e->action_dim > 0 && e->ensemble_size > 0 && e->ensemble_size <= 64 && e->hidden > 0,
                 "bad ensemble dims (ensemble_size must be 1..64)");
  DRPO_CHECK_ARG(e->norm_mean && e->norm_std && e->min_log_var && e->max_log_var && e->trunk0_w && e->trunk0_b && e->trunk1_w &&
                 e->trunk1_b && e->diff0_w && e->diff0_b && e->diff1_w && e->diff1_b && e->lvar0_w && e->lvar0_b && e->lvar1_w && e->lvar1_b,
                 "ensemble weight pointer is NULL");
  return DRPO_OK;
}

int64_t drpo_ensemble_workspace_bytes(const drpo_ensemble* ens, int64_t batch) {
  if (!ens || batch < 0) return -1;
  return ens_scratch_floats(*ens, batch) * 4 + 16 * 256 + (cu::ens_bf16_supported(*ens) ? cu::ens_bf16_ws_bytes(*ens) : 0);
}
// DRPO_PREC_BF16 of the standalone ensemble entry points: the fused tcgen05 member chain of ens_umma.cu
static int ens_bf16_call(const drpo_ensemble* ens, int m0, int m1, int per_member, const float* states, const float* actions, int64_t batch,
                         float* means, float* log_vars, const drpo_noise* noise, float* next_states, float* rewards, void* workspace,
                         int64_t workspace_bytes, void* stream) {
  DRPO_CHECK_ARG(workspace && workspace_bytes >= 4096, "ensemble forward (bf16): workspace too small");
  int* err_flag = (int*)workspace;
  DRPO_CUDA_OK(cudaMemsetAsync(err_flag, 0, 16, (cudaStream_t)stream));
  int rc = cu::ens_bf16_run(*ens, m0, m1, per_member, states, actions, batch, means, log_vars, noise, next_states, rewards,
                            (char*)workspace + 4096, workspace_bytes - 4096, stream, err_flag);
  if (rc) return rc;
  return publish_status(err_flag, 1, nullptr, stream);
}

int drpo_ensemble_forward(const drpo_ensemble* ens, int32_t member, int32_t per_member_inputs, const float* states,
                          const float* actions, int64_t batch, float* means, float* log_vars, int32_t precision, void* workspace,
                          int64_t workspace_bytes, void* stream) {
  int rc = check_ens(ens); if (rc) return rc;
  DRPO_CHECK_ARG(member >= -1 && member < ens->ensemble_size, "member %d out of range", member);
  DRPO_CHECK_ARG(batch >= 0 && (batch == 0 || (states && actions && means && log_vars)), "drpo_ensemble_forward: bad arguments");
  DRPO_CHECK_ARG(precision == DRPO_PREC_FP32 || precision == DRPO_PREC_BF16, "drpo_ensemble_forward: DRPO_PREC_FP32 or DRPO_PREC_BF16");
  if (batch == 0) return DRPO_OK;
  if (precision == DRPO_PREC_BF16)
    return ens_bf16_call(ens, member < 0 ? 0 : member, member < 0 ? ens->ensemble_size : member + 1, per_member_inputs && member < 0, states, actions,
                         batch, means, log_vars, nullptr, nullptr, nullptr, workspace, workspace_bytes, stream);
  Arena ar(workspace, workspace_bytes);
  EnsScratch w = ens_scratch(ar, *ens, batch);
  if (!ar.ok()) { set_error("drpo_ensemble_forward: workspace too small"); return DRPO_ERR_WORKSPACE; }
  const int S = ens->state_dim, A = ens->action_dim, O = S + 1;
  const int m0 = member < 0 ? 0 : member, m1 = member < 0 ? ens->ensemble_size : member + 1;
  for (int m = m0; m < m1; ++m) {
    const int64_t slot = member < 0 ? m : 0;
    const float* s = states + (per_member_inputs && member < 0 ? slot * batch * S : 0);
    const float* a = actions + (per_member_inputs && member < 0 ? slot * batch * A : 0);
    if ((rc = ens_member_raw(*ens, m, s, a, (int)batch, w, nullptr, stream))) return rc;
    DRPO_LAUNCH(ens_head_kernel, grid_for(batch * O), 256, 0, stream, w.dd, w.lr, s, ens->min_log_var, ens->max_log_var,
                means + slot * batch * O, log_vars + slot * batch * O, batch, S);
  }
  return DRPO_OK;
}

int drpo_ensemble_sample(const drpo_ensemble* ens, int32_t member, const float* states, const float* actions, int64_t batch,
                         const drpo_noise* noise, float* next_states, float* rewards, int32_t precision, void* workspace,
                         int64_t workspace_bytes, void* stream) {
  int rc = check_ens(ens); if (rc) return rc;
  DRPO_CHECK_ARG(member >= 0 && member < ens->ensemble_size, "member %d out of range", member);
  DRPO_CHECK_ARG(batch >= 0 && (batch == 0 || (states && actions && noise && next_states && rewards)), "drpo_ensemble_sample: bad arguments");
  DRPO_CHECK_ARG(precision == DRPO_PREC_FP32 || precision == DRPO_PREC_BF16, "drpo_ensemble_sample: DRPO_PREC_FP32 or DRPO_PREC_BF16");
  if (batch == 0) return DRPO_OK;
  if (precision == DRPO_PREC_BF16)
    return ens_bf16_call(ens, member, member + 1, 0, states, actions, batch, nullptr, nullptr, noise, next_states, rewards, workspace,
                         workspace_bytes, stream);
  Arena ar(workspace, workspace_bytes);
  EnsScratch w = ens_scratch(ar, *ens, batch);
  if (!ar.ok()) { set_error("drpo_ensemble_sample: workspace too small"); return DRPO_ERR_WORKSPACE; }
  const int S = ens->state_dim;
  if ((rc = ens_member_raw(*ens, member, states, actions, (int)batch, w, nullptr, stream))) return rc;
  NoiseView nv = make_noise(noise->eps, noise->row_stride, noise->seed, noise->stream_tag, noise->step);
  DRPO_LAUNCH(ens_sample_kernel, grid_for(batch * (S + 1)), 256, 0, stream, w.dd, w.lr, states, ens->min_log_var, ens->max_log_var, nv,
              (const int32_t*)nullptr, next_states, rewards, batch, S, (const int*)nullptr);
  return DRPO_OK;
}

static int check_mlp3(const drpo_mlp3* m, const char* what) {
  DRPO_CHECK_ARG(m && m->l0.w && m->l0.b && m->l1.w && m->l1.b && m->l2.w && m->l2.b, "%s: weight pointer is NULL", what);
  DRPO_CHECK_ARG(m->l0.out_dim == m->l1.in_dim && m->l1.out_dim == m->l2.in_dim && m->l0.in_dim > 0 && m->l2.out_dim > 0,
                 "%s: inconsistent layer dims", what);
  return DRPO_OK;
}
static int check_qc(const drpo_qc* q, const char* what) {
  DRPO_CHECK_ARG(q && q->trunk0.w && q->trunk1.w && q->mean0.w && q->mean1.w && q->lstd0.w && q->lstd1.w && q->trunk0.b && q->trunk1.b &&
                 q->mean0.b && q->mean1.b && q->lstd0.b && q->lstd1.b, "%s: weight pointer is NULL", what);
  return DRPO_OK;
}

int64_t drpo_policy_workspace_bytes(const drpo_mlp3* actor, int64_t batch) {
  if (!actor || batch < 0) return -1;
  return pol_scratch_floats(*actor, batch) * 4 + 8 * 256;
}

int drpo_policy_act(const drpo_mlp3* actor, const float* states, int64_t batch, int32_t eval_mode, const drpo_noise* noise,
                    float* actions, float* log_prob, int32_t precision, void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = check_mlp3(actor, "drpo_policy_act"); if (rc) return rc;
  DRPO_CHECK_ARG(batch >= 0 && (batch == 0 || (states && actions && (eval_mode || noise))), "drpo_policy_act: bad arguments");
  DRPO_CHECK_ARG(actor->l2.out_dim % 2 == 0, "actor output dim must be 2*action_dim");
  DRPO_CHECK_ARG(precision == DRPO_PREC_FP32, "drpo_policy_act: only DRPO_PREC_FP32 here");
  if (batch == 0) return DRPO_OK;
  Arena ar(workspace, workspace_bytes);
  PolScratch w = pol_scratch(ar, *actor, batch);
  if (!ar.ok()) { set_error("drpo_policy_act: workspace too small"); return DRPO_ERR_WORKSPACE; }
  if ((rc = mlp3_fwd(*actor, states, actor->l0.in_dim, (int)batch, ACT_RELU, w.hA, w.hB, w.out, nullptr, stream))) return rc;
  NoiseView nv = noise ? make_noise(noise->eps, noise->row_stride, noise->seed, noise->stream_tag, noise->step) : make_noise(nullptr, 0, 0, 0, 0);
  DRPO_LAUNCH(policy_head_kernel, grid_for(batch), 256, 0, stream, w.out, nv, (const int32_t*)nullptr, eval_mode, actions, log_prob, batch,
              actor->l2.out_dim / 2, (const int*)nullptr);
  return DRPO_OK;
}

int64_t drpo_qc_workspace_bytes(int64_t batch, int32_t hidden) {
  return (batch * (4LL * hidden + 2 * DRPO_MAX_CON + 64)) * 4 + 16 * 256;
}

int drpo_qc_forward(const drpo_qc* qc, const float* states, const float* actions, int64_t batch, int32_t state_dim, int32_t action_dim,
                    int32_t con_dim, int32_t mode, float std_ratio, const drpo_noise* noise, float* out_mean, float* out_std,
                    float* out_sample, void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = check_qc(qc, "drpo_qc_forward"); if (rc) return rc;
  DRPO_CHECK_ARG(batch >= 0 && (batch == 0 || (states && actions)) && mode >= 0 && mode <= 2, "drpo_qc_forward: bad arguments");
  DRPO_CHECK_ARG(qc->trunk0.in_dim == state_dim + action_dim && qc->mean1.out_dim == con_dim, "drpo_qc_forward: dims do not match the network");
  DRPO_CHECK_ARG(mode == 0 ? out_mean != nullptr : out_sample != nullptr, "drpo_qc_forward: output pointer is NULL");
  DRPO_CHECK_ARG(mode != 2 || noise, "drpo_qc_forward: sample mode needs noise");
  if (batch == 0) return DRPO_OK;
  const int H = qc->trunk0.out_dim, D = state_dim + action_dim, C = con_dim;
  Arena ar(workspace, workspace_bytes);
  float* sa = ar.take<float>(batch * D);
  QcActs a; a.t1 = ar.take<float>(batch * H); a.t2 = ar.take<float>(batch * H); a.m1 = ar.take<float>(batch * H); a.l1 = a.m1;
  a.mean_raw = ar.take<float>(batch * C); a.ls_raw = ar.take<float>(batch * C);
  if (!ar.ok()) { set_error("drpo_qc_forward: workspace too small"); return DRPO_ERR_WORKSPACE; }
  DRPO_LAUNCH(cat2_kernel, grid_for(batch * D), 256, 0, stream, states, actions, sa, batch, state_dim, action_dim);
  // the two heads share scratch (l1 aliases m1): run them one after the other
  if ((rc = linear_fwd(sa, D, qc->trunk0, a.t1, H, (int)batch, ACT_RELU, nullptr, stream))) return rc;
  if ((rc = linear_fwd(a.t1, H, qc->trunk1, a.t2, H, (int)batch, ACT_RELU, nullptr, stream))) return rc;
  if ((rc = linear_fwd(a.t2, H, qc->mean0, a.m1, H, (int)batch, ACT_RELU, nullptr, stream))) return rc;
  if ((rc = linear_fwd(a.m1, H, qc->mean1, a.mean_raw, C, (int)batch, ACT_NONE, nullptr, stream))) return rc;
  if (mode != 0) {
    if ((rc = linear_fwd(a.t2, H, qc->lstd0, a.l1, H, (int)batch, ACT_RELU, nullptr, stream))) return rc;
    if ((rc = linear_fwd(a.l1, H, qc->lstd1, a.ls_raw, C, (int)batch, ACT_NONE, nullptr, stream))) return rc;
  }
  NoiseView nv = noise ? make_noise(noise->eps, noise->row_stride, noise->seed, noise->stream_tag, noise->step) : make_noise(nullptr, 0, 0, 0, 0);
  DRPO_LAUNCH(qc_head_kernel, grid_for(batch * C), 256, 0, stream, a.mean_raw, a.ls_raw, mode, std_ratio, nv, (int64_t)0, out_mean, out_std,
              out_sample, batch, C);
  return DRPO_OK;
}

int64_t drpo_shield_workspace_bytes(const drpo_mlp3* actor, int64_t n, int32_t state_dim, int32_t action_dim, int32_t con_dim,
                                    int32_t hidden) {
  if (!actor || n < 0) return -1;
  const int64_t rows = (int64_t)SHIELD_MIX * n;
  return drpo_policy_workspace_bytes(actor, n) + drpo_qc_workspace_bytes(rows, hidden) +
         (2 * n * action_dim + rows * (state_dim + action_dim + con_dim)) * 4 + 8 * 256;
}

int drpo_shield_act(const drpo_shield_args* a) {
  DRPO_CHECK_ARG(a, "drpo_shield_act: NULL args");
  int rc;
  if ((rc = check_mlp3(a->actor, "drpo_shield_act(actor)"))) return rc;
  const int S = a->state_dim, A = a->action_dim, C = a->con_dim;
  const int64_t n = a->n;
  DRPO_CHECK_ARG(a->shield_type >= DRPO_SHIELD_NONE && a->shield_type <= DRPO_SHIELD_LINEAR, "drpo_shield_act: unknown shield type %d", a->shield_type);
  DRPO_CHECK_ARG(n >= 0 && (n == 0 || (a->states && a->actions)) && (a->eval_perf || a->noise_perf), "drpo_shield_act: bad arguments");
  DRPO_CHECK_ARG(a->actor->l0.in_dim == S && a->actor->l2.out_dim == 2 * A, "drpo_shield_act: actor dims do not match");
  DRPO_CHECK_ARG(C >= 1 && C <= DRPO_MAX_CON, "drpo_shield_act: con_dim out of range");
  if (n == 0) return DRPO_OK;
  DRPO_CHECK_ARG(a->path >= 0 && a->path <= 2, "drpo_shield_act: path must be 0 (auto), 1 (latency kernels) or 2 (batched)");
  const int n_mix_all = a->shield_type == DRPO_SHIELD_LINEAR ? SHIELD_MIX : (a->shield_type == DRPO_SHIELD_SAFE ? 1 : 0);
  const bool fits = a->actor->l0.out_dim <= SHIELD_FUSED_THREADS && a->actor->l1.out_dim <= SHIELD_FUSED_THREADS && S + A <= SHIELD_FUSED_THREADS &&
                    (n_mix_all == 0 || (a->qc && a->qc->trunk0.out_dim <= SHIELD_FUSED_THREADS && a->actor_safe &&
                                        a->actor_safe->l0.out_dim <= SHIELD_FUSED_THREADS && a->actor_safe->l1.out_dim <= SHIELD_FUSED_THREADS));
  DRPO_CHECK_ARG(a->path != 1 || fits, "drpo_shield_act: the latency kernels need hidden widths <= %d", SHIELD_FUSED_THREADS);
  const bool fused = a->path == 1 || (a->path == 0 && fits && (int64_t)std::max(n_mix_all, 1) * n <= SHIELD_FUSED_MAX_ROWS);
  if (a->shield_type == DRPO_SHIELD_NONE && !fused) {
    if ((rc = drpo_policy_act(a->actor, a->states, n, a->eval_perf, a->noise_perf, a->actions, nullptr, DRPO_PREC_FP32, a->workspace,
                              a->workspace_bytes, a->stream))) return rc;
    if (a->choice) DRPO_CUDA_OK(cudaMemsetAsync(a->choice, 0, n * sizeof(int32_t), (cudaStream_t)a->stream));
    return DRPO_OK;
  }
  if (a->shield_type != DRPO_SHIELD_NONE) {
    if ((rc = check_mlp3(a->actor_safe, "drpo_shield_act(actor_safe)"))) return rc;
    if ((rc = check_qc(a->qc, "drpo_shield_act"))) return rc;
    DRPO_CHECK_ARG(a->actor_safe->l0.in_dim == S && a->actor_safe->l2.out_dim == 2 * A, "drpo_shield_act: actor_safe dims do not match");
    DRPO_CHECK_ARG(a->qc->trunk0.in_dim == S + A && a->qc->mean1.out_dim == C, "drpo_shield_act: constraint critic dims do not match");
  }
  if (fused) {
    Arena ar(a->workspace, a->workspace_bytes);
    ShieldFusedArgs f;
    f.actor = *a->actor; f.actor_safe = n_mix_all ? *a->actor_safe : *a->actor; if (n_mix_all) f.qc = *a->qc; else f.qc = drpo_qc{};
    f.states = a->states; f.n = (int)n; f.S = S; f.A = A; f.C = C; f.n_mix = n_mix_all; f.eval_perf = a->eval_perf; f.uncertainty = a->uncertainty;
    f.std_ratio = a->std_ratio; f.threshold = a->threshold; f.ratios = shield_ratios();
    f.noise = a->noise_perf ? make_noise(a->noise_perf->eps, a->noise_perf->row_stride, a->noise_perf->seed, a->noise_perf->stream_tag, a->noise_perf->step)
                            : make_noise(nullptr, 0, 0, 0, 0);
    f.a_perf = ar.take<float>(n * A); f.a_safe = ar.take<float>(n * A); f.qrow = ar.take<float>((int64_t)std::max(n_mix_all, 1) * n);
    f.done_cnt = n_mix_all ? ar.take<int32_t>(n) : nullptr;
    f.actions = a->actions; f.qc_perf = a->qc_perf; f.choice = a->choice;
    if (!ar.ok()) { set_error("drpo_shield_act: workspace too small"); return DRPO_ERR_WORKSPACE; }
    DRPO_LAUNCH(shield_policy_kernel, dim3((unsigned)n, n_mix_all ? 2 : 1), SHIELD_FUSED_THREADS, 0, a->stream, f);
    if (n_mix_all) DRPO_LAUNCH(shield_qc_select_kernel, (unsigned)(n_mix_all * n), SHIELD_FUSED_THREADS, 0, a->stream, f);
    else if (a->choice) DRPO_CUDA_OK(cudaMemsetAsync(a->choice, 0, n * sizeof(int32_t), (cudaStream_t)a->stream));
    return DRPO_OK;
  }
  if ((rc = check_mlp3(a->actor_safe, "drpo_shield_act(actor_safe)"))) return rc;
  if ((rc = check_qc(a->qc, "drpo_shield_act"))) return rc;
  DRPO_CHECK_ARG(a->actor_safe->l0.in_dim == S && a->actor_safe->l2.out_dim == 2 * A, "drpo_shield_act: actor_safe dims do not match");
  const int n_mix = a->shield_type == DRPO_SHIELD_LINEAR ? SHIELD_MIX : 1;
  const int64_t rows = (int64_t)n_mix * n;
  const int H = a->qc->trunk0.out_dim;
  Arena ar(a->workspace, a->workspace_bytes);
  float* a_perf = ar.take<float>(n * A); float* a_safe = ar.take<float>(n * A);
  float* cand_s = ar.take<float>(rows * S); float* cand_a = ar.take<float>(rows * A); float* q = ar.take<float>(rows * C);
  const int64_t pol_bytes = drpo_policy_workspace_bytes(a->actor, n), qc_bytes = drpo_qc_workspace_bytes(rows, H);
  char* sub = ar.take<char>(std::max(pol_bytes, qc_bytes));
  if (!ar.ok()) { set_error("drpo_shield_act: workspace too small"); return DRPO_ERR_WORKSPACE; }
  cudaStream_t stream = (cudaStream_t)a->stream;
  if ((rc = drpo_policy_act(a->actor, a->states, n, a->eval_perf, a->noise_perf, a_perf, nullptr, DRPO_PREC_FP32, sub, pol_bytes, stream))) return rc;
  if ((rc = drpo_policy_act(a->actor_safe, a->states, n, 1, nullptr, a_safe, nullptr, DRPO_PREC_FP32, sub, pol_bytes, stream))) return rc;
  DRPO_LAUNCH(shield_candidates_kernel, grid_for(rows * (S + A)), 256, 0, stream, a->states, a_perf, a_safe, shield_ratios(), n_mix, n, S, A,
              cand_s, cand_a);
  if ((rc = drpo_qc_forward(a->qc, cand_s, cand_a, rows, S, A, C, a->uncertainty ? 1 : 0, a->std_ratio, nullptr, a->uncertainty ? nullptr : q, nullptr, q, sub, qc_bytes,
                            stream))) return rc;
  DRPO_LAUNCH(shield_select_kernel, grid_for(n), 256, 0, stream, q, cand_a, a_safe, a->threshold, n_mix, n, A, C, a->actions, a->qc_perf,
              a->choice);
  return DRPO_OK;
}

static int check_buffer(const drpo_buffer* b, const char* what) {
  DRPO_CHECK_ARG(b && b->states && b->actions && b->next_states && b->rewards && b->dones && b->violations && b->constraint_values,
                 "%s: buffer component pointer is NULL", what);
  DRPO_CHECK_ARG(b->capacity > 0 && b->state_dim > 0 && b->action_dim > 0 && b->con_dim > 0, "%s: bad buffer dims", what);
  return DRPO_OK;
}

int drpo_buffer_gather(const drpo_buffer* real, const drpo_buffer* virt, const int64_t* idx, int64_t n_real, int64_t n_total,
                       float reward_scale, float alive_bonus, float constraint_scale, float constraint_offset, const drpo_batch* out,
                       void* stream) {
  int rc;
  DRPO_CHECK_ARG(n_real >= 0 && n_total >= n_real && out && idx, "drpo_buffer_gather: bad arguments");
  if (n_real > 0 && (rc = check_buffer(real, "drpo_buffer_gather(real)"))) return rc;
  if (n_total > n_real && (rc = check_buffer(virt, "drpo_buffer_gather(virt)"))) return rc;
  if (n_total == 0) return DRPO_OK;
  const drpo_buffer* r = n_real > 0 ? real : virt; const drpo_buffer* v = n_total > n_real ? virt : real;
  DRPO_CHECK_ARG(r->state_dim == v->state_dim && r->action_dim == v->action_dim && r->con_dim == v->con_dim, "buffers disagree on dims");
  DRPO_LAUNCH(buffer_gather_kernel, grid_for(n_total * r->state_dim), 256, 0, stream, *r, *v, idx, n_real, n_total, reward_scale, alive_bonus,
              constraint_scale, constraint_offset, *out);
  return DRPO_OK;
}

static int check_rollout(const drpo_rollout_args* a) {
  DRPO_CHECK_ARG(a && a->actor && a->ensemble && a->env, "drpo_rollout: NULL argument");
  int rc;
  if ((rc = check_mlp3(a->actor, "drpo_rollout(actor)"))) return rc;
  if ((rc = check_ens(a->ensemble))) return rc;
  if ((rc = check_env(a->env))) return rc;
  if ((rc = check_buffer(&a->virt, "drpo_rollout(virt)"))) return rc;
  const int S = a->ensemble->state_dim, A = a->ensemble->action_dim;
  DRPO_CHECK_ARG(a->env->state_dim == S && a->actor->l0.in_dim == S && a->actor->l2.out_dim == 2 * A, "drpo_rollout: actor/ensemble/env dims disagree");
  DRPO_CHECK_ARG(a->virt.state_dim == S && a->virt.action_dim == A && a->virt.con_dim == a->env->con_dim, "drpo_rollout: buffer dims disagree");
  DRPO_CHECK_ARG(a->batch >= 0 && a->horizon >= 1 && a->member_idx_host && a->step_counts && a->virt.pointer, "drpo_rollout: bad arguments");
  DRPO_CHECK_ARG(a->batch < (1LL << 31) && a->traj_id_offset + a->batch < (1LL << 31), "drpo_rollout: batch too large for 32-bit trajectory ids");
  // SampleBuffer.extend asserts batch <= capacity (src/sampling.py:131); the rollout may append up to batch*horizon rows
  DRPO_CHECK_ARG(a->batch * a->horizon <= a->virt.capacity, "drpo_rollout: batch*horizon (%lld) exceeds the buffer capacity (%lld)",
                 (long long)(a->batch * a->horizon), (long long)a->virt.capacity);
  for (int t = 0; t < a->horizon; ++t)
    DRPO_CHECK_ARG(a->member_idx_host[t] >= 0 && a->member_idx_host[t] < a->ensemble->ensemble_size, "member index out of range at step %d", t);
  DRPO_CHECK_ARG((a->eps_policy == nullptr) == (a->eps_model == nullptr), "give both or neither of eps_policy / eps_model");
  return DRPO_OK;
}

int64_t drpo_rollout_workspace_bytes(const drpo_rollout_args* a) {
  if (!a || !a->actor || !a->ensemble || !a->env) return -1;
  int64_t fp32 = rollout_ws_bytes_fp32(*a);
  int64_t bf16 = umma_rollout_ws_bytes(*a);
  return fp32 > bf16 ? fp32 : bf16;
}

int drpo_rollout(const drpo_rollout_args* a) {
  int rc = check_rollout(a); if (rc) return rc;
  if (a->batch == 0) {
    DRPO_CUDA_OK(cudaMemsetAsync(a->step_counts, 0, sizeof(int32_t) * (a->horizon + 1), (cudaStream_t)a->stream));
    return DRPO_OK;
  }
  DRPO_CHECK_ARG(!a->init_ready_flags || (a->precision == DRPO_PREC_BF16 && a->init_rows_per_flag_log2 >= 7 && a->init_rows_per_flag_log2 <= 30),
                 "drpo_rollout: streamed start states need DRPO_PREC_BF16 and row blocks of 2^7..2^30 rows");
  if (a->precision == DRPO_PREC_FP32) return rollout_fp32(*a);
  if (a->precision == DRPO_PREC_BF16) return umma_rollout(*a);
  set_error("drpo_rollout: unknown precision %d", a->precision);
  return DRPO_ERR_ARG;
}

int drpo_debug_rollout_layer(const drpo_rollout_args* a, int32_t layer, float* out) {
  int rc = check_rollout(a); if (rc) return rc;
  DRPO_CHECK_ARG(layer >= 0 && out && a->batch > 0, "drpo_debug_rollout_layer: bad arguments");
  return umma_debug_layer(*a, layer, out);
}

int64_t drpo_critic_workspace_bytes(int64_t batch, int32_t state_dim, int32_t action_dim, int32_t con_dim, int32_t hidden) {
  const int64_t a = critic_ws_bytes(batch, state_dim, action_dim, con_dim, hidden);
  const int64_t b = hidden == 256 && state_dim + action_dim <= 64 ? cu::critic_ws_bytes(batch, state_dim, action_dim, con_dim) + 4096 : 0;
  return a > b ? a : b;
}

int drpo_critic_step(const drpo_critic_args* a) {
  DRPO_CHECK_ARG(a, "drpo_critic_step: NULL args");
  int rc;
  if ((rc = check_mlp3(a->actor, "drpo_critic_step(actor)"))) return rc;
  if ((rc = check_mlp3(a->actor_safe, "drpo_critic_step(actor_safe)"))) return rc;
  for (int i = 0; i < 2; ++i) {
    if ((rc = check_mlp3(&a->q[i], "drpo_critic_step(q)"))) return rc;
    if ((rc = check_mlp3(&a->q_target[i], "drpo_critic_step(q_target)"))) return rc;
  }
  if ((rc = check_qc(&a->qc, "drpo_critic_step(qc)"))) return rc;
  if ((rc = check_qc(&a->qc_target, "drpo_critic_step(qc_target)"))) return rc;
  DRPO_CHECK_ARG(a->batch_size >= 1 && a->global_batch_size >= a->batch_size, "drpo_critic_step: bad batch sizes");
  DRPO_CHECK_ARG(a->con_dim >= 1 && a->con_dim <= DRPO_MAX_CON, "drpo_critic_step: bad con_dim");
  DRPO_CHECK_ARG(a->params && a->grads && a->adam_m && a->adam_v && a->target_params && a->losses && a->log_alpha, "drpo_critic_step: NULL arena");
  DRPO_CHECK_ARG(a->q[0].l0.in_dim == a->state_dim + a->action_dim && a->qc.mean1.out_dim == a->con_dim, "drpo_critic_step: dims disagree");
  DRPO_CHECK_ARG(a->q[0].l0.out_dim <= 256 && a->q[0].l0.in_dim + 1 <= 320, "drpo_critic_step: hidden/input dims too large for the split-K scratch");
  DRPO_CHECK_ARG((a->phases & ~3) == 0 && a->phases != 0, "drpo_critic_step: bad phases");
  if (a->phases & 1) {
    const drpo_batch& b = a->batch;
    DRPO_CHECK_ARG(b.obs && b.act && b.next_obs && b.rew && b.done && b.cv, "drpo_critic_step: NULL batch tensor");
  }
  DRPO_CHECK_ARG(a->precision == DRPO_PREC_FP32 || a->precision == DRPO_PREC_BF16 || a->precision == DRPO_PREC_TF32,
                 "drpo_critic_step: unknown precision %d", a->precision);
  if (a->precision == DRPO_PREC_BF16) {
    // fused tcgen05 path: phase 1 = pack + fused forward/loss/dX kernel + dW kernel + gradient assembly; phase 2 is shared
    drpo_critic_args b = *a;
    if (a->phases & 1) {
      DRPO_CHECK_ARG(a->workspace_bytes >= 4096, "drpo_critic_step: workspace too small");
      int* err_flag = (int*)a->workspace;                       // first 4 KB of the workspace: in-kernel watchdog flag
      b.workspace = (char*)a->workspace + 4096; b.workspace_bytes = a->workspace_bytes - 4096;
      DRPO_CUDA_OK(cudaMemsetAsync(err_flag, 0, 16, (cudaStream_t)a->stream));
      if ((rc = cu::critic_phase1(b, err_flag))) return rc;
      if ((rc = publish_status(err_flag, 1, a->losses + DRPO_LOSS_ERR_SLOT, a->stream))) return rc;
    }
    if (a->phases & 2) { b = *a; b.phases = 2; return critic_step_fp32(b); }
    return DRPO_OK;
  }
  // DRPO_PREC_TF32 = library tensor-core mode: the dense contractions run as TF32 tensor-op GEMMs (cuBLAS), rest unchanged
  g_gemm_mode = a->precision == DRPO_PREC_TF32 ? 1 : 0;
  rc = critic_step_fp32(*a);
  g_gemm_mode = 0;
  return rc;
}

static bool solver_dims_fused(int state_dim, int action_dim, int con_dim, int hidden) {
  return hidden == 256 && state_dim + action_dim <= 64 && state_dim + 1 <= 64 && (action_dim == 1 || action_dim == 2) &&
         (con_dim == 1 || con_dim == 2 || con_dim == 4);
}
int64_t drpo_multiplier_workspace_bytes(int64_t batch, int32_t state_dim, int32_t action_dim, int32_t con_dim, int32_t hidden) {
  const int64_t a = mult_ws_bytes(batch, state_dim, action_dim, con_dim, hidden);
  const int64_t b = solver_dims_fused(state_dim, action_dim, con_dim, hidden) ? cu::solver_ws_bytes(0, batch, state_dim, action_dim, con_dim) + 4096 : 0;
  return a > b ? a : b;
}

int drpo_multiplier_step(const drpo_multiplier_args* a) {
  DRPO_CHECK_ARG(a, "drpo_multiplier_step: NULL args");
  int rc;
  if ((rc = check_mlp3(a->actor, "drpo_multiplier_step(actor)"))) return rc;
  if ((rc = check_mlp3(a->actor_safe, "drpo_multiplier_step(actor_safe)"))) return rc;
  if ((rc = check_mlp3(&a->lam, "drpo_multiplier_step(lam)"))) return rc;
  if ((rc = check_qc(a->qc, "drpo_multiplier_step(qc)"))) return rc;
  DRPO_CHECK_ARG(a->batch_size >= 1 && a->global_batch_size >= a->batch_size && a->obs, "drpo_multiplier_step: bad batch");
  DRPO_CHECK_ARG(a->lam.l0.in_dim == a->state_dim + 1 && a->lam.l2.out_dim == 1, "drpo_multiplier_step: multiplier dims disagree");
  DRPO_CHECK_ARG(a->params && a->grads && a->adam_m && a->adam_v && a->losses, "drpo_multiplier_step: NULL arena");
  DRPO_CHECK_ARG((a->phases & ~3) == 0 && a->phases != 0, "drpo_multiplier_step: bad phases");
  DRPO_CHECK_ARG(a->precision == DRPO_PREC_FP32 || a->precision == DRPO_PREC_BF16 || a->precision == DRPO_PREC_TF32,
                 "drpo_multiplier_step: unknown precision %d", a->precision);
  if (a->precision == DRPO_PREC_BF16) {
    // fused tcgen05 path (solver_umma.cu): phase 1 = pack + fused forward/loss/dX kernel + dW kernel + gradient assembly; phase 2 shared
    drpo_multiplier_args b = *a;
    if (a->phases & 1) {
      DRPO_CHECK_ARG(a->workspace_bytes >= 4096, "drpo_multiplier_step: workspace too small");
      int* err_flag = (int*)a->workspace;
      b.workspace = (char*)a->workspace + 4096; b.workspace_bytes = a->workspace_bytes - 4096;
      DRPO_CUDA_OK(cudaMemsetAsync(err_flag, 0, 16, (cudaStream_t)a->stream));
      if ((rc = cu::multiplier_phase1(b, err_flag))) return rc;
      if ((rc = publish_status(err_flag, 1, a->losses + DRPO_LOSS_ERR_SLOT, a->stream))) return rc;
    }
    if (a->phases & 2) { b = *a; b.phases = 2; return multiplier_step_fp32(b); }
    return DRPO_OK;
  }
  g_gemm_mode = a->precision == DRPO_PREC_TF32 ? 1 : 0;          // library tensor-core mode: TF32 tensor-op GEMMs (cuBLAS)
  rc = multiplier_step_fp32(*a);
  g_gemm_mode = 0;
  return rc;
}

int64_t drpo_actor_workspace_bytes(int64_t batch, int32_t state_dim, int32_t action_dim, int32_t con_dim, int32_t hidden) {
  const int64_t a = actor_ws_bytes(batch, state_dim, action_dim, con_dim, hidden);
  const int64_t b = solver_dims_fused(state_dim, action_dim, con_dim, hidden) ? cu::solver_ws_bytes(1, batch, state_dim, action_dim, con_dim) + 4096 : 0;
  return a > b ? a : b;
}

int drpo_actor_step(const drpo_actor_args* a) {
  DRPO_CHECK_ARG(a, "drpo_actor_step: NULL args");
  int rc;
  if ((rc = check_mlp3(&a->actor, "drpo_actor_step(actor)"))) return rc;
  if ((rc = check_mlp3(&a->actor_safe, "drpo_actor_step(actor_safe)"))) return rc;
  if ((rc = check_mlp3(a->q, "drpo_actor_step(q)"))) return rc;
  if ((rc = check_mlp3(a->lam, "drpo_actor_step(lam)"))) return rc;
  if ((rc = check_qc(a->qc, "drpo_actor_step(qc)"))) return rc;
  DRPO_CHECK_ARG(a->batch_size >= 1 && a->global_batch_size >= a->batch_size && a->obs, "drpo_actor_step: bad batch");
  DRPO_CHECK_ARG(a->con_dim >= 1 && a->con_dim <= DRPO_MAX_CON, "drpo_actor_step: bad con_dim");
  DRPO_CHECK_ARG(a->actor.l0.in_dim == a->state_dim && a->actor.l2.out_dim == 2 * a->action_dim && a->actor_safe.l0.in_dim == a->state_dim &&
                     a->actor_safe.l2.out_dim == 2 * a->action_dim && a->q->l0.in_dim == a->state_dim + a->action_dim && a->q->l2.out_dim == 1 &&
                     a->qc->trunk0.in_dim == a->state_dim + a->action_dim && a->qc->mean1.out_dim == a->con_dim &&
                     a->lam->l0.in_dim == a->state_dim + 1 && a->lam->l2.out_dim == 1, "drpo_actor_step: network dims disagree");
  DRPO_CHECK_ARG(a->actor.l0.out_dim == a->q->l0.out_dim && a->actor.l0.out_dim == a->qc->trunk0.out_dim && a->actor.l0.out_dim == a->lam->l0.out_dim &&
                     a->actor.l0.out_dim <= 256, "drpo_actor_step: the nets must share one hidden width <= 256");
  DRPO_CHECK_ARG(a->params_actor && a->grads_actor && a->m_actor && a->v_actor && a->params_safe && a->grads_safe && a->m_safe && a->v_safe &&
                     a->log_alpha && a->alpha_m && a->alpha_v && a->losses, "drpo_actor_step: NULL arena");
  DRPO_CHECK_ARG((a->phases & ~3) == 0 && a->phases != 0, "drpo_actor_step: bad phases");
  DRPO_CHECK_ARG(a->precision == DRPO_PREC_FP32 || a->precision == DRPO_PREC_BF16 || a->precision == DRPO_PREC_TF32,
                 "drpo_actor_step: unknown precision %d", a->precision);
  if (a->precision == DRPO_PREC_BF16) {
    // fused tcgen05 path (solver_umma.cu); phase 2 (clips, three Adam steps) is shared with the fp32 path
    drpo_actor_args b = *a;
    if (a->phases & 1) {
      DRPO_CHECK_ARG(a->workspace_bytes >= 4096, "drpo_actor_step: workspace too small");
      int* err_flag = (int*)a->workspace;
      b.workspace = (char*)a->workspace + 4096; b.workspace_bytes = a->workspace_bytes - 4096;
      DRPO_CUDA_OK(cudaMemsetAsync(err_flag, 0, 16, (cudaStream_t)a->stream));
      if ((rc = cu::actor_phase1(b, err_flag))) return rc;
      if ((rc = publish_status(err_flag, 1, a->losses + DRPO_LOSS_ERR_SLOT, a->stream))) return rc;
    }
    if (a->phases & 2) { b = *a; b.phases = 2; return actor_step_fp32(b); }
    return DRPO_OK;
  }
  g_gemm_mode = a->precision == DRPO_PREC_TF32 ? 1 : 0;
  rc = actor_step_fp32(*a);
  g_gemm_mode = 0;
  return rc;
}

int64_t drpo_ensemble_train_workspace_bytes(const drpo_ensemble* ens, int64_t rows_per_member) {
  if (!ens) return -1;
  return ens_train_ws_bytes(rows_per_member, ens->state_dim, ens->action_dim, ens->hidden, ens->ensemble_size);
}

int drpo_ensemble_train_step(const drpo_ensemble_train_args* a) {
  DRPO_CHECK_ARG(a, "drpo_ensemble_train_step: NULL args");
  int rc;
  if ((rc = check_ens(&a->ens))) return rc;
  DRPO_CHECK_ARG(a->states && a->actions && a->targets && a->n_rows >= 1 && a->losses, "drpo_ensemble_train_step: bad batch");
  DRPO_CHECK_ARG((a->phases & ~7) == 0 && a->phases != 0 && !((a->phases & 4) && (a->phases & 3)), "drpo_ensemble_train_step: bad phases");
  DRPO_CHECK_ARG((a->phases & 4) || (a->params && a->grads && a->adam_m && a->adam_v && a->n_params > 0), "drpo_ensemble_train_step: NULL arena");
  DRPO_CHECK_ARG(a->ens.hidden <= 256 && a->ens.state_dim + a->ens.action_dim + 1 <= 320, "drpo_ensemble_train_step: dims too large for the split-K scratch");
  if (a->phases & 3) {
    const float* lo = a->params; const float* hi = a->params + a->n_params;
    const float* ptrs[] = {a->ens.trunk0_w, a->ens.trunk0_b, a->ens.trunk1_w, a->ens.trunk1_b, a->ens.diff0_w, a->ens.diff0_b, a->ens.diff1_w, a->ens.diff1_b,
                           a->ens.lvar0_w, a->ens.lvar0_b, a->ens.lvar1_w, a->ens.lvar1_b, a->ens.min_log_var, a->ens.max_log_var};
    for (const float* q : ptrs) DRPO_CHECK_ARG(q >= lo && q < hi, "drpo_ensemble_train_step: a trainable tensor lies outside the parameter arena");
  }
  DRPO_CHECK_ARG(a->precision == DRPO_PREC_FP32 || a->precision == DRPO_PREC_BF16 || a->precision == DRPO_PREC_TF32,
                 "drpo_ensemble_train_step: unknown precision %d", a->precision);
  g_gemm_mode = a->precision == DRPO_PREC_TF32 ? 1 : 0;          // library GEMMs only on explicit request; DRPO_PREC_BF16 = the fp32 kernels here
  rc = ensemble_train_step(*a);
  g_gemm_mode = 0;
  return rc;
}

int drpo_debug_critic_rows(float* rows) { cu::critic_set_debug_rows(rows); return DRPO_OK; }
int drpo_debug_solver_rows(float* rows) { cu::solver_set_debug_rows(rows); return DRPO_OK; }
int drpo_debug_critic_prof(int64_t* stamps) { cu::critic_set_prof((long long*)stamps); return DRPO_OK; }

int drpo_debug_critic_dw(const void* a_oct, const void* b_oct, int32_t b_octets, int64_t rows_padded, int32_t ksplit, float* partial,
                         float* out, void* stream) {
  DRPO_CHECK_ARG(a_oct && b_oct && partial && out && b_octets >= 2 && b_octets <= 32 && (b_octets & 1) == 0 && rows_padded % 128 == 0 &&
                     ksplit >= 1 && ksplit <= rows_padded / 64, "drpo_debug_critic_dw: bad arguments");
  int* err_flag = (int*)partial;                                // first 16 bytes of the partial scratch
  DRPO_CUDA_OK(cudaMemsetAsync(err_flag, 0, 16, (cudaStream_t)stream));
  int rc = cu::critic_debug_dw(a_oct, b_oct, b_octets, rows_padded, ksplit, partial + 4, out, err_flag, stream);
  if (rc) return rc;
  if (rc == DRPO_OK) rc = publish_status(err_flag, 1, nullptr, stream);
  return DRPO_OK;
}

}  // extern "C"
