/*
 * drpo_b200.h — C ABI of libdrpo_sm100.so, the B200 (sm_100a) implementation of DRPO's hot path.
 *
 * The reference (ManUtdMoon/Distributional-Reachability-Policy-Optimization) is pure Python/PyTorch and has
 * NO FFI/plugin interface.  Each entry point below therefore replaces a *Python method* of the
 * reference; the citation next to it is the reference file:line whose behaviour it reproduces.  The host-side
 * mirror of those methods (same names / arguments / error behaviour) lives in drpo_b200/*.py and binds these
 * symbols with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; tensors are contiguous row-major fp32,
 *     masks are uint8 (torch.bool storage), nn.Linear weights are [out,in] (y = x W^T + b) exactly as in the
 *     reference state_dict, so checkpoints interchange;
 *   - the library owns no device memory: parameters, replay buffers and workspaces are allocated by the caller (PyTorch).  Its only
 *     allocations are two pinned, device-mapped host status words (cudaHostAlloc on first use, see drpo_kernel_status) and, in
 *     DRPO_PREC_TF32 only, one cuBLAS handle bound to the first device that uses it;
 *     `*_workspace_bytes` tells how much scratch a call needs;
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); no call synchronises the device
 *     (drpo_kernel_status and drpo_timing_read are the exceptions: they exist to synchronise).  The launches themselves are
 *     capturable, but calls are NOT advertised as CUDA-graph capturable: the first call of each kind sets function attributes
 *     (cudaFuncSetAttribute) and the first bf16 call allocates the status words;
 *   - return value 0 = ok, negative = error (message from drpo_last_error(), thread-local); nothing throws across
 *     the boundary;
 *   - not re-entrant per device; one process per GPU (torchrun), called from the single training thread.
 */
#ifndef DRPO_B200_H
#define DRPO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRPO_ABI_VERSION 2

/* Every update step's `losses` array holds DRPO_LOSSES_LEN floats.  Slot DRPO_LOSS_ERR_SLOT is the step's watchdog flag: phase 1
 * (bit0) writes 0, or 1 when a fused tcgen05 kernel of that phase reported an in-kernel pipeline time-out; phase 2 (bit1) applies
 * NO parameter update when the slot is non-zero.  Multi-GPU callers all-reduce (sum) the slot together with the gradients (the
 * Python mirror keeps `losses` in the tail of the gradient arena: ONE NCCL message), so a time-out on any rank vetoes the step on
 * every rank and replicas stay identical.  The host sees the time-out through drpo_kernel_status / drpo_kernel_status_peek. */
#define DRPO_LOSSES_LEN 16
#define DRPO_LOSS_ERR_SLOT 15

enum { DRPO_OK = 0, DRPO_ERR_ARG = -1, DRPO_ERR_CUDA = -2, DRPO_ERR_WORKSPACE = -3, DRPO_ERR_UNSUPPORTED = -4 };

/* arithmetic mode of the dense layers */
enum { DRPO_PREC_FP32 = 0,   /* fp32 FMA path, parity <= 1e-5 relative vs the reference */
       DRPO_PREC_BF16 = 1,   /* hand-written tcgen05/TMEM path (bf16 operands, fp32 accumulate), parity <= 2e-2:
                                drpo_rollout = fused rollout-step kernel; drpo_critic_step = fused forward/loss/dX kernel + dW
                                kernel; drpo_multiplier_step = as DRPO_PREC_TF32 */
       DRPO_PREC_TF32 = 2 }; /* library tensor-core mode of drpo_critic_step / drpo_multiplier_step: fp32 path with the dense
                                contractions as TF32 tensor-op GEMMs (cuBLAS / cuBLASLt) */

/* ------------------------------------------------------------------------------------------------------------
 * Env hooks: check_done / check_violation / get_constraint_values
 *   point-robot  src/env/point_robot.py:96-130
 *   bounded      src/env/poles/constraints.py:90-132,203-204,216-247 used by
 *                src/env/poles/inverted_pendulum.py:79-121 (cartpole) and src/env/quadrotor/quadrotor.py:83-158
 *   tracking     src/env/tracking/pyth_veh3dofconti_surrcstr_data.py:253-338
 * The arithmetic is fp64 on fp32 inputs (numpy semantics), result cast to fp32 (src/torch_util.py:20-22).
 * ---------------------------------------------------------------------------------------------------------- */
enum { DRPO_ENV_POINT_ROBOT = 0, DRPO_ENV_BOUNDED = 1, DRPO_ENV_TRACKING = 2 };
#define DRPO_MAX_ACTIVE 4
#define DRPO_MAX_DONE_DIMS 4
#define DRPO_MAX_HAZARDS 4
#define DRPO_MAX_CON 8

typedef struct drpo_env_params {
  int32_t kind;
  int32_t state_dim;
  int32_t con_dim;
  /* point robot */
  int32_t n_hazards;
  double hazard_xy[DRPO_MAX_HAZARDS][2];
  double hazard_size;
  double goal_xy[2];
  double goal_size;
  float xy_bound;
  /* bounded: cv = [ -x[d_i] + lower_i ..., x[d_i] - upper_i ... ]; violation = any(cv > 0) */
  int32_t n_active;
  int32_t active_dims[DRPO_MAX_ACTIVE];
  double lower[DRPO_MAX_ACTIVE];
  double upper[DRPO_MAX_ACTIVE];
  /* bounded: done = violation | any(|x[d_j]| > done_thr_j), compared in fp32 */
  int32_t n_done_dims;
  int32_t done_dims[DRPO_MAX_DONE_DIMS];
  float done_thr[DRPO_MAX_DONE_DIMS];
  /* tracking */
  int32_t surr_veh_num;
  int32_t surr_start;
  double veh_length;
  double veh_width;
} drpo_env_params;

/* replaces env.check_done / check_violation / get_constraint_values as called at src/smbpo.py:63-65,238-240 */
int drpo_hooks_eval(const drpo_env_params* env, const float* states, int64_t n,
                    uint8_t* done, uint8_t* violation, float* constraint_values /* [n,con_dim] */, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Networks (weights are the reference's own tensors)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct drpo_linear {       /* one nn.Linear / one member slice of a BatchedLinear */
  const float* w;                  /* [out,in] */
  const float* b;                  /* [out]    */
  int32_t in_dim, out_dim;
} drpo_linear;

typedef struct drpo_ensemble {     /* BatchedGaussianEnsemble, src/dynamics.py:55-103 */
  int32_t state_dim, action_dim, ensemble_size, hidden;
  const float* norm_mean;          /* state_normalizer.mean [S]            */
  const float* norm_std;           /* state_normalizer.std  [S]            */
  const float* min_log_var;        /* [S+1] */
  const float* max_log_var;        /* [S+1] */
  const float* trunk0_w;  const float* trunk0_b;   /* [E,H,S+A] , [E,H] */
  const float* trunk1_w;  const float* trunk1_b;   /* [E,H,H]   , [E,H] */
  const float* diff0_w;   const float* diff0_b;    /* [E,H,H]   , [E,H] */
  const float* diff1_w;   const float* diff1_b;    /* [E,S+1,H] , [E,S+1] */
  const float* lvar0_w;   const float* lvar0_b;
  const float* lvar1_w;   const float* lvar1_b;
  /* optional bf16 tcgen05 image of the weights, built by drpo_ensemble_pack_bf16 (NULL => fp32 only) */
  const void* packed_bf16;
} drpo_ensemble;

typedef struct drpo_mlp3 {         /* Linear-act-Linear-act-Linear: actor, Q_i, multiplier (src/torch_util.py:190-211) */
  drpo_linear l0, l1, l2;
  const void* packed_bf16;         /* optional tcgen05 image (drpo_mlp3_pack_bf16) */
} drpo_mlp3;

typedef struct drpo_qc {           /* ConstraintCritic, src/ssac.py:46-92 */
  drpo_linear trunk0, trunk1, mean0, mean1, lstd0, lstd1;
} drpo_qc;

/* Noise: either injected tensors (parity runs) or the in-kernel Philox4x32-10 stream (throughput runs).
 * Philox counter = (row id, column, stream tag, step); key = seed -> results do not depend on how rows are sharded. */
typedef struct drpo_noise {
  const float* eps;                /* injected N(0,1) draws, or NULL to use Philox */
  int64_t row_stride;              /* elements between consecutive rows of eps */
  uint64_t seed;
  uint32_t stream_tag;             /* distinguishes the draws of one call */
  uint32_t step;
} drpo_noise;

/* Fill out[n,cols] with the exact N(0,1) values the kernels would draw for (seed, tag, step, row ids).  Lets the
 * CPU oracle consume the same noise as a Philox-mode run. row_ids may be NULL (= 0..n-1). */
int drpo_philox_normal(float* out, int64_t n, int32_t cols, const int32_t* row_ids, uint64_t seed,
                       uint32_t stream_tag, uint32_t step, void* stream);

/* BatchedGaussianEnsemble._forward1 (src/dynamics.py:112-122) when member >= 0, _forward_all (:124-134) when
 * member == -1 (states/actions are then [E,B,*] if per_member_inputs, else [B,*] shared by all members as in
 * means() :206-210).  Outputs means/log_vars [B,S+1] or [E,B,S+1]. */
int64_t drpo_ensemble_workspace_bytes(const drpo_ensemble* ens, int64_t batch);
int drpo_ensemble_forward(const drpo_ensemble* ens, int32_t member, int32_t per_member_inputs,
                          const float* states, const float* actions, int64_t batch,
                          float* means, float* log_vars, int32_t precision,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* BatchedGaussianEnsemble.sample (src/dynamics.py:198-203) with the member already picked on the host
 * (random.choice(self._elite_inds), :199): next_states [B,S], rewards [B]. */
int drpo_ensemble_sample(const drpo_ensemble* ens, int32_t member, const float* states, const float* actions,
                         int64_t batch, const drpo_noise* noise, float* next_states, float* rewards,
                         int32_t precision, void* workspace, int64_t workspace_bytes, void* stream);

/* SquashedGaussianPolicy: TorchPolicy.act (src/policy.py:77-80) + _distr (:89-97) + SquashedGaussian
 * (src/squashed_gaussian.py:7-16).  eval_mode != 0 -> tanh(mu).  log_prob (optional, may be NULL) is
 * Independent(SquashedGaussian).log_prob of the drawn action with the cached pre-tanh value (src/ssac.py:286-288). */
int64_t drpo_policy_workspace_bytes(const drpo_mlp3* actor, int64_t batch);
int drpo_policy_act(const drpo_mlp3* actor, const float* states, int64_t batch, int32_t eval_mode,
                    const drpo_noise* noise, float* actions, float* log_prob, int32_t precision,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Device replay buffer (ConstraintSafetySampleBuffer, src/sampling.py:12-151,215-229): ring of 7 components.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct drpo_buffer {
  float* states;            /* [capacity,S] */
  float* actions;           /* [capacity,A] */
  float* next_states;       /* [capacity,S] */
  float* rewards;           /* [capacity]   */
  uint8_t* dones;           /* [capacity]   */
  uint8_t* violations;      /* [capacity]   */
  float* constraint_values; /* [capacity] if C==1 else [capacity,C] */
  int64_t* pointer;         /* device scalar, monotone; slot = pointer % capacity (src/sampling.py:128-145) */
  int64_t capacity;
  int32_t state_dim, action_dim, con_dim;
} drpo_buffer;

/* SampleBuffer.sample (src/sampling.py:147-151) fused with SMBPO.update_solver's minibatch assembly
 * (src/smbpo.py:253-270): rows [0,n_real) gathered from `real` at indices idx[0..n_real), the rest from `virt`;
 * rewards*reward_scale+alive_bonus; cv*constraint_scale (+offset where > 0).  Outputs are the 7 batch tensors. */
typedef struct drpo_batch {
  float* obs; float* act; float* next_obs; float* rew; uint8_t* done; uint8_t* viol; float* cv;
} drpo_batch;
int drpo_buffer_gather(const drpo_buffer* real, const drpo_buffer* virt, const int64_t* idx, int64_t n_real,
                       int64_t n_total, float reward_scale, float alive_bonus, float constraint_scale,
                       float constraint_offset, const drpo_batch* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * SMBPO.rollout (src/smbpo.py:229-249): H x (policy sample -> member sample -> hooks -> store -> compact).
 * Transitions are written step-major, survivor-order-preserving, straight into `virt` at
 * (pointer + row) % capacity; `pointer` is advanced on the device; no host synchronisation.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct drpo_rollout_args {
  const drpo_mlp3* actor;
  const drpo_ensemble* ensemble;
  const drpo_env_params* env;
  const float* initial_states;     /* [B0,S] */
  int64_t batch;                   /* B0 (this rank's shard) */
  int64_t traj_id_offset;          /* global id of row 0 (multi-GPU sharding: noise is keyed by global id) */
  int32_t horizon;
  const int32_t* member_idx_host;  /* HOST array [horizon]: the elite picked for each step (src/dynamics.py:199) */
  /* noise: injected (parity) eps_policy [H,B_total,A], eps_model [H,B_total,S+1] indexed by global traj id, or Philox */
  const float* eps_policy;
  const float* eps_model;
  int64_t eps_batch_stride;        /* B_total */
  uint64_t seed;
  drpo_buffer virt;                /* destination ring */
  int32_t* step_counts;            /* device [horizon+1]: rows stored at each step; [horizon] = total */
  int32_t precision;
  void* workspace;
  int64_t workspace_bytes;
  void* stream;
  /* Optional streaming of the start states (DRPO_PREC_BF16 only; NULL = off): `initial_states` may still be in flight from the
   * host on ANOTHER stream, in row blocks of 2^init_rows_per_flag_log2 rows; the copy of block j is followed (same stream) by a
   * DMA write of a non-zero int32 to init_ready_flags[j].  The first step's kernel polls the flag of a row block before it
   * reads it, so the host-to-device transfer overlaps that step instead of preceding it.  The flags must be written by copy
   * engines (cudaMemcpyAsync from pinned memory), never by a kernel: the rollout kernel occupies every SM while it waits. */
  const int32_t* init_ready_flags;
  int32_t init_rows_per_flag_log2;
} drpo_rollout_args;

int64_t drpo_rollout_workspace_bytes(const drpo_rollout_args* args);
int drpo_rollout(const drpo_rollout_args* args);
/* Debug aid for the DRPO_PREC_BF16 path: runs step 0 only and dumps the fp32 TMEM accumulator of dense layer `layer`
 * (0..8 = actor L0,L1,L2, member trunk0, trunk1, diff0, diff1, logvar0, logvar1) as out[batch, out_dim]. */
int drpo_debug_rollout_layer(const drpo_rollout_args* args, int32_t layer, float* out);

/* ------------------------------------------------------------------------------------------------------------
 * SSAC.update_critic (src/ssac.py:437-456 with compute_target :284-294, compute_cons_target :338-362,
 * cons_critic_loss_given_target :415-427) in DRPO mode (reachability, qc_under_uncertainty, distributional_qc):
 * forward of all passes, hand-written backward, two grad-norm clips, Adam (coupled L2), EMA of both targets.
 * Trainable parameters live in ONE flat fp32 arena `params` ([Q1 | Q2 | Qc] in state_dict order) with same-size
 * arenas for grads, Adam m/v and targets; the drpo_mlp3/drpo_qc structs point into them.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct drpo_adam {
  double lr;           /* from the host-side CosineAnnealingLR (src/ssac.py:204-208) */
  double beta1, beta2, eps, weight_decay;   /* doubles: torch derives its fp32 scalars from python floats */
  int32_t step;        /* 1-based step count of THIS update */
} drpo_adam;

typedef struct drpo_critic_args {
  /* batch (already reward/constraint-scaled, src/smbpo.py:261-270) */
  drpo_batch batch;
  int64_t batch_size;          /* local rows */
  int64_t global_batch_size;   /* rows over all ranks: losses/gradients are normalised by this */
  int32_t state_dim, action_dim, con_dim;
  /* frozen nets */
  const drpo_mlp3* actor;
  const drpo_mlp3* actor_safe;
  /* trainable nets + targets: pointers into the arenas below */
  drpo_mlp3 q[2];
  drpo_mlp3 q_target[2];
  drpo_qc qc;
  drpo_qc qc_target;
  float* params;  float* grads;  float* adam_m;  float* adam_v;  float* target_params;
  int64_t n_params_q;          /* params of Q1+Q2 (first clip group) */
  int64_t n_params_qc;         /* params of Qc (second clip group) */
  const float* log_alpha;      /* device scalar (src/ssac.py:225-226) */
  /* noise: injected eps_actor [B,A], eps_safe [B,A], eps_qc [B,C] or Philox(seed, step) */
  const float* eps_actor; const float* eps_safe; const float* eps_qc;
  uint64_t seed; uint32_t noise_step; int64_t row_id_offset;
  /* hyper-parameters (src/ssac.py:116,133,142,156) */
  double discount, tau, grad_norm, qc_td_bound;
  drpo_adam adam;
  /* phases: bit0 = forward+backward (fills grads, losses), bit1 = clip+Adam+EMA.  Multi-GPU callers run bit0,
   * all-reduce `grads` and `losses[0..1]`, `losses[15]` (NCCL sum; one message when `losses` is the tail of the gradient arena),
   * then run bit1. */
  int32_t phases;
  float* losses;               /* device [DRPO_LOSSES_LEN]: loss_Q, loss_C, grad-norm Q, grad-norm Qc, ..., [15] watchdog flag */
  int32_t precision;
  void* workspace; int64_t workspace_bytes; void* stream;
} drpo_critic_args;

int64_t drpo_critic_workspace_bytes(int64_t batch, int32_t state_dim, int32_t action_dim, int32_t con_dim,
                                    int32_t hidden);
int drpo_critic_step(const drpo_critic_args* args);
/* Debug aids of the DRPO_PREC_BF16 critic step (tests): `rows` != NULL makes the next phase-1 calls write per-row
 * intermediates [batch,16] = a1[2], log-prob, a2[2], Q1', Q2', Qc' sample, Q1, Q2, Qc mean, Qc raw log-std, dL/dQ1, dL/dQ2,
 * dL/dmean, dL/dlogstd (first constraint).  drpo_debug_critic_dw runs the split-K weight-gradient kernel on one operand
 * pair given in the slab-octet layout ([rows_padded/64][features/8][64][8] bf16): out[256, 8*b_octets] = dH^T H. */
int drpo_debug_critic_rows(float* rows);
/* profiling aid: device int64 [24][32] clock stamps per dense op (issuer: operands ready, MMAs issued, per-chunk waits; last epilogue group:
 * accumulator full, epilogue done) of CTA 0's second tile */
int drpo_debug_critic_prof(int64_t* stamps);
int drpo_debug_critic_dw(const void* a_oct, const void* b_oct, int32_t b_octets, int64_t rows_padded, int32_t ksplit,
                         float* partial, float* out, void* stream);
/* Debug aid of the DRPO_PREC_BF16 multiplier / actor steps (tests): per-row intermediates [batch,16] of the next phase-1 calls.
 * multiplier step: a[0], a[A-1], Qc_ub(obs,a), penalty, a_safe[0], safe_Qc, lambda-net output, lambda, dL/d(output).
 * actor step: a_eval[0], safe_Qc, lambda, a[0], log-prob, Q_k, Qc_ub(obs,a), dQ-path dL/da[0], total dL/da[0], dL/dmu[0],
 * dL/d(raw log-std)[0], a_safe'[0], Qc_ub(obs,a_safe'), dL/da_safe'[0], dL/dmu_safe[0], dL/d(raw log-std)_safe[0]. */
int drpo_debug_solver_rows(float* rows);

/* ------------------------------------------------------------------------------------------------------------
 * SSAC.update_multiplier (src/ssac.py:529-578) in DRPO mode (mlp_multiplier): actor rsample, two Qc passes with
 * the Phi^-1(beta) shift mu + std_ratio*sigma (:85), lambda net forward/backward, clip, Adam.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct drpo_multiplier_args {
  const float* obs; int64_t batch_size; int64_t global_batch_size;
  int32_t state_dim, action_dim, con_dim;
  const drpo_mlp3* actor; const drpo_mlp3* actor_safe; const drpo_qc* qc;
  drpo_mlp3 lam;               /* trainable, pointers into params */
  float* params; float* grads; float* adam_m; float* adam_v; int64_t n_params;
  const float* eps_actor; uint64_t seed; uint32_t noise_step; int64_t row_id_offset;
  double std_ratio, constraint_threshold, penalty_lb, penalty_ub, upper_bound, lam_epsilon, grad_norm;
  drpo_adam adam;
  int32_t phases;
  float* losses;               /* device [DRPO_LOSSES_LEN]: loss, grad norm, ..., [15] watchdog flag */
  int32_t precision;
  void* workspace; int64_t workspace_bytes; void* stream;
} drpo_multiplier_args;

int64_t drpo_multiplier_workspace_bytes(int64_t batch, int32_t state_dim, int32_t action_dim, int32_t con_dim,
                                        int32_t hidden);
int drpo_multiplier_step(const drpo_multiplier_args* args);

/* ------------------------------------------------------------------------------------------------------------
 * SSAC.update_actor_and_alpha (src/ssac.py:458-527) in DRPO mode ("next" row 1): performance actor
 * loss mean(alpha*log_prob - Q_k + lambda*Qc_ub) with Q_k = critic.random_choice (:41-43, index chosen by the caller),
 * Qc_ub = max_c(mean + std_ratio*std) (:85, :588-600), lambda = multiplier(obs, safe Qc_ub) (no gradient); temperature
 * loss -alpha*mean(log_prob + target_entropy); safe-actor loss mean(Qc_ub(obs, a_safe')).  Backward through the frozen
 * critics to the actions, the squashed-Gaussian rsample and both policy nets; grad-norm clips; three Adam steps.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct drpo_actor_args {
  const float* obs; int64_t batch_size; int64_t global_batch_size;
  int32_t state_dim, action_dim, con_dim;
  drpo_mlp3 actor; drpo_mlp3 actor_safe;   /* trainable: pointers into params_actor / params_safe */
  const drpo_mlp3* q;                      /* the critic random.choice picked */
  const drpo_qc* qc; const drpo_mlp3* lam; /* frozen */
  float* params_actor; float* grads_actor; float* m_actor; float* v_actor; int64_t n_actor;
  float* params_safe; float* grads_safe; float* m_safe; float* v_safe; int64_t n_safe;
  float* log_alpha; float* alpha_m; float* alpha_v;    /* device scalars */
  /* noise: injected eps_actor [B,A], eps_safe [B,A] (the two rsample draws) or Philox(seed, step) */
  const float* eps_actor; const float* eps_safe; uint64_t seed; uint32_t noise_step; int64_t row_id_offset;
  double std_ratio, multiplier_ub, grad_norm, target_entropy;
  drpo_adam adam_actor, adam_alpha, adam_safe;
  /* phases: bit0 = forward+backward (fills grads_actor, grads_safe, losses[0..2] and losses[5]), bit1 = clips + Adam.
   * Multi-GPU callers all-reduce the two gradient arenas and losses[0..2], losses[5] between the phases. */
  int32_t phases;
  float* losses;   /* device [DRPO_LOSSES_LEN]: actor loss, alpha loss, safe-actor loss, actor grad norm, -, d alpha_loss/d log_alpha, safe grad norm, -, ..., [15] watchdog flag */
  int32_t precision;                       /* DRPO_PREC_FP32; DRPO_PREC_BF16 = fused tcgen05 kernel (hidden 256, S+A <= 64); DRPO_PREC_TF32 = library TF32 GEMMs */
  void* workspace; int64_t workspace_bytes; void* stream;
} drpo_actor_args;

int64_t drpo_actor_workspace_bytes(int64_t batch, int32_t state_dim, int32_t action_dim, int32_t con_dim, int32_t hidden);
int drpo_actor_step(const drpo_actor_args* args);

/* ------------------------------------------------------------------------------------------------------------
 * BatchedGaussianEnsemble.fit's training iteration (src/dynamics.py:143-170; SURVEY.md section 8f "next" row 2): compute_loss =
 * sum over members of the Gaussian NLL on the member's contiguous block of the batch (_rebatch :136-141, a remainder is
 * dropped) + log_var_bound_weight * (sum max_log_var - sum min_log_var); backward; Adam with coupled L2 on every trainable
 * tensor.  Trainable parameters (trunk, both heads, the two log-var bounds) live in ONE flat fp32 arena `params`; `ens`
 * points into it.  phases: bit0 = forward + backward (grads, losses), bit1 = Adam, bit2 = forward only (per-member NLL; with
 * shared_rows != 0 every member scores the same n_rows rows: the holdout ranking at the end of fit, :172-186).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct drpo_ensemble_train_args {
  drpo_ensemble ens;
  float* params; float* grads; float* adam_m; float* adam_v; int64_t n_params;
  const float* states; const float* actions; const float* targets;   /* [n_rows,S], [n_rows,A], [n_rows,S+1] = [next_state, reward] */
  int64_t n_rows; int32_t shared_rows;
  double log_var_bound_weight;
  drpo_adam adam;
  int32_t phases;
  float* losses;               /* device [1 + ensemble_size]: compute_loss, then every member's NLL */
  int32_t precision;           /* DRPO_PREC_FP32 (also what DRPO_PREC_BF16 runs here: no fused kernel for the training iteration); DRPO_PREC_TF32 = library TF32 GEMMs */
  void* workspace; int64_t workspace_bytes; void* stream;
} drpo_ensemble_train_args;

int64_t drpo_ensemble_train_workspace_bytes(const drpo_ensemble* ens, int64_t rows_per_member);
int drpo_ensemble_train_step(const drpo_ensemble_train_args* args);

/* ConstraintCritic.forward (src/ssac.py:64-92): mode 0 -> mean; 1 -> mean + std_ratio*std (uncertainty=True);
 * 2 -> (mean, std, mean + clamp(eps,-2,2)*std) (sample=True).  out_mean/out_std/out_sample are [B,C]. */
int64_t drpo_qc_workspace_bytes(int64_t batch, int32_t hidden);
int drpo_qc_forward(const drpo_qc* qc, const float* states, const float* actions, int64_t batch,
                    int32_t state_dim, int32_t action_dim, int32_t con_dim, int32_t mode, float std_ratio,
                    const drpo_noise* noise, float* out_mean, float* out_std, float* out_sample,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Safety shield: the action-selection block of SMBPO.step_generator (src/smbpo.py:124-136) and of
 * sample_episodes_batched (src/sampling.py:420-439).
 *   shield_type DRPO_SHIELD_NONE   : actions = actor.act(states)
 *               DRPO_SHIELD_SAFE   : actions = where(_get_qc(Qc(s, a_perf)) > threshold, actor_safe.act(s, eval), a_perf)
 *               DRPO_SHIELD_LINEAR : actions = a_safe; for i in 0..10: mix = a_safe*r_i + a_perf*(1-r_i), r_i = (10-i)/10;
 *                                    actions = where(_get_qc(Qc(s, mix)) <= threshold, mix, actions)   (one Qc pass over 11*n rows)
 * eval_perf = 0 samples the performance action with noise_perf (the training step); the safe actor always acts in eval
 * mode (both call sites).  uncertainty != 0 evaluates Qc as mean + std_ratio*std (src/ssac.py:85; the training step passes
 * distributional_qc), else the mean head (the evaluation sampler).
 * Outputs: actions [n,A]; optional qc_perf [n] (_get_qc of the performance action) and choice [n] (SAFE: 1 = safe action
 * taken; LINEAR: the selected i, -1 = none passed and the safe action stands; NONE: 0).
 * ---------------------------------------------------------------------------------------------------------- */
enum { DRPO_SHIELD_NONE = 0, DRPO_SHIELD_SAFE = 1, DRPO_SHIELD_LINEAR = 2 };
typedef struct drpo_shield_args {
  const drpo_mlp3* actor;
  const drpo_mlp3* actor_safe;
  const drpo_qc* qc;
  const float* states;             /* [n,S] */
  int64_t n;
  int32_t state_dim, action_dim, con_dim;
  int32_t shield_type;
  int32_t eval_perf;
  int32_t uncertainty;
  float std_ratio;
  float threshold;
  const drpo_noise* noise_perf;    /* needed when eval_perf == 0 */
  float* actions;                  /* out [n,A] */
  float* qc_perf;                  /* out [n] or NULL */
  int32_t* choice;                 /* out [n] or NULL */
  int32_t path;                    /* 0 = auto: the two-launch latency kernels up to 256 candidate rows (1 state per training
                                      step, 10 evaluation envs), else the batched GEMM path; 1 / 2 force one of them */
  void* workspace; int64_t workspace_bytes;
  void* stream;
} drpo_shield_args;
int64_t drpo_shield_workspace_bytes(const drpo_mlp3* actor, int64_t n, int32_t state_dim, int32_t action_dim, int32_t con_dim,
                                    int32_t hidden);
int drpo_shield_act(const drpo_shield_args* args);

/* misc */
const char* drpo_last_error(void);
int drpo_abi_version(void);
/* Watchdog status of the DRPO_PREC_BF16 (tcgen05) kernels: 0 = ok, else the code of the first in-kernel wait that timed out (a
 * pipeline-protocol bug: the kernel reports it and runs to completion instead of hanging the GPU).  The status is STICKY for the
 * life of the process.  A flagged rollout leaves the ring pointer where it was and reports zero transitions; a flagged update
 * step applies no parameter update (DRPO_LOSS_ERR_SLOT).  drpo_kernel_status synchronises the device first;
 * drpo_kernel_status_peek only reads the pinned words (what has been reported so far).  Message via drpo_last_error(). */
int drpo_kernel_status(void);
int drpo_kernel_status_peek(void);
/* Measurement aid (bench.py's roofline): while enabled, every launch of the fused rollout step kernel is bracketed by CUDA
 * events on the caller's stream, and a third event follows the step's HBM-bound satellites (hooks + ring store, compaction);
 * drpo_timing_read synchronises them and returns the summed kernel duration, the launch count and (optional) the summed
 * duration of the satellites since the last drpo_timing_enable call. */
void drpo_timing_enable(int32_t on);
int drpo_timing_read(double* total_ms_host, int64_t* launches_host, double* satellites_ms_host);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t drpo_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DRPO_B200_H */
