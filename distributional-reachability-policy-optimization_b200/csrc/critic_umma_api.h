// Interface between the C-ABI translation unit and the fused tcgen05 SSAC update steps (critic_umma.cu: critic step;
// solver_umma.cu: multiplier and actor steps).
#pragma once
#include "../../include/drpo_b200.h"
namespace drpo {
namespace cu {
int64_t critic_ws_bytes(int64_t B, int S, int A, int C);
// phase 1 (forward, losses, backward) of drpo_critic_step in DRPO_PREC_BF16: fills args.grads and args.losses[0..1]
int critic_phase1(const drpo_critic_args& a, int* err_flag);
// tests: per-row intermediates [B,16] written by the next critic_phase1 (NULL = off)
void critic_set_debug_rows(float* p);
void critic_set_prof(long long* p);     // profiling aid: per-op clock stamps [24][4] of block 0, second tile
// tests: dW[256, 8*b_octets] of one (dH, H) operand pair in the octet layout
int critic_debug_dw(const void* a_oct, const void* b_oct, int b_octets, int64_t rows_padded, int ksplit, float* partial, float* out,
                    int* err_flag, void* stream);
// DRPO_PREC_BF16 multiplier / actor steps (solver_umma.cu).  mode: 0 = multiplier step, 1 = actor step
int64_t solver_ws_bytes(int mode, int64_t B, int S, int A, int C);
// phase 1 of drpo_multiplier_step: fills args.grads and args.losses[0]
int multiplier_phase1(const drpo_multiplier_args& a, int* err_flag);
// phase 1 of drpo_actor_step: fills args.grads_actor, args.grads_safe and args.losses[0..2], [5]
int actor_phase1(const drpo_actor_args& a, int* err_flag);
// tests: per-row intermediates [B,16] written by the next multiplier_phase1 / actor_phase1 (NULL = off)
void solver_set_debug_rows(float* p);
// DRPO_PREC_BF16 ensemble forward / sample outside the rollout (ens_umma.cu): members m0..m1-1 on `batch` rows; noise == NULL:
// means / log_vars [m1-m0, batch, S+1]; noise != NULL (one member): next_states [batch,S], rewards [batch]
bool ens_bf16_supported(const drpo_ensemble& e);
int64_t ens_bf16_ws_bytes(const drpo_ensemble& e);
int ens_bf16_run(const drpo_ensemble& e, int m0, int m1, int per_member_inputs, const float* states, const float* actions, int64_t batch,
                 float* means, float* log_vars, const drpo_noise* noise, float* next_states, float* rewards, void* workspace,
                 int64_t workspace_bytes, void* stream, int* err_flag);
}  // namespace cu
}  // namespace drpo
