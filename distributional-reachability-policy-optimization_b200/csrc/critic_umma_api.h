// Interface between the C-ABI translation unit and the fused tcgen05 critic step (critic_umma.cu).
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
}  // namespace cu
}  // namespace drpo
