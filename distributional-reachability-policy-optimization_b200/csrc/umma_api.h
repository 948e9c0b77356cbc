// Interface between the C-ABI translation unit and the tcgen05 (DRPO_PREC_BF16) translation unit.
#pragma once
#include "../../include/drpo_b200.h"
namespace drpo {
int64_t umma_rollout_ws_bytes(const drpo_rollout_args& a);
int umma_rollout(const drpo_rollout_args& a);
int umma_debug_layer(const drpo_rollout_args& a, int layer, float* out);
// optional kernel timing for bench.py's roofline: CUDA events around every fused step-kernel launch
void umma_timing_enable(int on);
int umma_timing_read(double* total_ms, int64_t* launches, double* tail_ms);
}  // namespace drpo
