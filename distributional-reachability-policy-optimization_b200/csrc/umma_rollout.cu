#include "common.cuh"
#include "umma_api.h"
namespace drpo {
int64_t umma_rollout_ws_bytes(const drpo_rollout_args&) { return 0; }
int umma_rollout(const drpo_rollout_args&) { set_error("bf16 rollout not built yet"); return DRPO_ERR_UNSUPPORTED; }
}
