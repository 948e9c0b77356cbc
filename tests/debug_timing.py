"""Debug aid (not a test): per-slab clock stamps of CTA 0 for the bf16 rollout step kernel."""
import sys, torch
sys.path.insert(0, "/root/repo")
import drpo_b200
from drpo_b200 import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 4
S, A, C = 12, 2, 2
dev = torch.device("cuda:0")
cfg = drpo_b200.SMBPO.Config(); cfg.rollout_batch_size, cfg.horizon, cfg.buffer_max = B, 1, B * 2
alg = drpo_b200.SMBPO(cfg, drpo_b200.device_env("quadrotor"), device=dev)
alg.model_ensemble.load_state_dict(synthetic.make_ensemble_weights(1, S, A)); alg.model_ensemble._elite_inds = [0, 1, 2, 3, 4]
alg.solver.load_state_dict(synthetic.make_ssac_weights(2, S, A, C), strict=False)
alg.rollout_precision = drpo_b200.PREC_BF16
init = synthetic.make_start_states("quadrotor", B, 3).to(dev)
for _ in range(2):
    out = alg.rollout(alg.actor, initial_states=init, member_idx=[0], _debug_layer=100)
torch.cuda.synchronize()
st = out.flatten().view(torch.int32).cpu().numpy().astype("int64").reshape(-1, 8)[:128].reshape(4, 32, 8)
names = ["P1"] * 4 + ["P2"] * 4 + ["P3"] + ["M1"] * 4 + ["M2"] * 4 + ["D0"] * 4 + ["L0"] * 4 + ["D1", "L1"]
for tile in range(1, 3):
    t0 = st[tile, 0, 6]
    print(f"tile {tile}: prologue {st[tile,0,7]-t0} cycles; next tile's prologue starts at {st[tile+1,0,6]-t0}")
    print(" chunk      loop_top drained weights token issued | epi_wait acc_ready epi_done")
    for c in range(27):
        r = st[tile, c] - t0
        top = r[6] if c > 0 else -1
        wts = r[7] if c > 0 else -1
        print(f" {c:2d} {names[c]:3s} {top:8d} {r[3]:8d} {wts:8d} {r[4]:8d} {r[5]:8d} | {r[0]:8d} {r[1]:8d} {r[2]:8d}")
