"""Profiling aid (not a test): clock stamps of CTA 0 for the fused rollout step kernel (rollout_fused.cuh, debug build): the issue
time line of every weight block, the four hidden-epilogue groups per hidden layer and the output group, for the CTA's 2nd and 3rd tile."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
st = out.flatten().view(torch.int32).cpu().numpy().astype("int64")[:4 * 64 * 8].reshape(4, 64, 8)
bl = ["L0", "L1b", "L1a0", "L1a1", "L1a2", "L1a3", "L2", "T0", "T1a0", "T1a1", "T1a2", "T1a3", "D0a0", "D0a1", "D0a2", "D0a3",
      "V0a0", "V0a1", "V0a2", "V0a3", "D1", "V1"]
hl = ["L0", "L1", "T0", "T1", "D0", "V0"]
for tile in (1, 2):
    t0 = st[tile, 0, 0]
    print(f"=== tile {tile} (cycles relative to L0's weights ready) ; next tile's L0 at {st[tile + 1, 0, 0] - t0 if tile < 3 else -1}")
    print(" block     ring_full  cnt_done  token  issued")
    for b, name in enumerate(bl):
        r = st[tile, b] - t0
        print(f"  {b:2d} {name:5s} {r[0]:8d} {r[1]:8d} {r[2]:8d} {r[3]:8d}")
    print(" hidden layer / group: wait_begin  acc_full  drained  act_published")
    for lh, name in enumerate(hl):
        for j in range(4):
            r = st[tile, 32 + 4 * lh + j] - t0
            print(f"  {name:3s} g{j} {r[4]:8d} {r[5]:8d} {r[6]:8d} {r[7]:8d}")
    r = st[tile, 60] - t0
    print(f" output group: head_full {r[0]} xm_published {r[1]} next_prologue_done {r[2]} diff_full {r[3]} lvar_full {r[4]} stores_done {r[5]}")
