"""Profiling aid (not a test): clock stamps of CTA 0 of the two-tiles-in-flight rollout step kernel (rollout_pipe.cuh, debug build):
per slot the issue time line of the nine jobs, the six hidden epilogues and the output group, for the CTA's 2nd and 3rd iteration."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drpo_b200
from drpo_b200 import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 2 * 4
wl = sys.argv[2] if len(sys.argv) > 2 else "quadrotor"
env_name, S, A, C = synthetic.WORKLOADS[wl]
dev = torch.device("cuda:0")
cfg = drpo_b200.SMBPO.Config(); cfg.rollout_batch_size, cfg.horizon, cfg.buffer_max = B, 1, B * 2
alg = drpo_b200.SMBPO(cfg, drpo_b200.device_env(env_name), device=dev)
alg.model_ensemble.load_state_dict(synthetic.make_ensemble_weights(1, S, A)); alg.model_ensemble._elite_inds = [0, 1, 2, 3, 4]
alg.solver.load_state_dict(synthetic.make_ssac_weights(2, S, A, C), strict=False)
alg.rollout_precision = drpo_b200.PREC_BF16
init = synthetic.make_start_states(wl, B, 3).to(dev)
for _ in range(2):
    out = alg.rollout(alg.actor, initial_states=init, member_idx=[0], _debug_layer=100)
torch.cuda.synchronize()
st = out.flatten().view(torch.int32).cpu().numpy().astype("int64")[:4 * 64 * 8].reshape(4, 64, 8)
jobs = ["P0", "P1", "P2", "T0", "T1", "D0", "D1", "V0", "V1"]
hid = ["P0", "P1", "T0", "T1", "D0", "V0"]
for it in (1, 2):
    t0 = st[it, 0, 0]
    nxt = st[it + 1, 0, 0] - t0 if it < 3 else -1
    print(f"=== iteration {it} (cycles relative to slot 0 / P0 waits done); next iteration's P0 at {nxt}")
    print(" job   slot: waits_done weights_there issued | cycles waiting for weights")
    for j, name in enumerate(jobs):
        for s in (0, 1):
            r = st[it, s * 32 + j] - t0
            print(f"  {name} s{s}: {r[0]:8d} {r[1]:8d} {r[2]:8d} | {st[it, s * 32 + j][3]:6d}")
    print(" hidden epilogue / slot: wait_begin acc_full chunk_done published")
    for i, name in enumerate(hid):
        for s in (0, 1):
            r = st[it, s * 32 + 10 + i] - t0
            print(f"  {name} s{s}: {r[4]:8d} {r[5]:8d} {r[7]:8d} {r[6]:8d}")
    for s in (0, 1):
        r = st[it, s * 32 + 20] - t0
        print(f" output group s{s}: head_full {r[0]} xm_published {r[1]} next_prologue_done {r[2]} diff_full {r[3]} lvar_full {r[4]} stores_done {r[5]}")
