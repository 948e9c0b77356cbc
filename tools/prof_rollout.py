"""Profiling driver (not a test): a few bf16 rollout steps of the quadrotor workload, sized for ncu."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drpo_b200
from drpo_b200 import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 8
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S, A, C = 12, 2, 2
dev = torch.device("cuda:0")
cfg = drpo_b200.SMBPO.Config(); cfg.rollout_batch_size, cfg.horizon, cfg.buffer_max = B, H, B * H + 1024
alg = drpo_b200.SMBPO(cfg, drpo_b200.device_env("quadrotor"), device=dev)
alg.model_ensemble.load_state_dict(synthetic.make_ensemble_weights(1, S, A)); alg.model_ensemble._elite_inds = [0, 1, 2, 3, 4]
alg.solver.load_state_dict(synthetic.make_ssac_weights(2, S, A, C), strict=False)
alg.rollout_precision = drpo_b200.PREC_BF16
init = synthetic.make_start_states("quadrotor", B, 3).to(dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(4):
    alg.virt_buffer._pointer.zero_()
    if it == 3: ev0.record()
    view = alg.rollout(alg.actor, initial_states=init, member_idx=[i % 5 for i in range(H)])
ev1.record(); torch.cuda.synchronize()
print(f"rollout B={B} H={H}: {ev0.elapsed_time(ev1):.3f} ms, counts {view.counts()[:3]}...")
