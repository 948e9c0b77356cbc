"""Profiling driver (not a test): bf16 rollout of a small workload (cartpole-move 100k x 10) to expose fixed per-rollout costs."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drpo_b200
from drpo_b200 import synthetic
wl = sys.argv[1] if len(sys.argv) > 1 else "cartpole-move"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
H = 10
env_name, S, A, C = synthetic.WORKLOADS[wl]
dev = torch.device("cuda:0")
cfg = drpo_b200.SMBPO.Config(); cfg.rollout_batch_size, cfg.horizon, cfg.buffer_max = B, H, B * H + 1024
alg = drpo_b200.SMBPO(cfg, drpo_b200.device_env(env_name), device=dev)
alg.model_ensemble.load_state_dict(synthetic.make_ensemble_weights(1, S, A)); alg.model_ensemble._elite_inds = [0, 1, 2, 3, 4]
alg.solver.load_state_dict(synthetic.make_ssac_weights(2, S, A, C), strict=False)
alg.rollout_precision = drpo_b200.PREC_BF16
init = synthetic.make_start_states(wl, B, 3).to(dev)
for it in range(5):
    alg.virt_buffer._pointer.zero_()
    view = alg.rollout(alg.actor, initial_states=init, member_idx=[i % 5 for i in range(H)])
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); ev0.record()
for it in range(10):
    alg.virt_buffer._pointer.zero_()
    view = alg.rollout(alg.actor, initial_states=init, member_idx=[i % 5 for i in range(H)])
t1 = time.perf_counter(); ev1.record(); torch.cuda.synchronize()
print(f"{wl} B={B}: {ev0.elapsed_time(ev1)/10:.3f} ms per rollout (host enqueue {1e3*(t1-t0)/10:.3f} ms), counts {view.counts()[:3]}")
