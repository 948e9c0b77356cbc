"""Profiling aid (not a test): host time of SSAC.update_critic split into the C entry point (table building + launches) and the Python
wrapper around it, at a small shard (device time hidden: the host runs ahead until the queue fills)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drpo_b200
from drpo_b200 import synthetic, _lib
S, A, C = synthetic.WORKLOADS["tracking"][1:]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")
cfg = drpo_b200.SSAC.Config(); cfg.batch_size = B; cfg.constraint_critic_cfg.std_ratio = 1.0
solver = drpo_b200.SSAC(cfg, S, A, C, 10, 100, 1000, 10, 5.0, device=dev)
solver.load_state_dict(synthetic.make_ssac_weights(43567, S, A, C), strict=False)
solver.precision = drpo_b200.PREC_BF16
batch = [t.to(dev) for t in synthetic.make_critic_batch("tracking", B, 49283)]
lib = _lib.load()
orig = lib.drpo_critic_step
acc = [0.0, 0]
def timed(a):
    t0 = time.perf_counter(); r = orig(a); acc[0] += time.perf_counter() - t0; acc[1] += 1; return r
lib.drpo_critic_step = timed
for _ in range(20): solver.update_critic(*batch)
torch.cuda.synchronize(); acc[0] = 0.0; acc[1] = 0
n, host = 25, 0.0
for i in range(n):
    torch.cuda.synchronize()                       # empty queue: the 8 calls below measure pure enqueue cost, no back-pressure
    t0 = time.perf_counter()
    for _ in range(8):
        solver.update_critic(*batch)
    host += time.perf_counter() - t0
torch.cuda.synchronize()
print(f"update_critic B={B}: {1e6 * host / (8 * n):.1f} us of host time per call, of which the C entry point {1e6 * acc[0] / acc[1]:.1f} us "
      f"(table building + {lib.drpo_launch_count() // max(acc[1] + 20, 1)} launches)")
