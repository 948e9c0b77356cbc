"""Profiling driver (not a test): SSAC critic updates at B = 65536 (tracking dims), tensor-core mode."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drpo_b200
from drpo_b200 import synthetic
S, A, C = synthetic.WORKLOADS["tracking"][1:]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
cfg = drpo_b200.SSAC.Config(); cfg.batch_size = B; cfg.constraint_critic_cfg.std_ratio = 1.0
solver = drpo_b200.SSAC(cfg, S, A, C, 10, 100, 1000, 10, 5.0, device=dev)
solver.load_state_dict(synthetic.make_ssac_weights(43567, S, A, C), strict=False)
solver.precision = drpo_b200.PREC_BF16 if os.environ.get("PREC", "bf16") == "bf16" else drpo_b200.PREC_FP32
batch = [t.to(dev) for t in synthetic.make_critic_batch("tracking", B, 49283)]
for _ in range(3): solver.update_critic(*batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): lq, lc = solver.update_critic(*batch)
e1.record(); torch.cuda.synchronize()
print(f"critic B={B}: {e0.elapsed_time(e1)/n:.3f} ms/update, loss_q {float(lq):.4f} loss_c {float(lc):.4f}")
