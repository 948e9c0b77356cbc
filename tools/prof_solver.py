"""Profiling driver (not a test): SSAC actor / multiplier updates at B rows (tracking dims), bf16 fused path (csrc/solver_umma.cu)."""
import os, sys, torch
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
solver.precision = {"bf16": drpo_b200.PREC_BF16, "fp32": drpo_b200.PREC_FP32, "tf32": drpo_b200.PREC_TF32}[os.environ.get("PREC", "bf16")]
obs = synthetic.make_critic_batch("tracking", B, 49283)[0].to(dev)
for name, fn in (("actor", lambda: solver.update_actor_and_alpha(obs)), ("multiplier", lambda: solver.update_multiplier(obs))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name} B={B}: {e0.elapsed_time(e1)/n:.3f} ms/update, losses {out.flatten().tolist()}")
