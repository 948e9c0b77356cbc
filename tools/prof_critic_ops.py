"""Profiling driver (not a test): per-op clock stamps of the fused critic kernel (CTA 0, second tile)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drpo_b200
from drpo_b200 import synthetic, _lib
S, A, C = synthetic.WORKLOADS["tracking"][1:]
B = 65536
dev = torch.device("cuda:0")
cfg = drpo_b200.SSAC.Config(); cfg.batch_size = B; cfg.constraint_critic_cfg.std_ratio = 1.0
solver = drpo_b200.SSAC(cfg, S, A, C, 10, 100, 1000, 10, 5.0, device=dev)
solver.load_state_dict(synthetic.make_ssac_weights(43567, S, A, C), strict=False)
solver.precision = drpo_b200.PREC_BF16
batch = [t.to(dev) for t in synthetic.make_critic_batch("tracking", B, 49283)]
for _ in range(2): solver.update_critic(*batch)
stamps = torch.zeros(24, 32, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.drpo_debug_critic_prof(stamps.data_ptr())
solver.update_critic(*batch); torch.cuda.synchronize()
lib.drpo_debug_critic_prof(None)
s = stamps.cpu()
t0 = int(s[0, 0])
names = ["actor0", "actor1", "qt1.0", "qt1.1", "qt2.0", "qt2.1", "safe0", "safe1", "qct.t0", "qct.t1", "qct.m0", "qct.l0",
         "q1.0", "q1.1", "q1.bwd", "q2.0", "q2.1", "q2.bwd", "qc.t0", "qc.t1", "qc.m0", "qc.l0", "qc.dt2", "qc.dt1"]
print("op        ready   issued(+)  accfull3(+)  epi_done(+)   op_total")
for o in range(24):
    r, i = int(s[o, 0]), int(s[o, 1]); f, d = int(s[o, 28]), int(s[o, 30])
    nxt = int(s[o + 1, 0]) if o < 23 else d
    ch = " ".join(f"{int(s[o, 4 + c]) - r:5d}" for c in range(4) if int(s[o, 4 + c]))
    print(f"{names[o]:8s} {r - t0:8d} {i - r:8d} {f - r:10d} {max(d - r, -1):10d} {nxt - r:10d}  {ch}   w@{int(s[o, 8]) - r:6d} " + " ".join(f"g{g}[{int(s[o, 16 + 4 * g]) - r:5d} {int(s[o, 17 + 4 * g]) - r:5d} {int(s[o, 18 + 4 * g]) - r:5d}]" for g in range(4)))
