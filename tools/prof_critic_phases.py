"""Profiling driver (not a test): where one SSAC critic update goes at a per-GPU shard of B rows (tracking dims, bf16 fused path).
CUDA events around phase 1 (pack + fused kernel + dW + gradient assembly), the gradient message (an all-reduce when launched under
torchrun, else skipped) and phase 2 (norms, clip, Adam, EMA); host time = wall clock of the un-synchronised Python call.
  python tools/prof_critic_phases.py [B=8192] [n=50]           (single GPU: the shard a rank sees at N = 65536 / B)
  torchrun --nproc-per-node N tools/prof_critic_phases.py ...   (adds the NCCL all-reduce of the 1.4 MB gradient arena)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drpo_b200
from drpo_b200 import synthetic

S, A, C = synthetic.WORKLOADS["tracking"][1:]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
cfg = drpo_b200.SSAC.Config(); cfg.batch_size = B; cfg.constraint_critic_cfg.std_ratio = 1.0
solver = drpo_b200.SSAC(cfg, S, A, C, 10, 100, 1000, 10, 5.0, device=dev)
solver.load_state_dict(synthetic.make_ssac_weights(43567, S, A, C), strict=False)
solver.precision = drpo_b200.PREC_BF16
solver._global_batch_override = B * world
full = synthetic.make_critic_batch("tracking", B * world, 49283)
batch = [t[rank * B:(rank + 1) * B].to(dev) for t in full]
grad = solver.critic_optimizer.grad_full if hasattr(solver.critic_optimizer, "grad_full") else solver.critic_optimizer.grad


def ev():
    return torch.cuda.Event(enable_timing=True)


for _ in range(5):
    solver.update_critic(*batch, phases=1)
    if world > 1:
        dist.all_reduce(grad)
    solver.update_critic(*batch, phases=2)
torch.cuda.synchronize()
acc = [0.0, 0.0, 0.0]
host = [0.0, 0.0, 0.0]
for _ in range(n):
    e = [ev() for _ in range(4)]
    t0 = time.perf_counter(); e[0].record()
    solver.update_critic(*batch, phases=1)
    e[1].record(); t1 = time.perf_counter()
    if world > 1:
        dist.all_reduce(grad)
    e[2].record(); t2 = time.perf_counter()
    solver.update_critic(*batch, phases=2)
    e[3].record(); t3 = time.perf_counter()
    torch.cuda.synchronize()
    for i in range(3):
        acc[i] += e[i].elapsed_time(e[i + 1])
    host[0] += (t1 - t0) * 1e3; host[1] += (t2 - t1) * 1e3; host[2] += (t3 - t2) * 1e3
# back-to-back updates (what bench.py times): device time per update when the host runs ahead
torch.cuda.synchronize()
e0, e1 = ev(), ev()
solver.data_parallel = world > 1
t0 = time.perf_counter(); e0.record()
for _ in range(n):
    solver.update_critic(*batch)
e1.record(); th = time.perf_counter() - t0
torch.cuda.synchronize()
if rank == 0:
    print(f"critic shard B={B} (world {world}): phase1 {acc[0]/n*1e3:.0f} us | all-reduce {acc[1]/n*1e3:.0f} us | phase2 {acc[2]/n*1e3:.0f} us  (device, synchronised per update)")
    print(f"   host (enqueue) time: phase1 {host[0]/n*1e3:.0f} us | all-reduce {host[1]/n*1e3:.0f} us | phase2 {host[2]/n*1e3:.0f} us")
    print(f"   back to back: {e0.elapsed_time(e1)/n*1e3:.0f} us per update on the device, {th/n*1e6:.0f} us of host time per update")
if world > 1:
    dist.destroy_process_group()
