"""Profiling driver (not a test): standalone ensemble forward (all 7 members, shared inputs = BatchedGaussianEnsemble.means) at
B rows, fp32 FFMA path vs the bf16 fused tcgen05 member chain (csrc/ens_umma.cu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drpo_b200
from drpo_b200 import synthetic
wl = sys.argv[1] if len(sys.argv) > 1 else "quadrotor"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
_, S, A, C = synthetic.WORKLOADS[wl]
dev = torch.device("cuda:0")
ens = drpo_b200.BatchedGaussianEnsemble(drpo_b200.BatchedGaussianEnsemble.Config(), S, A, device=dev)
ens.load_state_dict(synthetic.make_ensemble_weights(64578, S, A), strict=True)
s = synthetic.make_start_states(wl, B, 4354).to(dev)
a = torch.rand(B, A, device=dev) * 2 - 1
flops = 2 * (200 * (S + A) + 3 * 200 * 200 + 2 * 200 * (S + 1)) * 7 * B
for name, prec in (("fp32", drpo_b200.PREC_FP32), ("bf16", drpo_b200.PREC_BF16)):
    ens.forward_precision = prec
    for _ in range(3): ens.means(s, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ens.means(s, a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"ensemble.means {wl} B={B} x 7 members [{name}]: {ms:.3f} ms, {flops / ms / 1e9:.1f} TFLOP/s algorithmic")
