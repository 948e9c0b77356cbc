import sys, time, torch
sys.path.insert(0, "/root/repo")
import drpo_b200
from drpo_b200 import synthetic
S, A, C = 12, 2, 2
dev = torch.device("cuda:0")
cfg = drpo_b200.SSAC.Config()
solver = drpo_b200.SSAC(cfg, S, A, C, 10, 100, 300, 10, 10.0, device=dev)
solver.load_state_dict(synthetic.make_ssac_weights(3, S, A, C), strict=False)
s1 = torch.randn(1, S, device=dev); s10 = torch.randn(10, S, device=dev)
def t(fn, k=500):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): r = fn()
    if torch.is_tensor(r): r.cpu()
    return (time.perf_counter() - t0) / k * 1e6
print("ensure_arenas us", t(lambda: solver._ensure_arenas()))
print("shield 1 state us", t(lambda: solver.shield_act(s1, eval=False, shield_type="safe", uncertainty=True)))
print("shield linear 10 us", t(lambda: solver.shield_act(s10, eval=True, shield_type="linear")))
print("policy.act 1 us", t(lambda: solver.actor.act(s1, True)))
