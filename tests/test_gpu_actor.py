"""drpo_actor_step (SSAC.update_actor_and_alpha, src/ssac.py:458-527 - SURVEY.md section 8f "next" row 1) through the C ABI:
three consecutive updates against the reference's golden vectors, and a larger batch against the oracle (raw gradients,
grad norms, losses, parameters, log_alpha)."""
import pytest
import torch

from oracle import drpo_oracle as O
from tests.util import assert_close, dev, make_ssac, to_dev
from tests.test_gpu_parity import _golden_after

pytestmark = pytest.mark.gpu
RTOL = 1e-5      # fp32 path: "within 1e-5 relative" (BASELINE.json north_star)


def _arena_views(module, arena, prefix):
    from drpo_b200.ssac import _arena_offsets
    named = list(module.named_parameters())
    offs, _ = _arena_offsets([p for _, p in named])
    return {f"{prefix}{k}": arena[o:o + p.numel()].view(p.shape) for (k, p), o in zip(named, offs)}


@pytest.mark.parametrize("tag,S,A,C", [("point_robot", 11, 2, 1), ("cartpole", 4, 1, 4)])
def test_actor_update_vs_golden(golden, tag, S, A, C):
    g = golden("actor")
    w = O.make_ssac_weights(int(g[f"{tag}.seed"]), S, A, C)
    B = g[f"{tag}.it0.obs"].shape[0]
    solver = make_ssac(w, S, A, C, B)
    keys = [k for k in w if k.startswith(("actor.", "actor_safe."))]
    for it in range(3):
        assert solver.actor_optimizer.param_groups[0]["lr"] == pytest.approx(float(g[f"{tag}.lrs"][it]), rel=1e-12)
        losses = solver.update_actor_and_alpha(to_dev(g[f"{tag}.it{it}.obs"]),
                                               noise=(to_dev(g[f"{tag}.it{it}.eps_actor"]), to_dev(g[f"{tag}.it{it}.eps_safe"])),
                                               q_index=int(g[f"{tag}.it{it}.q_index"]))
        assert_close(losses, g[f"{tag}.it{it}.losses"], RTOL, f"losses it{it}")
        assert float(solver.log_alpha) == pytest.approx(float(g[f"{tag}.it{it}.log_alpha_after"]), rel=1e-5)
        _golden_after(solver.state_dict(), g, f"{tag}.it{it}", keys, 2e-5, f"it{it}")


@pytest.mark.parametrize("S,A,C,B,prec,tol", [(51, 2, 1, 4096, "fp32", 5e-5), (12, 2, 2, 1000, "fp32", 5e-5), (51, 2, 1, 4096, "tf32", 2e-2)])
def test_actor_update_vs_oracle(S, A, C, B, prec, tol):
    import drpo_b200
    w = O.make_ssac_weights(71, S, A, C)
    solver = make_ssac(w, S, A, C, B)
    solver.precision = {"fp32": drpo_b200.PREC_FP32, "tf32": drpo_b200.PREC_TF32}[prec]
    wo = {k: v.clone() for k, v in w.items()}
    g = torch.Generator().manual_seed(72)
    obs = torch.randn(B, S, generator=g)
    noise = (torch.randn(B, A, generator=g), torch.randn(B, A, generator=g))
    la = torch.tensor(0.0)
    adams = {k: O.AdamState() for k in ("actor", "alpha", "safe")}
    losses, aux = O.actor_update(wo, obs, noise, O.SSACHyper(), la, 1, C, -float(A), adams, dict(actor=8e-5, alpha=8e-5, safe=8e-5))
    got = solver.update_actor_and_alpha(to_dev(obs), noise=tuple(to_dev(n) for n in noise), q_index=1)
    assert_close(got, torch.stack(losses), max(tol, 2e-5), "losses")
    assert_close(solver._actor_losses[3], aux["grad_norm_actor"], max(tol, 2e-5), "actor grad norm")
    assert_close(solver._actor_losses[6], aux["grad_norm_safe"], max(tol, 2e-5), "safe-actor grad norm")
    assert_close(solver._actor_losses[5], aux["grad_alpha"], max(tol, 2e-5), "d alpha_loss / d log_alpha")
    views = {**_arena_views(solver.actor, solver.actor_optimizer.grad, "actor."),
             **_arena_views(solver.actor_safe, solver.actor_safe_optimizer.grad, "actor_safe.")}
    for k, want in aux["grads_raw"].items():
        # the actor gradient is a batch sum of signed terms that went through two frozen critics: fp32 summation order (split-K
        # on the GPU, torch's blocked sums on the CPU) shows in the entries that cancel: bulk within `tol`, all within 10x
        assert_close(views[k], want, tol, f"grad {k}", max_outlier_frac=5e-2 if prec == "fp32" else 1e-2)
        assert_close(views[k], want, 10 * tol, f"grad {k} (bound)")
    sd = solver.state_dict()
    if prec == "fp32":
        for k in wo:
            if k.startswith(("actor.", "actor_safe.")):
                assert_close(sd[k], wo[k], 2e-5, f"param {k}", max_outlier_frac=2e-3)
        assert float(solver.log_alpha) == pytest.approx(float(la), rel=1e-5)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_rollout_and_update_end_to_end(prec):
    """SMBPO.rollout_and_update (src/smbpo.py:281-291): model rollout into the virtual ring, then critic / actor / multiplier
    updates on minibatches gathered on the device - every step through the C ABI, nothing in eager torch."""
    import drpo_b200
    from drpo_b200 import _lib, synthetic
    lib = _lib.load()
    cfg = drpo_b200.SMBPO.Config()
    cfg.buffer_max, cfg.horizon, cfg.rollout_batch_size, cfg.solver_updates_per_step = 1 << 16, 5, 2000, 10
    cfg.sac_cfg.batch_size = 512
    alg = drpo_b200.SMBPO(cfg, drpo_b200.device_env("quadrotor"), device=dev())
    env_name, S, A, C = synthetic.WORKLOADS["quadrotor"]
    alg.model_ensemble.load_state_dict(synthetic.make_ensemble_weights(64578, S, A), strict=True)
    alg.model_ensemble._elite_inds = [0, 1, 2, 3, 4]
    alg.solver.load_state_dict(synthetic.make_ssac_weights(219803, S, A, C), strict=False)
    p = {"fp32": drpo_b200.PREC_FP32, "bf16": drpo_b200.PREC_BF16}[prec]
    alg.rollout_precision = p; alg.solver.precision = p
    init = synthetic.make_start_states("quadrotor", 4000, 4354).to(dev())
    g = torch.Generator().manual_seed(1)
    n = init.shape[0]
    alg.replay_buffer.extend(states=init, actions=to_dev(torch.rand(n, A, generator=g) * 2 - 1), next_states=init + 0.01,
                             rewards=to_dev(torch.randn(n, generator=g)), dones=to_dev(torch.rand(n, generator=g) < 0.05),
                             violations=to_dev(torch.rand(n, generator=g) < 0.05), constraint_values=to_dev(torch.randn(n, C, generator=g) - 0.5))
    before = {k: v.clone() for k, v in alg.solver.state_dict().items()}
    l0 = lib.drpo_launch_count()
    alg.rollout_and_update()
    torch.cuda.synchronize()
    _lib.check_kernel_status("rollout_and_update")
    assert lib.drpo_launch_count() - l0 > 100
    assert len(alg.virt_buffer) > 2000                    # several rollout steps were stored
    after = alg.solver.state_dict()
    for prefix in ("critic.", "constraint_critic.", "actor.", "actor_safe.", "multiplier."):
        moved = max(float((after[k] - before[k]).abs().max()) for k in after if k.startswith(prefix))
        assert moved > 0, prefix
    assert all(torch.isfinite(v).all() for v in after.values() if v.is_floating_point())
    assert float(alg.solver.log_alpha) != 0.0


def test_actor_row_sharding_matches_full_batch():
    """Phase 1 on two row shards (global normaliser) sums to the full-batch gradients and losses: what the data-parallel
    all-reduce between the two phases of drpo_actor_step relies on."""
    S, A, C, B = 12, 2, 2, 1000
    w = O.make_ssac_weights(5, S, A, C)
    g = torch.Generator().manual_seed(9)
    obs = torch.randn(B, S, generator=g)
    noise = (torch.randn(B, A, generator=g), torch.randn(B, A, generator=g))
    full = make_ssac(w, S, A, C, B)
    lf = full.update_actor_and_alpha(to_dev(obs), noise=tuple(to_dev(n) for n in noise), q_index=0, phases=1)
    ga, gs, galpha = full.actor_optimizer.grad.clone(), full.actor_safe_optimizer.grad.clone(), full._actor_losses[5].clone()
    ta, ts, tl, tal = torch.zeros_like(ga), torch.zeros_like(gs), torch.zeros(3, device=dev()), torch.zeros((), device=dev())
    for lo, hi in ((0, 300), (300, 1000)):
        part = make_ssac(w, S, A, C, hi - lo)
        part._global_batch_override = B
        lp = part.update_actor_and_alpha(to_dev(obs[lo:hi]), noise=tuple(to_dev(n[lo:hi]) for n in noise), q_index=0, phases=1)
        ta += part.actor_optimizer.grad; ts += part.actor_safe_optimizer.grad; tl += lp; tal += part._actor_losses[5]
    assert_close(ta, ga, 5e-5, "actor gradient"); assert_close(ts, gs, 5e-5, "safe-actor gradient")
    assert_close(tl, lf, 2e-5, "losses"); assert_close(tal, galpha, 2e-5, "alpha gradient")
