"""Parity of the CUDA path (through the C ABI) against the oracle and the golden vectors generated from the reference.
Tolerances (BASELINE.json north_star): integer/bool masks and transition counts bit-exact; fp32 path values within 1e-5
relative; bf16 tensor-core path within 2e-2."""
import numpy as np
import pytest
import torch

from oracle import drpo_oracle as O
from tests.util import assert_close, dev, make_ensemble, make_ssac, oracle_spec_to_device_env, to_dev

pytestmark = pytest.mark.gpu
T = torch.from_numpy
RTOL = 1e-5

SPECS = {"point_robot": O.env_point_robot(), "cartpole": O.env_cartpole(), "quadrotor": O.env_quadrotor(),
         "tracking1": O.env_tracking(10, 1), "tracking4": O.env_tracking(10, 4), "safetygym60": O.env_safetygym60()}


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["point_robot", "cartpole", "quadrotor"])
def test_hooks_bit_exact_vs_golden(golden, tag):
    g = golden("hooks")
    env = oracle_spec_to_device_env(SPECS[tag])
    d, v, cv = env.evaluate(to_dev(g[f"{tag}.states"]))
    assert np.array_equal(d.cpu().numpy(), g[f"{tag}.done"])
    assert np.array_equal(v.cpu().numpy(), g[f"{tag}.viol"])
    want = g[f"{tag}.cv"]
    got = cv.cpu().numpy().reshape(want.shape)
    assert np.array_equal(got.view(np.uint32) | (np.isnan(got) * 0xFFFFFFFF).astype(np.uint32),
                          want.view(np.uint32) | (np.isnan(want) * 0xFFFFFFFF).astype(np.uint32)), "constraint values not bit-identical"


@pytest.mark.parametrize("tag", ["tracking1", "tracking4"])
def test_hooks_tracking(golden, tag):
    """The tracking hook uses numpy's fp32 sin/cos, which CUDA's sinf/cosf do not reproduce bit for bit:
    values within 1e-5, masks equal wherever |cv| exceeds that margin; done is exact (comparisons only)."""
    g = golden("hooks")
    env = oracle_spec_to_device_env(SPECS[tag])
    d, v, cv = env.evaluate(to_dev(g[f"{tag}.states"]))
    assert np.array_equal(d.cpu().numpy(), g[f"{tag}.done"])
    want = g[f"{tag}.cv"]
    got = cv.cpu().numpy()
    assert np.abs(got - want).max() <= 1e-5
    clear = np.abs(want) > 1e-5
    assert np.array_equal(v.cpu().numpy()[clear], g[f"{tag}.viol"][clear])


@pytest.mark.parametrize("tag", ["point_robot", "cartpole", "quadrotor"])
def test_hooks_random_large(tag):
    """1M random states around the constraint boundaries: masks bit-exact against the oracle."""
    spec = SPECS[tag]
    gen = np.random.RandomState(5)
    s = (gen.randn(1 << 20, spec.state_dim) * 1.5).astype(np.float32)
    env = oracle_spec_to_device_env(spec)
    d, v, cv = env.evaluate(to_dev(s))
    od, ov, ocv = O.hooks(spec, s)
    assert np.array_equal(d.cpu().numpy(), od) and np.array_equal(v.cpu().numpy(), ov)
    assert np.array_equal(cv.cpu().numpy().reshape(ocv.shape), ocv)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,S,A", [("point_robot", 11, 2), ("cartpole", 4, 1), ("quadrotor", 12, 2)])
def test_ensemble_vs_golden(golden, tag, S, A):
    g = golden("ensemble")
    w = O.make_ensemble_weights(int(g[f"{tag}.seed"]), S, A)
    ens = make_ensemble(w, S, A)
    s, a, eps = to_dev(g[f"{tag}.states"]), to_dev(g[f"{tag}.actions"]), to_dev(g[f"{tag}.eps"])
    m, lv = ens._forward1(s, a, int(g[f"{tag}.member"]))
    assert_close(m, g[f"{tag}.means"], RTOL, "means"); assert_close(lv, g[f"{tag}.log_vars"], RTOL, "log_vars")
    ens._elite_inds = [int(g[f"{tag}.member"])]
    ns, r = ens.sample(s, a, eps=eps)
    assert_close(ns, g[f"{tag}.next_states"], RTOL, "next_states"); assert_close(r, g[f"{tag}.rewards"], RTOL, "rewards")
    ms, mr = ens.means(s, a)
    assert_close(ms, g[f"{tag}.means_all_s"], RTOL, "means()"); assert_close(mr, g[f"{tag}.means_all_r"], RTOL, "means() r")
    ens._elite_inds = [6, 0, 2, 5, 1]
    es, er = ens.elite_samples(s, a, eps=to_dev(g[f"{tag}.eps_elite"]))
    assert_close(es, g[f"{tag}.elite_s"], RTOL, "elite_samples"); assert_close(er, g[f"{tag}.elite_r"], RTOL, "elite r")
    # _forward_all with per-member inputs
    sE = s.repeat(7, 1, 1) + torch.arange(7, device=s.device).view(7, 1, 1) * 0.01
    aE = a.repeat(7, 1, 1)
    mE, lvE = ens._forward_all(sE, aE)
    om, olv = O.ensemble_forward_all(w, sE.cpu(), aE.cpu())
    assert_close(mE, om, RTOL, "_forward_all means"); assert_close(lvE, olv, RTOL, "_forward_all log_vars")


def test_ensemble_ragged_sizes():
    """Batch sizes that are not tile multiples, including 1 and 0."""
    S, A = 12, 2
    w = O.make_ensemble_weights(7, S, A)
    ens = make_ensemble(w, S, A)
    g = torch.Generator().manual_seed(1)
    for B in (0, 1, 63, 65, 1000):
        s, a = torch.randn(B, S, generator=g), torch.rand(B, A, generator=g) * 2 - 1
        m, lv = ens._forward1(to_dev(s), to_dev(a), 2)
        om, olv = O.ensemble_forward1(w, s, a, 2)
        assert_close(m, om, RTOL, f"means B={B}"); assert_close(lv, olv, RTOL, f"lv B={B}")


@pytest.mark.parametrize("tag,S,A,C", [("point_robot", 11, 2, 1), ("cartpole", 4, 1, 4)])
def test_policy_vs_golden(golden, tag, S, A, C):
    g = golden("policy")
    w = O.make_ssac_weights(int(g[f"{tag}.seed"]), S, A, C)
    solver = make_ssac(w, S, A, C, 64)
    s, eps = to_dev(g[f"{tag}.states"]), to_dev(g[f"{tag}.eps"])
    a, lp = solver.actor.act_with_log_prob(s, eval=False, eps=eps, want_log_prob=True)
    assert_close(a, g[f"{tag}.actions"], RTOL, "actions"); assert_close(lp, g[f"{tag}.log_prob"], RTOL, "log_prob")
    assert_close(solver.actor_safe.act(s, eval=True), g[f"{tag}.eval_actions_safe"], RTOL, "eval actions")


def test_qc_forward_modes():
    S, A, C = 12, 2, 2
    w = O.make_ssac_weights(11, S, A, C)
    solver = make_ssac(w, S, A, C, 64)
    g = torch.Generator().manual_seed(2)
    s, a, eps = torch.randn(300, S, generator=g), torch.rand(300, A, generator=g) * 2 - 1, torch.randn(300, C, generator=g) * 1.5
    mu, sd = O.qc_forward(w, "constraint_critic.", s, a)
    qc = solver.constraint_critic
    assert_close(qc(to_dev(s), to_dev(a)), mu, RTOL, "mean")
    assert_close(qc(to_dev(s), to_dev(a), uncertainty=True), mu + 2.0 * sd, RTOL, "uncertainty shift")
    m2, s2, smp = qc(to_dev(s), to_dev(a), sample=True, eps=to_dev(eps))
    assert_close(m2, mu, RTOL); assert_close(s2, sd, RTOL); assert_close(smp, mu + eps.clamp(-2, 2) * sd, RTOL, "sample")


# ---------------------------------------------------------------------------------------------------------------
def _run_rollout(alg, init, H, eps_p, eps_m, members):
    alg.horizon = H
    view = alg.rollout(alg.actor, initial_states=to_dev(init), noise=(to_dev(eps_p), to_dev(eps_m)), member_idx=members)
    torch.cuda.synchronize()
    return view


def _make_alg(spec, wm, ws, B, capacity=None):
    import drpo_b200
    cfg = drpo_b200.SMBPO.Config()
    cfg.rollout_batch_size = B
    cfg.buffer_max = capacity or max(B * 16, 4096)
    env = oracle_spec_to_device_env(spec)
    env.action_dim = wm["trunk.0.weight"].shape[2] - spec.state_dim
    alg = drpo_b200.SMBPO(cfg, env, device=dev())
    alg.model_ensemble.load_state_dict(wm, strict=True)
    alg.model_ensemble._elite_inds = [0, 1, 2, 3, 4]
    alg.solver.load_state_dict(ws, strict=False)
    return alg


def test_rollout_vs_golden(golden):
    """The reference's own SMBPO.rollout on PointRobot (golden) vs drpo_rollout: counts and masks bit-exact."""
    g = golden("rollout")
    wm = O.make_ensemble_weights(int(g["seed_model"]), 11, 2, diff_scale=float(g["diff_scale"]))
    ws = O.make_ssac_weights(int(g["seed_ssac"]), 11, 2, 1)
    H = len(g["members"])
    alg = _make_alg(O.env_point_robot(), wm, ws, g["init"].shape[0])
    view = _run_rollout(alg, T(g["init"]), H, T(g["eps_policy"]), T(g["eps_model"]), [int(m) for m in g["members"]])
    assert view.counts() == g["counts"].tolist()
    assert len(view) == int(g["counts"].sum()) == len(alg.virt_buffer)
    out = view.get(as_dict=True)
    assert np.array_equal(out["dones"].cpu().numpy(), g["out.dones"])
    assert np.array_equal(out["violations"].cpu().numpy(), g["out.violations"])
    for k in ("states", "actions", "next_states", "rewards", "constraint_values"):
        assert_close(out[k], g[f"out.{k}"], RTOL, k)


@pytest.mark.parametrize("tag,S,A,C,B", [("cartpole", 4, 1, 4, 3000), ("quadrotor", 12, 2, 2, 5000), ("point_robot", 11, 2, 1, 2500),
                                         ("tracking1", 51, 2, 1, 1500), ("safetygym60", 60, 2, 1, 1500)])
def test_rollout_vs_oracle(tag, S, A, C, B):
    """Free-running H=10 rollout with injected noise: per-step alive counts and total transition count equal, masks
    bit-exact, values within 1e-5; a divergence would have to be explained by a sub-tolerance margin at a boundary."""
    spec, H = SPECS[tag], 10
    wm = O.make_ensemble_weights(31, S, A, diff_scale=0.05)
    ws = O.make_ssac_weights(32, S, A, C)
    g = torch.Generator().manual_seed(33)
    init = torch.randn(B, S, generator=g) * 0.3
    if tag == "quadrotor":
        init[:, 2] = 0.55 + 0.9 * torch.rand(B, generator=g)
    eps_p, eps_m = torch.randn(H, B, A, generator=g), torch.randn(H, B, S + 1, generator=g)
    members = [int(x) for x in torch.randint(0, 5, (H,), generator=g)]
    ref, counts, _ = O.rollout(ws, wm, spec, init, H, eps_p, eps_m, members)
    alg = _make_alg(spec, wm, ws, B)
    view = _run_rollout(alg, init, H, eps_p, eps_m, members)
    assert view.counts()[:len(counts)] == counts and sum(view.counts()) == sum(counts)
    out = view.get(as_dict=True)
    assert torch.equal(out["dones"].cpu(), ref["dones"]) and torch.equal(out["violations"].cpu(), ref["violations"])
    for k in ("states", "actions", "next_states", "rewards", "constraint_values"):
        assert_close(out[k], ref[k], RTOL, f"{tag}.{k}")


def test_rollout_ring_wrap_and_append():
    """Two rollouts into a small ring whose pointer wraps (SampleBuffer.extend semantics, src/sampling.py:128-145)."""
    S, A, C, B, H = 4, 1, 4, 500, 4
    spec = SPECS["cartpole"]
    wm, ws = O.make_ensemble_weights(41, S, A, diff_scale=0.05), O.make_ssac_weights(42, S, A, C)
    g = torch.Generator().manual_seed(43)
    cap = B * H + 100
    alg = _make_alg(spec, wm, ws, B, capacity=cap)
    refs = []
    for it in range(3):
        init = torch.randn(B, S, generator=g) * 0.1
        eps_p, eps_m = torch.randn(H, B, A, generator=g), torch.randn(H, B, S + 1, generator=g)
        members = [1, 0, 3, 2]
        ref, counts, _ = O.rollout(ws, wm, spec, init, H, eps_p, eps_m, members)
        view = _run_rollout(alg, init, H, eps_p, eps_m, members)
        assert sum(view.counts()) == sum(counts)
        out = view.get(as_dict=True)
        assert torch.equal(out["dones"].cpu(), ref["dones"])
        assert_close(out["next_states"], ref["next_states"], RTOL)
        refs.append(ref)
    total = sum(len(r["rewards"]) for r in refs)
    assert int(alg.virt_buffer._pointer) == total and total > cap          # wrapped
    allr = torch.cat([r["rewards"] for r in refs])
    assert_close(alg.virt_buffer.get("rewards"), allr[-cap:], RTOL, "ring contents after wrap")


def test_rollout_all_done_and_philox():
    """(a) every trajectory terminates at step 0 -> later steps store nothing; (b) Philox mode: the oracle fed with the
    exported noise stream reproduces the rollout, and the result does not depend on how start states are sharded."""
    import drpo_b200
    from drpo_b200 import _lib
    S, A, C, B, H = 4, 1, 4, 700, 5
    spec = SPECS["cartpole"]
    wm, ws = O.make_ensemble_weights(51, S, A, diff_scale=0.05), O.make_ssac_weights(52, S, A, C)
    alg = _make_alg(spec, wm, ws, B)
    init = torch.full((B, S), 5.0)                              # far outside the bounds: done at once
    z = torch.zeros(H, B, 8)
    view = _run_rollout(alg, init, H, z[:, :, :A].contiguous(), z[:, :, :S + 1].contiguous(), [0] * H)
    assert view.counts() == [B, 0, 0, 0, 0] and len(view) == B

    g = torch.Generator().manual_seed(53)
    init = torch.randn(B, S, generator=g) * 0.1
    alg2 = _make_alg(spec, wm, ws, B)
    alg2.horizon = H
    members = [2, 4, 0, 1, 3]
    view = alg2.rollout(alg2.actor, initial_states=to_dev(init), member_idx=members)
    seed = alg2.rollout_seed + alg2._rollouts_done
    lib = _lib.load()
    eps_p, eps_m = torch.empty(H, B, A, device=dev()), torch.empty(H, B, S + 1, device=dev())
    for t in range(H):
        _lib.check(lib.drpo_philox_normal(eps_p[t].data_ptr(), B, A, None, seed, 1, t, None), "philox")
        _lib.check(lib.drpo_philox_normal(eps_m[t].data_ptr(), B, S + 1, None, seed, 2, t, None), "philox")
    torch.cuda.synchronize()
    assert abs(float(eps_m.mean())) < 0.02 and abs(float(eps_m.std()) - 1) < 0.02
    ref, counts, _ = O.rollout(ws, wm, spec, init, H, eps_p.cpu(), eps_m.cpu(), members)
    assert view.counts()[:len(counts)] == counts
    out = view.get(as_dict=True)
    assert torch.equal(out["dones"].cpu(), ref["dones"])
    assert_close(out["next_states"], ref["next_states"], RTOL, "philox rollout")
    # sharded: second half of the start states as "rank 1 of 2" must reproduce rows of the full run
    alg3 = _make_alg(spec, wm, ws, B // 2)
    alg3.horizon, alg3.shard_rank, alg3.shard_world = H, 1, 2
    alg3._rollouts_done = alg2._rollouts_done - 1
    v3 = alg3.rollout(alg3.actor, initial_states=to_dev(init[B // 2:]), member_idx=members)
    n0 = v3.counts()[0]
    assert n0 == B // 2
    assert_close(v3.get("next_states")[:n0], out["next_states"][B // 2:B], RTOL, "shard independence")


# ---------------------------------------------------------------------------------------------------------------
def _golden_after(sd, g, prefix, keys, rtol, what):
    worst = 0.0
    for k in keys:
        full = f"{prefix}.after.{k}"
        if full in g:
            worst = max(worst, assert_close(sd[k], g[full], rtol, f"{what} {k}", max_outlier_frac=2e-3))
        else:
            worst = max(worst, assert_close(sd[k].flatten()[:64], g[full + "#head"], rtol, f"{what} {k} head", max_outlier_frac=0.02))
            v = sd[k].double()
            assert_close(torch.stack([v.sum(), v.abs().sum()]).cpu(), g[full + "#sum"], 1e-4, f"{what} {k} sums")
    return worst


@pytest.mark.parametrize("tag,S,A,C", [("point_robot", 11, 2, 1), ("cartpole", 4, 1, 4), ("tracking", 51, 2, 1)])
def test_critic_update_vs_golden(golden, tag, S, A, C):
    """Three consecutive SSAC.update_critic calls of the reference (golden) vs drpo_critic_step."""
    g = golden("critic")
    w = O.make_ssac_weights(int(g[f"{tag}.seed"]), S, A, C)
    B = g[f"{tag}.it0.states"].shape[0]
    solver = make_ssac(w, S, A, C, B, std_ratio=float(g[f"{tag}.std_ratio"]))
    keys = [k for k in w if k.startswith(("critic", "constraint_critic"))]
    for it in range(3):
        batch = [to_dev(g[f"{tag}.it{it}.{n}"]) for n in O.COMPONENTS]
        noise = tuple(to_dev(g[f"{tag}.it{it}.{n}"]) for n in ("eps_actor", "eps_safe", "eps_qc"))
        assert solver.critic_optimizer.param_groups[0]["lr"] == pytest.approx(float(g[f"{tag}.lrs"][it]), rel=1e-12)
        lq, lc = solver.update_critic(*batch, noise=noise)
        assert_close(lq, g[f"{tag}.it{it}.loss_q"], RTOL, "loss_q"); assert_close(lc, g[f"{tag}.it{it}.loss_c"], RTOL, "loss_c")
        _golden_after(solver.state_dict(), g, f"{tag}.it{it}", keys, 2e-5, f"it{it}")


@pytest.mark.parametrize("S,A,C,B", [(51, 2, 1, 4096), (12, 2, 2, 1000)])
def test_critic_update_vs_oracle(S, A, C, B):
    """Larger batch: raw gradients (before clip), grad norms, losses, updated parameters and EMA targets."""
    w = O.make_ssac_weights(61, S, A, C)
    solver = make_ssac(w, S, A, C, B)
    wo = {k: v.clone() for k, v in w.items()}
    g = torch.Generator().manual_seed(62)
    obs = torch.randn(B, S, generator=g); act = torch.rand(B, A, generator=g) * 2 - 1
    nobs = obs + 0.1 * torch.randn(B, S, generator=g); rew = torch.randn(B, generator=g)
    done = torch.rand(B, generator=g) < 0.1; viol = torch.rand(B, generator=g) < 0.1
    cv = (torch.randn(B, generator=g) - 0.5) if C == 1 else (torch.randn(B, C, generator=g) - 0.5)
    batch = [obs, act, nobs, rew, done, viol, cv]
    noise = (torch.randn(B, A, generator=g), torch.randn(B, A, generator=g), torch.randn(*cv.shape, generator=g))
    adam = O.AdamState()
    lq, lc, aux = O.critic_update(wo, batch, noise, O.SSACHyper(), 0.0, adam, 3e-4)
    glq, glc = solver.update_critic(*[to_dev(b) for b in batch], noise=tuple(to_dev(n) for n in noise))
    assert_close(glq, lq, RTOL, "loss_q"); assert_close(glc, lc, RTOL, "loss_c")
    assert_close(solver._losses[2], aux["grad_norm_q"], 2e-5, "grad norm Q"); assert_close(solver._losses[3], aux["grad_norm_c"], 2e-5, "grad norm Qc")
    # raw gradients: the arena layout is state_dict order of critic then constraint_critic
    gviews = solver.critic_arena_views(solver.critic_optimizer.grad)
    for k in [k for k in w if k.startswith(("critic.", "constraint_critic."))]:
        assert_close(gviews[k], aux["grads_raw"][k], 5e-5, f"grad {k}")
    sd = solver.state_dict()
    for k in wo:
        if k.startswith(("critic", "constraint_critic")):
            assert_close(sd[k], wo[k], 2e-5, f"param {k}", max_outlier_frac=2e-3)


@pytest.mark.parametrize("tag,S,A,C", [("point_robot", 11, 2, 1), ("cartpole", 4, 1, 4)])
def test_multiplier_update_vs_golden(golden, tag, S, A, C):
    g = golden("multiplier")
    w = O.make_ssac_weights(int(g[f"{tag}.seed"]), S, A, C)
    B = g[f"{tag}.it0.obs"].shape[0]
    solver = make_ssac(w, S, A, C, B, penalty_lb=float(g[f"{tag}.penalty_lb"]))
    keys = [k for k in w if k.startswith("multiplier")]
    for it in range(3):
        loss = solver.update_multiplier(to_dev(g[f"{tag}.it{it}.obs"]), eps=to_dev(g[f"{tag}.it{it}.eps"]))
        assert_close(loss, g[f"{tag}.it{it}.loss"], RTOL, "multiplier loss")
        _golden_after(solver.state_dict(), g, f"{tag}.it{it}", keys, 2e-5, f"it{it}")


def test_buffer_gather_matches_reference_assembly():
    """drpo_buffer_gather == SampleBuffer.sample + update_solver's scaling (src/smbpo.py:253-270)."""
    import drpo_b200
    spec = SPECS["quadrotor"]
    cfg = drpo_b200.SMBPO.Config()
    cfg.buffer_max, cfg.reward_scale, cfg.alive_bonus, cfg.constraint_scale, cfg.constraint_offset = 4096, 2.0, 2.0, 10.0, 0.5
    cfg.sac_cfg.batch_size, cfg.real_fraction = 512, 0.25
    alg = drpo_b200.SMBPO(cfg, oracle_spec_to_device_env(spec), device=dev())
    g = torch.Generator().manual_seed(3)
    for buf, n in ((alg.replay_buffer, 1000), (alg.virt_buffer, 3000)):
        buf.extend(states=to_dev(torch.randn(n, 12, generator=g)), actions=to_dev(torch.randn(n, 2, generator=g)),
                   next_states=to_dev(torch.randn(n, 12, generator=g)), rewards=to_dev(torch.randn(n, generator=g)),
                   dones=to_dev(torch.rand(n, generator=g) < 0.2), violations=to_dev(torch.rand(n, generator=g) < 0.2),
                   constraint_values=to_dev(torch.randn(n, 2, generator=g)))
    torch.manual_seed(9)
    got = alg.sample_batch()
    torch.manual_seed(9)
    n_real = 128
    ir = torch.randint(1000, [n_real], device=dev()); iv = torch.randint(3000, [512 - n_real], device=dev())
    comb = [torch.cat([alg.replay_buffer._bufs[n][ir], alg.virt_buffer._bufs[n][iv]]) for n in O.COMPONENTS]
    want = O.preprocess_batch([c.cpu() for c in comb], 2.0, 2.0, 10.0, 0.5)
    for a, b, n in zip(got, want, O.COMPONENTS):
        assert torch.equal(a.cpu(), b), n


@pytest.mark.parametrize("S,A,C,B", [(51, 2, 1, 4096), (12, 2, 2, 1000)])
def test_critic_and_multiplier_tensor_mode_vs_oracle(S, A, C, B):
    """Library tensor-core mode of the SSAC steps (PREC_TF32: TF32 tensor-op GEMMs, fp32 everywhere else) against the fp32 oracle
    within the 2e-2 the north star allows for the reduced-precision GEMM path; losses much tighter."""
    import drpo_b200
    w = O.make_ssac_weights(61, S, A, C)
    solver = make_ssac(w, S, A, C, B)
    solver.precision = drpo_b200.PREC_TF32
    wo = {k: v.clone() for k, v in w.items()}
    g = torch.Generator().manual_seed(62)
    obs = torch.randn(B, S, generator=g); act = torch.rand(B, A, generator=g) * 2 - 1
    nobs = obs + 0.1 * torch.randn(B, S, generator=g); rew = torch.randn(B, generator=g)
    done = torch.rand(B, generator=g) < 0.1; viol = torch.rand(B, generator=g) < 0.1
    cv = (torch.randn(B, generator=g) - 0.5) if C == 1 else (torch.randn(B, C, generator=g) - 0.5)
    batch = [obs, act, nobs, rew, done, viol, cv]
    noise = (torch.randn(B, A, generator=g), torch.randn(B, A, generator=g), torch.randn(*cv.shape, generator=g))
    lq, lc, aux = O.critic_update(wo, batch, noise, O.SSACHyper(), 0.0, O.AdamState(), 3e-4)
    glq, glc = solver.update_critic(*[to_dev(b) for b in batch], noise=tuple(to_dev(n) for n in noise))
    assert_close(glq, lq, 5e-3, "loss_q (tf32)"); assert_close(glc, lc, 5e-3, "loss_c (tf32)")
    assert_close(solver._losses[2], aux["grad_norm_q"], 2e-2, "grad norm Q (tf32)")
    assert_close(solver._losses[3], aux["grad_norm_c"], 2e-2, "grad norm Qc (tf32)")
    gviews = solver.critic_arena_views(solver.critic_optimizer.grad)
    for k in [k for k in w if k.startswith(("critic.", "constraint_critic."))]:
        # tf32 products summed over the batch: entries that cancel to ~0 carry the largest relative error
        assert_close(gviews[k], aux["grads_raw"][k], 2e-2, f"grad {k} (tf32)", max_outlier_frac=1e-2)
        assert_close(gviews[k], aux["grads_raw"][k], 2e-1, f"grad {k} (tf32, outlier bound)")
    eps = torch.randn(B, A, generator=g)
    lm, _ = O.multiplier_update(wo, obs, eps, O.SSACHyper(), C, O.AdamState(), 3e-4)
    glm = solver.update_multiplier(to_dev(obs), eps=to_dev(eps))
    assert_close(glm, lm, 2e-2, "multiplier loss (tf32)")
