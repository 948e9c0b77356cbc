"""The oracle (oracle/drpo_oracle.py) against the golden vectors generated from the unmodified
reference by oracle/make_golden.py.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import drpo_oracle as O

T = torch.from_numpy


def close(a, b, rtol=1e-6, atol=1e-6):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


@pytest.mark.parametrize("tag,spec", [("point_robot", O.env_point_robot()), ("cartpole", O.env_cartpole()),
                                      ("quadrotor", O.env_quadrotor()), ("tracking1", O.env_tracking(10, 1)),
                                      ("tracking4", O.env_tracking(10, 4))])
def test_hooks_bit_exact(golden, tag, spec):
    g = golden("hooks")
    d, v, cv = O.hooks(spec, g[f"{tag}.states"])
    assert np.array_equal(d, g[f"{tag}.done"])
    assert np.array_equal(v, g[f"{tag}.viol"])
    assert np.array_equal(cv, g[f"{tag}.cv"], equal_nan=True)


def test_hooks_known_answers():
    """SURVEY.md §8c known-answer vectors (fp64-on-fp32 semantics)."""
    f = np.float32
    s = np.zeros((4, 11), f)
    s[0, :2] = (0.4, -0.4); s[1, :2] = (0.4, np.nextafter(f(-0.4), f(0))); s[2, :2] = (2.2, 2.0); s[3, :2] = (3.0000002, 0)
    d, v, cv = O.hooks(O.env_point_robot(), s)
    assert v.tolist() == [True, False, False, False] and d.tolist() == [False, False, True, True]
    assert cv[0] == f(5.9604645663569045e-09) and cv[1] == f(-2.3841858265427618e-08)
    c = np.zeros((4, 4), f)
    c[0, 0] = 0.9; c[1, 0] = np.nextafter(f(0.9), f(2)); c[2, 1] = -0.2; c[3, 3] = np.inf
    d, v, cv = O.hooks(O.env_cartpole(), c)
    assert v.tolist() == [False, True, True, False] and np.isnan(cv[3]).all()
    assert cv[2, 1] == f(2.9802322831784522e-09)


@pytest.mark.parametrize("tag,S,A", [("point_robot", 11, 2), ("cartpole", 4, 1), ("quadrotor", 12, 2)])
def test_ensemble(golden, tag, S, A):
    g = golden("ensemble")
    w = O.make_ensemble_weights(int(g[f"{tag}.seed"]), S, A)
    assert O.weights_checksum(w) == pytest.approx(float(g[f"{tag}.wsum"]), rel=1e-12)
    s, a, eps = T(g[f"{tag}.states"]), T(g[f"{tag}.actions"]), T(g[f"{tag}.eps"])
    m, lv = O.ensemble_forward1(w, s, a, int(g[f"{tag}.member"]))
    close(m, g[f"{tag}.means"]); close(lv, g[f"{tag}.log_vars"])
    ns, r = O.ensemble_sample(w, s, a, int(g[f"{tag}.member"]), eps)
    close(ns, g[f"{tag}.next_states"]); close(r, g[f"{tag}.rewards"])
    ms, mr = O.ensemble_means(w, s, a)
    close(ms, g[f"{tag}.means_all_s"]); close(mr, g[f"{tag}.means_all_r"])
    es, er = O.ensemble_elite_samples(w, s, a, [6, 0, 2, 5, 1], T(g[f"{tag}.eps_elite"]))
    close(es, g[f"{tag}.elite_s"]); close(er, g[f"{tag}.elite_r"])


@pytest.mark.parametrize("tag,S,A,C", [("point_robot", 11, 2, 1), ("cartpole", 4, 1, 4)])
def test_policy(golden, tag, S, A, C):
    g = golden("policy")
    w = O.make_ssac_weights(int(g[f"{tag}.seed"]), S, A, C)
    assert O.weights_checksum(w) == pytest.approx(float(g[f"{tag}.wsum"]), rel=1e-12)
    a, x, mu, std = O.policy_act(w, "actor.", T(g[f"{tag}.states"]), T(g[f"{tag}.eps"]))
    close(a, g[f"{tag}.actions"])
    close(O.squashed_log_prob(mu, std, x), g[f"{tag}.log_prob"], rtol=1e-5, atol=1e-5)
    ae, _, _, _ = O.policy_act(w, "actor_safe.", T(g[f"{tag}.states"]), None)
    close(ae, g[f"{tag}.eval_actions_safe"])


def test_rollout(golden):
    g = golden("rollout")
    wm = O.make_ensemble_weights(int(g["seed_model"]), 11, 2, diff_scale=float(g["diff_scale"]))
    ws = O.make_ssac_weights(int(g["seed_ssac"]), 11, 2, 1)
    res, counts, _ = O.rollout(ws, wm, O.env_point_robot(), T(g["init"]), len(g["members"]),
                               T(g["eps_policy"]), T(g["eps_model"]), [int(m) for m in g["members"]])
    assert counts == g["counts"].tolist()
    assert np.array_equal(res["dones"].numpy(), g["out.dones"])
    assert np.array_equal(res["violations"].numpy(), g["out.violations"])
    for k in ("states", "actions", "next_states", "rewards", "constraint_values"):
        close(res[k], g[f"out.{k}"])


def _check_after(w, g, prefix, keys):
    for k in keys:
        full = f"{prefix}.after.{k}"
        if full in g:
            close(w[k], g[full], rtol=2e-6, atol=2e-7)
        else:
            v = w[k].double()
            close([v.sum().item(), v.abs().sum().item()], g[full + "#sum"], rtol=1e-6, atol=1e-6)
            close(w[k].flatten()[:64], g[full + "#head"], rtol=2e-6, atol=2e-7)


@pytest.mark.parametrize("tag,S,A,C", [("point_robot", 11, 2, 1), ("cartpole", 4, 1, 4), ("tracking", 51, 2, 1)])
def test_critic_update(golden, tag, S, A, C):
    g = golden("critic")
    w = O.make_ssac_weights(int(g[f"{tag}.seed"]), S, A, C)
    hp = O.SSACHyper(std_ratio=float(g[f"{tag}.std_ratio"]))
    adam = O.AdamState()
    lrs = g[f"{tag}.lrs"]
    keys = [k for k in w if k.startswith(("critic", "constraint_critic"))]
    for it in range(3):
        batch = [T(g[f"{tag}.it{it}.{n}"]) for n in O.COMPONENTS]
        noise = tuple(T(g[f"{tag}.it{it}.{n}"]) for n in ("eps_actor", "eps_safe", "eps_qc"))
        lq, lc, _ = O.critic_update(w, batch, noise, hp, float(g[f"{tag}.log_alpha"]), adam, float(lrs[it]))
        close(lq, g[f"{tag}.it{it}.loss_q"]); close(lc, g[f"{tag}.it{it}.loss_c"])
        _check_after(w, g, f"{tag}.it{it}", keys)
        nxt = O.cosine_lr(float(lrs[it]), it + 1, int(g[f"{tag}.T_max"]), 8e-5, 3e-4)
        assert nxt == pytest.approx(float(lrs[it + 1]), rel=1e-12)


@pytest.mark.parametrize("tag,S,A,C", [("point_robot", 11, 2, 1), ("cartpole", 4, 1, 4)])
def test_multiplier_update(golden, tag, S, A, C):
    g = golden("multiplier")
    w = O.make_ssac_weights(int(g[f"{tag}.seed"]), S, A, C)
    hp = O.SSACHyper(penalty_lb=float(g[f"{tag}.penalty_lb"]))
    adam = O.AdamState()
    keys = [k for k in w if k.startswith("multiplier")]
    for it in range(3):
        loss, _ = O.multiplier_update(w, T(g[f"{tag}.it{it}.obs"]), T(g[f"{tag}.it{it}.eps"]), hp, C, adam,
                                      float(g[f"{tag}.lrs"][it]))
        close(loss, g[f"{tag}.it{it}.loss"], rtol=1e-6)
        _check_after(w, g, f"{tag}.it{it}", keys)


@pytest.mark.parametrize("tag,S,A,C", [("point_robot", 11, 2, 1), ("cartpole", 4, 1, 4)])
def test_actor_update(golden, tag, S, A, C):
    """Oracle restatement of SSAC.update_actor_and_alpha (src/ssac.py:458-527) vs three consecutive updates of the reference:
    the three losses, the actor / safe-actor parameters and log_alpha."""
    g = golden("actor")
    w = O.make_ssac_weights(int(g[f"{tag}.seed"]), S, A, C)
    la = torch.tensor(0.0)
    adams = {k: O.AdamState() for k in ("actor", "alpha", "safe")}
    keys = [k for k in w if k.startswith(("actor.", "actor_safe."))]
    for it in range(3):
        lrs = dict(actor=float(g[f"{tag}.lrs"][it]), alpha=float(g[f"{tag}.alpha_lr"]), safe=float(g[f"{tag}.lrs"][it]))
        losses, _ = O.actor_update(w, T(g[f"{tag}.it{it}.obs"]), (T(g[f"{tag}.it{it}.eps_actor"]), T(g[f"{tag}.it{it}.eps_safe"])),
                                   O.SSACHyper(), la, int(g[f"{tag}.it{it}.q_index"]), C, float(g[f"{tag}.target_entropy"]), adams, lrs)
        close(torch.stack(losses), g[f"{tag}.it{it}.losses"], rtol=1e-6)
        assert float(la) == pytest.approx(float(g[f"{tag}.it{it}.log_alpha_after"]), rel=1e-6)
        _check_after(w, g, f"{tag}.it{it}", keys)


@pytest.mark.parametrize("tag,S,A", [("point_robot", 11, 2), ("quadrotor", 12, 2)])
def test_ensemble_fit(golden, tag, S, A):
    """Oracle restatement of BatchedGaussianEnsemble.fit's loop (src/dynamics.py:143-189, src/normalization.py:14-21): normaliser
    fit, three Adam steps on compute_loss (one batch with a remainder), holdout losses and elite ranking vs the reference."""
    g = golden("ensemble_fit")
    w = O.make_ensemble_weights(int(g[f"{tag}.seed"]), S, A)
    states, actions, targets = T(g[f"{tag}.states"]), T(g[f"{tag}.actions"]), T(g[f"{tag}.targets"])
    O.normalizer_fit(w, states)
    close(w["state_normalizer.mean"], g[f"{tag}.norm_mean"], rtol=1e-6); close(w["state_normalizer.std"], g[f"{tag}.norm_std"], rtol=1e-6)
    adam = O.AdamState()
    keys = [k for k in w if k.startswith(O.ENSEMBLE_TRAINABLE)]
    for it in range(3):
        idx = torch.as_tensor(g[f"{tag}.idx{it}"])
        loss, _ = O.ensemble_train_step(w, states[idx], actions[idx], targets[idx], adam)
        close(loss, g[f"{tag}.losses"][it], rtol=1e-6)
        _check_after(w, g, f"{tag}.it{it}", keys)
    hold = torch.as_tensor(g[f"{tag}.holdout_idx"])
    elites, losses = O.ensemble_holdout_ranking(w, states[hold], actions[hold], targets[hold])
    close(losses, g[f"{tag}.holdout_losses"], rtol=1e-5)
    assert elites == [int(x) for x in g[f"{tag}.elites"]]


# ---- safety shield ----------------------------------------------------------------------------
@pytest.mark.parametrize("tag,S,A,C", [("point_robot", 11, 2, 1), ("cartpole", 4, 1, 4)])
@pytest.mark.parametrize("shield_type", ["safe", "linear", "none"])
def test_shield_eval_bit_exact(golden, tag, S, A, C, shield_type):
    """sample_episodes_batched's action selection (src/sampling.py:420-439), recorded from the reference itself."""
    g = golden("shield")
    w = O.make_ssac_weights(int(g[f"{tag}.seed"]), S, A, C)
    assert abs(O.weights_checksum(w) - float(g[f"{tag}.wsum"])) < 1e-6
    a, q, choice = O.shield_actions(w, T(g[f"{tag}.states"]), C, shield_type, float(g[f"{tag}.threshold"]))
    assert np.array_equal(a.numpy(), g[f"{tag}.{shield_type}.actions"])
    close(q, g[f"{tag}.qc_perf"])
    if shield_type == "safe":
        assert 0 < int(choice.sum()) < len(choice)              # both branches are exercised


def test_shield_training_step(golden):
    """SMBPO.step_generator's shield (src/smbpo.py:124-136) over 48 consecutive real-env steps of the reference."""
    g = golden("shield")
    w = O.make_ssac_weights(int(g["step.seed"]), 11, 2, 1)
    a, q, choice = O.shield_actions(w, T(g["step.states"]), 1, "safe", float(g["step.threshold"]), eps_perf=T(g["step.eps"]),
                                    uncertainty=True, std_ratio=float(g["step.std_ratio"]))
    close(q, g["step.qc"], rtol=1e-5, atol=1e-5)
    margin = np.abs(g["step.qc"] - g["step.threshold"]) > 1e-4     # the reference ran row by row: decisions away from the threshold
    close(a[T(margin)], g["step.actions"][margin], rtol=1e-5, atol=1e-5)
    assert 0 < int(choice.sum()) < len(choice)
