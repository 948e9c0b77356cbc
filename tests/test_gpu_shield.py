"""drpo_shield_act through the C ABI against the reference-generated golden vectors and the oracle."""
import numpy as np
import pytest
import torch

from oracle import drpo_oracle as O
from tests.util import assert_close, make_ssac, to_dev

pytestmark = pytest.mark.gpu
RTOL = 1e-5
MARGIN = 1e-4          # rows whose Qc sits this close to the threshold may legitimately take the other branch in fp32


PATHS = [1, 2]         # 1 = two-launch latency kernels, 2 = batched GEMM path (0 = auto picks by size)


def _compare_with_oracle(w, solver, s, C, st, thr, path=0, **kw):
    a_o, q_o, ch_o = O.shield_actions(w, s, C, st, thr, eps_perf=kw.get("eps"), uncertainty=kw.get("uncertainty", False),
                                      std_ratio=float(solver.constraint_critic.std_ratio))
    a, q, ch = solver.shield_act(to_dev(s), eval=kw.get("eps") is None, shield_type=st, safe_shield_threshold=thr,
                                 uncertainty=kw.get("uncertainty", False), eps=None if kw.get("eps") is None else to_dev(kw["eps"]),
                                 return_info=True, path=path)
    if st != "none":
        assert_close(q, q_o, RTOL, f"{st}: qc of the performance action")
    # rows where every candidate's decision is clear of the threshold must take exactly the oracle's branch
    A = a_o.shape[1]
    a_perf = O.policy_act(w, "actor.", s, kw.get("eps"))[0]
    a_safe = O.policy_act(w, "actor_safe.", s, None)[0]
    clear = torch.ones(len(s), dtype=torch.bool)
    if st != "none":
        for i in (range(11) if st == "linear" else [10]):
            r = (10 - i) / 10
            mix = a_safe * r + a_perf * (1 - r)
            mean, std = O.qc_forward(w, "constraint_critic.", s, mix, need_std=kw.get("uncertainty", False))
            qi = O.get_qc(mean + float(solver.constraint_critic.std_ratio) * std if kw.get("uncertainty", False) else mean, C)
            clear &= (qi - thr).abs() > MARGIN
    assert clear.float().mean() > 0.8
    assert torch.equal(ch.cpu()[clear], ch_o[clear]), f"{st}: branch choice differs on clear rows"
    assert_close(a.cpu()[clear], a_o[clear], RTOL, f"{st}: actions")
    return a, ch, clear


@pytest.mark.parametrize("tag,S,A,C", [("point_robot", 11, 2, 1), ("cartpole", 4, 1, 4)])
@pytest.mark.parametrize("shield_type", ["safe", "linear", "none"])
@pytest.mark.parametrize("path", PATHS)
def test_shield_eval_vs_golden(golden, tag, S, A, C, shield_type, path):
    g = golden("shield")
    w = O.make_ssac_weights(int(g[f"{tag}.seed"]), S, A, C)
    solver = make_ssac(w, S, A, C, 64)
    s, thr = torch.from_numpy(g[f"{tag}.states"]), float(g[f"{tag}.threshold"])
    a, ch, clear_all = _compare_with_oracle(w, solver, s, C, shield_type, thr, path=path)
    # the golden actions come from the reference's own sample_episodes_batched
    ref = torch.from_numpy(g[f"{tag}.{shield_type}.actions"])
    q_ref = torch.from_numpy(g[f"{tag}.qc_perf"])
    if shield_type == "linear":
        same = (a.cpu() - ref).abs().max(dim=1)[0] <= 1e-5
        assert same.float().mean() > 0.9                # near-threshold candidates aside (checked exactly against the oracle above)
    else:
        clear = (q_ref - thr).abs() > MARGIN if shield_type == "safe" else torch.ones(len(s), dtype=torch.bool)
        assert_close(a.cpu()[clear], ref[clear], RTOL, "actions vs the reference")
    if shield_type == "safe":
        assert 0 < int(ch.sum()) < len(ch)
    # the module-level mirror of the sampler's block
    from drpo_b200.sampling import shielded_actions
    a2 = shielded_actions(solver, to_dev(s), eval=True, safe_shield_threshold=thr, shield_type=shield_type)
    assert_close(a2.cpu()[clear_all], a.cpu()[clear_all], RTOL, "auto path")


@pytest.mark.parametrize("path", PATHS)
def test_shield_training_step_vs_golden(golden, path):
    """SMBPO.step_generator's shield: sampled performance action, Qc = mean + std_ratio*std, rows one at a time and batched."""
    g = golden("shield")
    S, A, C = 11, 2, 1
    w = O.make_ssac_weights(int(g["step.seed"]), S, A, C)
    solver = make_ssac(w, S, A, C, 64, std_ratio=float(g["step.std_ratio"]))
    s, eps, thr = torch.from_numpy(g["step.states"]), torch.from_numpy(g["step.eps"]), float(g["step.threshold"])
    a, ch, _ = _compare_with_oracle(w, solver, s, C, "safe", thr, path=path, eps=eps, uncertainty=True)
    clear = torch.from_numpy(np.abs(g["step.qc"] - thr) > MARGIN)
    assert_close(a.cpu()[clear], torch.from_numpy(g["step.actions"])[clear], RTOL, "actions vs the reference's step_generator")
    assert 0 < int(ch.sum()) < len(ch)
    for r in range(0, len(s), 7):                                        # act1-style single rows
        a1 = solver.shield_act(to_dev(s[r:r + 1]), eval=False, shield_type="safe", safe_shield_threshold=thr, uncertainty=True,
                               eps=to_dev(eps[r:r + 1]), path=path)
        if clear[r]:
            assert_close(a1.cpu(), torch.from_numpy(g["step.actions"][r:r + 1]), RTOL, f"row {r}")


def test_shield_small_batches_take_the_latency_path_and_agree_with_the_batched_one():
    """10 evaluation envs / 1 training state: the auto path (two launches) against the forced batched path and the launch count."""
    from drpo_b200 import _lib
    S, A, C = 12, 2, 2
    w = O.make_ssac_weights(6, S, A, C)
    solver = make_ssac(w, S, A, C, 64)
    g = torch.Generator().manual_seed(10)
    s = to_dev(torch.randn(10, S, generator=g))
    lib = _lib.load()
    for st, unc in (("linear", False), ("safe", True), ("none", False)):
        q = solver.shield_act(s, shield_type=st, uncertainty=unc, return_info=True, path=2)[1]
        thr = float(q.median()) if st != "none" else 0.0
        l0 = lib.drpo_launch_count()
        a0, q0, c0 = solver.shield_act(s, shield_type=st, safe_shield_threshold=thr, uncertainty=unc, return_info=True)
        assert lib.drpo_launch_count() - l0 == (1 if st == "none" else 2)
        a2, q2, c2 = solver.shield_act(s, shield_type=st, safe_shield_threshold=thr, uncertainty=unc, return_info=True, path=2)
        if st != "none":
            assert_close(q0, q2, RTOL, "qc")
        assert torch.equal(c0, c2)
        assert_close(a0, a2, RTOL, f"{st}: actions")


@pytest.mark.parametrize("path", PATHS)
def test_shield_edge_cases(path):
    S, A, C = 12, 2, 2
    w = O.make_ssac_weights(5, S, A, C)
    solver = make_ssac(w, S, A, C, 64)
    empty = solver.shield_act(torch.zeros(0, S, device="cuda"), shield_type="linear", path=path)
    assert empty.shape == (0, A)
    g = torch.Generator().manual_seed(9)
    s = torch.randn(1000, S, generator=g)
    a_perf = solver.actor.act(to_dev(s), eval=True)
    a_safe = solver.actor_safe.act(to_dev(s), eval=True)
    # threshold above every Qc: nothing is shielded ("safe") / the performance action (i = 10) always wins ("linear")
    for st in ("safe", "linear"):
        a, q, ch = solver.shield_act(to_dev(s), shield_type=st, safe_shield_threshold=1e9, return_info=True, path=path)
        assert_close(a, a_perf, RTOL); assert int((ch != (0 if st == "safe" else 10)).sum()) == 0
    # threshold below every Qc: the safe action stands
    for st in ("safe", "linear"):
        a, q, ch = solver.shield_act(to_dev(s), shield_type=st, safe_shield_threshold=-1e9, return_info=True, path=path)
        assert_close(a, a_safe, RTOL); assert int((ch != (1 if st == "safe" else -1)).sum()) == 0
    # NaN state: Qc is NaN -> "safe" keeps the performance action (NaN > thr is False), "linear" keeps the safe one
    s_nan = s[:4].clone(); s_nan[1, 3] = float("nan")
    _, q, ch = solver.shield_act(to_dev(s_nan), shield_type="safe", safe_shield_threshold=0.0, return_info=True, path=path)
    assert torch.isnan(q[1]) and int(ch[1]) == 0
    _, q, ch = solver.shield_act(to_dev(s_nan), shield_type="linear", safe_shield_threshold=0.0, return_info=True, path=path)
    assert int(ch[1]) == -1


def test_smbpo_shielded_act1_mirrors_the_training_step(golden):
    """SMBPO.shielded_act1 (src/smbpo.py:124-136) on single states of the reference's recorded episode; safe_shield off = act1."""
    import drpo_b200
    from tests.util import dev
    g = golden("shield")
    S, A, C = 11, 2, 1
    cfg = drpo_b200.SMBPO.Config()
    cfg.buffer_max, cfg.safe_shield_threshold = 4096, float(g["step.threshold"])
    alg = drpo_b200.SMBPO(cfg, drpo_b200.device_env("point-robot"), device=dev())
    alg.solver.load_state_dict(O.make_ssac_weights(int(g["step.seed"]), S, A, C), strict=False)
    assert float(alg.solver.constraint_critic.std_ratio) == float(g["step.std_ratio"])
    s, eps = to_dev(g["step.states"]), to_dev(g["step.eps"])
    clear = np.abs(g["step.qc"] - float(g["step.threshold"])) > MARGIN
    for r in range(0, len(s), 5):
        a = alg.shielded_act1(s[r], eps=eps[r:r + 1])
        assert a.shape == (A,)
        if clear[r]:
            assert_close(a.cpu(), torch.from_numpy(g["step.actions"][r]), RTOL, f"step {r}")
    alg.safe_shield = False
    a = alg.shielded_act1(s[0], eps=eps[0:1])
    assert_close(a, alg.actor.act(s[0:1], False, eps[0:1])[0], RTOL, "unshielded act1")
