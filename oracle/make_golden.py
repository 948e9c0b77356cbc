"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
It imports the reference through oracle/ref_shim.py, loads seeded synthetic weights
(oracle.drpo_oracle.make_*_weights) into the reference's own modules, injects noise by
patching the four RNG entry points the hot path uses (torch.normal,
torch.randn_like, torch.distributions.normal._standard_normal, random.choice), runs the
reference functions and stores inputs + outputs.  It also prints the oracle-vs-reference error
for every case so a drift is visible at generation time.
"""
import math
import os
import pathlib
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402
from oracle import drpo_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


class NoiseTape:
    """FIFO of injected draws; each entry is a callable(shape)->tensor or a plain value."""

    def __init__(self):
        self.normal, self.randn_like, self.std_normal, self.choice = [], [], [], []
        self._orig = None

    def __enter__(self):
        import torch.distributions.normal as tdn
        self._orig = (torch.normal, torch.randn_like, tdn._standard_normal, random.choice)
        tape = self

        def normal(mean, std, *a, **k):
            eps = tape.normal.pop(0)
            eps = eps(mean.shape) if callable(eps) else eps
            return eps * std + mean                      # ATen normal_out: out.normal_().mul_(std).add_(mean)

        def randn_like(x, *a, **k):
            eps = tape.randn_like.pop(0)
            if eps is None:
                return torch.zeros_like(x)
            return eps(x.shape) if callable(eps) else eps

        def std_normal(shape, dtype, device):
            eps = tape.std_normal.pop(0)
            return eps(shape) if callable(eps) else eps

        def choice(seq):
            v = tape.choice.pop(0)
            return seq[v] if isinstance(v, int) and not isinstance(seq[0], int) else v

        torch.normal, torch.randn_like, tdn._standard_normal, random.choice = normal, randn_like, std_normal, choice
        return self

    def __exit__(self, *exc):
        import torch.distributions.normal as tdn
        torch.normal, torch.randn_like, tdn._standard_normal, random.choice = self._orig


def t2n(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def maxrel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


# ----------------------------------------------------------------------------------------------
def reference_hooks(spec: O.EnvSpec):
    """The reference's own hook code for each env kind (real classes where importable)."""
    if spec.name == "point-robot":
        from src.env.point_robot import PointRobot
        env = PointRobot()
        return env.check_done, env.check_violation, env.get_constraint_values
    if spec.kind == "bounded":
        # SafeInvertedPendulumEnv / QuadrotorWrapperEnv need MuJoCo / safe-control-gym (absent): use the
        # reference's BoundedConstraint class and follow inverted_pendulum.py:79-121 / quadrotor.py:83-158.
        from src.env.poles.constraints import BoundedConstraint, ConstrainedVariableType
        cons = BoundedConstraint(spec.state_dim, lower_bounds=list(spec.lower), upper_bounds=list(spec.upper),
                                 constrained_variable=ConstrainedVariableType.STATE,
                                 active_dims=list(spec.active_dims))
        if spec.name == "cartpole-move":
            return cons.is_violated, cons.is_violated, lambda s: np.squeeze(cons.get_value(s))

        def check_done(states):
            x_thr, z_thr, th = spec.done_thr[0], spec.done_thr[1], 85 * math.pi / 180
            x, z, theta = states[..., 0], states[..., 2], states[..., 4]
            done = (x < -x_thr) + (x > x_thr) + (z < -z_thr) + (z > z_thr) + (theta < -th) + (theta > th)
            return np.logical_or(done, cons.is_violated(states))
        return check_done, cons.is_violated, lambda s: np.squeeze(cons.get_value(s))
    if spec.kind == "tracking":
        from src.env.tracking.pyth_veh3dofconti_surrcstr_data import SimuVeh3dofcontiSurrCstr
        env = SimuVeh3dofcontiSurrCstr(pre_horizon=10, surr_veh_num=spec.surr_veh_num)
        return env.check_done, env.check_violation, env.get_constraint_values
    raise ValueError(spec)


def hook_inputs(spec: O.EnvSpec, n: int, seed: int) -> np.ndarray:
    g = np.random.RandomState(seed)
    s = g.randn(n, spec.state_dim).astype(np.float32)
    f32 = np.float32
    if spec.kind == "point_robot":
        s[:, :2] = g.uniform(-3.3, 3.3, size=(n, 2)).astype(f32)
        edge = [(0.4, -0.4), (0.4, np.nextafter(f32(-0.4), f32(0))), (2.2, 1.9), (3.0000002, 0.0), (3.0, 0.0),
                (-3.0, 3.0), (2.2, 2.5), (0.4, -2.0), (-0.4, 0.4), (-0.4, 2.0), (1.2, -1.2), (-1.2, 1.2)]
        for i, (x, y) in enumerate(edge):
            s[i, 0], s[i, 1] = f32(x), f32(y)
    elif spec.kind == "bounded":
        for d, lo, hi in zip(spec.active_dims, spec.lower, spec.upper):
            s[:, d] = g.uniform(lo - 0.3 * (hi - lo), hi + 0.3 * (hi - lo), size=n).astype(f32)
            k = 4 * list(spec.active_dims).index(d)
            s[k + 0, d], s[k + 1, d] = f32(lo), np.nextafter(f32(lo), f32(-9))
            s[k + 2, d], s[k + 3, d] = f32(hi), np.nextafter(f32(hi), f32(9))
        for j, (d, thr) in enumerate(zip(spec.done_dims, spec.done_thr)):
            s[20 + 4 * j + 0, d], s[20 + 4 * j + 1, d] = f32(thr), np.nextafter(f32(thr), f32(9))
            s[20 + 4 * j + 2, d], s[20 + 4 * j + 3, d] = f32(-thr), np.nextafter(f32(-thr), f32(-9))
        s[40, 0] = np.inf
        s[41, spec.state_dim - 1] = np.nan
        s[42, spec.state_dim - 1] = -np.inf
    elif spec.kind == "tracking":
        s[:, 0] = g.uniform(-6, 6, n)
        s[:, 1] = g.uniform(-2.5, 2.5, n)
        s[:, 2] = g.uniform(-3.5, 3.5, n)
        s[:, 6] = g.uniform(-3.2, 3.2, n)
        for v in range(spec.surr_veh_num):
            b = spec.surr_start + 4 * v
            s[:, b] = g.uniform(-10, 10, n)
            s[:, b + 1] = g.uniform(-5, 5, n)
            s[:, b + 2] = g.uniform(-3.2, 3.2, n)
    return s


def gen_hooks():
    out = {}
    for tag, spec in [("point_robot", O.env_point_robot()), ("cartpole", O.env_cartpole()),
                      ("quadrotor", O.env_quadrotor()), ("tracking1", O.env_tracking(10, 1)),
                      ("tracking4", O.env_tracking(10, 4))]:
        s = hook_inputs(spec, 512, seed=11)
        cd, cvio, gcv = reference_hooks(spec)
        from src.torch_util import torchify
        done = torchify(cd(s)).numpy()
        viol = torchify(cvio(s)).numpy()
        with np.errstate(invalid="ignore"):
            cv = torchify(gcv(s)).numpy()
        od, ov, ocv = O.hooks(spec, s)
        assert np.array_equal(done, od) and np.array_equal(viol, ov), tag
        assert np.array_equal(cv, ocv, equal_nan=True), (tag, np.abs(cv - ocv).max())
        print(f"hooks[{tag}]: oracle == reference bit-exact; done {done.mean():.3f} viol {viol.mean():.3f}")
        out.update({f"{tag}.states": s, f"{tag}.done": done, f"{tag}.viol": viol, f"{tag}.cv": cv})
    np.savez_compressed(os.path.join(GOLD, "hooks.npz"), **out)


# ----------------------------------------------------------------------------------------------
def build_reference_ensemble(w, S, A):
    from src.dynamics import BatchedGaussianEnsemble
    ens = BatchedGaussianEnsemble(BatchedGaussianEnsemble.Config(), S, A)
    ens.load_state_dict(w, strict=True)
    return ens


def gen_ensemble():
    out = {}
    for tag, S, A, seed in [("point_robot", 11, 2, 101), ("cartpole", 4, 1, 102), ("quadrotor", 12, 2, 103)]:
        w = O.make_ensemble_weights(seed, S, A)
        ens = build_reference_ensemble(w, S, A)
        g = torch.Generator().manual_seed(seed + 1)
        B = 96
        s = torch.randn(B, S, generator=g)
        a = torch.rand(B, A, generator=g) * 2 - 1
        eps = torch.randn(B, S + 1, generator=g)
        with torch.no_grad():
            m_ref, lv_ref = ens._forward1(s, a, 3)
            ens._elite_inds = [0, 1, 2, 3, 4]
            with NoiseTape() as tape:
                tape.choice.append(3)
                tape.randn_like.append(eps)
                ns_ref, r_ref = ens.sample(s, a)
            ms_ref, mr_ref = ens.means(s, a)
            eps_e = torch.randn(5, B, S + 1, generator=g)
            ens._elite_inds = [6, 0, 2, 5, 1]
            with NoiseTape() as tape:
                tape.randn_like.append(eps_e)
                es_ref, er_ref = ens.elite_samples(s, a)
        m, lv = O.ensemble_forward1(w, s, a, 3)
        ns, r = O.ensemble_sample(w, s, a, 3, eps)
        ms, mr = O.ensemble_means(w, s, a)
        es, er = O.ensemble_elite_samples(w, s, a, [6, 0, 2, 5, 1], eps_e)
        print(f"ensemble[{tag}]: oracle-vs-ref maxrel means {maxrel(m, m_ref):.2e} lv {maxrel(lv, lv_ref):.2e} "
              f"sample {maxrel(ns, ns_ref):.2e} means_all {maxrel(ms, ms_ref):.2e} elite {maxrel(es, es_ref):.2e}")
        out.update(t2n({f"{tag}.seed": seed, f"{tag}.wsum": O.weights_checksum(w), f"{tag}.states": s,
                        f"{tag}.actions": a, f"{tag}.eps": eps, f"{tag}.member": 3, f"{tag}.means": m_ref,
                        f"{tag}.log_vars": lv_ref, f"{tag}.next_states": ns_ref, f"{tag}.rewards": r_ref,
                        f"{tag}.means_all_s": ms_ref, f"{tag}.means_all_r": mr_ref, f"{tag}.eps_elite": eps_e,
                        f"{tag}.elite_s": es_ref, f"{tag}.elite_r": er_ref}))
    np.savez_compressed(os.path.join(GOLD, "ensemble.npz"), **out)


# ----------------------------------------------------------------------------------------------
class _FakeEnv:
    """Env stand-in handing the reference's SSAC/SMBPO the hook triple of `spec`."""

    def __init__(self, spec):
        self.spec = spec
        self.con_dim = spec.con_dim
        self.check_done, self.check_violation, self.get_constraint_values = reference_hooks(spec)


def build_reference_ssac(w, S, A, C, B, spec, std_ratio=2.0, penalty_lb=-1.0):
    from src.ssac import SSAC
    cfg = SSAC.Config()
    cfg.batch_size = B
    cfg.constraint_critic_cfg.std_ratio = std_ratio
    cfg.penalty_lb = penalty_lb
    solver = SSAC(cfg, S, A, C, 10, 100, 300, 10, 10.0, lambda: _FakeEnv(spec), None)
    missing, unexpected = solver.load_state_dict(w, strict=False)
    assert not unexpected and all(m.startswith("total_updates") for m in missing), (missing, unexpected)
    return solver


def make_batch(g, S, A, C, B, spec):
    obs = torch.randn(B, S, generator=g)
    act = torch.rand(B, A, generator=g) * 2 - 1
    nobs = obs + 0.1 * torch.randn(B, S, generator=g)
    rew = torch.randn(B, generator=g)
    done = torch.rand(B, generator=g) < 0.1
    viol = torch.rand(B, generator=g) < 0.1
    cv = torch.randn(B, generator=g) - 0.5 if C == 1 else torch.randn(B, C, generator=g) - 0.5
    return [obs, act, nobs, rew, done, viol, cv]


def gen_policy():
    out = {}
    for tag, S, A, C, seed in [("point_robot", 11, 2, 1, 201), ("cartpole", 4, 1, 4, 202)]:
        w = O.make_ssac_weights(seed, S, A, C)
        solver = build_reference_ssac(w, S, A, C, 64, O.env_point_robot() if C == 1 else O.env_cartpole())
        g = torch.Generator().manual_seed(seed + 1)
        s = torch.randn(80, S, generator=g) * 2
        eps = torch.randn(80, A, generator=g)
        with NoiseTape() as tape:
            tape.normal.append(eps)
            with torch.no_grad():
                distr = solver.actor.distr(s)
                a_ref = distr.sample()
                lp_ref = distr.log_prob(a_ref)
            a_eval_ref = solver.actor_safe.act(s, eval=True)
        a, x, mu, std = O.policy_act(w, "actor.", s, eps)
        lp = O.squashed_log_prob(mu, std, x)
        a_eval, _, _, _ = O.policy_act(w, "actor_safe.", s, None)
        print(f"policy[{tag}]: oracle-vs-ref maxrel act {maxrel(a, a_ref):.2e} logp {maxrel(lp, lp_ref):.2e} "
              f"eval {maxrel(a_eval, a_eval_ref):.2e}")
        out.update(t2n({f"{tag}.seed": seed, f"{tag}.wsum": O.weights_checksum(w), f"{tag}.states": s,
                        f"{tag}.eps": eps, f"{tag}.actions": a_ref, f"{tag}.log_prob": lp_ref,
                        f"{tag}.eval_actions_safe": a_eval_ref}))
    np.savez_compressed(os.path.join(GOLD, "policy.npz"), **out)


def gen_rollout():
    """The real SMBPO.rollout on the real PointRobot hooks (src/smbpo.py:229-249)."""
    from src.env.point_robot import PointRobot
    from src.env.torch_wrapper import TorchWrapper
    from src.log import default_log as log
    from src.smbpo import SMBPO
    from src.checkpoint import CheckpointableData
    log.setup(pathlib.Path(tempfile.mkdtemp()))
    S, A, C, B0, H = 11, 2, 1, 192, 6
    cfg = SMBPO.Config()
    cfg.rollout_batch_size, cfg.horizon = B0, H
    alg = SMBPO(cfg, lambda id=None: TorchWrapper(PointRobot(id=id)), CheckpointableData(), 10)
    wm = O.make_ensemble_weights(301, S, A, diff_scale=0.3)
    ws = O.make_ssac_weights(302, S, A, C)
    alg.model_ensemble.load_state_dict(wm, strict=True)
    alg.solver.load_state_dict({k: v for k, v in ws.items()}, strict=False)
    alg.model_ensemble._elite_inds = [0, 1, 2, 3, 4]
    g = torch.Generator().manual_seed(303)
    init = torch.randn(B0, S, generator=g)
    init[:, 0] = torch.rand(B0, generator=g) * 5.6 - 2.8
    init[:, 1] = torch.rand(B0, generator=g) * 5.6 - 2.8
    eps_p = torch.randn(H, B0, A, generator=g)
    eps_m = torch.randn(H, B0, S + 1, generator=g)
    members = [int(x) for x in torch.randint(0, 5, (H,), generator=g)]

    state = {"ids": torch.arange(B0), "t": 0}
    orig_check_done = alg.check_done

    def check_done(states):
        d = orig_check_done(states)
        state["pending"] = state["ids"][~d]
        return d
    alg.check_done = check_done
    with NoiseTape() as tape:
        for t in range(H):
            tape.normal.append(lambda shape, t=t: eps_p[t][state["ids"]])
            tape.choice.append(members[t])

            def model_noise(shape, t=t):
                e = eps_m[t][state["ids"]]
                return e
            tape.randn_like.append(model_noise)
        # advance alive ids after each step's check_done: hook into get_constraint_value (called last)
        orig_gcv = alg.get_constraint_value

        def gcv(states):
            r = orig_gcv(states)
            state["ids"] = state["pending"]
            return r
        alg.get_constraint_value = gcv
        buf = alg.rollout(alg.actor, initial_states=init)
    ref = buf.get(as_dict=True)
    virt = alg.virt_buffer.get(as_dict=True)
    assert all(torch.equal(ref[k], virt[k]) for k in ref)
    res, counts, _ = O.rollout(ws, wm, O.env_point_robot(), init, H, eps_p, eps_m, members)
    assert len(ref["states"]) == sum(counts), (len(ref["states"]), counts)
    assert torch.equal(res["dones"], ref["dones"]) and torch.equal(res["violations"], ref["violations"])
    print(f"rollout[point_robot]: counts {counts}; masks equal; oracle-vs-ref maxrel next_states "
          f"{maxrel(res['next_states'], ref['next_states']):.2e} cv {maxrel(res['constraint_values'], ref['constraint_values']):.2e}")
    out = t2n({"seed_model": 301, "seed_ssac": 302, "wsum_model": O.weights_checksum(wm),
               "wsum_ssac": O.weights_checksum(ws), "init": init, "eps_policy": eps_p, "eps_model": eps_m,
               "members": members, "counts": counts, "diff_scale": 0.3})
    out.update({f"out.{k}": v.numpy() for k, v in ref.items()})
    np.savez_compressed(os.path.join(GOLD, "rollout.npz"), **out)


def _summ(w, prefixes):
    """Compact per-tensor summary: full small tensors, (sum, abs-sum, first 8) of the big ones."""
    out = {}
    for k, v in w.items():
        if k.startswith(prefixes):
            v = v.detach()
            if v.numel() <= 1024:
                out[k] = v.numpy().copy()
            else:
                out[k + "#sum"] = np.array([v.double().sum().item(), v.double().abs().sum().item()])
                out[k + "#head"] = v.flatten()[:64].numpy().copy()
    return out


def gen_critic():
    out = {}
    for tag, S, A, C, B, seed, spec, sr in [("point_robot", 11, 2, 1, 64, 401, O.env_point_robot(), 2.0),
                                            ("cartpole", 4, 1, 4, 48, 402, O.env_cartpole(), 2.0),
                                            ("tracking", 51, 2, 1, 32, 403, O.env_tracking(10, 1), 1.0)]:
        w = O.make_ssac_weights(seed, S, A, C)
        solver = build_reference_ssac(w, S, A, C, B, spec, std_ratio=sr)
        wo = {k: v.clone() for k, v in w.items()}
        adam = O.AdamState()
        hp = O.SSACHyper(std_ratio=sr)
        g = torch.Generator().manual_seed(seed + 1)
        lr = solver.critic_optimizer.param_groups[0]["lr"]
        lrs = [lr]
        for it in range(3):
            batch = make_batch(g, S, A, C, B, spec)
            shape_c = (B,) if C == 1 else (B, C)
            noise = (torch.randn(B, A, generator=g), torch.randn(B, A, generator=g), torch.randn(*shape_c, generator=g))
            with NoiseTape() as tape:
                tape.normal += [noise[0], noise[1]]
                tape.randn_like += [noise[2], None]
                lq_ref, lc_ref = solver.update_critic(*batch)
            lq, lc, aux = O.critic_update(wo, batch, noise, hp, float(solver.log_alpha), adam, lrs[-1])
            lrs.append(O.cosine_lr(lrs[-1], it + 1, solver.updates_per_training, solver.critic_lr_end, solver.critic_lr))
            assert abs(lrs[-1] - solver.critic_optimizer.param_groups[0]["lr"]) < 1e-12
            sd = solver.state_dict()
            err = max(maxrel(wo[k], sd[k]) for k in wo if k.startswith(("critic", "constraint_critic")))
            print(f"critic[{tag}] it{it}: loss_q ref {lq_ref:.6f} oracle {lq:.6f}; loss_c ref {lc_ref:.6f} oracle {lc:.6f}; "
                  f"params maxrel {err:.2e}; gnorm {aux['grad_norm_q']:.3f}/{aux['grad_norm_c']:.3f}")
            out.update(t2n({f"{tag}.it{it}.{n}": x for n, x in zip(O.COMPONENTS, batch)}))
            out.update(t2n({f"{tag}.it{it}.eps_actor": noise[0], f"{tag}.it{it}.eps_safe": noise[1],
                            f"{tag}.it{it}.eps_qc": noise[2], f"{tag}.it{it}.loss_q": lq_ref,
                            f"{tag}.it{it}.loss_c": lc_ref}))
            out.update({f"{tag}.it{it}.after.{k}": v for k, v in
                        _summ(sd, ("critic", "constraint_critic")).items()})
        out.update(t2n({f"{tag}.seed": seed, f"{tag}.wsum": O.weights_checksum(w), f"{tag}.lrs": lrs,
                        f"{tag}.log_alpha": float(solver.log_alpha), f"{tag}.std_ratio": sr,
                        f"{tag}.T_max": solver.updates_per_training}))
    np.savez_compressed(os.path.join(GOLD, "critic.npz"), **out)


def gen_multiplier():
    out = {}
    for tag, S, A, C, B, seed, spec, plb in [("point_robot", 11, 2, 1, 64, 501, O.env_point_robot(), -5.0),
                                             ("cartpole", 4, 1, 4, 48, 502, O.env_cartpole(), -1.0)]:
        w = O.make_ssac_weights(seed, S, A, C)
        solver = build_reference_ssac(w, S, A, C, B, spec, penalty_lb=plb)
        wo = {k: v.clone() for k, v in w.items()}
        adam = O.AdamState()
        hp = O.SSACHyper(penalty_lb=plb)
        g = torch.Generator().manual_seed(seed + 1)
        lrs = [solver.multiplier_optimizer.param_groups[0]["lr"]]
        for it in range(3):
            obs = torch.randn(B, S, generator=g)
            eps = torch.randn(B, A, generator=g)
            with NoiseTape() as tape:
                tape.std_normal.append(eps)
                tape.randn_like += [None, None]
                # the reference returns nothing: recompute its loss first on a copy of the noise
                loss_ref = solver.multiplier_loss(obs).detach()
            with NoiseTape() as tape:
                tape.std_normal.append(eps)
                tape.randn_like += [None, None]
                solver.update_multiplier(obs)
            lo, aux = O.multiplier_update(wo, obs, eps, hp, C, adam, lrs[-1])
            lrs.append(O.cosine_lr(lrs[-1], it + 1, solver.lam_updates_num, solver.multiplier_lr_end, solver.multiplier_lr))
            assert abs(lrs[-1] - solver.multiplier_optimizer.param_groups[0]["lr"]) < 1e-12
            sd = solver.state_dict()
            err = max(maxrel(wo[k], sd[k]) for k in wo if k.startswith("multiplier"))
            print(f"multiplier[{tag}] it{it}: loss ref {loss_ref:.6f} oracle {lo:.6f}; params maxrel {err:.2e}; "
                  f"unsafe frac {(aux['safe_qc'] > 0).float().mean():.2f}")
            out.update(t2n({f"{tag}.it{it}.obs": obs, f"{tag}.it{it}.eps": eps, f"{tag}.it{it}.loss": loss_ref}))
            out.update({f"{tag}.it{it}.after.{k}": v for k, v in _summ(sd, ("multiplier",)).items()})
        out.update(t2n({f"{tag}.seed": seed, f"{tag}.wsum": O.weights_checksum(w), f"{tag}.lrs": lrs,
                        f"{tag}.penalty_lb": plb, f"{tag}.T_max": solver.lam_updates_num}))
    np.savez_compressed(os.path.join(GOLD, "multiplier.npz"), **out)


def gen_actor():
    """SSAC.update_actor_and_alpha (src/ssac.py:507-527): three consecutive updates of the reference with injected rsample
    draws and critic choices; the oracle restatement runs alongside and must land on the same parameters."""
    out = {}
    for tag, S, A, C, B, seed, spec in [("point_robot", 11, 2, 1, 64, 601, O.env_point_robot()),
                                        ("cartpole", 4, 1, 4, 48, 602, O.env_cartpole())]:
        w = O.make_ssac_weights(seed, S, A, C)
        solver = build_reference_ssac(w, S, A, C, B, spec)
        if not isinstance(solver.target_entropy, (int, float)):       # Config() built by hand leaves the Optional(float) marker: the
            solver.target_entropy = -A                                # config loader resolves it to None -> -dim(A) (src/ssac.py:229-230)
        wo = {k: v.clone() for k, v in w.items()}
        la = torch.tensor(float(solver.log_alpha.detach()))
        adams = {k: O.AdamState() for k in ("actor", "alpha", "safe")}
        hp = O.SSACHyper()
        g = torch.Generator().manual_seed(seed + 1)
        lr_a = [solver.actor_optimizer.param_groups[0]["lr"]]
        for it in range(3):
            obs = torch.randn(B, S, generator=g)
            eps_a, eps_s = torch.randn(B, A, generator=g), torch.randn(B, A, generator=g)
            q_index = int(torch.randint(2, (1,), generator=g))

            def tape_up(tape):
                tape.std_normal += [eps_a, eps_s]
                tape.choice.append(q_index)
                tape.randn_like += [None, None, None]
            with NoiseTape() as tape:
                tape_up(tape)
                ref_losses = [l.detach().clone() for l in solver.actor_loss(obs, include_alpha=True)]
            with NoiseTape() as tape:
                tape_up(tape)
                solver.update_actor_and_alpha(obs)
            lrs = dict(actor=lr_a[-1], alpha=solver.actor_lr, safe=lr_a[-1])
            losses, aux = O.actor_update(wo, obs, (eps_a, eps_s), hp, la, q_index, C, float(solver.target_entropy), adams, lrs)
            lr_a.append(O.cosine_lr(lr_a[-1], it + 1, solver.actor_updates_num, solver.actor_lr_end, solver.actor_lr))
            assert abs(lr_a[-1] - solver.actor_optimizer.param_groups[0]["lr"]) < 1e-12
            assert abs(lr_a[-1] - solver.actor_safe_optimizer.param_groups[0]["lr"]) < 1e-12
            sd = solver.state_dict()
            err = max(maxrel(wo[k], sd[k]) for k in wo if k.startswith(("actor.", "actor_safe.")))
            print(f"actor[{tag}] it{it}: losses ref {[round(float(l), 6) for l in ref_losses]} oracle {[round(float(l), 6) for l in losses]}; "
                  f"params maxrel {err:.2e}; log_alpha ref {float(solver.log_alpha):.8f} oracle {float(la):.8f}; "
                  f"gnorm {aux['grad_norm_actor']:.3f}/{aux['grad_norm_safe']:.4f}")
            out.update(t2n({f"{tag}.it{it}.obs": obs, f"{tag}.it{it}.eps_actor": eps_a, f"{tag}.it{it}.eps_safe": eps_s,
                            f"{tag}.it{it}.q_index": q_index, f"{tag}.it{it}.losses": torch.stack(ref_losses),
                            f"{tag}.it{it}.log_alpha_after": float(solver.log_alpha)}))
            out.update({f"{tag}.it{it}.after.{k}": v for k, v in _summ(sd, ("actor.", "actor_safe.")).items()})
        out.update(t2n({f"{tag}.seed": seed, f"{tag}.wsum": O.weights_checksum(w), f"{tag}.lrs": lr_a,
                        f"{tag}.alpha_lr": solver.actor_lr, f"{tag}.target_entropy": float(solver.target_entropy),
                        f"{tag}.T_max": solver.actor_updates_num}))
    np.savez_compressed(os.path.join(GOLD, "actor.npz"), **out)


def gen_ensemble_fit():
    """The training loop of BatchedGaussianEnsemble.fit (src/dynamics.py:155-189) with its three torch.randint draws replaced by
    fixed index tensors: normaliser fit, three Adam steps on compute_loss, holdout ranking."""
    out = {}
    for tag, S, A, seed in [("point_robot", 11, 2, 701), ("quadrotor", 12, 2, 703)]:
        w = O.make_ensemble_weights(seed, S, A)
        ens = build_reference_ensemble(w, S, A)
        E, Bm = ens.ensemble_size, 32
        g = torch.Generator().manual_seed(seed + 1)
        n = 600
        states = torch.randn(n, S, generator=g) * 2 + 0.5
        actions = torch.rand(n, A, generator=g) * 2 - 1
        next_states = states + 0.05 * torch.randn(n, S, generator=g)
        rewards = torch.randn(n, generator=g)
        targets = torch.cat([next_states, rewards.unsqueeze(1)], dim=1)
        ens.state_normalizer.fit(states)
        wo = {k: v.clone() for k, v in w.items()}
        O.normalizer_fit(wo, states)
        adam = O.AdamState()
        idxs = [torch.randint(n, [E * Bm + (3 if it == 1 else 0)], generator=g) for it in range(3)]     # one batch with a remainder
        losses = []
        for it, idx in enumerate(idxs):
            loss = ens.compute_loss(states[idx], actions[idx], targets[idx])
            ens.optimizer.zero_grad(); loss.backward(); ens.optimizer.step()
            lo, _ = O.ensemble_train_step(wo, states[idx], actions[idx], targets[idx], adam)
            sd = ens.state_dict()
            err = max(maxrel(wo[k], sd[k]) for k in wo)
            print(f"ensemble_fit[{tag}] it{it}: loss ref {float(loss):.6f} oracle {float(lo):.6f}; params maxrel {err:.2e}")
            losses.append(float(loss))
            out.update({f"{tag}.it{it}.after.{k}": v for k, v in _summ(sd, ("trunk.", "diff_head.", "log_var_head.", "min_log_var", "max_log_var")).items()})
        hold = torch.randint(n, [ens.holdout_size], generator=g)
        hi = hold.repeat(E, 1)
        ref_h = ens._mse_loss(states[hi], actions[hi], targets[hi], enable_grad=False)
        elites, oh = O.ensemble_holdout_ranking(wo, states[hold], actions[hold], targets[hold], ens.num_elites)
        print(f"ensemble_fit[{tag}] holdout maxrel {maxrel(oh, ref_h):.2e} elites ref {torch.argsort(ref_h)[:ens.num_elites].tolist()} oracle {elites}")
        out.update(t2n({f"{tag}.seed": seed, f"{tag}.states": states, f"{tag}.actions": actions, f"{tag}.targets": targets,
                        f"{tag}.idx0": idxs[0], f"{tag}.idx1": idxs[1], f"{tag}.idx2": idxs[2], f"{tag}.losses": torch.tensor(losses),
                        f"{tag}.norm_mean": ens.state_normalizer.mean, f"{tag}.norm_std": ens.state_normalizer.std,
                        f"{tag}.holdout_idx": hold, f"{tag}.holdout_losses": ref_h,
                        f"{tag}.elites": torch.tensor(torch.argsort(ref_h)[:ens.num_elites].tolist())}))
    np.savez_compressed(os.path.join(GOLD, "ensemble_fit.npz"), **out)


def gen_shield():
    """The safety shield of the batched evaluation sampler (src/sampling.py:420-439, run through the reference's own
    sample_episodes_batched on a one-step recording env) and of the training step (src/smbpo.py:124-136, run through the
    reference's own SMBPO.step_generator on the real PointRobot with the model fit / rollout / updates stubbed out)."""
    import gym
    from src.env.batch import BaseBatchedEnv
    from src.sampling import sample_episodes_batched
    out = {}
    for tag, S, A, C, seed in [("point_robot", 11, 2, 1, 801), ("cartpole", 4, 1, 4, 802)]:
        w = O.make_ssac_weights(seed, S, A, C)
        solver = build_reference_ssac(w, S, A, C, 64, O.env_point_robot() if C == 1 else O.env_cartpole())
        g = torch.Generator().manual_seed(seed + 1)
        n = 96
        states = torch.randn(n, S, generator=g) * 1.5

        class Proto:
            observation_space = gym.spaces.Box(-np.inf, np.inf, shape=(S,))
            action_space = gym.spaces.Box(-1.0, 1.0, shape=(A,))
            con_dim = C
            _max_episode_steps = 1

        class OneStep(BaseBatchedEnv):
            def _reset_index(self, index):
                return states[int(index)].clone()

            def _step(self, actions):
                self.taken = actions.clone()
                k = len(actions)
                return states.clone(), torch.zeros(k), torch.ones(k, dtype=torch.bool), [{"violation": False}] * k

        with torch.no_grad():
            a_perf = solver.actor.act(states, eval=True)
            q_perf = solver._get_qc(solver.constraint_critic(states, a_perf))
        thr = float(q_perf.median())                    # half of the performance actions are "dangerous"
        out.update(t2n({f"{tag}.seed": seed, f"{tag}.wsum": O.weights_checksum(w), f"{tag}.states": states,
                        f"{tag}.threshold": thr, f"{tag}.qc_perf": q_perf}))
        for st in ("safe", "linear", "none"):
            env = OneStep(Proto(), n)
            with torch.no_grad():
                sample_episodes_batched(env, solver, n, eval=True, safe_shield_threshold=thr, shield_type=st)
            a_o, q_o, ch = O.shield_actions(w, states, C, st, thr)
            assert torch.equal(a_o, env.taken) or maxrel(a_o, env.taken) < 1e-6, (tag, st, maxrel(a_o, env.taken))
            print(f"shield[{tag},{st}]: oracle-vs-ref actions maxrel {maxrel(a_o, env.taken):.2e} bit-equal "
                  f"{bool(torch.equal(a_o, env.taken))}; choices {np.bincount(ch.numpy() + 1, minlength=12).tolist()}")
            out.update(t2n({f"{tag}.{st}.actions": env.taken}))

    # ---- training-step shield: the reference's own step_generator on the real PointRobot ----
    from src.env.point_robot import PointRobot
    from src.env.torch_wrapper import TorchWrapper
    from src.log import default_log as log
    from src.smbpo import SMBPO
    from src.checkpoint import CheckpointableData
    log.setup(pathlib.Path(tempfile.mkdtemp()))
    S, A, C, T = 11, 2, 1, 48
    cfg = SMBPO.Config()
    cfg.buffer_min = 0
    alg = SMBPO(cfg, lambda id=None: TorchWrapper(PointRobot(id=id)), CheckpointableData(), 10)
    ws = O.make_ssac_weights(803, S, A, C)
    alg.solver.load_state_dict({k: v for k, v in ws.items()}, strict=False)
    alg.update_models = lambda steps: None
    alg.rollout_and_update = lambda: None
    g = torch.Generator().manual_seed(804)
    eps = torch.randn(T, A, generator=g)
    seen = {"s": [], "a": [], "q": []}
    real_step, real_get_qc = alg.real_env.step, alg.solver._get_qc

    def step(action):
        seen["a"].append(action.clone())
        return real_step(action)

    def get_qc(x):
        q = real_get_qc(x)
        seen["q"].append(q.clone())
        return q
    alg.real_env.step, alg.solver._get_qc = step, get_qc
    # threshold = the median shielded Qc of a first pass, so that both branches are taken
    def run(thr):
        alg.safe_shield_threshold = thr
        for k in seen:
            seen[k].clear()
        random.seed(11); np.random.seed(11); torch.manual_seed(11)
        alg.steps_sampled.fill_(0)
        with NoiseTape() as tape, torch.no_grad():
            for t in range(T):
                tape.normal.append(eps[t:t + 1])
                tape.randn_like.append(None)
            orig_reset = alg.real_env.reset
            gen = alg.step_generator()
            states = []
            # the generator reads `state` right before acting: record it through policy.act1's input
            orig_act = alg.actor.act

            def act(s, eval):
                states.append(s[0].clone())
                return orig_act(s, eval)
            alg.actor.act = act
            for t in range(T):
                next(gen)
            alg.actor.act = orig_act
        return torch.stack(states), torch.stack(seen["a"]), torch.cat([q.reshape(1) for q in seen["q"]])
    _, _, q0 = run(1e9)
    thr = float(q0.median())
    st_s, st_a, st_q = run(thr)
    a_o, q_o, ch = O.shield_actions(ws, st_s, C, "safe", thr, eps_perf=eps, uncertainty=True, std_ratio=alg.solver.constraint_critic.std_ratio)
    print(f"shield[step_generator]: oracle-vs-ref actions maxrel {maxrel(a_o, st_a):.2e} qc {maxrel(q_o, st_q):.2e}; "
          f"shielded {int(ch.sum())}/{T}")
    out.update(t2n({"step.seed": 803, "step.wsum": O.weights_checksum(ws), "step.states": st_s, "step.eps": eps,
                    "step.threshold": thr, "step.actions": st_a, "step.qc": st_q,
                    "step.std_ratio": float(alg.solver.constraint_critic.std_ratio)}))
    np.savez_compressed(os.path.join(GOLD, "shield.npz"), **out)


if __name__ == "__main__":
    ref_shim.import_reference()
    torch.set_num_threads(4)
    os.makedirs(GOLD, exist_ok=True)
    gen_hooks()
    gen_ensemble()
    gen_policy()
    gen_rollout()
    gen_critic()
    gen_multiplier()
    gen_actor()
    gen_ensemble_fit()
    gen_shield()
    print("golden vectors written to", GOLD)
