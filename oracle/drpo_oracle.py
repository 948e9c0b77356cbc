"""TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of DRPO's hot path.

This is NOT the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package (``drpo_b200``) never does and fails loudly when its CUDA library is missing.

Parity status: **pinned against the reference itself** — ``oracle/make_golden.py`` imports
the unmodified reference from ``/root/reference`` (through ``oracle/ref_shim.py``), runs its
own ``BatchedGaussianEnsemble``, ``SquashedGaussianPolicy``, ``SMBPO.rollout``,
``SSAC.update_critic`` and ``SSAC.update_multiplier`` on seeded weights with injected noise and
stores the results in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
below against those vectors.  (The reference ships no tests or golden vectors of its own,
SURVEY.md §4.)

Arithmetic is numpy float64 for the env hooks (the reference's hooks are numpy) and torch CPU
fp32 for the networks (the reference's arithmetic lives in torch 2.11 / numpy 2.3, pinned by this
image).  Every function cites the reference file:line it restates.  Weights are passed as plain
``dict[str, Tensor]`` keyed like the reference ``state_dict``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
W = Dict[str, Tensor]

# deterministic synthetic weights: one definition shared by make_golden.py, the tests and bench.py
from drpo_b200.synthetic import (make_ensemble_weights, make_ssac_weights, weights_checksum)  # noqa: E402,F401

# ----------------------------------------------------------------------------------------------
# env hooks: check_done / check_violation / get_constraint_values  (numpy, fp64 on fp32 inputs)
# ----------------------------------------------------------------------------------------------

@dataclass
class EnvSpec:
    """Device-form description of one env's hooks."""
    kind: str                       # 'point_robot' | 'bounded' | 'tracking'
    state_dim: int
    con_dim: int
    # point_robot (src/env/point_robot.py:11-14)
    hazards: Tuple[Tuple[float, float], ...] = ((0.4, -1.2), (-0.4, 1.2))
    hazard_size: float = 0.8
    goal: Tuple[float, float] = (2.2, 2.2)
    goal_size: float = 0.3
    xy_bound: float = 3.0
    # bounded (src/env/poles/constraints.py:216-247): cv = [-x[d]+lb, x[d]-ub]
    active_dims: Tuple[int, ...] = ()
    lower: Tuple[float, ...] = ()
    upper: Tuple[float, ...] = ()
    # extra done rule |x[d]| > thr, compared in float32 (quadrotor.py:96-112); done |= violation
    done_dims: Tuple[int, ...] = ()
    done_thr: Tuple[float, ...] = ()
    # tracking (pyth_veh3dofconti_surrcstr_data.py:55-72)
    surr_veh_num: int = 4
    veh_length: float = 4.8
    veh_width: float = 2.0
    surr_start: int = 0
    name: str = ""


def env_point_robot(S: int = 11) -> EnvSpec:
    return EnvSpec(kind="point_robot", state_dim=S, con_dim=1, name="point-robot")


def env_cartpole() -> EnvSpec:
    # src/env/poles/inverted_pendulum.py:11-37: x in [-0.9,0.9], theta in [-0.2,0.2]; done == violation (:79-87)
    return EnvSpec(kind="bounded", state_dim=4, con_dim=4, active_dims=(0, 1),
                   lower=(-0.9, -0.2), upper=(0.9, 0.2), name="cartpole-move")


def env_quadrotor(x_threshold: float = 2.0, z_threshold: float = 2.0) -> EnvSpec:
    # src/env/quadrotor/quadrotor.py:46-58,83-114 + constrained_tracking_reset.yaml:61-70
    return EnvSpec(kind="bounded", state_dim=12, con_dim=2, active_dims=(2,), lower=(0.5,), upper=(1.5,),
                   done_dims=(0, 2, 4), done_thr=(x_threshold, z_threshold, 85 * math.pi / 180),
                   name="quadrotor")


def env_tracking(pre_horizon: int = 10, surr_veh_num: int = 1) -> EnvSpec:
    S = 6 + 1 + 4 * pre_horizon + 4 * surr_veh_num
    return EnvSpec(kind="tracking", state_dim=S, con_dim=1, surr_veh_num=surr_veh_num,
                   surr_start=6 + 1 + 4 * pre_horizon, name="tracking")


def env_safetygym60() -> EnvSpec:
    """Synthetic stand-in (the safetygym env is not in this branch): point-robot-style
    hazard distance on dims 0:2 of a 60-dim observation."""
    return EnvSpec(kind="point_robot", state_dim=60, con_dim=1, name="safetygym-point-synthetic")


def _bounded_cv(spec: EnvSpec, states: np.ndarray) -> np.ndarray:
    # src/env/poles/constraints.py:67,186-187,203-204,240-245: x @ filter.T @ A.T - b, all float64
    dim = len(spec.active_dims)
    filt = np.eye(spec.state_dim)[list(spec.active_dims)]
    A = np.vstack((-np.eye(dim), np.eye(dim)))
    b = np.hstack((-np.array(spec.lower, ndmin=1), np.array(spec.upper, ndmin=1)))
    with np.errstate(all="ignore"):     # non-finite states give NaN rows, exactly as in the reference
        return states @ filt.transpose() @ A.transpose() - b


def hooks(spec: EnvSpec, states: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(done bool[n], violation bool[n], constraint_values float32 [n] or [n,C]).

    The final float32 cast is ``torchify`` (src/torch_util.py:20-22) as applied at
    src/smbpo.py:63-65."""
    states = np.asarray(states)
    assert states.dtype == np.float32 and states.ndim == 2
    if spec.kind == "point_robot":
        # src/env/point_robot.py:96-130
        min_dist = np.full(states.shape[0], float("inf"))
        for hz in spec.hazards:
            vec = np.array(hz, dtype=np.float64) - states[:, :2]
            dist = np.linalg.norm(vec, axis=1)
            min_dist = np.minimum(dist, min_dist)
        cv = spec.hazard_size - min_dist
        viol = cv > 0
        b = np.float32(spec.xy_bound)
        oob = (states[:, 0] < -b) | (states[:, 0] > b) | (states[:, 1] < -b) | (states[:, 1] > b)
        goal = np.linalg.norm(states[:, :2] - np.array(spec.goal, dtype=np.float64), axis=1) <= spec.goal_size
        done = oob | goal
    elif spec.kind == "bounded":
        cv = _bounded_cv(spec, states)                       # [n, 2*dim] float64
        viol = np.any(np.greater(cv, 0.0), axis=-1)          # constraints.py:108-132 (strict=False)
        done = viol.copy()                                   # inverted_pendulum.py:79-82 ; quadrotor.py:112-114
        for d, thr in zip(spec.done_dims, spec.done_thr):
            t32 = np.float32(thr)                            # python float is a weak scalar (NEP 50)
            done = done | (states[:, d] < -t32) | (states[:, d] > t32)
    elif spec.kind == "tracking":
        # src/env/tracking/pyth_veh3dofconti_surrcstr_data.py:253-338 (float32 until the last line)
        done = (np.abs(states[:, 0]) > 5) | (np.abs(states[:, 1]) > 2) | (np.abs(states[:, 2]) > np.pi)
        d = (spec.veh_length - spec.veh_width) / 2
        r = np.sqrt(2) / 2 * spec.veh_width
        ego = np.array([[d, 0], [-d, 0]], dtype=np.float32)
        phis = states[:, 6]
        c = np.expand_dims(np.cos(phis), -1)
        s = np.expand_dims(np.sin(phis), -1)
        surr = states[:, spec.surr_start:].reshape(-1, spec.surr_veh_num, 4)
        xs, ys, ph = surr[:, :, 0], surr[:, :, 1], surr[:, :, 2]
        xe = xs * c + ys * s
        ye = -xs * s + ys * c
        centers = np.stack((np.stack((xe + d * np.cos(ph), ye + d * np.sin(ph)), axis=2),
                            np.stack((xe - d * np.cos(ph), ye - d * np.sin(ph)), axis=2)), axis=2)
        ego = ego[np.newaxis, np.newaxis, ...]
        ds = [np.linalg.norm(ego[..., i, :] - centers[..., j, :], axis=-1) for i in (0, 1) for j in (0, 1)]
        min_dist = np.min(np.min(np.stack(ds, axis=1), axis=-1), axis=-1)
        cv = 2 * r - min_dist
        viol = cv > 0
    else:
        raise ValueError(spec.kind)
    return np.asarray(done, dtype=bool), np.asarray(viol, dtype=bool), np.asarray(cv).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# networks
# ----------------------------------------------------------------------------------------------

def _linear_seq(w: W, prefix: str, x: Tensor, idx, act, out_act=None, member: Optional[int] = None) -> Tensor:
    for li, i in enumerate(idx):
        Wt, b = w[f"{prefix}{i}.weight"], w[f"{prefix}{i}.bias"]
        if member is not None:
            Wt, b = Wt[member], b[member]
        x = F.linear(x, Wt, b)
        if li < len(idx) - 1:
            x = act(x)
        elif out_act is not None:
            x = out_act(x)
    return x


def ensemble_forward1(w: W, states: Tensor, actions: Tensor, index: int) -> Tuple[Tensor, Tensor]:
    """src/dynamics.py:112-122 (+ src/normalization.py:23-24, unbatched_forward :258-264)."""
    norm = (states - w["state_normalizer.mean"]) / (w["state_normalizer.std"] + 1e-6)
    x = torch.cat([norm, actions], dim=-1)
    h = _linear_seq(w, "trunk.", x, (0, 2), F.silu, F.silu, member=index)
    diffs = _linear_seq(w, "diff_head.", h, (0, 2), F.silu, member=index)
    means = diffs + torch.cat([states, torch.zeros([states.shape[0], 1])], dim=1)
    lv = _linear_seq(w, "log_var_head.", h, (0, 2), F.silu, member=index)
    lv = w["max_log_var"] - F.softplus(w["max_log_var"] - lv)
    lv = w["min_log_var"] + F.softplus(lv - w["min_log_var"])
    return means, lv


def ensemble_forward_all(w: W, states: Tensor, actions: Tensor) -> Tuple[Tensor, Tensor]:
    """src/dynamics.py:124-134 with BatchedLinear (:49-52); states/actions are [E,B,*]."""
    def bl(prefix, i, x):
        return torch.bmm(x, w[f"{prefix}{i}.weight"].transpose(1, 2)) + w[f"{prefix}{i}.bias"].unsqueeze(1)
    norm = (states - w["state_normalizer.mean"]) / (w["state_normalizer.std"] + 1e-6)
    x = torch.cat([norm, actions], dim=-1)
    h = F.silu(bl("trunk.", 2, F.silu(bl("trunk.", 0, x))))
    diffs = bl("diff_head.", 2, F.silu(bl("diff_head.", 0, h)))
    E, B = states.shape[0], states.shape[1]
    means = diffs + torch.cat([states, torch.zeros([E, B, 1])], dim=-1)
    lv = bl("log_var_head.", 2, F.silu(bl("log_var_head.", 0, h)))
    lv = w["max_log_var"] - F.softplus(w["max_log_var"] - lv)
    lv = w["min_log_var"] + F.softplus(lv - w["min_log_var"])
    return means, lv


def ensemble_sample(w: W, states: Tensor, actions: Tensor, index: int, eps: Tensor) -> Tuple[Tensor, Tensor]:
    """src/dynamics.py:198-203 with the member pick and randn_like injected."""
    means, lv = ensemble_forward1(w, states, actions, index)
    stds = torch.exp(lv).sqrt()
    samples = means + stds * eps
    return samples[:, :-1], samples[:, -1]


def ensemble_means(w: W, states: Tensor, actions: Tensor) -> Tuple[Tensor, Tensor]:
    """src/dynamics.py:206-210."""
    E = w["trunk.0.weight"].shape[0]
    means, _ = ensemble_forward_all(w, states.repeat(E, 1, 1), actions.repeat(E, 1, 1))
    return means[:, :, :-1], means[:, :, -1]


def ensemble_elite_samples(w: W, states: Tensor, actions: Tensor, elites: List[int], eps: Tensor):
    """src/dynamics.py:218-234; eps is [len(elites), B, S+1]."""
    E = w["trunk.0.weight"].shape[0]
    means, lv = ensemble_forward_all(w, states.repeat(E, 1, 1), actions.repeat(E, 1, 1))
    means, lv = means[elites, ...], lv[elites, ...]
    samples = means + torch.exp(lv).sqrt() * eps
    return samples[:, :, :-1], samples[:, :, -1]


ENSEMBLE_TRAINABLE = ("trunk.", "diff_head.", "log_var_head.", "min_log_var", "max_log_var")


def ensemble_mse_loss(w: W, states: Tensor, actions: Tensor, targets: Tensor) -> Tensor:
    """BatchedGaussianEnsemble._mse_loss (src/dynamics.py:236-253): per-member Gaussian NLL [E]; inputs are [E,B,*]."""
    means, log_vars = ensemble_forward_all(w, states, actions)
    inv_vars = torch.exp(-log_vars)
    squared_errors = torch.mean((targets - means) ** 2 * inv_vars, dim=(-2, -1))
    log_dets = torch.mean(log_vars, dim=(-2, -1))
    return squared_errors + log_dets


def ensemble_compute_loss(w: W, states: Tensor, actions: Tensor, targets: Tensor, log_var_bound_weight: float = 0.01) -> Tensor:
    """compute_loss (src/dynamics.py:143-153): rows are dealt to the members in contiguous blocks (_rebatch :136-141), a remainder
    that does not divide by the ensemble size is dropped."""
    E = w["trunk.0.weight"].shape[0]
    n = targets.shape[0] - targets.shape[0] % E
    rb = lambda x: x[:n].reshape(E, n // E, *x.shape[1:])
    return torch.sum(ensemble_mse_loss(w, rb(states), rb(actions), rb(targets))) + \
        log_var_bound_weight * (w["max_log_var"].sum() - w["min_log_var"].sum())


def ensemble_train_step(w: W, states: Tensor, actions: Tensor, targets: Tensor, adam: AdamState, lr: float = 1e-3,
                        weight_decay: float = 1e-4):
    """One iteration of BatchedGaussianEnsemble.fit's loop (src/dynamics.py:164-170): compute_loss, backward, Adam (coupled L2
    1e-4 on every trainable tensor incl. the log-var bounds, :93-101).  In place on ``w``."""
    names = [k for k in w if k.startswith(ENSEMBLE_TRAINABLE)]
    for k in names:
        w[k].requires_grad_(True)
        w[k].grad = None
    loss = ensemble_compute_loss(w, states, actions, targets)
    loss.backward()
    grads = {k: w[k].grad.detach().clone() for k in names}
    for k in names:
        w[k].requires_grad_(False)
        w[k].grad = None
    with torch.no_grad():
        adam_step({k: w[k] for k in names}, {k: g.clone() for k, g in grads.items()}, adam, lr, weight_decay)
    return loss.detach(), dict(grads_raw=grads)


def normalizer_fit(w: W, states: Tensor):
    """Normalizer.fit (src/normalization.py:14-21): unbiased std, std < 1e-6 -> 1."""
    w["state_normalizer.mean"] = states.mean(dim=0)
    std = states.std(dim=0)
    std[std < 1e-6] = 1.0
    w["state_normalizer.std"] = std


def ensemble_holdout_ranking(w: W, states: Tensor, actions: Tensor, targets: Tensor, num_elites: int = 5):
    """End of fit (src/dynamics.py:172-186): every member scores the same holdout rows; the best ``num_elites`` become the elites."""
    E = w["trunk.0.weight"].shape[0]
    rep = lambda x: x.unsqueeze(0).repeat(E, *([1] * x.dim()))
    with torch.no_grad():
        losses = ensemble_mse_loss(w, rep(states), rep(actions), rep(targets))
    return torch.argsort(losses)[:num_elites].tolist(), losses


LOG_STD_BOUNDS = (-6.0, 4.0)          # src/policy.py:85


def policy_mu_std(w: W, prefix: str, states: Tensor) -> Tuple[Tensor, Tensor]:
    """src/policy.py:89-96 on the actor MLP of src/ssac.py:184-186."""
    out = _linear_seq(w, prefix + "net.", states, (0, 2, 4), F.relu)
    mu, raw = out.chunk(2, dim=-1)
    lo, hi = LOG_STD_BOUNDS
    log_std = lo + (hi - lo) * torch.sigmoid(raw)
    return mu, log_std.exp() * 1.0


def policy_act(w: W, prefix: str, states: Tensor, eps: Optional[Tensor]) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """src/policy.py:77-80.  eps=None -> eval action tanh(mu) (src/squashed_gaussian.py:12-16).
    Returns (action, pre-tanh x, mu, std); torch.normal(mu,std) == eps*std + mu."""
    mu, std = policy_mu_std(w, prefix, states)
    x = mu if eps is None else eps * std + mu
    return torch.tanh(x), x, mu, std


def squashed_log_prob(mu: Tensor, std: Tensor, x: Tensor) -> Tensor:
    """Independent(SquashedGaussian).log_prob with TanhTransform's cached pre-image x
    (torch/distributions/transformed_distribution.py log_prob; normal.py log_prob;
    transforms.py TanhTransform.log_abs_det_jacobian) as used at src/ssac.py:286-288."""
    ladj = 2.0 * (math.log(2.0) - x - F.softplus(-2.0 * x))
    base = -((x - mu) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))
    return ((0 - ladj) + base).sum(-1)


def q_forward(w: W, prefix: str, s: Tensor, a: Tensor) -> List[Tensor]:
    """CriticEnsemble.all (src/ssac.py:31-33)."""
    sa = torch.cat([s, a], -1)
    return [_linear_seq(w, f"{prefix}qs.{i}.", sa, (0, 2, 4), F.relu).squeeze(1) for i in range(2)]


QC_LOG_STD_MIN, QC_LOG_STD_MAX = -4.0, 4.0     # src/ssac.py:51-52


def qc_forward(w: W, prefix: str, s: Tensor, a: Tensor, need_std: bool = True):
    """ConstraintCritic.forward (src/ssac.py:64-78): returns (mean, std) with shapes [B] (C==1) or [B,C]."""
    sa = torch.cat([s, a], -1)
    h = _linear_seq(w, prefix + "trunk.", sa, (0, 2), F.relu, F.relu)
    mean = _linear_seq(w, prefix + "mean_head.", h, (0, 2), F.relu)
    C = mean.shape[-1]
    if C == 1:
        mean = mean.squeeze(1)
    if not need_std:
        return mean, None
    ls = _linear_seq(w, prefix + "log_std_head.", h, (0, 2), F.relu)
    if C == 1:
        ls = ls.squeeze(1)
    ls = QC_LOG_STD_MAX - F.softplus(QC_LOG_STD_MAX - ls)
    ls = QC_LOG_STD_MIN + F.softplus(ls - QC_LOG_STD_MIN)
    return mean, ls.exp()


def get_qc(qc: Tensor, C: int) -> Tensor:
    """SSAC._get_qc (src/ssac.py:588-600)."""
    return torch.max(qc, dim=-1)[0] if C > 1 else qc


def multiplier_forward(w: W, s: Tensor, qc: Tensor, ub: float = 50.0) -> Tensor:
    """MLPMultiplier.forward (src/ssac.py:107-111)."""
    x = torch.cat([s, qc.unsqueeze(-1)], -1)
    raw = _linear_seq(w, "multiplier.lam.", x, (0, 2, 4), torch.tanh).squeeze(1)
    return ub / 2.0 * (1.0 + torch.tanh(raw / ub * 2))


# ----------------------------------------------------------------------------------------------
# safety shield  (src/smbpo.py:124-136 training step; src/sampling.py:420-439 batched evaluation)
# ----------------------------------------------------------------------------------------------

def shield_actions(w: W, states: Tensor, C: int, shield_type: str, threshold: float, eps_perf: Optional[Tensor] = None,
                   uncertainty: bool = False, std_ratio: float = 2.0):
    """The action-selection block of sample_episodes_batched (src/sampling.py:420-439; eval=True, Qc = mean head) and of
    SMBPO.step_generator (src/smbpo.py:124-136: performance action SAMPLED with ``eps_perf``, Qc with
    ``uncertainty=distributional_qc`` = mean + std_ratio*std, safe action in eval mode; that block is shield_type "safe").
    Returns (actions, qc of the performance action, choice): choice = 1/0 (safe / performance action) for "safe"; the last
    mixing step i in 0..10 whose Qc is <= threshold for "linear" (-1: none, the safe action stands; ratio = (10-i)/10)."""
    a_perf = policy_act(w, "actor.", states, eps_perf)[0]

    def qc_of(a):
        mean, std = qc_forward(w, "constraint_critic.", states, a, need_std=uncertainty)
        return get_qc(mean + torch.mul(std_ratio, std) if uncertainty else mean, C)       # src/ssac.py:85 ; :588-600
    qcs = qc_of(a_perf)
    a_safe = policy_act(w, "actor_safe.", states, None)[0]
    A = a_perf.shape[1]
    if shield_type == "safe":
        danger = (qcs > threshold).tile((A, 1)).t()
        return torch.where(danger, a_safe, a_perf), qcs, danger[:, 0].to(torch.int32)
    if shield_type == "linear":
        actions, choice = a_safe, torch.full((len(states),), -1, dtype=torch.int32)
        for i in range(11):
            ratio = (10 - i) / 10
            mix = a_safe * ratio + a_perf * (1 - ratio)
            safe = qc_of(mix) <= threshold
            actions = torch.where(safe.tile((A, 1)).t(), mix, actions)
            choice = torch.where(safe, torch.full_like(choice, i), choice)
        return actions, qcs, choice
    return a_perf, qcs, torch.zeros(len(states), dtype=torch.int32)


# ----------------------------------------------------------------------------------------------
# rollout  (src/smbpo.py:229-249)
# ----------------------------------------------------------------------------------------------

COMPONENTS = ("states", "actions", "next_states", "rewards", "dones", "violations", "constraint_values")


def rollout(w_ssac: W, w_model: W, spec: EnvSpec, initial_states: Tensor, horizon: int,
            eps_policy: Tensor, eps_model: Tensor, member_idx: List[int], policy_prefix: str = "actor."):
    """Branched H-step rollout with injected noise indexed by ORIGINAL trajectory id:
    eps_policy [H,B0,A], eps_model [H,B0,S+1], member_idx [H] (the host-side random.choice of
    src/dynamics.py:199).  Returns (dict of the 7 step-major, survivor-ordered components,
    per-step row counts, per-step trajectory ids)."""
    out = {k: [] for k in COMPONENTS}
    counts, ids_per_step = [], []
    states = initial_states
    ids = torch.arange(initial_states.shape[0])
    with torch.no_grad():
        for t in range(horizon):
            actions, _, _, _ = policy_act(w_ssac, policy_prefix, states, eps_policy[t][ids])
            next_states, rewards = ensemble_sample(w_model, states, actions, member_idx[t], eps_model[t][ids])
            d, v, cv = hooks(spec, next_states.numpy())
            d, v, cv = torch.from_numpy(d), torch.from_numpy(v), torch.from_numpy(cv)
            for k, x in zip(COMPONENTS, (states, actions, next_states, rewards, d, v, cv)):
                out[k].append(x)
            counts.append(int(states.shape[0]))
            ids_per_step.append(ids)
            cont = ~d
            if cont.sum() == 0:
                break
            states, ids = next_states[cont], ids[cont]
    return {k: torch.cat(v) for k, v in out.items()}, counts, ids_per_step


# ----------------------------------------------------------------------------------------------
# optimiser pieces (torch.optim.Adam / clip_grad_norm_ / CosineAnnealingLR / update_ema restated)
# ----------------------------------------------------------------------------------------------

@dataclass
class AdamState:
    step: int = 0
    m: Dict[str, Tensor] = field(default_factory=dict)
    v: Dict[str, Tensor] = field(default_factory=dict)


def clip_grad_norm(grads: List[Tensor], max_norm: float) -> float:
    """torch.nn.utils.clip_grad_norm_ (torch 2.11): total = ||(||g_i||_2)_i||_2, g *= min(1, max/(total+1e-6))."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g, 2.0) for g in grads]), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return float(total)


def adam_step(params: W, grads: W, st: AdamState, lr: float, wd: float = 1e-4,
              b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam single-tensor formulation with coupled L2 weight decay (src/defaults.py:4,
    src/ssac.py:199-203)."""
    st.step += 1
    t = st.step
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    for k, p in params.items():
        g = grads[k]
        if wd != 0:
            g = g.add(p, alpha=wd)
        m = st.m.setdefault(k, torch.zeros_like(p))
        v = st.v.setdefault(k, torch.zeros_like(p))
        m.lerp_(g, 1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-(lr / bc1))


def cosine_lr(lr_prev: float, t: int, T: int, eta_min: float, base_lr: float) -> float:
    """torch 2.11 CosineAnnealingLR recursive form; t = epoch index AFTER the step (>=1)."""
    if t == 0:
        return base_lr
    if (t - 1 - T) % (2 * T) == 0:
        return lr_prev + (base_lr - eta_min) * (1 - math.cos(math.pi / T)) / 2
    return (1 + math.cos(math.pi * t / T)) / (1 + math.cos(math.pi * (t - 1) / T)) * (lr_prev - eta_min) + eta_min


def update_ema(target: W, source: W, tprefix: str, sprefix: str, rate: float):
    """src/torch_util.py:223-226."""
    for k, p in source.items():
        if k.startswith(sprefix):
            tk = tprefix + k[len(sprefix):]
            target[tk].copy_(rate * p + (1 - rate) * target[tk])


# ----------------------------------------------------------------------------------------------
# SSAC critic step (src/ssac.py:284-302,338-362,415-456) and multiplier step (:529-578)
# ----------------------------------------------------------------------------------------------

@dataclass
class SSACHyper:
    discount: float = 0.99
    tau: float = 0.005
    grad_norm: float = 5.0
    qc_td_bound: float = 5.0
    std_ratio: float = 2.0
    weight_decay: float = 1e-4
    constraint_threshold: float = 0.0
    penalty_lb: float = -1.0
    penalty_ub: float = 100.0
    multiplier_ub: float = 50.0
    lam_epsilon: float = 1.0


CRITIC_PREFIXES = ("critic.", "constraint_critic.")


def critic_losses(w: W, batch, noise, hp: SSACHyper, log_alpha: float):
    """Returns (loss_Q, loss_C) as autograd scalars.  batch = the 7 components in COMPONENT order
    (already reward/constraint-scaled, src/smbpo.py:261-270); noise = (eps_actor [B,A],
    eps_safe [B,A], eps_qc [B] or [B,C])."""
    obs, act, nobs, rew, done, viol, cv = batch
    eps_a, eps_s, eps_qc = noise
    C = 1 if cv.dim() == 1 else cv.shape[1]
    with torch.no_grad():
        # compute_target  src/ssac.py:284-294
        a1, x1, mu1, std1 = policy_act(w, "actor.", nobs, eps_a)
        logp = squashed_log_prob(mu1, std1, x1)
        q1t, q2t = q_forward(w, "critic_target.", nobs, a1)
        nv = torch.min(q1t, q2t) - math.exp(log_alpha) * logp
        q = rew + hp.discount * (1.0 - done.float()) * nv
        # compute_cons_target  src/ssac.py:338-362
        a2, _, _, _ = policy_act(w, "actor_safe.", nobs, eps_s)
        mt, st = qc_forward(w, "constraint_critic_target.", nobs, a2)
        nqc = mt + torch.clamp(eps_qc, -2.0, 2.0) * st
        qc_mean, _ = qc_forward(w, "constraint_critic.", obs, act, need_std=False)
        nonterm = (1.0 - hp.discount) * cv + hp.discount * torch.maximum(cv, nqc)
        dones = done.tile((C, 1)).t().squeeze().float() if done.shape[0] > 1 or C > 1 else done.float()
        tu = nonterm * (1 - dones) + cv * dones
        tb = torch.clamp(tu - qc_mean, min=-hp.qc_td_bound, max=hp.qc_td_bound) + qc_mean
    q1, q2 = q_forward(w, "critic.", obs, act)
    loss_q = (F.mse_loss(q1, q) + F.mse_loss(q2, q)) / 2          # pythonic_mean, src/ssac.py:296-298
    mu, sd = qc_forward(w, "constraint_critic.", obs, act)         # src/ssac.py:416-423
    loss_c = torch.mean(torch.pow(mu - tu, 2) / (2 * torch.pow(sd.detach(), 2))
                        + torch.pow(mu.detach() - tb, 2) / (2 * torch.pow(sd, 2)) + torch.log(sd))
    return loss_q, loss_c, dict(q_target=q, target_u=tu, target_b=tb, q1=q1.detach(), q2=q2.detach(),
                                qc_mean=mu.detach(), qc_std=sd.detach())


def critic_update(w: W, batch, noise, hp: SSACHyper, log_alpha: float, adam: AdamState, lr: float):
    """One SSAC.update_critic (src/ssac.py:437-456): in-place on ``w``.  Returns the two losses, the
    (pre-clip) grad norms and intermediate targets."""
    names = [k for k in w if k.startswith(CRITIC_PREFIXES)]
    for k in names:
        w[k].requires_grad_(True)
        w[k].grad = None
    loss_q, loss_c, aux = critic_losses(w, batch, noise, hp, log_alpha)
    (loss_q + loss_c).backward()
    grads = {k: w[k].grad.detach().clone() for k in names}
    for k in names:
        w[k].requires_grad_(False)
        w[k].grad = None
    aux["grads_raw"] = {k: g.clone() for k, g in grads.items()}
    n_q = clip_grad_norm([grads[k] for k in names if k.startswith("critic.")], hp.grad_norm)
    n_c = clip_grad_norm([grads[k] for k in names if k.startswith("constraint_critic.")], hp.grad_norm)
    with torch.no_grad():
        adam_step({k: w[k] for k in names}, grads, adam, lr, hp.weight_decay)
        update_ema(w, w, "critic_target.", "critic.", hp.tau)
        update_ema(w, w, "constraint_critic_target.", "constraint_critic.", hp.tau)
    aux["grad_norm_q"], aux["grad_norm_c"] = n_q, n_c
    return loss_q.detach(), loss_c.detach(), aux


def multiplier_loss(w: W, obs: Tensor, eps: Tensor, hp: SSACHyper, C: int):
    """SSAC.multiplier_loss (src/ssac.py:529-568), DRPO mode (mlp_multiplier, reachability,
    distributional_qc).  eps is the rsample noise [B,A]."""
    with torch.no_grad():
        a, _, _, _ = policy_act(w, "actor.", obs, eps)
        mu, sd = qc_forward(w, "constraint_critic.", obs, a)
        qc_a = get_qc(mu + hp.std_ratio * sd, C)                   # uncertainty=True, src/ssac.py:85
        penalty = torch.clamp(qc_a - hp.constraint_threshold, min=hp.penalty_lb, max=hp.penalty_ub)
        a_s, _, _, _ = policy_act(w, "actor_safe.", obs, None)
        mu_s, sd_s = qc_forward(w, "constraint_critic.", obs, a_s)
        safe_qc = get_qc(mu_s + hp.std_ratio * sd_s, C)
    lams = multiplier_forward(w, obs, safe_qc, hp.multiplier_ub)
    lams_safe = torch.mul(safe_qc <= 0, lams)
    lams_unsafe = torch.mul(safe_qc > 0, lams)
    loss = -0.5 * torch.mean(torch.mul(lams_safe, penalty)) + \
        F.mse_loss(lams_unsafe, (safe_qc > 0) * (hp.multiplier_ub - hp.lam_epsilon))
    return loss, dict(lams=lams.detach(), safe_qc=safe_qc, penalty=penalty)


def multiplier_update(w: W, obs: Tensor, eps: Tensor, hp: SSACHyper, C: int, adam: AdamState, lr: float):
    """SSAC.update_multiplier (src/ssac.py:570-578)."""
    names = [k for k in w if k.startswith("multiplier.")]
    for k in names:
        w[k].requires_grad_(True)
        w[k].grad = None
    loss, aux = multiplier_loss(w, obs, eps, hp, C)
    loss.backward()
    grads = {k: w[k].grad.detach().clone() for k in names}
    for k in names:
        w[k].requires_grad_(False)
        w[k].grad = None
    aux["grads_raw"] = {k: g.clone() for k, g in grads.items()}
    aux["grad_norm"] = clip_grad_norm(list(grads.values()), hp.grad_norm)
    with torch.no_grad():
        adam_step({k: w[k] for k in names}, grads, adam, lr, hp.weight_decay)
    return loss.detach(), aux


# ----------------------------------------------------------------------------------------------
# SSAC actor / alpha / safe-actor step (src/ssac.py:458-527)  -- SURVEY.md §8(f) "next" row 1
# ----------------------------------------------------------------------------------------------

def actor_losses(w: W, obs: Tensor, noise, hp: SSACHyper, log_alpha: Tensor, q_index: int, C: int,
                 target_entropy: float):
    """SSAC.actor_loss (src/ssac.py:458-505), DRPO mode (reachability, distributional_qc, mlp_multiplier,
    autotune_alpha, use_log_alpha_loss=False).  noise = (eps_actor [B,A], eps_safe [B,A]): the two rsample
    draws; q_index = the critic random.choice picked (src/ssac.py:41-43).  Returns the three losses
    [actor, alpha, actor_safe] as autograd scalars."""
    eps_a, eps_s = noise
    a, x, mu, std = policy_act(w, "actor.", obs, eps_a)                         # distr.rsample()
    log_prob = squashed_log_prob(mu, std, x)
    actor_q = q_forward(w, "critic.", obs, a)[q_index]                          # critic.random_choice
    alpha = log_alpha.exp()
    uncstr = torch.mean(alpha.detach() * log_prob - actor_q)
    m_a, s_a = qc_forward(w, "constraint_critic.", obs, a)
    actor_qc = get_qc(m_a + hp.std_ratio * s_a, C)                              # uncertainty=True, src/ssac.py:85
    with torch.no_grad():
        a_eval, _, _, _ = policy_act(w, "actor_safe.", obs, None)
        m_s, s_s = qc_forward(w, "constraint_critic.", obs, a_eval)
        safe_qc = get_qc(m_s + hp.std_ratio * s_s, C)
        lams = multiplier_forward(w, obs, safe_qc, hp.multiplier_ub)
    cstr = torch.mean(torch.mul(lams, actor_qc))
    a2, _, _, _ = policy_act(w, "actor_safe.", obs, eps_s)                      # distr_safe.rsample()
    m_2, s_2 = qc_forward(w, "constraint_critic.", obs, a2)
    actor_safe_loss = torch.mean(get_qc(m_2 + hp.std_ratio * s_2, C))
    alpha_loss = -alpha * torch.mean(log_prob.detach() + target_entropy)
    return [uncstr + cstr, alpha_loss, actor_safe_loss], dict(log_prob=log_prob.detach(), lams=lams, actor_qc=actor_qc.detach())


def actor_update(w: W, obs: Tensor, noise, hp: SSACHyper, log_alpha: Tensor, q_index: int, C: int, target_entropy: float,
                 adams: Dict[str, AdamState], lrs: Dict[str, float]):
    """SSAC.update_actor_and_alpha (src/ssac.py:507-527): three losses, three Adam optimisers (actor and
    actor_safe with coupled L2 1e-4 and grad-norm clip 5, log_alpha without either); in place on ``w`` and
    ``log_alpha`` (a 0-dim tensor).  The three backward passes of the reference touch disjoint parameter
    sets (actor / log_alpha / actor_safe; alpha and lams are detached in the actor loss), so one backward of
    their sum yields the same gradients."""
    names_a = [k for k in w if k.startswith("actor.")]
    names_s = [k for k in w if k.startswith("actor_safe.")]
    for k in names_a + names_s:
        w[k].requires_grad_(True)
        w[k].grad = None
    la = log_alpha.detach().clone().requires_grad_(True)
    losses, aux = actor_losses(w, obs, noise, hp, la, q_index, C, target_entropy)
    (losses[0] + losses[1] + losses[2]).backward()
    grads = {k: w[k].grad.detach().clone() for k in names_a + names_s}
    g_alpha = la.grad.detach().clone()
    for k in list(w):
        w[k].requires_grad_(False)
        w[k].grad = None
    aux["grads_raw"] = {k: g.clone() for k, g in grads.items()}
    aux["grad_alpha"] = g_alpha.clone()
    aux["grad_norm_actor"] = clip_grad_norm([grads[k] for k in names_a], hp.grad_norm)
    aux["grad_norm_safe"] = clip_grad_norm([grads[k] for k in names_s], hp.grad_norm)
    with torch.no_grad():
        adam_step({k: w[k] for k in names_a}, grads, adams["actor"], lrs["actor"], hp.weight_decay)
        adam_step({"log_alpha": log_alpha}, {"log_alpha": g_alpha}, adams["alpha"], lrs["alpha"], 0.0)
        adam_step({k: w[k] for k in names_s}, grads, adams["safe"], lrs["safe"], hp.weight_decay)
    return [l.detach() for l in losses], aux


def preprocess_batch(batch, reward_scale: float, alive_bonus: float, constraint_scale: float,
                     constraint_offset: float):
    """src/smbpo.py:261-270."""
    b = list(batch)
    if reward_scale != 0:
        b[3] = b[3] * reward_scale
    if alive_bonus != 0:
        b[3] = b[3] + alive_bonus
    b[6] = b[6] * constraint_scale
    b[6] = b[6] + (b[6] > 0).float() * constraint_offset
    return b
