"""Seeded synthetic weights and start states (pure torch CPU; no CUDA, no oracle).

One definition shared by ``bench.py`` (both arms), the tests and ``oracle/make_golden.py`` so that every party works on
bit-identical parameters without fixtures having to store ~1 M floats.  Key names and shapes are the reference's
``state_dict`` layout."""
import math
from typing import Dict, List

import torch

Tensor = torch.Tensor
W = Dict[str, Tensor]



def _randn(gen, *shape, scale=1.0):
    return torch.randn(*shape, generator=gen, dtype=torch.float32) * scale


def make_ensemble_weights(seed: int, S: int, A: int, E: int = 7, hidden: int = 200,
                          diff_scale: float = 0.01) -> W:
    """Weights with the reference's key names/shapes (src/dynamics.py:70-101).

    Not the reference's initialiser: a seeded generator so that fixtures need not store 0.9 M
    parameters.  ``diff_scale`` shrinks the last diff-head layer so trajectories survive several
    steps."""
    g = torch.Generator().manual_seed(seed)
    D, O = S + A, S + 1
    w = {
        "min_log_var": torch.full([O], -10.0) + _randn(g, O, scale=0.1),
        "max_log_var": torch.full([O], 1.0) + _randn(g, O, scale=0.1),
        "state_normalizer.mean": _randn(g, S, scale=0.3),
        "state_normalizer.std": 0.5 + torch.rand(S, generator=g),
        "trunk.0.weight": _randn(g, E, hidden, D, scale=1.0 / math.sqrt(D)),
        "trunk.0.bias": _randn(g, E, hidden, scale=0.1),
        "trunk.2.weight": _randn(g, E, hidden, hidden, scale=1.0 / math.sqrt(hidden)),
        "trunk.2.bias": _randn(g, E, hidden, scale=0.1),
        "diff_head.0.weight": _randn(g, E, hidden, hidden, scale=1.0 / math.sqrt(hidden)),
        "diff_head.0.bias": _randn(g, E, hidden, scale=0.1),
        "diff_head.2.weight": _randn(g, E, O, hidden, scale=diff_scale / math.sqrt(hidden)),
        "diff_head.2.bias": _randn(g, E, O, scale=diff_scale),
        "log_var_head.0.weight": _randn(g, E, hidden, hidden, scale=1.0 / math.sqrt(hidden)),
        "log_var_head.0.bias": _randn(g, E, hidden, scale=0.1),
        "log_var_head.2.weight": _randn(g, E, O, hidden, scale=1.0 / math.sqrt(hidden)),
        "log_var_head.2.bias": _randn(g, E, O, scale=0.5) - 7.0,
    }
    return w


def _mlp_weights(g, prefix: str, dims: List[int], idx=(0, 2, 4), out_scale=1.0) -> W:
    w = {}
    for li, (i, o) in enumerate(zip(dims[:-1], dims[1:])):
        last = li == len(dims) - 2
        w[f"{prefix}{idx[li]}.weight"] = _randn(g, o, i, scale=(out_scale if last else 1.0) * math.sqrt(2.0 / (i + o)))
        w[f"{prefix}{idx[li]}.bias"] = _randn(g, o, scale=0.05)
    return w


def make_ssac_weights(seed: int, S: int, A: int, C: int, hidden: int = 256) -> W:
    """All SSAC networks (src/ssac.py:184-197,233-234) under the reference's state_dict keys."""
    g = torch.Generator().manual_seed(seed)
    w: W = {}
    w.update(_mlp_weights(g, "actor.net.", [S, hidden, hidden, 2 * A]))
    w.update(_mlp_weights(g, "actor_safe.net.", [S, hidden, hidden, 2 * A]))
    for q in range(2):
        w.update(_mlp_weights(g, f"critic.qs.{q}.", [S + A, hidden, hidden, 1]))
    w.update(_mlp_weights(g, "constraint_critic.trunk.", [S + A, hidden, hidden], idx=(0, 2)))
    w.update(_mlp_weights(g, "constraint_critic.mean_head.", [hidden, hidden, C], idx=(0, 2)))
    w.update(_mlp_weights(g, "constraint_critic.log_std_head.", [hidden, hidden, C], idx=(0, 2)))
    w.update(_mlp_weights(g, "multiplier.lam.", [S + 1, hidden, hidden, 1]))
    # centre mu + std_ratio*sigma around 0 so both the safe and unsafe branches of the losses are exercised
    w["constraint_critic.mean_head.2.bias"] -= 0.5
    w["constraint_critic.log_std_head.2.bias"] -= 1.5
    for k in list(w.keys()):
        if k.startswith("critic."):
            w["critic_target." + k[len("critic."):]] = w[k].clone() + 0.01 * _randn(g, *w[k].shape)
        if k.startswith("constraint_critic."):
            w["constraint_critic_target." + k[len("constraint_critic."):]] = \
                w[k].clone() + 0.01 * _randn(g, *w[k].shape)
    return w


def weights_checksum(w: W) -> float:
    return float(sum(v.double().abs().sum().item() * (i + 1) for i, (k, v) in enumerate(sorted(w.items()))))




# ----------------------------------------------------------------------------------------------
# synthetic start states / minibatches
# ----------------------------------------------------------------------------------------------
WORKLOADS = {
    # name: (env registry name, S, A, C)
    "point-robot": ("point-robot", 11, 2, 1),
    "cartpole-move": ("cartpole-move", 4, 1, 4),
    "quadrotor": ("quadrotor", 12, 2, 2),
    "tracking": ("tracking", 51, 2, 1),
    "safetygym-point-synthetic": ("safetygym-point-synthetic", 60, 2, 1),
}


def make_start_states(workload: str, n: int, seed: int) -> Tensor:
    g = torch.Generator().manual_seed(seed)
    U = lambda lo, hi: lo + (hi - lo) * torch.rand(n, generator=g)
    _, S, _, _ = WORKLOADS[workload]
    s = torch.randn(n, S, generator=g)
    if workload == "quadrotor":       # constrained_tracking_reset.yaml:14-43 (z shifted by init_z = 1)
        s[:, 0], s[:, 1], s[:, 2], s[:, 3], s[:, 4], s[:, 5] = U(-1.5, 1.5), U(-1, 1), U(0.5, 1.5), U(-1, 1), U(-0.2, 0.2), U(-0.1, 0.1)
    elif workload == "cartpole-move":
        s[:, 0], s[:, 1] = U(-0.8, 0.8), U(-0.15, 0.15)
        s[:, 2:] *= 0.5
    elif workload in ("point-robot", "safetygym-point-synthetic"):
        s[:, 0], s[:, 1] = U(-3.0, 3.0), U(-3.0, 3.0)
    elif workload == "tracking":
        s[:, 0], s[:, 1], s[:, 2] = U(-5, 5), U(-2, 2), U(-math.pi, math.pi)
        s[:, 47], s[:, 48], s[:, 49], s[:, 50] = U(-10, 10), U(-5, 5), U(-math.pi, math.pi), U(-2, 2)
    return s.contiguous()


def make_critic_batch(workload: str, n: int, seed: int) -> List[Tensor]:
    """The 7 batch components in COMPONENT_NAMES order (actions U(-1,1), rewards N(0,1), done~Bern(0.02), cv~N(-0.5,1))."""
    _, S, A, C = WORKLOADS[workload]
    g = torch.Generator().manual_seed(seed)
    obs = make_start_states(workload, n, seed + 1)
    act = torch.rand(n, A, generator=g) * 2 - 1
    nobs = obs + 0.05 * torch.randn(n, S, generator=g)
    rew = torch.randn(n, generator=g)
    done = torch.rand(n, generator=g) < 0.02
    viol = torch.rand(n, generator=g) < 0.02
    cv = torch.randn(n, generator=g) - 0.5 if C == 1 else torch.randn(n, C, generator=g) - 0.5
    return [obs, act, nobs, rew, done, viol, cv]
