"""Device form of the env hook triple ``check_done / check_violation / get_constraint_values``.

The reference calls the env's *numpy* hooks from inside the rollout (src/smbpo.py:63-65,238-240), paying three
device->host->device round trips per step.  Here every supported env is described by a small parameter struct
(``drpo_env_params`` in include/drpo_b200.h) evaluated inside the CUDA kernels with numpy's fp64-on-fp32 semantics.
"""
import math
from dataclasses import dataclass
from typing import Tuple

import torch

from . import _lib


@dataclass
class DeviceEnv:
    kind: str                       # 'point_robot' | 'bounded' | 'tracking'
    state_dim: int
    con_dim: int
    name: str = ""
    action_dim: int = 2
    # point robot: src/env/point_robot.py:11-14
    hazards: Tuple[Tuple[float, float], ...] = ((0.4, -1.2), (-0.4, 1.2))
    hazard_size: float = 0.8
    goal: Tuple[float, float] = (2.2, 2.2)
    goal_size: float = 0.3
    xy_bound: float = 3.0
    # bounded: src/env/poles/constraints.py:216-247
    active_dims: Tuple[int, ...] = ()
    lower: Tuple[float, ...] = ()
    upper: Tuple[float, ...] = ()
    done_dims: Tuple[int, ...] = ()
    done_thr: Tuple[float, ...] = ()
    # tracking: src/env/tracking/pyth_veh3dofconti_surrcstr_data.py:55-72
    surr_veh_num: int = 1
    surr_start: int = 0
    veh_length: float = 4.8
    veh_width: float = 2.0

    def to_params(self) -> "_lib.EnvParams":
        p = _lib.EnvParams()
        p.kind = {"point_robot": _lib.ENV_POINT_ROBOT, "bounded": _lib.ENV_BOUNDED, "tracking": _lib.ENV_TRACKING}[self.kind]
        p.state_dim, p.con_dim = self.state_dim, self.con_dim
        p.n_hazards = len(self.hazards)
        for i, (x, y) in enumerate(self.hazards):
            p.hazard_xy[i][0], p.hazard_xy[i][1] = x, y
        p.hazard_size, p.goal_size, p.xy_bound = self.hazard_size, self.goal_size, self.xy_bound
        p.goal_xy[0], p.goal_xy[1] = self.goal
        p.n_active = len(self.active_dims)
        for i, d in enumerate(self.active_dims):
            p.active_dims[i], p.lower[i], p.upper[i] = d, self.lower[i], self.upper[i]
        p.n_done_dims = len(self.done_dims)
        for i, d in enumerate(self.done_dims):
            p.done_dims[i], p.done_thr[i] = d, self.done_thr[i]
        p.surr_veh_num, p.surr_start = self.surr_veh_num, self.surr_start
        p.veh_length, p.veh_width = self.veh_length, self.veh_width
        return p

    # --- the reference's hook names, on device tensors -------------------------------------------------------
    def evaluate(self, states: torch.Tensor):
        """(done bool[n], violation bool[n], constraint_values f32 [n] or [n,C]) for CUDA float32 states [n,S]."""
        lib = _lib.load()
        assert states.dim() == 2 and states.shape[1] == self.state_dim and states.dtype == torch.float32
        states = states.contiguous()
        n = states.shape[0]
        done = torch.empty(n, dtype=torch.bool, device=states.device)
        viol = torch.empty(n, dtype=torch.bool, device=states.device)
        cv = torch.empty((n, self.con_dim), dtype=torch.float32, device=states.device)
        params = self.to_params()
        _lib.check(lib.drpo_hooks_eval(params, _lib.ptr(states), n, _lib.ptr(done), _lib.ptr(viol), _lib.ptr(cv),
                                       _lib.stream_ptr()), "drpo_hooks_eval")
        return done, viol, (cv.squeeze(1) if self.con_dim == 1 else cv)

    def check_done(self, states):
        return self.evaluate(states)[0]

    def check_violation(self, states):
        return self.evaluate(states)[1]

    def get_constraint_values(self, states):
        return self.evaluate(states)[2]

    get_constrained_values = get_constraint_values      # README spelling (README.md:39)


def point_robot(state_dim: int = 11) -> DeviceEnv:
    return DeviceEnv(kind="point_robot", state_dim=state_dim, con_dim=1, name="point-robot")


def cartpole() -> DeviceEnv:
    """src/env/poles/inverted_pendulum.py:11-37; done == violation (:79-87)."""
    return DeviceEnv(kind="bounded", state_dim=4, con_dim=4, active_dims=(0, 1), lower=(-0.9, -0.2), upper=(0.9, 0.2),
                     name="cartpole-move", action_dim=1)


def quadrotor(x_threshold: float = 2.0, z_threshold: float = 2.0) -> DeviceEnv:
    """src/env/quadrotor/quadrotor.py:46-58,83-114; thresholds come from the external safe-control-gym fork."""
    return DeviceEnv(kind="bounded", state_dim=12, con_dim=2, active_dims=(2,), lower=(0.5,), upper=(1.5,),
                     done_dims=(0, 2, 4), done_thr=(x_threshold, z_threshold, 85 * math.pi / 180), name="quadrotor")


def tracking(pre_horizon: int = 10, surr_veh_num: int = 1) -> DeviceEnv:
    return DeviceEnv(kind="tracking", state_dim=6 + 1 + 4 * pre_horizon + 4 * surr_veh_num, con_dim=1,
                     surr_veh_num=surr_veh_num, surr_start=6 + 1 + 4 * pre_horizon, name="tracking")


def safetygym_point_synthetic() -> DeviceEnv:
    """Synthetic 60-dim stand-in: the safetygym env lives on another branch of the reference (README.md:14)."""
    return DeviceEnv(kind="point_robot", state_dim=60, con_dim=1, name="safetygym-point-synthetic")


REGISTRY = {
    "point-robot": point_robot, "cartpole-move": cartpole, "cartpole-upright": cartpole, "quadrotor": quadrotor,
    "tracking": tracking, "safetygym-point-synthetic": safetygym_point_synthetic,
}


def device_env(name: str, **kwargs) -> DeviceEnv:
    if name not in REGISTRY:
        raise KeyError(f"no device hooks registered for env '{name}' (have: {sorted(REGISTRY)})")
    return REGISTRY[name](**kwargs)
