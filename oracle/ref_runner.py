"""TEST / BASELINE INFRASTRUCTURE ONLY — drives the UNMODIFIED reference (imported through ``oracle/ref_shim.py``) on the
CPU for ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs and for the drop-in test.

Nothing here is on the product path.  The reference's own classes run: ``src.smbpo.SMBPO.rollout`` (src/smbpo.py:229-249),
``src.ssac.SSAC.update_critic`` (src/ssac.py:437-456), ``update_multiplier`` (:570-578), ``update_actor_and_alpha`` (:507-527).
Envs whose simulators are not installable (MuJoCo cartpole, safe-control-gym quadrotor) are replaced by a stand-in that carries
the reference's own ``BoundedConstraint`` hook code (src/env/poles/constraints.py), exactly as ``oracle/make_golden.py`` does.
"""
import math
import pathlib
import tempfile

import numpy as np
import torch

from oracle import ref_shim
from oracle import drpo_oracle as O

_SPECS = {"quadrotor": O.env_quadrotor, "cartpole-move": O.env_cartpole, "point-robot": O.env_point_robot,
          "safetygym-point-synthetic": O.env_safetygym60, "tracking": lambda: O.env_tracking(10, 1)}


def available() -> bool:
    return ref_shim.reference_available()


def _reference_hooks(spec):
    """The reference's own hook code per env kind (same construction as oracle/make_golden.py:reference_hooks)."""
    if spec.name == "point-robot":
        from src.env.point_robot import PointRobot
        env = PointRobot()
        return env.check_done, env.check_violation, env.get_constraint_values
    if spec.kind == "bounded":
        from src.env.poles.constraints import BoundedConstraint, ConstrainedVariableType
        cons = BoundedConstraint(spec.state_dim, lower_bounds=list(spec.lower), upper_bounds=list(spec.upper),
                                 constrained_variable=ConstrainedVariableType.STATE, active_dims=list(spec.active_dims))
        if spec.name == "cartpole-move":                                  # src/env/poles/inverted_pendulum.py:79-121
            return cons.is_violated, cons.is_violated, lambda s: np.squeeze(cons.get_value(s))

        def check_done(states):                                           # src/env/quadrotor/quadrotor.py:83-114
            thr = list(spec.done_thr) + [85 * math.pi / 180] * 3
            done = np.zeros(states.shape[:-1], dtype=bool)
            for d, t in zip(spec.done_dims, thr):
                done = done | (states[..., d] < -t) | (states[..., d] > t)
            return np.logical_or(done, cons.is_violated(states))
        return check_done, cons.is_violated, lambda s: np.squeeze(cons.get_value(s))
    if spec.kind == "tracking":
        from src.env.tracking.pyth_veh3dofconti_surrcstr_data import SimuVeh3dofcontiSurrCstr
        env = SimuVeh3dofcontiSurrCstr(pre_horizon=10, surr_veh_num=spec.surr_veh_num)
        return env.check_done, env.check_violation, env.get_constraint_values
    # point-robot style synthetic hazards at other state dims (safetygym-60): the oracle's numpy restatement of the same formulas
    return (lambda s: O.hooks(spec, s)[0]), (lambda s: O.hooks(spec, s)[1]), (lambda s: O.hooks(spec, s)[2])


class HookEnv:
    """Env stand-in handing the reference's SMBPO / SSAC the dims and the hook triple of ``spec`` (no simulator)."""
    _max_episode_steps = 1000

    def __init__(self, spec, action_dim, id=None):
        import gym
        self.spec_, self.con_dim = spec, spec.con_dim
        self.observation_space = gym.spaces.Box(-np.inf, np.inf, shape=(spec.state_dim,), dtype=np.float32)
        self.action_space = gym.spaces.Box(-1.0, 1.0, shape=(action_dim,), dtype=np.float32)
        self.check_done, self.check_violation, self.get_constraint_values = _reference_hooks(spec)

    def reset(self):
        return np.zeros(self.spec_.state_dim, dtype=np.float32)

    def step(self, action):
        raise RuntimeError("HookEnv has no simulator: only the model-rollout / update paths of the reference are driven")

    def seed(self, seed=None):
        return [seed]


def make_reference_smbpo(workload, S, A, C, B0, H, w_model, w_ssac, critic_batch=256, std_ratio=2.0, device="cpu"):
    """The reference's own SMBPO with the given seeded weights loaded (CPU unless bench.py's reference-on-GPU leg asks otherwise)."""
    ref_shim.import_reference(device)
    from src.checkpoint import CheckpointableData
    from src.env.torch_wrapper import TorchWrapper
    from src.log import default_log as log
    from src.smbpo import SMBPO
    if getattr(log, "dir", None) is None:
        log.setup(pathlib.Path(tempfile.mkdtemp()))
    spec = _SPECS[workload]()
    cfg = SMBPO.Config()
    cfg.rollout_batch_size, cfg.horizon = B0, H
    cfg.buffer_max = max(B0 * H + 1024, 4096)
    cfg.sac_cfg.batch_size = critic_batch
    cfg.sac_cfg.constraint_critic_cfg.std_ratio = std_ratio
    cfg.sac_cfg.target_entropy = -float(A)                                # the CLI resolves the Optional placeholder (src/config.py:16-30)
    if workload == "point-robot":
        from src.env.point_robot import PointRobot
        factory = lambda id=None: TorchWrapper(PointRobot(id=id))
    else:
        factory = lambda id=None: TorchWrapper(HookEnv(spec, A, id=id))
    alg = SMBPO(cfg, factory, CheckpointableData(), 10)
    alg.to(device)
    alg.model_ensemble.load_state_dict(w_model, strict=True)
    alg.solver.load_state_dict(dict(w_ssac), strict=False)
    alg.model_ensemble._elite_inds = [0, 1, 2, 3, 4]
    return alg


def rollout_runner(workload, S, A, C, B0, H, w_model, w_ssac, init, device="cpu"):
    """callable() -> transitions written by ONE reference SMBPO.rollout of ``init`` (virt_buffer reset before every call)."""
    alg = make_reference_smbpo(workload, S, A, C, B0, H, w_model, w_ssac, device=device)

    def run():
        alg.virt_buffer._pointer = 0 if isinstance(alg.virt_buffer._pointer, int) else alg.virt_buffer._pointer * 0
        buf = alg.rollout(alg.actor, initial_states=init)
        return len(buf)
    return run


def critic_runner(S, A, C, B, w_ssac, batch, std_ratio=1.0):
    """(update_critic, update_actor, update_multiplier) callables on the reference's own SSAC at minibatch ``B``."""
    ref_shim.import_reference()
    from src.ssac import SSAC
    spec = O.env_tracking(10, 1) if S == 51 else O.env_point_robot(S)
    cfg = SSAC.Config()
    cfg.batch_size = B                                                    # src/ssac.py:356-361 asserts on self.batch_size
    cfg.constraint_critic_cfg.std_ratio = std_ratio
    cfg.target_entropy = -float(A)                                        # the CLI resolves the Optional placeholder (src/config.py:16-30)
    solver = SSAC(cfg, S, A, C, 10, 100, 1000, 10, 5.0, lambda: HookEnv(spec, A), None)
    solver.load_state_dict(dict(w_ssac), strict=False)
    batch = [t.clone() for t in batch]
    return (lambda: solver.update_critic(*batch)), (lambda: solver.update_actor_and_alpha(batch[0])), \
           (lambda: solver.update_multiplier(batch[0]))
