"""Drop-in behind the reference's own driver.

``make_dropin(ref_smbpo_module)`` returns a subclass of the REFERENCE's ``src.smbpo.SMBPO`` in which the hot path is the B200
one and everything else is the reference's unmodified code:

  reference code that keeps running     ``setup``, ``step_generator`` (real-env stepping, safety shield call sites, the runtime
                                        asserts of src/smbpo.py:158-163), ``epoch``, ``evaluate``, ``log_statistics``, ``update_models``,
                                        ``rollout_and_update``, ``UniformPolicy``, logging, checkpointing, ``main.py``
  replaced (same names / signatures)    ``model_ensemble`` -> drpo_b200.BatchedGaussianEnsemble, ``solver`` -> drpo_b200.SSAC (the
                                        reference's ctor instantiates both by module-level name, src/smbpo.py:67-71),
                                        ``replay_buffer`` / ``virt_buffer`` -> drpo_b200.ConstraintSafetySampleBuffer (``_create_buffer``),
                                        ``rollout`` -> drpo_rollout, ``update_solver`` -> drpo_buffer_gather + drpo_critic_step /
                                        drpo_actor_step / drpo_multiplier_step

Use from ``main.py``:   ``from src import smbpo as ref; SMBPO = drpo_b200.dropin.make_dropin(ref)``.
"""
from . import _lib
from .dynamics import BatchedGaussianEnsemble
from .envs import DeviceEnv, device_env
from .sampling import ConstraintSafetySampleBuffer
from .smbpo import SMBPO as _B200SMBPO
from .ssac import SSAC


def _device_env_of(real_env):
    """Device form of the env hooks (env-kind enum + parameter struct) for the reference env object: its own ``device_env``
    attribute if it carries one, else the registry entry of its class."""
    env = real_env
    for _ in range(8):                                         # unwrap gym wrappers
        de = getattr(env, "device_env", None)
        if isinstance(de, DeviceEnv):
            return de
        name = type(env).__name__
        reg = {"PointRobot": "point-robot", "SafeInvertedPendulumEnv": "cartpole-move", "QuadrotorWrapperEnv": "quadrotor",
               "SimuVeh3dofcontiSurrCstr": "tracking"}.get(name)
        if reg is not None:
            kw = {}
            if reg == "tracking":
                kw = dict(pre_horizon=getattr(env, "pre_horizon", 10), surr_veh_num=getattr(env, "surr_veh_num", 1))
            return device_env(reg, **kw)
        if not hasattr(env, "env"):
            break
        env = env.env
    raise RuntimeError(f"no device form of the env hooks is registered for {type(real_env).__name__}: give the env a "
                       "`device_env` attribute (drpo_b200.envs.DeviceEnv)")


def make_dropin(ref, precision=_lib.PREC_BF16, device=None):
    """``ref`` = the reference's imported ``src.smbpo`` module."""
    import torch

    class DropInSMBPO(ref.SMBPO):
        def __init__(self, config, env_factory, data, epochs):
            dev = torch.device(device if device is not None else "cuda")
            object.__setattr__(self, "_b200_device", dev)
            saved = (ref.BatchedGaussianEnsemble, ref.SSAC)
            # src/smbpo.py:67-71 constructs both learners by these module-level names with the reference's own argument lists
            ref.BatchedGaussianEnsemble = lambda cfg, S, A: BatchedGaussianEnsemble(cfg, S, A, device=dev)
            ref.SSAC = lambda *a: SSAC(*a, device=dev)
            try:
                super().__init__(config, env_factory, data, epochs)
            finally:
                ref.BatchedGaussianEnsemble, ref.SSAC = saved
            self._env_params = _device_env_of(self.real_env).to_params()
            self._ws = _lib.Workspace()
            self.rollout_precision = precision
            self.solver.precision = precision
            self.rollout_seed, self._rollouts_done = 0x0DDB411, 0
            self.shard_rank, self.shard_world = 0, 1

        def _create_buffer(self, capacity):
            # a plain (non-nn.Module-registered) attribute, like the reference's DummyModuleWrapper: hidden from state_dict
            return _BufferHandle(ConstraintSafetySampleBuffer(self.state_dim, self.action_dim, capacity, con_dim=self.con_dim,
                                                              device=self._b200_device))

        rollout = _B200SMBPO.rollout
        sample_batch = _B200SMBPO.sample_batch
        update_solver = _B200SMBPO.update_solver
        noise_state = _B200SMBPO.noise_state
        load_noise_state = _B200SMBPO.load_noise_state

    DropInSMBPO.__name__ = "SMBPO"
    return DropInSMBPO


class _BufferHandle:
    """Keeps the replay buffers out of the parent's state_dict (the role of the reference's DummyModuleWrapper, src/torch_util.py:116-133)
    while forwarding everything to the device ring buffer."""

    def __init__(self, buf):
        self.__dict__["_buf"] = buf

    def __getattr__(self, name):
        return getattr(self.__dict__["_buf"], name)

    def __setattr__(self, name, value):
        setattr(self.__dict__["_buf"], name, value)

    def __len__(self):
        return len(self.__dict__["_buf"])
