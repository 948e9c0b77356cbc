"""The hot-path half of the reference's algorithm driver (src/smbpo.py): ``rollout``, ``update_solver`` and
``rollout_and_update`` with the reference's names, arguments and side effects.  Real-env stepping, evaluation, logging
and checkpoint cadence stay with the reference's own driver (SURVEY.md §2 rows 2,12,16: out of scope)."""
import ctypes
import random

import torch
from torch import nn

from . import _lib
from .config import BaseConfig, Configurable
from .dynamics import BatchedGaussianEnsemble
from .envs import DeviceEnv, device_env
from .sampling import ConstraintSafetySampleBuffer, RolloutView
from .ssac import SSAC


class SMBPO(Configurable, nn.Module):
    class Config(BaseConfig):
        sac_cfg = SSAC.Config()
        model_cfg = BatchedGaussianEnsemble.Config()
        model_initial_steps = 10000
        model_steps = 2000
        model_update_period = 250
        save_trajectories = False
        horizon = 10
        alive_bonus = 1.0
        buffer_min = 5000
        buffer_max = 10**6
        steps_per_epoch = 1000
        rollout_batch_size = 100
        solver_updates_per_step = 10
        real_fraction = 0.1
        action_clip_gap = 1e-6
        reward_scale = 1.
        mode = 'train'
        constraint_scale = 10.
        constraint_offset = 0.
        safe_shield = True
        safe_shield_threshold = -0.1
        eval_shield_threshold = -0.05
        eval_shield_type = "linear"

    def __init__(self, config, env, data=None, epochs=100, device=None):
        """``env`` is a DeviceEnv, a registered env name, or an env_factory whose product carries ``.device_env``/``.name``."""
        Configurable.__init__(self, config)
        nn.Module.__init__(self)
        device = torch.device(device if device is not None else "cuda")
        if isinstance(env, str):
            env = device_env(env)
        elif callable(env) and not isinstance(env, DeviceEnv):
            real = env()
            self.real_env = real
            env = getattr(real, "device_env", None) or device_env(getattr(real, "name"))
        assert isinstance(env, DeviceEnv)
        self.env = env
        self.data = data
        self.state_dim, self.con_dim = env.state_dim, env.con_dim
        self.action_dim = env.action_dim
        self._env_params = env.to_params()
        self.check_done, self.check_violation, self.get_constraint_value = env.check_done, env.check_violation, env.get_constraint_values

        self.model_ensemble = BatchedGaussianEnsemble(self.model_cfg, self.state_dim, self.action_dim, device=device)
        self.solver = SSAC(self.sac_cfg, self.state_dim, self.action_dim, self.con_dim, self.horizon, epochs,
                           self.steps_per_epoch, self.solver_updates_per_step, self.constraint_scale, None,
                           self.model_ensemble, device=device)
        # replay buffers are deliberately hidden from state_dict, like the reference's DummyModuleWrapper (src/smbpo.py:100-104)
        object.__setattr__(self, "replay_buffer", self._create_buffer(self.buffer_max, device))
        object.__setattr__(self, "virt_buffer", self._create_buffer(self.buffer_max, device))
        for name in ("episodes_sampled", "steps_sampled", "n_violations", "epochs_completed"):
            self.register_buffer(name, torch.tensor(0, device=device))
        self.recent_critic_losses, self.recent_cons_critic_losses = [], []
        self._ws = _lib.Workspace()
        self.rollout_precision = _lib.PREC_FP32
        self.rollout_seed = 0x0DDB411
        self._rollouts_done = 0
        self.shard_rank, self.shard_world = 0, 1      # multi-GPU rollout sharding (trajectory ids are global)

    @property
    def actor(self):
        return self.solver.actor

    @property
    def constraint_critic(self):
        return self.solver.constraint_critic

    @property
    def actor_safe(self):
        return self.solver.actor_safe

    def shielded_act1(self, state, eps=None):
        """The action of one training step (src/smbpo.py:124-136): a sampled performance action, replaced by the safe actor's
        eval action when the shield is on and _get_qc(constraint_critic(s, a, uncertainty=distributional_qc)) exceeds
        safe_shield_threshold.  One drpo_shield_act call on a single row."""
        s = state.unsqueeze(0)
        if not self.safe_shield:
            return self.actor.act(s, False, eps)[0]
        return self.solver.shield_act(s, eval=False, shield_type="safe", safe_shield_threshold=self.safe_shield_threshold,
                                      uncertainty=self.solver.distributional_qc, eps=eps)[0]

    def _create_buffer(self, capacity, device=None):
        return ConstraintSafetySampleBuffer(self.state_dim, self.action_dim, capacity, con_dim=self.con_dim,
                                            device=device or self.virt_buffer.device)

    # ---- SMBPO.rollout  (src/smbpo.py:229-249) -------------------------------------------------------------------
    def rollout(self, policy, initial_states=None, noise=None, member_idx=None, _debug_layer=None):
        """Branched H-step model rollout.  Appends the transitions to ``self.virt_buffer`` (one write, on the device)
        and returns a view of them.  ``noise=(eps_policy [H,B,A], eps_model [H,B,S+1])`` injects the Gaussian draws
        indexed by original trajectory id (parity); ``member_idx`` overrides the per-step elite picks."""
        lib = _lib.load()
        _lib.peek_kernel_status("SMBPO.rollout")
        if initial_states is None:
            states = self.replay_buffer.get('states')
            idx = torch.randperm(states.shape[0], device=states.device)[:self.rollout_batch_size]
            initial_states = states.index_select(0, idx)
        initial_states = initial_states.contiguous().float()
        B, H = initial_states.shape[0], self.horizon
        # Host start states (bf16 path): the host-to-device transfer runs in row blocks on a side stream, every block followed by a
        # DMA write of its ready flag, and the first step's kernel waits per block - the copy overlaps that step
        flags = copy_stream = None
        if initial_states.device.type == "cpu" and self.virt_buffer.device.type == "cuda":
            dev = self.virt_buffer.device
            if self.rollout_precision == _lib.PREC_BF16 and initial_states.is_pinned() and B >= (1 << 16) and _debug_layer is None:
                host, shift = initial_states, 17                                    # 131 072-row blocks: few host calls per transfer
                nblk = (B + (1 << shift) - 1) >> shift
                initial_states = torch.empty_like(host, device=dev)
                flags = torch.zeros(nblk, dtype=torch.int32, device=dev)
                if getattr(self, "_flag_ones", None) is None or self._flag_ones.numel() < nblk:
                    self._flag_ones = torch.ones(max(nblk, 64), dtype=torch.int32).pin_memory()
                    self._copy_stream = torch.cuda.Stream(device=dev)
                copy_stream = self._copy_stream
                copy_stream.wait_stream(torch.cuda.current_stream(dev))             # the flags are zero before any copy starts
                with torch.cuda.stream(copy_stream):
                    for j in range(nblk):
                        lo, hi = j << shift, min((j + 1) << shift, B)
                        initial_states[lo:hi].copy_(host[lo:hi], non_blocking=True)
                        flags[j:j + 1].copy_(self._flag_ones[j:j + 1], non_blocking=True)     # copy engine, not a kernel
            else:
                initial_states = initial_states.to(dev, non_blocking=True)
        if member_idx is None:        # one host-side random.choice per step, as BatchedGaussianEnsemble.sample does (:199)
            member_idx = [random.choice(self.model_ensemble._elite_inds) for _ in range(H)]
        members = (ctypes.c_int32 * H)(*[int(m) for m in member_idx])
        ring = self.virt_buffer
        counts = torch.zeros(H + 1, dtype=torch.int32, device=initial_states.device)
        start = ring._pointer.clone()
        actor_s, ens_s = policy.as_struct(), self.model_ensemble.as_struct()
        a = _lib.RolloutArgs()
        a.actor, a.ensemble, a.env = ctypes.pointer(actor_s), ctypes.pointer(ens_s), ctypes.pointer(self._env_params)
        a.initial_states, a.batch, a.horizon = _lib.ptr(initial_states), B, H
        # global id of this shard's row 0 (even row split by default; uneven splits set traj_id_offset explicitly)
        off = getattr(self, "_traj_id_offset_override", None)
        a.traj_id_offset = self.shard_rank * B if off is None else int(off)
        a.member_idx_host = members
        keep = None
        if noise is not None:
            keep = [n.contiguous().float() for n in noise]
            a.eps_policy, a.eps_model, a.eps_batch_stride = _lib.ptr(keep[0]), _lib.ptr(keep[1]), keep[0].shape[1]
        self._rollouts_done += 1
        a.seed = self.rollout_seed + self._rollouts_done
        a.virt, a.step_counts, a.precision = ring.as_struct(), _lib.ptr(counts), self.rollout_precision
        ws = self._ws.get(lib.drpo_rollout_workspace_bytes(a), initial_states.device)
        a.workspace, a.workspace_bytes, a.stream = _lib.ptr(ws), ws.numel(), _lib.stream_ptr()
        if flags is not None:
            a.init_ready_flags, a.init_rows_per_flag_log2 = _lib.ptr(flags), 17
        if _debug_layer is not None:
            n_out = 2048 if _debug_layer == 100 else [policy.net[0].weight.shape[0], policy.net[2].weight.shape[0], policy.net[4].weight.shape[0],
                     self.model_ensemble.hidden_dim, self.model_ensemble.hidden_dim, self.model_ensemble.hidden_dim,
                     self.model_ensemble.hidden_dim, self.state_dim + 1, self.state_dim + 1][_debug_layer]
            out = torch.zeros((max(B, 8), n_out), device=initial_states.device)
            _lib.check(lib.drpo_debug_rollout_layer(a, _debug_layer, _lib.ptr(out)), "drpo_debug_rollout_layer")
            return out
        _lib.check(lib.drpo_rollout(a), "drpo_rollout")
        if copy_stream is not None:
            cur = torch.cuda.current_stream(initial_states.device)
            cur.wait_stream(copy_stream)                    # later work on this stream is ordered after the transfer as usual
            initial_states.record_stream(copy_stream); flags.record_stream(copy_stream)
        return RolloutView(ring, start, counts)

    # ---- SMBPO.update_solver  (src/smbpo.py:251-279) ---------------------------------------------------------------
    def sample_batch(self, batch_size=None):
        """replay/virtual mix + reward and constraint scaling in one gather kernel (src/smbpo.py:253-270)."""
        lib = _lib.load()
        _lib.peek_kernel_status("SMBPO.sample_batch")
        solver = self.solver
        B = batch_size or solver.batch_size
        n_real = int(self.real_fraction * B)
        dev = self.virt_buffer.device
        idx = torch.cat([torch.randint(max(len(self.replay_buffer), 1), [n_real], device=dev),
                         torch.randint(max(len(self.virt_buffer), 1), [B - n_real], device=dev)])
        S, A, C = self.state_dim, self.action_dim, self.con_dim
        out = [torch.empty((B, S), device=dev), torch.empty((B, A), device=dev), torch.empty((B, S), device=dev),
               torch.empty((B,), device=dev), torch.empty((B,), dtype=torch.bool, device=dev),
               torch.empty((B,), dtype=torch.bool, device=dev),
               torch.empty((B,) if C == 1 else (B, C), device=dev)]
        batch = _lib.Batch(*[_lib.ptr(t) for t in out])
        real_s, virt_s = self.replay_buffer.as_struct(), self.virt_buffer.as_struct()
        _lib.check(lib.drpo_buffer_gather(real_s, virt_s, _lib.ptr(idx), n_real, B, float(self.reward_scale), float(self.alive_bonus),
                                          float(self.constraint_scale), float(self.constraint_offset), batch, _lib.stream_ptr()),
                   "drpo_buffer_gather")
        return out

    def update_solver(self, update_actor=True, update_multiplier=False):
        samples = self.sample_batch()
        critic_loss, constraint_critic_loss = self.solver.update_critic(*samples)
        self.recent_critic_losses.append(critic_loss)
        self.recent_cons_critic_losses.append(constraint_critic_loss)
        if update_actor:
            self.solver.update_actor_and_alpha(samples[0])
        if update_multiplier:
            self.solver.update_multiplier(samples[0])

    def rollout_and_update(self):
        """src/smbpo.py:281-291."""
        self.rollout(self.actor)
        for step in range(self.solver_updates_per_step):
            self.update_solver(update_actor=(step % self.sac_cfg.actor_update_interval == 0),
                               update_multiplier=(step % self.sac_cfg.multiplier_update_interval == 0))

    # ---- Philox stream positions (not part of the reference's state_dict; a checkpointing caller saves them beside it) --------
    def noise_state(self):
        """Counters that key the in-kernel Philox streams.  A resumed run that restores them continues the streams instead of
        replaying them (the reference has no equivalent: torch's global generator is not checkpointed there either)."""
        sv = self.solver
        return dict(rollouts_done=self._rollouts_done, rollout_seed=self.rollout_seed, solver_seed=sv.noise_seed,
                    actor_step=sv.actor._noise_step, actor_safe_step=sv.actor_safe._noise_step,
                    qc_step=sv.constraint_critic._noise_step, ensemble_step=self.model_ensemble._noise_step,
                    opt_steps=[o.step_count for o in (sv.critic_optimizer, sv.multiplier_optimizer, sv.actor_optimizer, sv.actor_safe_optimizer)])

    def load_noise_state(self, st):
        sv = self.solver
        self._rollouts_done, self.rollout_seed, sv.noise_seed = st["rollouts_done"], st["rollout_seed"], st["solver_seed"]
        sv.actor._noise_step, sv.actor_safe._noise_step = st["actor_step"], st["actor_safe_step"]
        sv.constraint_critic._noise_step, self.model_ensemble._noise_step = st["qc_step"], st["ensemble_step"]
        for o, n in zip((sv.critic_optimizer, sv.multiplier_optimizer, sv.actor_optimizer, sv.actor_safe_optimizer), st["opt_steps"]):
            o.step_count = n

    def update_models(self, model_steps):
        """src/smbpo.py:214-227."""
        losses = self.model_ensemble.fit(self.replay_buffer, steps=model_steps)
        r = self.replay_buffer.get('rewards')
        self.solver.update_r_bounds(r.min().item() + self.alive_bonus, r.max().item() + self.alive_bonus)
        return losses
