"""Safe SAC learner with the reference's API (src/ssac.py).  ``update_critic`` and ``update_multiplier`` — the hot
path — run as hand-written CUDA behind the C ABI (forward, hand-written backward, clip, Adam, EMA); parameters live in
flat fp32 arenas that the ``nn.Parameter`` views alias, so ``state_dict`` round-trips the reference keys."""
import copy
import math
import random

import torch
from torch import nn
import torch.nn.functional as F

from . import _lib
from .config import BaseConfig, Configurable, Optional
from .policy import BasePolicy, SquashedGaussianPolicy


class Squeeze(nn.Module):
    def __init__(self, dim=None):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        return x.squeeze(dim=self.dim)


_ACTS = {"relu": nn.ReLU, "tanh": nn.Tanh, "identity": nn.Identity, "swish": nn.SiLU}


def mlp(dims, activation="relu", output_activation=None, squeeze_output=False, device=None):
    """src/torch_util.py:190-211: xavier-normal weights, zero biases."""
    layers = []
    for i in range(len(dims) - 2):
        layers += [nn.Linear(dims[i], dims[i + 1], device=device), _ACTS[activation]()]
    layers.append(nn.Linear(dims[-2], dims[-1], device=device))
    if output_activation is not None:
        layers.append(_ACTS[output_activation]())
    if squeeze_output and dims[-1] == 1:
        layers.append(Squeeze(1))
    net = nn.Sequential(*layers)
    for m in net:
        if isinstance(m, nn.Linear):
            nn.init.xavier_normal_(m.weight)
            nn.init.zeros_(m.bias)
    return net


def _mlp3_struct(seq):
    return _lib.Mlp3(_lib.linear_of(seq[0].weight.data, seq[0].bias.data), _lib.linear_of(seq[2].weight.data, seq[2].bias.data),
                     _lib.linear_of(seq[4].weight.data, seq[4].bias.data), None)


class CriticEnsemble(Configurable, nn.Module):
    class Config(BaseConfig):
        n_critics = 2
        hidden_layers = 2
        hidden_dim = 256

    def __init__(self, config, state_dim, action_dim, device=None):
        Configurable.__init__(self, config)
        nn.Module.__init__(self)
        if self.n_critics != 2 or self.hidden_layers != 2:
            raise NotImplementedError("the CUDA critic step implements the reference's twin 2-hidden-layer Q nets")
        dims = [state_dim + action_dim, *([self.hidden_dim] * self.hidden_layers), 1]
        self.qs = nn.ModuleList([mlp(dims, squeeze_output=True, device=device) for _ in range(self.n_critics)])

    def all(self, state, action):
        sa = torch.cat([state, action], -1)
        return [q(sa) for q in self.qs]

    def min(self, state, action):
        return torch.min(*self.all(state, action))

    def mean(self, state, action):
        qs = self.all(state, action)
        return sum(qs) / len(qs)

    def random_choice(self, state, action):
        sa = torch.cat([state, action], -1)
        return random.choice(self.qs)(sa)


class ConstraintCritic(Configurable, nn.Module):
    """Distributional reachability certificate (src/ssac.py:46-92): mean head + soft-clamped log-std head."""

    class Config(BaseConfig):
        trunk_layers = 2
        head_layers = 1
        hidden_dim = 256
        log_std_min = -4.
        log_std_max = 4.
        std_ratio = 2.

    def __init__(self, config, state_dim, action_dim, output_dim, output_activation=None, device=None):
        Configurable.__init__(self, config)
        nn.Module.__init__(self)
        if self.trunk_layers != 2 or self.head_layers != 1 or (self.log_std_min, self.log_std_max) != (-4., 4.):
            raise NotImplementedError("the CUDA constraint critic implements the reference's default topology")
        H = self.hidden_dim
        self.state_dim, self.action_dim, self.output_dim = state_dim, action_dim, output_dim
        self.trunk = mlp([state_dim + action_dim, H, H], output_activation="relu", device=device)
        self.mean_head = mlp([H, H, output_dim], squeeze_output=True, device=device)
        self.log_std_head = mlp([H, H, output_dim], squeeze_output=True, device=device)
        self._ws = _lib.Workspace()
        self.noise_seed = 0x0C0C
        self._noise_step = 0

    def as_struct(self) -> "_lib.Qc":
        t, m, l = self.trunk, self.mean_head, self.log_std_head
        lin = _lib.linear_of
        return _lib.Qc(lin(t[0].weight.data, t[0].bias.data), lin(t[2].weight.data, t[2].bias.data),
                       lin(m[0].weight.data, m[0].bias.data), lin(m[2].weight.data, m[2].bias.data),
                       lin(l[0].weight.data, l[0].bias.data), lin(l[2].weight.data, l[2].bias.data))

    def forward_torch(self, state, action, uncertainty=False, sample=False):
        """Differentiable torch form for callers that backpropagate through the constraint critic themselves (``actor_loss``); the
        update steps and every no-grad call go through the C ABI."""
        h = self.trunk(torch.cat([state, action], -1))
        mean = self.mean_head(h)
        if (not uncertainty) and (not sample):
            return mean
        ls = self.log_std_head(h)
        ls = self.log_std_max - F.softplus(self.log_std_max - ls)
        ls = self.log_std_min + F.softplus(ls - self.log_std_min)
        std = ls.exp()
        if uncertainty:
            return mean + self.std_ratio * std
        return mean, std, mean + torch.clamp(torch.randn_like(std), -2., 2.) * std

    def forward(self, state, action, uncertainty=False, sample=False, eps=None):
        assert not (uncertainty and sample), "Uncertainty bound and sample cannot be True simultaneously."
        # Inputs that carry autograd history get the differentiable torch form.  Every other call runs through the C ABI and its
        # outputs are DETACHED from the critic's weights: a caller that wants d(loss)/d(weights) must call forward_torch
        # explicitly (the reference's own weight-gradient users, cons_critic_loss_given_target and actor_loss, are
        # drpo_critic_step / drpo_actor_step here).
        if torch.is_grad_enabled() and (state.requires_grad or action.requires_grad):
            return self.forward_torch(state, action, uncertainty, sample)
        lib = _lib.load()
        state, action = state.contiguous().float(), action.contiguous().float()
        B, C = state.shape[0], self.output_dim
        mode = 1 if uncertainty else (2 if sample else 0)
        mean = torch.empty((B, C), device=state.device)
        std = torch.empty((B, C), device=state.device) if mode == 2 else None
        out = torch.empty((B, C), device=state.device) if mode else None
        self._noise_step += 1
        eps = eps.contiguous().float() if eps is not None else None        # a named local: the copy must outlive the launch
        noise = _lib.Noise(_lib.ptr(eps), C, self.noise_seed, 18, self._noise_step)
        ws = self._ws.get(lib.drpo_qc_workspace_bytes(B, self.hidden_dim), state.device)
        _lib.check(lib.drpo_qc_forward(self.as_struct(), _lib.ptr(state), _lib.ptr(action), B, self.state_dim, self.action_dim, C,
                                       mode, float(self.std_ratio), noise, _lib.ptr(mean), _lib.ptr(std), _lib.ptr(out),
                                       _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "drpo_qc_forward")
        sq = (lambda t: t.squeeze(1)) if C == 1 else (lambda t: t)
        if mode == 0:
            return sq(mean)
        if mode == 1:
            return sq(out)
        return sq(mean), sq(std), sq(out)


class MLPMultiplier(Configurable, nn.Module):
    class Config(BaseConfig):
        hidden_layers = 2
        hidden_dim = 256
        upper_bound = 50.

    def __init__(self, config, state_dim, device=None):
        Configurable.__init__(self, config)
        nn.Module.__init__(self)
        if self.hidden_layers != 2:
            raise NotImplementedError("the CUDA multiplier step implements the reference's 2-hidden-layer net")
        dims = [state_dim + 1, *([self.hidden_dim] * self.hidden_layers), 1]
        self.lam = mlp(dims, activation="tanh", output_activation="identity", squeeze_output=True, device=device)

    def forward(self, state, Qc):
        states_aug = torch.cat([state, Qc.unsqueeze(-1)], -1)
        return self.upper_bound / 2. * (1. + torch.tanh(self.lam(states_aug) / self.upper_bound * 2))


class CosineLR:
    """torch.optim.lr_scheduler.CosineAnnealingLR's recursive rule, kept on the host (src/ssac.py:204-208)."""

    def __init__(self, opt, T_max, eta_min):
        self.opt, self.T_max, self.eta_min, self.last_epoch = opt, max(int(T_max), 1), eta_min, 0
        self.base_lr = opt.param_groups[0]["lr"]

    def step(self):
        self.last_epoch += 1
        t, T, lr = self.last_epoch, self.T_max, self.opt.param_groups[0]["lr"]
        if (t - 1 - T) % (2 * T) == 0:
            lr = lr + (self.base_lr - self.eta_min) * (1 - math.cos(math.pi / T)) / 2
        else:
            lr = (1 + math.cos(math.pi * t / T)) / (1 + math.cos(math.pi * (t - 1) / T)) * (lr - self.eta_min) + self.eta_min
        self.opt.param_groups[0]["lr"] = lr

    def get_last_lr(self):
        return [self.opt.param_groups[0]["lr"]]


class FusedAdam:
    """State of the in-kernel Adam (torch.optim.Adam semantics, coupled L2): flat m/v arenas + step count."""

    def __init__(self, arena, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, grad_storage=None):
        """``grad_storage`` (optional): a flat tensor of ``arena.numel() + LOSSES_LEN`` floats to carve the gradient arena from.  The
        step's ``losses`` array (incl. the watchdog slot) is the TAIL of the gradient storage, so a data-parallel step all-reduces
        gradients, losses and the watchdog flag in ONE NCCL message (``grad_full``)."""
        self.param_groups = [dict(lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)]
        self.m, self.v = torch.zeros_like(arena), torch.zeros_like(arena)
        n = arena.numel()
        self.grad_full = grad_storage if grad_storage is not None else torch.zeros(n + _lib.LOSSES_LEN, dtype=arena.dtype, device=arena.device)
        self.grad, self.losses = self.grad_full[:n], self.grad_full[n:n + _lib.LOSSES_LEN]
        self.step_count = 0

    def to(self, device):
        if self.m.device != device:
            n = self.grad.numel()
            self.m, self.v, self.grad_full = self.m.to(device), self.v.to(device), self.grad_full.to(device)
            self.grad, self.losses = self.grad_full[:n], self.grad_full[n:n + _lib.LOSSES_LEN]

    def zero_grad(self):
        pass                                   # every gradient element is overwritten by the backward kernels

    def as_struct(self) -> "_lib.Adam":
        g = self.param_groups[0]
        return _lib.Adam(g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self.step_count)

    def state_dict(self):
        return dict(m=self.m, v=self.v, step=self.step_count, lr=self.param_groups[0]["lr"])

    def load_state_dict(self, sd):
        self.m.copy_(sd["m"]); self.v.copy_(sd["v"]); self.step_count = sd["step"]; self.param_groups[0]["lr"] = sd["lr"]


ARENA_ALIGN = 64      # floats (256 B): every parameter starts on the alignment the tensor-core GEMM kernels are chosen for


def _arena_offsets(params):
    offs, off = [], 0
    for p in params:
        off = (off + ARENA_ALIGN - 1) // ARENA_ALIGN * ARENA_ALIGN
        offs.append(off)
        off += p.numel()
    return offs, (off + ARENA_ALIGN - 1) // ARENA_ALIGN * ARENA_ALIGN


def _flatten_into_arena(params):
    """Move parameters into one flat fp32 tensor; each nn.Parameter becomes a view of it (state_dict order).  The gaps that
    align each parameter are zeros and stay zeros (zero gradient, zero weight decay), so norms, Adam and EMA are unaffected."""
    offs, n = _arena_offsets(params)
    arena = torch.zeros(n, dtype=torch.float32, device=params[0].device)
    for p, off in zip(params, offs):
        k = p.numel()
        arena[off:off + k].copy_(p.data.reshape(-1))
        p.data = arena[off:off + k].view(p.shape)
    return arena


def _arena_ok(arena, params):
    offs, n = _arena_offsets(params)
    if arena.numel() != n:
        return False
    for p, off in zip(params, offs):
        if p.data_ptr() != arena.data_ptr() + 4 * off or not p.is_contiguous():
            return False
    return True


class SSAC(Configurable, BasePolicy, nn.Module):
    class Config(BaseConfig):
        discount = 0.99
        init_alpha = 1.0
        autotune_alpha = True
        target_entropy = Optional(float)
        use_log_alpha_loss = False
        deterministic_backup = False
        critic_update_multiplier = 1
        actor_lr = 8e-5
        actor_lr_end = 4e-5
        critic_lr = 3e-4
        critic_lr_end = 8e-5
        multiplier_lr = 3e-4
        multiplier_lr_end = 1e-5
        critic_cfg = CriticEnsemble.Config()
        constraint_critic_cfg = ConstraintCritic.Config()
        mlp_multiplier_cfg = MLPMultiplier.Config()
        tau = 0.005
        actor_update_interval = 2
        batch_size = 256
        hidden_dim = 256
        hidden_layers = 2
        update_violation_cost = False
        grad_norm = 5.
        constraint_threshold = 0.
        constrained_fcn = 'reachability'
        mlp_multiplier = True
        penalty_lb = -1.0
        penalty_ub = 100.
        fixed_multiplier = 15.0
        multiplier_update_interval = 5
        lam_epsilon = 1.0
        qc_under_uncertainty = True
        qc_td_bound = 5.
        distributional_qc = True

    def __init__(self, config, state_dim, action_dim, con_dim, horizon, epochs, steps_per_epoch, solver_updates_per_step,
                 constraint_scale, env_factory=None, model_ensemble=None, optimizer_factory=torch.optim.Adam, device=None):
        Configurable.__init__(self, config)
        nn.Module.__init__(self)
        if self.use_log_alpha_loss:
            raise NotImplementedError("drpo_actor_step implements the reference default use_log_alpha_loss=False (alpha-coefficient "
                                      "temperature loss, src/ssac.py:497-500)")
        if not (self.constrained_fcn == 'reachability' and self.mlp_multiplier and self.qc_under_uncertainty
                and self.distributional_qc and not self.deterministic_backup):
            raise NotImplementedError("the CUDA SSAC implements DRPO mode: reachability + mlp_multiplier + "
                                      "qc_under_uncertainty + distributional_qc (run.sh:10-16)")
        device = torch.device(device if device is not None else "cuda")
        self.device_ = device
        self.state_dim, self.action_dim, self.con_dim, self.horizon = state_dim, action_dim, con_dim, horizon
        self.violation_cost = 0.0
        self.updates_per_training = epochs * steps_per_epoch * solver_updates_per_step
        self.lam_updates_num = int(self.updates_per_training / self.multiplier_update_interval)
        self.actor_updates_num = int(self.updates_per_training / self.actor_update_interval)
        self.model_ensemble = model_ensemble

        self.actor = SquashedGaussianPolicy(mlp([state_dim, *([self.hidden_dim] * self.hidden_layers), action_dim * 2], device=device))
        self.actor_safe = copy.deepcopy(self.actor)
        self.critic = CriticEnsemble(self.critic_cfg, state_dim, action_dim, device=device)
        self.critic_target = copy.deepcopy(self.critic)
        self.constraint_critic = ConstraintCritic(self.constraint_critic_cfg, state_dim, action_dim, output_dim=con_dim, device=device)
        self.constraint_critic_target = copy.deepcopy(self.constraint_critic)
        for p in [*self.critic_target.parameters(), *self.constraint_critic_target.parameters()]:
            p.requires_grad = False
        self.multiplier = MLPMultiplier(self.mlp_multiplier_cfg, state_dim, device=device)

        # ---- flat arenas (params / targets / grads / Adam m,v) --------------------------------------------------
        self._build_arenas()
        self.critic_optimizer = FusedAdam(self._critic_arena, self.critic_lr, weight_decay=1e-4)
        self.critic_lr_scheduler = CosineLR(self.critic_optimizer, self.updates_per_training, self.critic_lr_end)
        self.multiplier_optimizer = FusedAdam(self._mult_arena, self.multiplier_lr, weight_decay=1e-4)
        self.multiplier_lr_scheduler = CosineLR(self.multiplier_optimizer, self.lam_updates_num, self.multiplier_lr_end)

        # actors / alpha (SURVEY §8f "next" row 1): in-kernel Adam over flat arenas as well (src/ssac.py:210-228)
        # the two actor gradient arenas share ONE storage [actor | safe | losses]: one all-reduce message per data-parallel actor step
        na, ns = self._actor_arena.numel(), self._safe_arena.numel()
        self._actor_grad_joint = torch.zeros(na + ns + _lib.LOSSES_LEN, device=device)
        self.actor_optimizer = FusedAdam(self._actor_arena, self.actor_lr, weight_decay=1e-4,
                                         grad_storage=self._actor_grad_joint[:na + _lib.LOSSES_LEN])
        self.actor_optimizer.grad, self.actor_optimizer.losses = self._actor_grad_joint[:na], self._actor_grad_joint[na + ns:]
        self.actor_lr_scheduler = CosineLR(self.actor_optimizer, max(self.actor_updates_num, 1), self.actor_lr_end)
        self.actor_safe_optimizer = FusedAdam(self._safe_arena, self.actor_lr, weight_decay=1e-4,
                                              grad_storage=self._actor_grad_joint[na:])
        self.actor_safe_optimizer.grad, self.actor_safe_optimizer.losses = self._actor_grad_joint[na:na + ns], self._actor_grad_joint[na + ns:]
        self.actor_safe_lr_scheduler = CosineLR(self.actor_safe_optimizer, max(self.actor_updates_num, 1), self.actor_lr_end)
        self.log_alpha = torch.tensor(math.log(self.init_alpha), device=device, requires_grad=True)
        if self.autotune_alpha:
            self.alpha_optimizer = FusedAdam(torch.zeros(1, device=device), self.actor_lr, weight_decay=0.0)
        if self.target_entropy is None:
            self.target_entropy = -action_dim

        self.register_buffer('total_updates', torch.zeros([], device=device))
        self._ws = _lib.Workspace()
        self.noise_seed = 0xD2B0
        self.precision = _lib.PREC_FP32
        self._ws_need = {}
        self.data_parallel = False          # set True under torchrun: gradients are all-reduced between backward and Adam
        self._dp_synced = False             # replicas are made identical (rank 0's state) at the first data-parallel step
        self._dp_rng = random.Random(0xD2B0)  # critic pick of the actor step under data parallelism: identical on every rank
        self._structs = None

    # the steps' losses live in the tail of the gradient storage (FusedAdam): [0..] losses / norms, [15] watchdog flag
    @property
    def _losses(self):
        return self.critic_optimizer.losses

    @property
    def _mult_losses(self):
        return self.multiplier_optimizer.losses

    @property
    def _actor_losses(self):
        return self.actor_optimizer.losses

    def sync_replicas(self, src=0):
        """Make every data-parallel replica identical to rank ``src``: the five parameter arenas, both target nets, Adam m / v and
        step counts, log_alpha and its Adam state, learning rates, the noise seeds and the host RNG that picks the actor step's
        critic (the reference draws it with ``random.choice``, src/ssac.py:41-43).  Called automatically at the first
        data-parallel update; call it again after loading a checkpoint on one rank only."""
        dist = self._dist()
        if dist is None:
            return
        self._ensure_arenas()
        opts = [self.critic_optimizer, self.multiplier_optimizer, self.actor_optimizer, self.actor_safe_optimizer]
        if self.autotune_alpha:
            opts.append(self.alpha_optimizer)
        tensors = [self._critic_arena, self._target_arena, self._mult_arena, self._actor_arena, self._safe_arena, self.log_alpha.data]
        for o in opts:
            tensors += [o.m, o.v]
        for t in tensors:
            dist.broadcast(t, src=src)
        scheds = [self.critic_lr_scheduler, self.multiplier_lr_scheduler, self.actor_lr_scheduler, self.actor_safe_lr_scheduler]
        host = [dict(steps=[o.step_count for o in opts], lrs=[o.param_groups[0]["lr"] for o in opts],
                     epochs=[sc.last_epoch for sc in scheds], seed=self.noise_seed, rng=self._dp_rng.getstate(),
                     qc_seed=self.constraint_critic.noise_seed, actor_seed=self.actor.noise_seed)]
        dist.broadcast_object_list(host, src=src)
        h = host[0]
        for o, st, lr in zip(opts, h["steps"], h["lrs"]):
            o.step_count, o.param_groups[0]["lr"] = st, lr
        for sc, e in zip(scheds, h["epochs"]):
            sc.last_epoch = e
        self.noise_seed, self.constraint_critic.noise_seed, self.actor.noise_seed = h["seed"], h["qc_seed"], h["actor_seed"]
        self._dp_rng.setstate(h["rng"])
        self._dp_synced = True

    # ------------------------------------------------------------------------------------------------------
    def _critic_params(self):
        return [*self.critic.parameters(), *self.constraint_critic.parameters()]

    def _target_params(self):
        return [*self.critic_target.parameters(), *self.constraint_critic_target.parameters()]

    def _build_arenas(self):
        self._critic_arena = _flatten_into_arena(self._critic_params())
        self._target_arena = _flatten_into_arena(self._target_params())
        self._mult_arena = _flatten_into_arena(list(self.multiplier.parameters()))
        self._actor_arena = _flatten_into_arena(list(self.actor.parameters()))
        self._safe_arena = _flatten_into_arena(list(self.actor_safe.parameters()))
        # the two clip groups are contiguous (padded) ranges of the critic arena: [Q1 | Q2] then [Qc]
        offs, total = _arena_offsets(self._critic_params())
        self._n_q = offs[len(list(self.critic.parameters()))]
        self._n_qc = total - self._n_q
        self._structs = None

    def _apply(self, fn, *args, **kwargs):
        """``.to()`` / ``.cuda()`` / ``.float()`` replace parameter storage: re-validate the arenas and rebuild the structs on next use."""
        out = super()._apply(fn, *args, **kwargs)
        self._arenas_checked, self._structs = False, None
        return out

    def _arena_sentinels_ok(self):
        """First and last parameter of every arena still sit where the arena put them (a few data_ptr() calls instead of the ~60 of the
        full check, which cost ~0.1 ms per update - as much as the GPU work of a small-batch step)."""
        for arena, first, last in self._sentinels:
            if first.data_ptr() != arena.data_ptr() or last.data_ptr() + 4 * last.numel() > arena.data_ptr() + 4 * arena.numel() \
                    or last.data_ptr() < arena.data_ptr():
                return False
        return True

    def _ensure_arenas(self):
        if getattr(self, "_arenas_checked", False) and self._structs is not None and self._arena_sentinels_ok():
            return self._structs
        if not (_arena_ok(self._critic_arena, self._critic_params()) and _arena_ok(self._target_arena, self._target_params())
                and _arena_ok(self._mult_arena, list(self.multiplier.parameters()))
                and _arena_ok(self._actor_arena, list(self.actor.parameters()))
                and _arena_ok(self._safe_arena, list(self.actor_safe.parameters()))):
            old = (self.critic_optimizer, self.multiplier_optimizer, self.actor_optimizer, self.actor_safe_optimizer)
            self._build_arenas()
            if self._actor_grad_joint.device != self._actor_arena.device:
                na, ns = self._actor_arena.numel(), self._safe_arena.numel()
                self._actor_grad_joint = self._actor_grad_joint.to(self._actor_arena.device)
                for o in (self.actor_optimizer, self.actor_safe_optimizer):
                    o.m, o.v = o.m.to(self._actor_arena.device), o.v.to(self._actor_arena.device)
                    o.losses = self._actor_grad_joint[na + ns:]
                self.actor_optimizer.grad, self.actor_safe_optimizer.grad = self._actor_grad_joint[:na], self._actor_grad_joint[na:na + ns]
            for opt, arena in zip(old[:2], (self._critic_arena, self._mult_arena)):
                opt.to(arena.device)
        groups = ((self._critic_arena, self._critic_params()), (self._target_arena, self._target_params()),
                  (self._mult_arena, list(self.multiplier.parameters())), (self._actor_arena, list(self.actor.parameters())),
                  (self._safe_arena, list(self.actor_safe.parameters())))
        self._sentinels = [(arena, ps[0], ps[-1]) for arena, ps in groups]
        self._arenas_checked = True
        if self._structs is None:
            self._structs = dict(
                actor=self.actor.as_struct(), actor_safe=self.actor_safe.as_struct(),
                q=[_mlp3_struct(q) for q in self.critic.qs], qt=[_mlp3_struct(q) for q in self.critic_target.qs],
                qc=self.constraint_critic.as_struct(), qct=self.constraint_critic_target.as_struct(),
                lam=_mlp3_struct(self.multiplier.lam))
        return self._structs

    def critic_arena_views(self, arena):
        """name -> view of ``arena`` (the critic parameter / gradient / Adam arena), in state_dict order of critic.* then
        constraint_critic.* (the arena aligns every parameter to ARENA_ALIGN floats)."""
        named = [(f"critic.{k}", p) for k, p in self.critic.named_parameters()] + \
                [(f"constraint_critic.{k}", p) for k, p in self.constraint_critic.named_parameters()]
        offs, _ = _arena_offsets([p for _, p in named])
        return {k: arena[o:o + p.numel()].view(p.shape) for (k, p), o in zip(named, offs)}

    def invalidate_structs(self):
        """Call after replacing parameter storage by hand (``load_state_dict`` copies in place and needs nothing)."""
        self._structs, self._arenas_checked = None, False

    # ---- reference API -------------------------------------------------------------------------------------
    def act(self, states, eval):
        return self.actor.act(states, eval)

    @property
    def alpha(self):
        return self.log_alpha.exp()

    @property
    def violation_value(self):
        return -self.violation_cost / (1. - self.discount)

    def update_r_bounds(self, r_min, r_max):
        self.r_min, self.r_max = r_min, r_max
        if self.update_violation_cost:
            self.violation_cost = (r_max - r_min) / self.discount ** self.horizon - r_max

    def _get_qc(self, qc_con_dim):
        if self.con_dim > 1:
            assert qc_con_dim.size(-1) == self.con_dim
            return torch.max(qc_con_dim, dim=-1)[0]
        return qc_con_dim

    def shield_act(self, states, eval=True, shield_type="linear", safe_shield_threshold=-0.1, uncertainty=False, eps=None,
                   return_info=False, path=0):
        """The safety shield in one C-ABI call (drpo_shield_act): the action-selection block of sample_episodes_batched
        (src/sampling.py:420-439: ``eval=True``, Qc = mean head, shield_type "safe" / "linear" / anything else = none) and of
        SMBPO.step_generator (src/smbpo.py:124-136: ``eval=False`` samples the performance action, ``uncertainty`` =
        distributional_qc, shield_type "safe").  The safe actor always acts in eval mode, as at both call sites.
        ``return_info`` adds (_get_qc of the performance action, choice) - see include/drpo_b200.h.  ``path``: 0 = auto (two-launch
        latency kernels for small batches, batched GEMM path for large ones), 1 / 2 force one of them."""
        lib = _lib.load()
        states = states.contiguous().float()
        n, A = states.shape[0], self.action_dim
        actions = torch.empty((n, A), device=states.device)
        qc_perf = torch.empty((n,), device=states.device) if return_info else None
        choice = torch.empty((n,), dtype=torch.int32, device=states.device) if return_info else None
        a = _lib.ShieldArgs()
        # the network structs are cached with the update steps' (parameters are views of flat arenas): on a 1-row call building
        # them costs more than the two launches
        st = self._ensure_arenas()
        actor, safe, qc = st["actor"], st["actor_safe"], st["qc"]
        a.actor, a.actor_safe, a.qc = C_pointer(actor), C_pointer(safe), C_pointer(qc)
        a.states, a.n, a.state_dim, a.action_dim, a.con_dim = _lib.ptr(states), n, self.state_dim, A, self.con_dim
        a.shield_type = _lib.SHIELD_TYPES.get(shield_type, _lib.SHIELD_NONE)
        a.eval_perf, a.uncertainty = int(bool(eval)), int(bool(uncertainty))
        a.std_ratio, a.threshold = float(self.constraint_critic.std_ratio), float(safe_shield_threshold)
        self.actor._noise_step += 1
        eps = eps.contiguous().float() if eps is not None else None        # a named local: the copy must outlive the launch
        noise = _lib.Noise(_lib.ptr(eps), A, self.actor.noise_seed, 17, self.actor._noise_step)
        a.noise_perf = C_pointer(noise)
        a.actions, a.qc_perf, a.choice, a.path = _lib.ptr(actions), _lib.ptr(qc_perf), _lib.ptr(choice), int(path)
        ws = self._ws.get(lib.drpo_shield_workspace_bytes(a.actor, n, self.state_dim, A, self.con_dim, self.constraint_critic.hidden_dim),
                          states.device)
        a.workspace, a.workspace_bytes, a.stream = _lib.ptr(ws), ws.numel(), _lib.stream_ptr()
        _lib.check(lib.drpo_shield_act(a), "drpo_shield_act")
        return (actions, qc_perf, choice) if return_info else actions

    def _dist(self):
        import torch.distributed as dist
        if self.data_parallel and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
        return None

    def update_critic(self, obs, action, next_obs, reward, done, violation, constraint_value, noise=None, phases=None):
        """SSAC.update_critic (src/ssac.py:437-456).  ``noise`` = (eps_actor [B,A], eps_safe [B,A], eps_qc [B(,C)])
        injects the three Gaussian draws (parity); otherwise the in-kernel Philox stream is used.  ``phases`` (tests) runs only
        the given phases of drpo_critic_step (1 = forward/backward into the gradient arena, 2 = clip/Adam/EMA)."""
        lib = _lib.load()
        _lib.peek_kernel_status("SSAC.update_critic")
        st = self._ensure_arenas()
        dist = self._dist()
        if dist is not None and not self._dp_synced:
            self.sync_replicas()
        world = dist.get_world_size() if dist else 1
        rank = dist.get_rank() if dist else 0
        B = obs.shape[0]
        f = lambda t: t.contiguous().float()
        obs, action, next_obs, reward, cv = f(obs), f(action), f(next_obs), f(reward), f(constraint_value)
        done, violation = done.contiguous(), violation.contiguous()
        assert done.dtype == torch.bool and cv.numel() == B * self.con_dim
        opt = self.critic_optimizer
        opt.step_count += 1
        a = _lib.CriticArgs()
        a.batch = _lib.Batch(_lib.ptr(obs), _lib.ptr(action), _lib.ptr(next_obs), _lib.ptr(reward), _lib.ptr(done),
                             _lib.ptr(violation), _lib.ptr(cv))
        a.batch_size, a.global_batch_size = B, getattr(self, "_global_batch_override", None) or B * world
        a.state_dim, a.action_dim, a.con_dim = self.state_dim, self.action_dim, self.con_dim
        a.actor, a.actor_safe = C_pointer(st["actor"]), C_pointer(st["actor_safe"])
        a.q[0], a.q[1], a.q_target[0], a.q_target[1] = st["q"][0], st["q"][1], st["qt"][0], st["qt"][1]
        a.qc, a.qc_target = st["qc"], st["qct"]
        a.params, a.grads = _lib.ptr(self._critic_arena), _lib.ptr(opt.grad)
        a.adam_m, a.adam_v, a.target_params = _lib.ptr(opt.m), _lib.ptr(opt.v), _lib.ptr(self._target_arena)
        a.n_params_q, a.n_params_qc = self._n_q, self._n_qc
        a.log_alpha = _lib.ptr(self.log_alpha.data)
        if noise is not None:
            n0, n1, n2 = [f(x) for x in noise]
            a.eps_actor, a.eps_safe, a.eps_qc = _lib.ptr(n0), _lib.ptr(n1), _lib.ptr(n2)
        a.seed, a.noise_step, a.row_id_offset = self.noise_seed, opt.step_count, rank * B
        a.discount, a.tau, a.grad_norm, a.qc_td_bound = self.discount, self.tau, self.grad_norm, self.qc_td_bound
        a.adam = opt.as_struct()
        a.losses, a.precision = _lib.ptr(self._losses), self.precision
        need = self._ws_need.get(B)
        if need is None:                                    # (host time matters at the strong-scaled shard sizes: one C query per batch size)
            need = self._ws_need[B] = lib.drpo_critic_workspace_bytes(B, self.state_dim, self.action_dim, self.con_dim, self.hidden_dim)
        ws = self._ws.get(need, obs.device)
        a.workspace, a.workspace_bytes, a.stream = _lib.ptr(ws), ws.numel(), _lib.stream_ptr()
        if phases is not None:
            a.phases = phases
            _lib.check(lib.drpo_critic_step(a), "drpo_critic_step")
            return self._losses[0].clone(), self._losses[1].clone()
        if dist is None:
            a.phases = 3
            _lib.check(lib.drpo_critic_step(a), "drpo_critic_step")
        else:
            a.phases = 1
            _lib.check(lib.drpo_critic_step(a), "drpo_critic_step(forward/backward)")
            dist.all_reduce(opt.grad_full)                  # ONE NCCL sum over NVLink: gradients + losses + watchdog flag (the tail)
            a.phases = 2
            _lib.check(lib.drpo_critic_step(a), "drpo_critic_step(optimizer)")
        self.critic_lr_scheduler.step()
        return self._losses[:2].clone().unbind(0)           # (two 0-dim loss tensors, one copy kernel)

    def update_multiplier(self, obs, eps=None):
        """SSAC.update_multiplier (src/ssac.py:570-578)."""
        lib = _lib.load()
        _lib.peek_kernel_status("SSAC.update_multiplier")
        st = self._ensure_arenas()
        dist = self._dist()
        if dist is not None and not self._dp_synced:
            self.sync_replicas()
        world = dist.get_world_size() if dist else 1
        rank = dist.get_rank() if dist else 0
        obs = obs.contiguous().float()
        B = obs.shape[0]
        opt = self.multiplier_optimizer
        opt.step_count += 1
        a = _lib.MultiplierArgs()
        a.obs, a.batch_size, a.global_batch_size = _lib.ptr(obs), B, B * world
        a.state_dim, a.action_dim, a.con_dim = self.state_dim, self.action_dim, self.con_dim
        a.actor, a.actor_safe, a.qc, a.lam = C_pointer(st["actor"]), C_pointer(st["actor_safe"]), C_pointer(st["qc"]), st["lam"]
        a.params, a.grads, a.adam_m, a.adam_v = _lib.ptr(self._mult_arena), _lib.ptr(opt.grad), _lib.ptr(opt.m), _lib.ptr(opt.v)
        a.n_params = self._mult_arena.numel()
        if eps is not None:
            eps = eps.contiguous().float()
            a.eps_actor = _lib.ptr(eps)
        a.seed, a.noise_step, a.row_id_offset = self.noise_seed + 1, opt.step_count, rank * B
        a.std_ratio = self.constraint_critic.std_ratio
        a.constraint_threshold, a.penalty_lb, a.penalty_ub = self.constraint_threshold, self.penalty_lb, self.penalty_ub
        a.upper_bound, a.lam_epsilon, a.grad_norm = self.multiplier.upper_bound, self.lam_epsilon, self.grad_norm
        a.adam = opt.as_struct()
        a.losses, a.precision = _lib.ptr(self._mult_losses), self.precision
        ws = self._ws.get(lib.drpo_multiplier_workspace_bytes(B, self.state_dim, self.action_dim, self.con_dim, self.hidden_dim), obs.device)
        a.workspace, a.workspace_bytes, a.stream = _lib.ptr(ws), ws.numel(), _lib.stream_ptr()
        if dist is None:
            a.phases = 3
            _lib.check(lib.drpo_multiplier_step(a), "drpo_multiplier_step")
        else:
            a.phases = 1
            _lib.check(lib.drpo_multiplier_step(a), "drpo_multiplier_step(forward/backward)")
            dist.all_reduce(opt.grad_full)                  # gradients + loss + watchdog flag in one message
            a.phases = 2
            _lib.check(lib.drpo_multiplier_step(a), "drpo_multiplier_step(optimizer)")
        self.multiplier_lr_scheduler.step()
        return self._mult_losses[0].clone()

    # ---- SURVEY §8f "next" row 1: actor / alpha update ------------------------------------------------------------
    def actor_loss(self, obs, include_alpha=True):
        """src/ssac.py:458-505 as a differentiable expression (the reference exposes it and callers may backpropagate through the
        returned losses).  API mirror only: ``update_actor_and_alpha`` below does NOT use it - the update runs in drpo_actor_step."""
        def rsample(pol):
            mu, std = pol.mu_std(obs)
            x = mu + std * torch.randn_like(mu)
            lp = (-((x - mu) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))
                  - 2. * (math.log(2.) - x - F.softplus(-2. * x))).sum(-1)
            return torch.tanh(x), lp
        action, log_prob = rsample(self.actor)
        actor_Q = self.critic.random_choice(obs, action)
        alpha = self.alpha
        uncstr = torch.mean(alpha.detach() * log_prob - actor_Q)
        actor_Qc = self._get_qc(self.constraint_critic.forward_torch(obs, action, uncertainty=True))
        with torch.no_grad():
            action_safe = self.actor_safe.act(obs, eval=True)
            safe_Qc = self._get_qc(self.constraint_critic(obs, action_safe, uncertainty=True))
            lams = self.multiplier(obs, safe_Qc)
        cstr = torch.mean(torch.mul(lams, actor_Qc))
        a_s, _ = rsample(self.actor_safe)
        actor_safe_loss = torch.mean(self._get_qc(self.constraint_critic.forward_torch(obs, a_s, uncertainty=True)))
        losses = [uncstr + cstr]
        if include_alpha:
            coef = self.log_alpha if self.use_log_alpha_loss else alpha
            losses.append(-coef * torch.mean(log_prob.detach() + self.target_entropy))
        losses.append(actor_safe_loss)
        return losses

    def update_actor_and_alpha(self, obs, noise=None, q_index=None, phases=None):
        """SSAC.update_actor_and_alpha (src/ssac.py:507-527) through drpo_actor_step: performance actor, temperature and safe
        actor in one call.  ``noise`` = (eps_actor [B,A], eps_safe [B,A]) injects the two rsample draws (parity), otherwise the
        in-kernel Philox stream is used; ``q_index`` is the critic `random.choice` picks (src/ssac.py:41-43)."""
        assert self.autotune_alpha, "drpo_actor_step implements the autotune_alpha configuration of the reference"
        lib = _lib.load()
        _lib.peek_kernel_status("SSAC.update_actor_and_alpha")
        st = self._ensure_arenas()
        dist = self._dist()
        if dist is not None and not self._dp_synced:
            self.sync_replicas()
        world = dist.get_world_size() if dist else 1
        rank = dist.get_rank() if dist else 0
        obs = obs.contiguous().float()
        B = obs.shape[0]
        if q_index is None:
            # same draw as random.choice(self.qs) (src/ssac.py:41-43); under data parallelism from the replica-synchronised RNG, so
            # that every rank back-propagates through the SAME critic before the gradients are summed
            q_index = (self._dp_rng if dist is not None else random).choice(range(len(self.critic.qs)))
        oa, os_, oal = self.actor_optimizer, self.actor_safe_optimizer, self.alpha_optimizer
        for o in (oa, os_, oal):
            o.step_count += 1
        a = _lib.ActorArgs()
        a.obs, a.batch_size = _lib.ptr(obs), B
        a.global_batch_size = getattr(self, "_global_batch_override", None) or B * world
        a.state_dim, a.action_dim, a.con_dim = self.state_dim, self.action_dim, self.con_dim
        a.actor, a.actor_safe = st["actor"], st["actor_safe"]
        a.q, a.qc, a.lam = C_pointer(st["q"][q_index]), C_pointer(st["qc"]), C_pointer(st["lam"])
        a.params_actor, a.grads_actor, a.m_actor, a.v_actor = _lib.ptr(self._actor_arena), _lib.ptr(oa.grad), _lib.ptr(oa.m), _lib.ptr(oa.v)
        a.params_safe, a.grads_safe, a.m_safe, a.v_safe = _lib.ptr(self._safe_arena), _lib.ptr(os_.grad), _lib.ptr(os_.m), _lib.ptr(os_.v)
        a.n_actor, a.n_safe = self._actor_arena.numel(), self._safe_arena.numel()
        a.log_alpha, a.alpha_m, a.alpha_v = _lib.ptr(self.log_alpha.data), _lib.ptr(oal.m), _lib.ptr(oal.v)
        if noise is not None:
            n0, n1 = [x.contiguous().float() for x in noise]
            a.eps_actor, a.eps_safe = _lib.ptr(n0), _lib.ptr(n1)
        a.seed, a.noise_step, a.row_id_offset = self.noise_seed, oa.step_count, rank * B
        a.std_ratio, a.multiplier_ub = float(self.constraint_critic.std_ratio), float(self.multiplier.upper_bound)
        a.grad_norm, a.target_entropy = self.grad_norm, float(self.target_entropy)
        a.adam_actor, a.adam_alpha, a.adam_safe = oa.as_struct(), oal.as_struct(), os_.as_struct()
        a.losses, a.precision = _lib.ptr(self._actor_losses), self.precision
        ws = self._ws.get(lib.drpo_actor_workspace_bytes(B, self.state_dim, self.action_dim, self.con_dim, self.hidden_dim), obs.device)
        a.workspace, a.workspace_bytes, a.stream = _lib.ptr(ws), ws.numel(), _lib.stream_ptr()
        if phases is not None:
            a.phases = phases
            _lib.check(lib.drpo_actor_step(a), "drpo_actor_step")
            return self._actor_losses[:3].clone()
        if dist is None:
            a.phases = 3
            _lib.check(lib.drpo_actor_step(a), "drpo_actor_step")
        else:
            a.phases = 1
            _lib.check(lib.drpo_actor_step(a), "drpo_actor_step(forward/backward)")
            dist.all_reduce(self._actor_grad_joint)         # [actor grads | safe grads | losses + watchdog flag]: one message
            a.phases = 2
            _lib.check(lib.drpo_actor_step(a), "drpo_actor_step(optimizer)")
        self.actor_lr_scheduler.step()
        self.actor_safe_lr_scheduler.step()
        return self._actor_losses[:3].clone()


def C_pointer(struct):
    import ctypes
    return ctypes.pointer(struct)
