"""Probabilistic ensemble dynamics model with the reference's API (src/dynamics.py:26-253), forward/sample path on
hand-written CUDA.  Parameter names and shapes match the reference ``state_dict`` so checkpoints interchange."""
import random

import torch
import torch.nn as nn

from . import _lib
from .config import BaseConfig, Configurable


class BatchedLinear(nn.Module):
    """E stacked linear layers (src/dynamics.py:26-52): weight [E,out,in], bias [E,out]."""

    def __init__(self, ensemble_size, in_features, out_features, device=None):
        super().__init__()
        self.ensemble_size, self.in_features, self.out_features = ensemble_size, in_features, out_features
        self.weight = nn.Parameter(torch.empty(ensemble_size, out_features, in_features, device=device))
        self.bias = nn.Parameter(torch.empty(ensemble_size, out_features, device=device))
        # the reference's mlp() re-initialises every layer with xavier_normal_/zeros_ (src/torch_util.py:155-165,208)
        nn.init.xavier_normal_(self.weight)
        nn.init.zeros_(self.bias)

    def forward(self, input):
        assert input.dim() == 3 and input.shape[0] == self.ensemble_size
        return torch.bmm(input, self.weight.transpose(1, 2)) + self.bias.unsqueeze(1)


class Normalizer(nn.Module):
    """src/normalization.py:6-27."""

    def __init__(self, dim, epsilon=1e-6, device=None):
        super().__init__()
        self.dim, self.epsilon = dim, epsilon
        self.register_buffer("mean", torch.zeros(dim, device=device))
        self.register_buffer("std", torch.zeros(dim, device=device))

    def fit(self, X):
        assert torch.is_tensor(X) and X.dim() == 2 and X.shape[1] == self.dim
        self.mean.data.copy_(X.mean(dim=0))
        std = X.std(dim=0)
        std[std < 1e-6] = 1.0
        self.std.data.copy_(std)

    def forward(self, x):
        return (x - self.mean) / (self.std + self.epsilon)

    def unnormalize(self, normal_X):
        return self.mean + (self.std * normal_X)


def _batched_mlp(E, dims, output_activation, device):
    layers = []
    for i in range(len(dims) - 2):
        layers += [BatchedLinear(E, dims[i], dims[i + 1], device=device), nn.SiLU()]
    layers.append(BatchedLinear(E, dims[-2], dims[-1], device=device))
    if output_activation:
        layers.append(nn.SiLU())
    return nn.Sequential(*layers)


class BatchedGaussianEnsemble(Configurable, nn.Module):
    class Config(BaseConfig):
        ensemble_size = 7
        num_elites = 5
        hidden_dim = 200
        trunk_layers = 2
        head_hidden_layers = 1
        activation = 'swish'
        init_min_log_var = -10.0
        init_max_log_var = 1.0
        log_var_bound_weight = 0.01
        batch_size = 256
        learning_rate = 1e-3
        holdout_size = 256

    def __init__(self, config, state_dim, action_dim, device=None, optimizer_factory=torch.optim.Adam):
        Configurable.__init__(self, config)
        nn.Module.__init__(self)
        if self.trunk_layers != 2 or self.head_hidden_layers != 1 or self.activation != 'swish':
            raise NotImplementedError("the CUDA ensemble implements the reference's default topology "
                                      "(2 trunk layers, 1 hidden head layer, swish)")
        device = torch.device(device if device is not None else "cuda")
        self.state_dim, self.action_dim = state_dim, action_dim
        D, O, H, E = state_dim + action_dim, state_dim + 1, self.hidden_dim, self.ensemble_size
        self.min_log_var = nn.Parameter(torch.full([O], self.init_min_log_var, device=device))
        self.max_log_var = nn.Parameter(torch.full([O], self.init_max_log_var, device=device))
        self.state_normalizer = Normalizer(state_dim, device=device)
        self.trunk = _batched_mlp(E, [D, H, H], True, device)
        self.diff_head = _batched_mlp(E, [H, H, O], False, device)
        self.log_var_head = _batched_mlp(E, [H, H, O], False, device)
        # trainable tensors in one flat arena (views keep the reference's state_dict names / shapes): the in-kernel Adam of
        # drpo_ensemble_train_step streams over it (src/dynamics.py:93-101: lr 1e-3, coupled L2 1e-4 on every tensor)
        from .ssac import FusedAdam, _flatten_into_arena
        self._arena = _flatten_into_arena(self._trainable())
        self.optimizer = FusedAdam(self._arena, self.learning_rate, weight_decay=1e-4)
        self._fit_losses = torch.zeros(1 + E, device=device)
        self.precision = _lib.PREC_FP32
        # precision of the standalone forward entry points (_forward1 / _forward_all / means / sample / elite_samples):
        # PREC_FP32, or PREC_BF16 = the fused tcgen05 member chain (csrc/ens_umma.cu; hidden <= 256, state_dim + action_dim <= 64)
        self.forward_precision = _lib.PREC_FP32
        self.elite_inds = torch.randint(high=E, size=(self.num_elites,)).tolist()      # src/dynamics.py:105-106
        self._ws = _lib.Workspace()
        self.noise_seed = 0x5EEDD12A
        self._noise_step = 0

    def _trainable(self):
        return [*self.trunk.parameters(), *self.diff_head.parameters(), *self.log_var_head.parameters(), self.min_log_var, self.max_log_var]

    def _apply(self, fn, *args, **kwargs):
        """``.to()`` / ``.cuda()`` replace parameter storage: re-validate the arena and rebuild the cached struct on next use."""
        out = super()._apply(fn, *args, **kwargs)
        self._arena_checked, self._struct = False, None
        return out

    def _ensure_arena(self):
        from .ssac import _arena_ok, _flatten_into_arena
        if getattr(self, "_arena_checked", False):
            # steady state: the first and last trainable tensor still sit where the arena put them (two data_ptr() calls instead of
            # the full walk, which costs a sizeable part of a 0.4 ms training iteration)
            a0, n = self._arena.data_ptr(), 4 * self._arena.numel()
            first, last = self._sentinels
            if first.data_ptr() == a0 and a0 <= last.data_ptr() and last.data_ptr() + 4 * last.numel() <= a0 + n:
                return
        self._struct = None
        if not _arena_ok(self._arena, self._trainable()):
            self._arena = _flatten_into_arena(self._trainable())
            opt = self.optimizer
            if opt.m.device != self._arena.device:
                opt.m, opt.v, opt.grad = opt.m.to(self._arena.device), opt.v.to(self._arena.device), opt.grad.to(self._arena.device)
            self._fit_losses = self._fit_losses.to(self._arena.device)
        tr = self._trainable()
        self._sentinels, self._arena_checked = (tr[0], tr[-1]), True

    def arena_views(self, arena):
        """name -> view of ``arena`` (parameter / gradient / Adam arena) in the order of ``_trainable``."""
        from .ssac import _arena_offsets
        named = [(f"trunk.{k}", p) for k, p in self.trunk.named_parameters()] + \
                [(f"diff_head.{k}", p) for k, p in self.diff_head.named_parameters()] + \
                [(f"log_var_head.{k}", p) for k, p in self.log_var_head.named_parameters()] + \
                [("min_log_var", self.min_log_var), ("max_log_var", self.max_log_var)]
        offs, _ = _arena_offsets([p for _, p in named])
        return {k: arena[o:o + p.numel()].view(p.shape) for (k, p), o in zip(named, offs)}

    # ------------------------------------------------------------------------------------------------------
    @property
    def total_batch_size(self):
        return self.ensemble_size * self.batch_size

    def as_struct(self) -> "_lib.Ensemble":
        if getattr(self, "_struct", None) is not None and getattr(self, "_arena_checked", False):
            return self._struct
        p = _lib.ptr
        self._struct = _lib.Ensemble(
            self.state_dim, self.action_dim, self.ensemble_size, self.hidden_dim,
            p(self.state_normalizer.mean), p(self.state_normalizer.std), p(self.min_log_var.data), p(self.max_log_var.data),
            p(self.trunk[0].weight.data), p(self.trunk[0].bias.data), p(self.trunk[2].weight.data), p(self.trunk[2].bias.data),
            p(self.diff_head[0].weight.data), p(self.diff_head[0].bias.data), p(self.diff_head[2].weight.data),
            p(self.diff_head[2].bias.data), p(self.log_var_head[0].weight.data), p(self.log_var_head[0].bias.data),
            p(self.log_var_head[2].weight.data), p(self.log_var_head[2].bias.data), None)
        return self._struct

    def _workspace(self, lib, ens, batch, device):
        return self._ws.get(lib.drpo_ensemble_workspace_bytes(ens, batch), device)

    def _forward(self, states, actions, member, per_member):
        lib = _lib.load()
        states, actions = states.contiguous().float(), actions.contiguous().float()
        B = states.shape[-2]
        E = self.ensemble_size if member < 0 else 1
        O = self.state_dim + 1
        shape = (E, B, O) if member < 0 else (B, O)
        means = torch.empty(shape, device=states.device)
        log_vars = torch.empty(shape, device=states.device)
        ens = self.as_struct()
        ws = self._workspace(lib, ens, B, states.device)
        _lib.check(lib.drpo_ensemble_forward(ens, member, int(per_member), _lib.ptr(states), _lib.ptr(actions), B,
                                             _lib.ptr(means), _lib.ptr(log_vars), self.forward_precision, _lib.ptr(ws), ws.numel(),
                                             _lib.stream_ptr()), "drpo_ensemble_forward")
        return means, log_vars

    def _forward1(self, states, actions, index):
        """src/dynamics.py:112-122."""
        return self._forward(states, actions, int(index), False)

    def _forward_all(self, states, actions):
        """src/dynamics.py:124-134: states/actions [E,B,*]."""
        assert states.dim() == 3 and states.shape[0] == self.ensemble_size
        return self._forward(states, actions, -1, True)

    def sample(self, states, actions, eps=None):
        """src/dynamics.py:198-203.  The elite pick stays a host-side ``random.choice`` exactly as in the reference;
        ``eps`` (optional, [B,S+1]) injects the Gaussian draw, otherwise the in-kernel Philox stream is used."""
        lib = _lib.load()
        index = random.choice(self._elite_inds)
        states, actions = states.contiguous().float(), actions.contiguous().float()
        B = states.shape[0]
        next_states = torch.empty((B, self.state_dim), device=states.device)
        rewards = torch.empty((B,), device=states.device)
        self._noise_step += 1
        eps = eps.contiguous().float() if eps is not None else None        # a named local: the copy must outlive the launch
        noise = _lib.Noise(_lib.ptr(eps), self.state_dim + 1, self.noise_seed, 16, self._noise_step)
        ens = self.as_struct()
        ws = self._workspace(lib, ens, B, states.device)
        _lib.check(lib.drpo_ensemble_sample(ens, index, _lib.ptr(states), _lib.ptr(actions), B, noise, _lib.ptr(next_states),
                                            _lib.ptr(rewards), self.forward_precision, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                   "drpo_ensemble_sample")
        return next_states, rewards

    def means(self, states, actions):
        """src/dynamics.py:206-210 — without materialising ``states.repeat(E,1,1)``."""
        means, _ = self._forward(states, actions, -1, False)
        return means[:, :, :-1], means[:, :, -1]

    def mean(self, states, actions):
        next_state_means, reward_means = self.means(states, actions)
        return next_state_means.mean(dim=0), reward_means.mean(dim=0)

    def elite_samples(self, states, actions, eps=None):
        """src/dynamics.py:218-234."""
        means, log_vars = self._forward(states, actions, -1, False)
        means, log_vars = means[self._elite_inds, ...], log_vars[self._elite_inds, ...]
        stds = torch.exp(log_vars).sqrt()
        samples = means + stds * (torch.randn_like(means) if eps is None else eps)
        return samples[:, :, :-1], samples[:, :, -1]

    # ---- training of the ensemble: SURVEY.md §8f "next" row 2 -------------------------------------------------------------
    # _mse_loss / compute_loss are public in the reference and return differentiable losses: the torch expressions below mirror
    # them for callers that bring their own optimiser.  fit() / train_step() do NOT use them: they run drpo_ensemble_train_step.
    def _forward_all_torch(self, states, actions):
        x = torch.cat([self.state_normalizer(states), actions], dim=-1)
        h = self.trunk(x)
        diffs = self.diff_head(h)
        means = diffs + torch.cat([states, torch.zeros_like(states[..., :1])], dim=-1)
        lv = self.log_var_head(h)
        lv = self.max_log_var - nn.functional.softplus(self.max_log_var - lv)
        lv = self.min_log_var + nn.functional.softplus(lv - self.min_log_var)
        return means, lv

    def _mse_loss(self, states, actions, targets, enable_grad=True):
        with torch.set_grad_enabled(enable_grad):
            means, log_vars = self._forward_all_torch(states, actions)
            inv_vars = torch.exp(-log_vars)
            return torch.mean((targets - means) ** 2 * inv_vars, dim=(-2, -1)) + torch.mean(log_vars, dim=(-2, -1))

    def compute_loss(self, states, actions, targets):
        n = len(targets) - len(targets) % self.ensemble_size
        E = self.ensemble_size
        s, a, t = [x[:n].reshape(E, n // E, *x.shape[1:]) for x in (states, actions, targets)]
        return torch.sum(self._mse_loss(s, a, t)) + self.log_var_bound_weight * (self.max_log_var.sum() - self.min_log_var.sum())

    def train_step(self, states, actions, targets, phases=3):
        """One iteration of fit's loop (src/dynamics.py:164-170) through drpo_ensemble_train_step: compute_loss on the batch
        (members take contiguous blocks, a remainder is dropped), backward, Adam.  Returns compute_loss as a device scalar."""
        lib = _lib.load()
        self._ensure_arena()
        f = lambda t: t.contiguous().float()
        states, actions, targets = f(states), f(actions), f(targets)
        n = targets.shape[0]
        opt = self.optimizer
        if phases & 2:
            opt.step_count += 1
        a = _lib.EnsembleTrainArgs()
        a.ens = self.as_struct()
        a.params, a.grads, a.adam_m, a.adam_v, a.n_params = _lib.ptr(self._arena), _lib.ptr(opt.grad), _lib.ptr(opt.m), _lib.ptr(opt.v), self._arena.numel()
        a.states, a.actions, a.targets, a.n_rows, a.shared_rows = _lib.ptr(states), _lib.ptr(actions), _lib.ptr(targets), n, 0
        a.log_var_bound_weight, a.adam, a.phases = float(self.log_var_bound_weight), opt.as_struct(), phases
        a.losses, a.precision = _lib.ptr(self._fit_losses), self.precision
        ws = self._ws.get(lib.drpo_ensemble_train_workspace_bytes(a.ens, max(n // self.ensemble_size, 1)), states.device)
        a.workspace, a.workspace_bytes, a.stream = _lib.ptr(ws), ws.numel(), _lib.stream_ptr()
        _lib.check(lib.drpo_ensemble_train_step(a), "drpo_ensemble_train_step")
        return self._fit_losses[0]

    def holdout_losses(self, states, actions, targets):
        """Every member's NLL on the SAME rows (the holdout scoring at the end of fit, src/dynamics.py:172-181)."""
        lib = _lib.load()
        f = lambda t: t.contiguous().float()
        states, actions, targets = f(states), f(actions), f(targets)
        a = _lib.EnsembleTrainArgs()
        a.ens = self.as_struct()
        a.states, a.actions, a.targets, a.n_rows, a.shared_rows = _lib.ptr(states), _lib.ptr(actions), _lib.ptr(targets), targets.shape[0], 1
        a.log_var_bound_weight, a.phases = float(self.log_var_bound_weight), 4
        a.losses, a.precision = _lib.ptr(self._fit_losses), self.precision
        ws = self._ws.get(lib.drpo_ensemble_train_workspace_bytes(a.ens, targets.shape[0]), states.device)
        a.workspace, a.workspace_bytes, a.stream = _lib.ptr(ws), ws.numel(), _lib.stream_ptr()
        _lib.check(lib.drpo_ensemble_train_step(a), "drpo_ensemble_train_step(holdout)")
        return self._fit_losses[1:1 + self.ensemble_size].clone()

    def fit(self, buffer, steps=None, epochs=None, progress_bar=False, max_grad_norm=None, post_epoch_callback=None,
            post_step_callback=None, verbose=False, **kwargs):
        """src/dynamics.py:155-196: normaliser fit, then either ``steps`` Adam iterations on random minibatches followed by the
        holdout ranking of the members -> ``_elite_inds`` (:161-185), or ``epochs`` x ensemble_size shuffled passes over the
        buffer (:186-194 -> epochal_training, src/train.py:58-101; returns the per-epoch mean loss).  The index draws stay
        torch.randint / torch.randperm (as in the reference); every iteration is one drpo_ensemble_train_step; the losses are
        read back once per call (steps form) or once per epoch (epochs form)."""
        n = len(buffer)
        states, actions, next_states, rewards = buffer.get()[:4]
        self.state_normalizer.fit(states)
        targets = torch.cat([next_states, rewards.unsqueeze(1)], dim=1)
        if steps is not None:
            assert epochs is None, 'Cannot pass both steps and epochs'
            losses = torch.empty(steps, device=states.device)
            for i in range(steps):
                idx = torch.randint(n, [self.total_batch_size], device=states.device)
                losses[i] = self.train_step(states[idx], actions[idx], targets[idx])
            hold = torch.randint(n, [self.holdout_size], device=states.device)
            mse = self.holdout_losses(states[hold], actions[hold], targets[hold])
            self._elite_inds = torch.argsort(mse)[:self.num_elites].tolist()
            return losses.tolist()
        if epochs is None:
            raise ValueError('Must pass steps or epochs')
        if max_grad_norm is not None:
            raise NotImplementedError("the in-kernel Adam of drpo_ensemble_train_step has no gradient clip "
                                      "(BatchedGaussianEnsemble.fit never passes max_grad_norm, src/dynamics.py:190-192)")
        bs = self.total_batch_size
        n_batches = -(-n // bs)
        losses = []
        for epoch in range(self.ensemble_size * epochs):
            perm = torch.randperm(n).to(states.device)
            ep = torch.empty(n_batches, device=states.device)
            for b in range(n_batches):
                idx = perm[b * bs:min((b + 1) * bs, n)]
                if idx.numel() < self.ensemble_size:
                    raise ValueError(f"the last batch of an epoch has {idx.numel()} rows: fewer than one per member")
                ep[b] = self.train_step(states[idx], actions[idx], targets[idx])
                if post_step_callback is not None:
                    post_step_callback(epoch, b, n_batches)
            losses.append(float(ep.double().mean().item()))
            if post_epoch_callback is not None:
                post_epoch_callback(epoch + 1)
        return losses
