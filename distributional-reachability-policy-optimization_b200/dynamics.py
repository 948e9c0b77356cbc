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
        self.optimizer = optimizer_factory(
            [*self.trunk.parameters(), *self.diff_head.parameters(), *self.log_var_head.parameters(),
             self.min_log_var, self.max_log_var], lr=self.learning_rate, weight_decay=1e-4)
        self.elite_inds = torch.randint(high=E, size=(self.num_elites,)).tolist()      # src/dynamics.py:105-106
        self._ws = _lib.Workspace()
        self.noise_seed = 0x5EEDD12A
        self._noise_step = 0

    # ------------------------------------------------------------------------------------------------------
    @property
    def total_batch_size(self):
        return self.ensemble_size * self.batch_size

    def as_struct(self) -> "_lib.Ensemble":
        p = _lib.ptr
        return _lib.Ensemble(
            self.state_dim, self.action_dim, self.ensemble_size, self.hidden_dim,
            p(self.state_normalizer.mean), p(self.state_normalizer.std), p(self.min_log_var.data), p(self.max_log_var.data),
            p(self.trunk[0].weight.data), p(self.trunk[0].bias.data), p(self.trunk[2].weight.data), p(self.trunk[2].bias.data),
            p(self.diff_head[0].weight.data), p(self.diff_head[0].bias.data), p(self.diff_head[2].weight.data),
            p(self.diff_head[2].bias.data), p(self.log_var_head[0].weight.data), p(self.log_var_head[0].bias.data),
            p(self.log_var_head[2].weight.data), p(self.log_var_head[2].bias.data), None)

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
                                             _lib.ptr(means), _lib.ptr(log_vars), _lib.PREC_FP32, _lib.ptr(ws), ws.numel(),
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
        noise = _lib.Noise(_lib.ptr(eps.contiguous()) if eps is not None else None, self.state_dim + 1, self.noise_seed,
                           16, self._noise_step)
        ens = self.as_struct()
        ws = self._workspace(lib, ens, B, states.device)
        _lib.check(lib.drpo_ensemble_sample(ens, index, _lib.ptr(states), _lib.ptr(actions), B, noise, _lib.ptr(next_states),
                                            _lib.ptr(rewards), _lib.PREC_FP32, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
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

    # ---- training of the ensemble: SURVEY.md §8f "next" row 2 — host-side torch for now ------------------------
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

    def fit(self, buffer, steps=None, epochs=None, progress_bar=False, **kwargs):
        """src/dynamics.py:155-196 (steps form).  Not on the CUDA hot path yet: eager torch."""
        if steps is None:
            raise NotImplementedError("only fit(steps=...) is provided")
        n = len(buffer)
        states, actions, next_states, rewards = buffer.get()[:4]
        self.state_normalizer.fit(states)
        targets = torch.cat([next_states, rewards.unsqueeze(1)], dim=1)
        losses = []
        for _ in range(steps):
            idx = torch.randint(n, [self.total_batch_size], device=states.device)
            loss = self.compute_loss(states[idx], actions[idx], targets[idx])
            losses.append(loss.item())
            self.optimizer.zero_grad()
            loss.backward()
            self.optimizer.step()
        hold = torch.randint(n, [self.holdout_size], device=states.device).repeat(self.ensemble_size, 1)
        mse = self._mse_loss(states[hold], actions[hold], targets[hold], enable_grad=False)
        self._elite_inds = torch.argsort(mse)[:self.num_elites].tolist()
        return losses
