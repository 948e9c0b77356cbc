"""Tanh-squashed Gaussian actor with the reference's API (src/policy.py:61-99, src/squashed_gaussian.py)."""
import torch
from torch import nn

from . import _lib


class BasePolicy:
    def act(self, states, eval):
        raise NotImplementedError

    def act1(self, state, eval=False):
        return self.act(torch.unsqueeze(state, 0), eval)[0]


class SquashedGaussianPolicy(BasePolicy, nn.Module):
    """``net`` is Linear-ReLU-Linear-ReLU-Linear producing [mu, raw_log_std] (src/ssac.py:184-186)."""

    def __init__(self, net, log_std_bounds=(-6, 4), std_multiplier=1.0):
        nn.Module.__init__(self)
        if tuple(log_std_bounds) != (-6, 4) or std_multiplier != 1.0:
            raise NotImplementedError("the CUDA policy head implements the reference defaults (-6,4), 1.0")
        self.net = net
        self.log_std_bounds, self.std_multiplier = log_std_bounds, std_multiplier
        self._ws = _lib.Workspace()
        self.noise_seed = 0xAC7012
        self._noise_step = 0

    def as_struct(self) -> "_lib.Mlp3":
        n = self.net
        return _lib.Mlp3(_lib.linear_of(n[0].weight.data, n[0].bias.data), _lib.linear_of(n[2].weight.data, n[2].bias.data),
                         _lib.linear_of(n[4].weight.data, n[4].bias.data), None)

    def mu_std(self, states):
        """Differentiable torch form of _distr (src/policy.py:89-96) for callers that need autograd through the policy
        (``SSAC.actor_loss``); ``act`` and the update steps go through the C ABI."""
        mu, raw = self.net(states).chunk(2, dim=-1)
        log_std = -6.0 + 10.0 * torch.sigmoid(raw)
        return mu, log_std.exp()

    def act_with_log_prob(self, states, eval=False, eps=None, want_log_prob=False):
        lib = _lib.load()
        states = states.contiguous().float()
        B, A = states.shape[0], self.net[4].weight.shape[0] // 2
        actions = torch.empty((B, A), device=states.device)
        logp = torch.empty((B,), device=states.device) if want_log_prob else None
        self._noise_step += 1
        eps = eps.contiguous().float() if eps is not None else None        # a named local: the copy must outlive the launch
        noise = _lib.Noise(_lib.ptr(eps), A, self.noise_seed, 17, self._noise_step)
        s = self.as_struct()
        ws = self._ws.get(lib.drpo_policy_workspace_bytes(s, B), states.device)
        _lib.check(lib.drpo_policy_act(s, _lib.ptr(states), B, int(bool(eval)), noise, _lib.ptr(actions), _lib.ptr(logp),
                                       _lib.PREC_FP32, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "drpo_policy_act")
        return actions, logp

    def act(self, states, eval, eps=None):
        """TorchPolicy.act (src/policy.py:77-80): tanh(mu) when eval else a sample."""
        return self.act_with_log_prob(states, eval, eps)[0]
