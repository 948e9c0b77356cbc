"""Device ring buffers with the reference's semantics (src/sampling.py:12-151,215-229).

``ConstraintSafetySampleBuffer`` is the write target of the rollout kernels (they append straight into its storage
and advance ``_pointer`` on the device) and the read source of the critic step (``drpo_buffer_gather``)."""
import torch
from torch import nn

from . import _lib


class ConstraintSafetySampleBuffer(nn.Module):
    COMPONENT_NAMES = ("states", "actions", "next_states", "rewards", "dones", "violations", "constraint_values")

    def __init__(self, state_dim, action_dim, capacity, con_dim=1, device=None):
        super().__init__()
        device = torch.device(device if device is not None else "cuda")
        self.state_dim, self.action_dim, self.con_dim, self.capacity = state_dim, action_dim, con_dim, int(capacity)
        self.device = device
        self._bufs = {}
        self.register_buffer("_pointer", torch.zeros((), dtype=torch.long, device=device))
        comps = (("states", torch.float, [state_dim]), ("actions", torch.float, [action_dim]),
                 ("next_states", torch.float, [state_dim]), ("rewards", torch.float, []), ("dones", torch.bool, []),
                 ("violations", torch.bool, []),
                 ("constraint_values", torch.float, [] if con_dim == 1 else [con_dim]))      # src/sampling.py:226-228
        for name, dtype, shape in comps:
            buf = torch.zeros(self.capacity, *shape, dtype=dtype, device=device)
            self.register_buffer(f"_{name}", buf)
            self._bufs[name] = buf

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        self._bufs = {name: getattr(self, f"_{name}") for name in self.COMPONENT_NAMES}
        self.device = self._pointer.device
        return self

    def __len__(self):
        return min(int(self._pointer), self.capacity)

    def _get1(self, name):
        buf, ptr = self._bufs[name], int(self._pointer)
        if ptr <= self.capacity:
            return buf[:ptr]
        i = ptr % self.capacity
        return torch.cat([buf[i:], buf[:i]])

    def get(self, *names, device=None, as_dict=False):
        if len(names) == 0:
            names = self.COMPONENT_NAMES
        bufs = [self._get1(n) if device is None else self._get1(n).to(device) for n in names]
        if as_dict:
            return dict(zip(names, bufs))
        return bufs if len(bufs) > 1 else bufs[0]

    def append(self, **kwargs):
        assert set(kwargs.keys()) == set(self.COMPONENT_NAMES)
        i = int(self._pointer) % self.capacity
        for name in self.COMPONENT_NAMES:
            self._bufs[name][i] = torch.as_tensor(kwargs[name]).to(self.device)
        self._pointer += 1

    def extend(self, **kwargs):
        assert set(kwargs.keys()) == set(self.COMPONENT_NAMES)
        batch_size = len(list(kwargs.values())[0])
        assert batch_size <= self.capacity, "We do not support extending by more than buffer capacity"
        i = int(self._pointer) % self.capacity
        end = i + batch_size
        if end <= self.capacity:
            for name in self.COMPONENT_NAMES:
                self._bufs[name][i:end] = kwargs[name]
        else:
            fit, overflow = self.capacity - i, end - self.capacity
            for name in self.COMPONENT_NAMES:
                buf, arg = self._bufs[name], kwargs[name]
                buf[-fit:] = arg[:fit]
                buf[:overflow] = arg[-overflow:]
        self._pointer += batch_size

    def sample(self, batch_size, replace=True, device=None, include_indices=False):
        n = len(self)
        indices = torch.randint(n, [batch_size], device=self.device) if replace else \
            torch.randperm(n, device=self.device)[:batch_size]
        bufs = [self._bufs[name][indices] for name in self.COMPONENT_NAMES]
        return (bufs, indices) if include_indices else bufs

    def as_struct(self) -> "_lib.Buffer":
        b = self._bufs
        return _lib.Buffer(_lib.ptr(b["states"]), _lib.ptr(b["actions"]), _lib.ptr(b["next_states"]), _lib.ptr(b["rewards"]),
                           _lib.ptr(b["dones"]), _lib.ptr(b["violations"]), _lib.ptr(b["constraint_values"]),
                           _lib.ptr(self._pointer), self.capacity, self.state_dim, self.action_dim, self.con_dim)


class RolloutView:
    """What ``SMBPO.rollout`` returns: the rows this rollout appended to ``virt_buffer`` (the reference returns its
    scratch buffer, src/smbpo.py:232,249).  Rows are read back lazily from the ring."""
    COMPONENT_NAMES = ConstraintSafetySampleBuffer.COMPONENT_NAMES

    def __init__(self, ring, start_pointer, step_counts):
        # both stay on the device until somebody asks: the rollout itself never synchronises with the host
        self.ring, self._start, self.step_counts = ring, start_pointer, step_counts      # step_counts: device int32 [H+1]

    @property
    def start(self):
        return int(self._start)

    def _check(self):
        from . import _lib
        _lib.check_kernel_status("drpo_rollout")

    def __len__(self):
        n = int(self.step_counts[-1])
        self._check()
        return n

    def counts(self):
        c = [int(c) for c in self.step_counts[:-1].tolist()]
        self._check()
        return c

    def get(self, *names, device=None, as_dict=False):
        if len(names) == 0:
            names = self.COMPONENT_NAMES
        n = len(self)
        idx = (self.start + torch.arange(n, device=self.ring.device)) % self.ring.capacity
        bufs = [self.ring._bufs[name][idx] for name in names]
        if device is not None:
            bufs = [b.to(device) for b in bufs]
        if as_dict:
            return dict(zip(names, bufs))
        return bufs if len(bufs) > 1 else bufs[0]


def shielded_actions(policy, states, eval=False, safe_shield_threshold=-0.1, shield_type="linear"):
    """The per-step action selection of sample_episodes_batched (src/sampling.py:420-439): ``policy`` is the SSAC solver;
    without ``eval`` it is plain ``policy.act``; with it the performance action goes through the safety shield
    ("safe": switch to the safe actor where Qc > threshold; "linear": the mix of safe and performance action closest to
    the performance action whose Qc is <= threshold).  One drpo_shield_act call; stepping the env stays with the caller."""
    if not eval:
        return policy.act(states, eval=False)
    return policy.shield_act(states, eval=True, shield_type=shield_type, safe_shield_threshold=safe_shield_threshold)
