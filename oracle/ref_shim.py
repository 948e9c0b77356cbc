"""TEST INFRASTRUCTURE ONLY — import shim for the *unmodified* reference.

Used by ``oracle/make_golden.py`` (run in the build container, where
``/root/reference`` is mounted) to import the reference's own modules
(``src.dynamics``, ``src.ssac``, ``src.smbpo``, ``src.sampling``, ``src.policy``,
``src.env.point_robot``, ``src.env.poles.constraints``, the tracking env) so that golden
vectors can be generated from the reference itself.  Nothing of the reference is
edited or copied: we only inject stand-ins for three third-party packages the image
does not have (``gym``, ``h5py``, ``matplotlib``) into ``sys.modules``.

This file must never be imported by the product package or by the timed GPU region of ``bench.py``.  On the GPU box
``/root/reference`` does not exist; ``bench.py``'s CPU-baseline legs (``cpu_baseline`` / ``--impl reference``) and the drop-in
test resolve the staged copy ``baseline/_ref/`` instead (see ``_find_reference``).
"""
import os
import sys
import types

import numpy as np

def _find_reference() -> str:
    """$DRPO_REF, else /root/reference (build container), else the staged copy baseline/_ref/ (git-ignored, shipped to the GPU
    box by gpurun; written by oracle/stage_reference.py) - BASELINE.md §4."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for c in (os.environ.get("DRPO_REF"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if c and os.path.isdir(os.path.join(c, "src")):
            return c
    return os.environ.get("DRPO_REF", "/root/reference")


REFERENCE_ROOT = _find_reference()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


class _Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = dtype
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)
        return [seed]


class _Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        super().__init__(shape, dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)


class _Discrete(_Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = n

    def sample(self):
        return int(self._rng.randint(self.n))


class _Env:
    metadata = {}

    def seed(self, seed=None):
        return [seed]

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def close(self):
        pass


class _Wrapper(_Env):
    def __init__(self, env):
        self.__dict__["env"] = env

    def __getattr__(self, name):
        # forward unknown attributes (con_dim, observation_space, check_done …) to the wrapped env
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.__dict__["env"], name)

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, action):
        return self.env.step(action)


class _RescaleAction(_Wrapper):
    def __init__(self, env, a, b):
        super().__init__(env)
        self.a, self.b = a, b


def _install_stubs():
    if "gym" in sys.modules and getattr(sys.modules["gym"], "_drpo_stub", False):
        return
    gym = types.ModuleType("gym")
    gym._drpo_stub = True
    gym.Env, gym.Wrapper, gym.Space = _Env, _Wrapper, _Space
    gym.register = lambda *a, **k: None
    spaces = types.ModuleType("gym.spaces")
    spaces.Box, spaces.Discrete, spaces.Space = _Box, _Discrete, _Space
    gym.spaces = spaces
    wrappers = types.ModuleType("gym.wrappers")
    wrappers.RescaleAction = _RescaleAction
    gym.wrappers = wrappers
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    utils.seeding = seeding
    gym.utils = utils
    sys.modules.update({
        "gym": gym, "gym.spaces": spaces, "gym.wrappers": wrappers,
        "gym.utils": utils, "gym.utils.seeding": seeding,
    })
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    mpl = types.ModuleType("matplotlib")
    for sub in ("pyplot", "patches", "transforms", "figure", "axes", "colors", "lines", "animation"):
        m = types.ModuleType(f"matplotlib.{sub}")
        setattr(mpl, sub, m)
        sys.modules.setdefault(f"matplotlib.{sub}", m)
    sys.modules.setdefault("matplotlib", mpl)


def import_reference(device="cpu"):
    """Make ``import src.*`` resolve to the unmodified reference, forced onto ``device`` (the CPU for every parity / baseline use;
    "cuda" only for bench.py's clearly labelled reference-on-GPU number).  The reference binds its device at import time
    (src/torch_util.py:9), so the choice holds for the life of the process."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torch
    import src.torch_util as tu  # noqa: E402
    if getattr(tu, "_drpo_forced", None) not in (None, device):
        raise RuntimeError(f"the reference is already imported on {tu._drpo_forced}")
    tu.device = torch.device(device)
    tu._drpo_forced = device
    return sys.modules["src"]
