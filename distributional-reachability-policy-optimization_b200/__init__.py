"""drpo_b200 — B200-native (sm_100a) implementation of DRPO's data-parallel hot path.

Host-side mirror of the reference's Python API (src/dynamics.py, src/smbpo.py, src/ssac.py, env hooks) over the C ABI of
``libdrpo_sm100.so`` (include/drpo_b200.h).  There is no CPU or eager fallback for the hot path.
"""
from . import _lib                                                     # noqa: F401
from .config import BaseConfig, Configurable, Optional                # noqa: F401
from .envs import DeviceEnv, device_env                                # noqa: F401
from .sampling import ConstraintSafetySampleBuffer, RolloutView        # noqa: F401
from .dynamics import BatchedGaussianEnsemble, BatchedLinear, Normalizer   # noqa: F401
from .policy import SquashedGaussianPolicy                              # noqa: F401
from .ssac import SSAC, ConstraintCritic, CriticEnsemble, MLPMultiplier  # noqa: F401
from .smbpo import SMBPO                                                # noqa: F401

PREC_FP32, PREC_BF16, PREC_TF32 = _lib.PREC_FP32, _lib.PREC_BF16, _lib.PREC_TF32
