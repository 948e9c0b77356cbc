"""ctypes binding of libdrpo_sm100.so (include/drpo_b200.h).

There is NO fallback: if the shared library is missing, or a call returns an error, a RuntimeError is raised.
The library is built in-tree by ``__graft_entry__.build()`` / ``python -m drpo_b200.build``.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdrpo_sm100.so")

PREC_FP32, PREC_BF16, PREC_TF32 = 0, 1, 2
ENV_POINT_ROBOT, ENV_BOUNDED, ENV_TRACKING = 0, 1, 2
MAX_ACTIVE, MAX_DONE_DIMS, MAX_HAZARDS, MAX_CON = 4, 4, 4, 8

c_f32p = C.c_void_p  # device pointers travel as integers


class EnvParams(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("state_dim", C.c_int32), ("con_dim", C.c_int32), ("n_hazards", C.c_int32),
        ("hazard_xy", (C.c_double * 2) * MAX_HAZARDS), ("hazard_size", C.c_double), ("goal_xy", C.c_double * 2),
        ("goal_size", C.c_double), ("xy_bound", C.c_float),
        ("n_active", C.c_int32), ("active_dims", C.c_int32 * MAX_ACTIVE), ("lower", C.c_double * MAX_ACTIVE),
        ("upper", C.c_double * MAX_ACTIVE),
        ("n_done_dims", C.c_int32), ("done_dims", C.c_int32 * MAX_DONE_DIMS), ("done_thr", C.c_float * MAX_DONE_DIMS),
        ("surr_veh_num", C.c_int32), ("surr_start", C.c_int32), ("veh_length", C.c_double), ("veh_width", C.c_double),
    ]


class Linear(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p), ("in_dim", C.c_int32), ("out_dim", C.c_int32)]


class Ensemble(C.Structure):
    _fields_ = [
        ("state_dim", C.c_int32), ("action_dim", C.c_int32), ("ensemble_size", C.c_int32), ("hidden", C.c_int32),
        ("norm_mean", C.c_void_p), ("norm_std", C.c_void_p), ("min_log_var", C.c_void_p), ("max_log_var", C.c_void_p),
        ("trunk0_w", C.c_void_p), ("trunk0_b", C.c_void_p), ("trunk1_w", C.c_void_p), ("trunk1_b", C.c_void_p),
        ("diff0_w", C.c_void_p), ("diff0_b", C.c_void_p), ("diff1_w", C.c_void_p), ("diff1_b", C.c_void_p),
        ("lvar0_w", C.c_void_p), ("lvar0_b", C.c_void_p), ("lvar1_w", C.c_void_p), ("lvar1_b", C.c_void_p),
        ("packed_bf16", C.c_void_p),
    ]


class Mlp3(C.Structure):
    _fields_ = [("l0", Linear), ("l1", Linear), ("l2", Linear), ("packed_bf16", C.c_void_p)]


class Qc(C.Structure):
    _fields_ = [("trunk0", Linear), ("trunk1", Linear), ("mean0", Linear), ("mean1", Linear), ("lstd0", Linear),
                ("lstd1", Linear)]


class Noise(C.Structure):
    _fields_ = [("eps", C.c_void_p), ("row_stride", C.c_int64), ("seed", C.c_uint64), ("stream_tag", C.c_uint32),
                ("step", C.c_uint32)]


class Buffer(C.Structure):
    _fields_ = [("states", C.c_void_p), ("actions", C.c_void_p), ("next_states", C.c_void_p), ("rewards", C.c_void_p),
                ("dones", C.c_void_p), ("violations", C.c_void_p), ("constraint_values", C.c_void_p),
                ("pointer", C.c_void_p), ("capacity", C.c_int64), ("state_dim", C.c_int32), ("action_dim", C.c_int32),
                ("con_dim", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("act", C.c_void_p), ("next_obs", C.c_void_p), ("rew", C.c_void_p),
                ("done", C.c_void_p), ("viol", C.c_void_p), ("cv", C.c_void_p)]


class RolloutArgs(C.Structure):
    _fields_ = [
        ("actor", C.POINTER(Mlp3)), ("ensemble", C.POINTER(Ensemble)), ("env", C.POINTER(EnvParams)),
        ("initial_states", C.c_void_p), ("batch", C.c_int64), ("traj_id_offset", C.c_int64), ("horizon", C.c_int32),
        ("member_idx_host", C.POINTER(C.c_int32)), ("eps_policy", C.c_void_p), ("eps_model", C.c_void_p),
        ("eps_batch_stride", C.c_int64), ("seed", C.c_uint64), ("virt", Buffer), ("step_counts", C.c_void_p),
        ("precision", C.c_int32), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("stream", C.c_void_p),
        ("init_ready_flags", C.c_void_p), ("init_rows_per_flag_log2", C.c_int32),
    ]


class Adam(C.Structure):
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("weight_decay", C.c_double), ("step", C.c_int32)]


class CriticArgs(C.Structure):
    _fields_ = [
        ("batch", Batch), ("batch_size", C.c_int64), ("global_batch_size", C.c_int64),
        ("state_dim", C.c_int32), ("action_dim", C.c_int32), ("con_dim", C.c_int32),
        ("actor", C.POINTER(Mlp3)), ("actor_safe", C.POINTER(Mlp3)),
        ("q", Mlp3 * 2), ("q_target", Mlp3 * 2), ("qc", Qc), ("qc_target", Qc),
        ("params", C.c_void_p), ("grads", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p),
        ("target_params", C.c_void_p), ("n_params_q", C.c_int64), ("n_params_qc", C.c_int64), ("log_alpha", C.c_void_p),
        ("eps_actor", C.c_void_p), ("eps_safe", C.c_void_p), ("eps_qc", C.c_void_p),
        ("seed", C.c_uint64), ("noise_step", C.c_uint32), ("row_id_offset", C.c_int64),
        ("discount", C.c_double), ("tau", C.c_double), ("grad_norm", C.c_double), ("qc_td_bound", C.c_double),
        ("adam", Adam), ("phases", C.c_int32), ("losses", C.c_void_p), ("precision", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("stream", C.c_void_p),
    ]


class MultiplierArgs(C.Structure):
    _fields_ = [
        ("obs", C.c_void_p), ("batch_size", C.c_int64), ("global_batch_size", C.c_int64),
        ("state_dim", C.c_int32), ("action_dim", C.c_int32), ("con_dim", C.c_int32),
        ("actor", C.POINTER(Mlp3)), ("actor_safe", C.POINTER(Mlp3)), ("qc", C.POINTER(Qc)), ("lam", Mlp3),
        ("params", C.c_void_p), ("grads", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p),
        ("n_params", C.c_int64), ("eps_actor", C.c_void_p), ("seed", C.c_uint64), ("noise_step", C.c_uint32),
        ("row_id_offset", C.c_int64),
        ("std_ratio", C.c_double), ("constraint_threshold", C.c_double), ("penalty_lb", C.c_double),
        ("penalty_ub", C.c_double), ("upper_bound", C.c_double), ("lam_epsilon", C.c_double), ("grad_norm", C.c_double),
        ("adam", Adam), ("phases", C.c_int32), ("losses", C.c_void_p), ("precision", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("stream", C.c_void_p),
    ]


class ActorArgs(C.Structure):
    _fields_ = [
        ("obs", C.c_void_p), ("batch_size", C.c_int64), ("global_batch_size", C.c_int64),
        ("state_dim", C.c_int32), ("action_dim", C.c_int32), ("con_dim", C.c_int32),
        ("actor", Mlp3), ("actor_safe", Mlp3), ("q", C.POINTER(Mlp3)), ("qc", C.POINTER(Qc)), ("lam", C.POINTER(Mlp3)),
        ("params_actor", C.c_void_p), ("grads_actor", C.c_void_p), ("m_actor", C.c_void_p), ("v_actor", C.c_void_p), ("n_actor", C.c_int64),
        ("params_safe", C.c_void_p), ("grads_safe", C.c_void_p), ("m_safe", C.c_void_p), ("v_safe", C.c_void_p), ("n_safe", C.c_int64),
        ("log_alpha", C.c_void_p), ("alpha_m", C.c_void_p), ("alpha_v", C.c_void_p),
        ("eps_actor", C.c_void_p), ("eps_safe", C.c_void_p), ("seed", C.c_uint64), ("noise_step", C.c_uint32), ("row_id_offset", C.c_int64),
        ("std_ratio", C.c_double), ("multiplier_ub", C.c_double), ("grad_norm", C.c_double), ("target_entropy", C.c_double),
        ("adam_actor", Adam), ("adam_alpha", Adam), ("adam_safe", Adam),
        ("phases", C.c_int32), ("losses", C.c_void_p), ("precision", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("stream", C.c_void_p),
    ]


class EnsembleTrainArgs(C.Structure):
    _fields_ = [
        ("ens", Ensemble),
        ("params", C.c_void_p), ("grads", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p), ("n_params", C.c_int64),
        ("states", C.c_void_p), ("actions", C.c_void_p), ("targets", C.c_void_p), ("n_rows", C.c_int64), ("shared_rows", C.c_int32),
        ("log_var_bound_weight", C.c_double), ("adam", Adam), ("phases", C.c_int32), ("losses", C.c_void_p), ("precision", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("stream", C.c_void_p),
    ]


SHIELD_NONE, SHIELD_SAFE, SHIELD_LINEAR = 0, 1, 2
SHIELD_TYPES = {"none": SHIELD_NONE, "safe": SHIELD_SAFE, "linear": SHIELD_LINEAR}


class ShieldArgs(C.Structure):
    _fields_ = [
        ("actor", C.POINTER(Mlp3)), ("actor_safe", C.POINTER(Mlp3)), ("qc", C.POINTER(Qc)), ("states", C.c_void_p), ("n", C.c_int64),
        ("state_dim", C.c_int32), ("action_dim", C.c_int32), ("con_dim", C.c_int32), ("shield_type", C.c_int32),
        ("eval_perf", C.c_int32), ("uncertainty", C.c_int32), ("std_ratio", C.c_float), ("threshold", C.c_float),
        ("noise_perf", C.POINTER(Noise)), ("actions", C.c_void_p), ("qc_perf", C.c_void_p), ("choice", C.c_void_p), ("path", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("stream", C.c_void_p),
    ]


# every symbol include/drpo_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("drpo_last_error", C.c_char_p, []),
    ("drpo_abi_version", C.c_int, []),
    ("drpo_launch_count", C.c_int64, []),
    ("drpo_kernel_status", C.c_int, []),
    ("drpo_kernel_status_peek", C.c_int, []),
    ("drpo_timing_enable", None, [C.c_int32]),
    ("drpo_timing_read", C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    ("drpo_philox_normal", C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]),
    ("drpo_hooks_eval", C.c_int, [C.POINTER(EnvParams), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("drpo_ensemble_workspace_bytes", C.c_int64, [C.POINTER(Ensemble), C.c_int64]),
    ("drpo_ensemble_forward", C.c_int, [C.POINTER(Ensemble), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    ("drpo_ensemble_sample", C.c_int, [C.POINTER(Ensemble), C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(Noise),
                                       C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    ("drpo_policy_workspace_bytes", C.c_int64, [C.POINTER(Mlp3), C.c_int64]),
    ("drpo_policy_act", C.c_int, [C.POINTER(Mlp3), C.c_void_p, C.c_int64, C.c_int32, C.POINTER(Noise), C.c_void_p,
                                  C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    ("drpo_buffer_gather", C.c_int, [C.POINTER(Buffer), C.POINTER(Buffer), C.c_void_p, C.c_int64, C.c_int64, C.c_float,
                                     C.c_float, C.c_float, C.c_float, C.POINTER(Batch), C.c_void_p]),
    ("drpo_rollout_workspace_bytes", C.c_int64, [C.POINTER(RolloutArgs)]),
    ("drpo_rollout", C.c_int, [C.POINTER(RolloutArgs)]),
    ("drpo_debug_rollout_layer", C.c_int, [C.POINTER(RolloutArgs), C.c_int32, C.c_void_p]),
    ("drpo_critic_workspace_bytes", C.c_int64, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    ("drpo_critic_step", C.c_int, [C.POINTER(CriticArgs)]),
    ("drpo_debug_critic_rows", C.c_int, [C.c_void_p]),
    ("drpo_debug_critic_prof", C.c_int, [C.c_void_p]),
    ("drpo_debug_solver_rows", C.c_int, [C.c_void_p]),
    ("drpo_debug_critic_dw", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("drpo_multiplier_workspace_bytes", C.c_int64, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    ("drpo_multiplier_step", C.c_int, [C.POINTER(MultiplierArgs)]),
    ("drpo_ensemble_train_workspace_bytes", C.c_int64, [C.POINTER(Ensemble), C.c_int64]),
    ("drpo_ensemble_train_step", C.c_int, [C.POINTER(EnsembleTrainArgs)]),
    ("drpo_actor_workspace_bytes", C.c_int64, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    ("drpo_actor_step", C.c_int, [C.POINTER(ActorArgs)]),
    ("drpo_shield_workspace_bytes", C.c_int64, [C.POINTER(Mlp3), C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    ("drpo_shield_act", C.c_int, [C.POINTER(ShieldArgs)]),
    ("drpo_qc_workspace_bytes", C.c_int64, [C.c_int64, C.c_int32]),
    ("drpo_qc_forward", C.c_int, [C.POINTER(Qc), C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_float, C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_void_p]),
]

_lib = None


def load():
    """Load the shared library and bind every declared symbol.  Raises if it is missing: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). drpo_b200 has no CPU or eager fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = restype, argtypes
    if lib.drpo_abi_version() != 2:
        raise RuntimeError("libdrpo_sm100.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {load().drpo_last_error().decode()}")


def check_kernel_status(what):
    """Raise if the last bf16 rollout reported an in-kernel wait time-out (blocking; call where the host synchronises anyway)."""
    lib = load()
    code = lib.drpo_kernel_status()
    if code != 0:
        raise RuntimeError(f"{what}: {lib.drpo_last_error().decode()}")


def peek_kernel_status(what):
    """Non-blocking: raise if any earlier bf16 launch has reported a watchdog time-out (sticky pinned status words).  Called at
    the top of every update / rollout call, so a failure surfaces at the next call after the failing kernel completed."""
    lib = load()
    if lib.drpo_kernel_status_peek() != 0:
        raise RuntimeError(f"{what}: {lib.drpo_last_error().decode()}")


LOSSES_LEN, LOSS_ERR_SLOT = 16, 15


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("drpo_b200 kernels need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("drpo_b200 kernels need contiguous tensors")
    return t.data_ptr()


def stream_ptr():
    # the raw cudaStream_t of torch's current stream; the C accessor costs ~0.3 us where torch.cuda.current_stream().cuda_stream
    # builds a Stream object (~15 us per call - a tenth of a data-parallel update step's host time)
    get = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if get is not None:
        return get(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def linear_of(weight, bias):
    return Linear(ptr(weight), ptr(bias), weight.shape[-1], weight.shape[-2])


class Workspace:
    """Grow-only device scratch owned by PyTorch and lent to the library per call."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        return self.buf
