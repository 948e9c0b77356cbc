"""Helpers shared by the GPU parity tests: build product modules carrying the oracle's seeded weights."""
import numpy as np
import torch

from oracle import drpo_oracle as O


def dev():
    return torch.device("cuda:0")


def make_ensemble(w, S, A):
    import drpo_b200
    ens = drpo_b200.BatchedGaussianEnsemble(drpo_b200.BatchedGaussianEnsemble.Config(), S, A, device=dev())
    missing, unexpected = ens.load_state_dict(w, strict=True)
    ens._elite_inds = [0, 1, 2, 3, 4]
    return ens


def make_ssac(w, S, A, C, B, std_ratio=2.0, penalty_lb=-1.0):
    import drpo_b200
    cfg = drpo_b200.SSAC.Config()
    cfg.batch_size = B
    cfg.constraint_critic_cfg.std_ratio = std_ratio
    cfg.penalty_lb = penalty_lb
    solver = drpo_b200.SSAC(cfg, S, A, C, 10, 100, 300, 10, 10.0, device=dev())
    missing, unexpected = solver.load_state_dict(w, strict=False)
    assert not unexpected, unexpected
    assert all(m.startswith("total_updates") for m in missing), missing
    return solver


def oracle_spec_to_device_env(spec):
    from drpo_b200.envs import DeviceEnv
    return DeviceEnv(kind=spec.kind, state_dim=spec.state_dim, con_dim=spec.con_dim, name=spec.name,
                     hazards=tuple(spec.hazards), hazard_size=spec.hazard_size, goal=tuple(spec.goal), goal_size=spec.goal_size,
                     xy_bound=spec.xy_bound, active_dims=tuple(spec.active_dims), lower=tuple(spec.lower),
                     upper=tuple(spec.upper), done_dims=tuple(spec.done_dims), done_thr=tuple(spec.done_thr),
                     surr_veh_num=spec.surr_veh_num, surr_start=spec.surr_start, veh_length=spec.veh_length,
                     veh_width=spec.veh_width)


def assert_close(got, want, rtol=1e-5, what="", max_outlier_frac=0.0):
    """Scale-relative closeness: every element must satisfy |got-want| <= rtol*|want| + rtol*max|want|, i.e. the error is
    bounded by rtol of the element plus rtol of the tensor's largest magnitude (at most 2*rtol of the scale).  This is NOT a
    pure per-element relative error: values much smaller than the tensor's scale are held to the scale's tolerance.  NaNs must
    sit in the same places.  Returns max|err| / max|want|."""
    got = torch.as_tensor(got).detach().double().cpu()
    want = torch.as_tensor(want).detach().double().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if want.numel() == 0:
        return 0.0
    nan_g, nan_w = torch.isnan(got), torch.isnan(want)
    finite_w = want[~nan_w]
    scale = float(finite_w.abs().max()) if finite_w.numel() else 0.0
    err = (got - want).abs()
    tol = rtol * want.abs() + rtol * max(scale, 1e-30)
    bad = ((err > tol) & ~nan_g & ~nan_w) | (nan_g != nan_w)
    frac = float(bad.double().mean())
    ok = ~nan_g & ~nan_w
    rel = float(err[ok].max() / max(scale, 1e-30)) if bool(ok.any()) else 0.0
    assert frac <= max_outlier_frac, f"{what}: {int(bad.sum())}/{bad.numel()} elements off, max err/scale {rel:.3e} (rtol {rtol})"
    return rel


def to_dev(x):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(dev())
