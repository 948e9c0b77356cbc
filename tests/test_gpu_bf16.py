"""The DRPO_PREC_BF16 rollout (tcgen05/TMEM fused chain): layer-by-layer accumulator dumps against a torch emulation
of the same bf16-in / fp32-accumulate arithmetic, then the whole rollout against the fp32 oracle within 2e-2."""
import numpy as np
import pytest
import torch

from oracle import drpo_oracle as O
from tests.util import assert_close, dev, oracle_spec_to_device_env, to_dev

pytestmark = pytest.mark.gpu

LAYERS = ["actor.0", "actor.2", "actor.4", "trunk.0", "trunk.2", "diff.0", "lvar.0", "diff.2", "lvar.2"]


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def emulate(ws, wm, member, states, eps_p):
    """fp32 accumulators of the 9 dense layers with bf16-rounded weights/biases/activations (what the kernel computes,
    up to fp32 summation order and the tanh.approx SiLU)."""
    acc = []
    lin = lambda x, W, b: bf(x) @ bf(W).t() + bf(b)
    a0 = lin(states, ws["actor.net.0.weight"], ws["actor.net.0.bias"]); acc.append(a0)
    a1 = lin(torch.relu(a0), ws["actor.net.2.weight"], ws["actor.net.2.bias"]); acc.append(a1)
    a2 = lin(torch.relu(a1), ws["actor.net.4.weight"], ws["actor.net.4.bias"]); acc.append(a2)
    mu, raw = a2.chunk(2, dim=-1)
    std = torch.exp(-6.0 + 10.0 * torch.sigmoid(raw))
    action = torch.tanh(mu + std * eps_p)
    norm = (states - wm["state_normalizer.mean"]) / (wm["state_normalizer.std"] + 1e-6)
    x0 = torch.cat([norm, action], -1)
    W = lambda k: wm[k][member]
    t0 = lin(x0, W("trunk.0.weight"), W("trunk.0.bias")); acc.append(t0)
    t1 = lin(torch.nn.functional.silu(t0), W("trunk.2.weight"), W("trunk.2.bias")); acc.append(t1)
    h2 = torch.nn.functional.silu(t1)
    d0 = lin(h2, W("diff_head.0.weight"), W("diff_head.0.bias")); acc.append(d0)
    l0 = lin(h2, W("log_var_head.0.weight"), W("log_var_head.0.bias")); acc.append(l0)
    d1 = lin(torch.nn.functional.silu(d0), W("diff_head.2.weight"), W("diff_head.2.bias")); acc.append(d1)
    l1 = lin(torch.nn.functional.silu(l0), W("log_var_head.2.weight"), W("log_var_head.2.bias")); acc.append(l1)
    return acc, action


def _alg(spec, wm, ws, B, S, A):
    import drpo_b200
    cfg = drpo_b200.SMBPO.Config()
    cfg.rollout_batch_size, cfg.buffer_max = B, max(B * 16, 4096)
    env = oracle_spec_to_device_env(spec)
    env.action_dim = A
    alg = drpo_b200.SMBPO(cfg, env, device=dev())
    alg.model_ensemble.load_state_dict(wm)
    alg.model_ensemble._elite_inds = [0, 1, 2, 3, 4]
    alg.solver.load_state_dict(ws, strict=False)
    alg.rollout_precision = drpo_b200.PREC_BF16
    return alg


@pytest.mark.parametrize("tag,S,A,C,B", [("quadrotor", 12, 2, 2, 300), ("cartpole", 4, 1, 4, 128), ("tracking", 51, 2, 1, 1000)])
def test_layer_accumulators(tag, S, A, C, B):
    spec = {"quadrotor": O.env_quadrotor(), "cartpole": O.env_cartpole(), "tracking": O.env_tracking(10, 1)}[tag]
    wm, ws = O.make_ensemble_weights(71, S, A, diff_scale=0.05), O.make_ssac_weights(72, S, A, C)
    g = torch.Generator().manual_seed(73)
    init = torch.randn(B, S, generator=g) * 0.5
    eps_p, eps_m = torch.randn(1, B, A, generator=g), torch.randn(1, B, S + 1, generator=g)
    member = 3
    want, _ = emulate(ws, wm, member, init, eps_p[0])
    alg = _alg(spec, wm, ws, B, S, A)
    alg.horizon = 1
    report = []
    for l, name in enumerate(LAYERS):
        got = alg.rollout(alg.actor, initial_states=to_dev(init), noise=(to_dev(eps_p), to_dev(eps_m)), member_idx=[member], _debug_layer=l)
        torch.cuda.synchronize()
        w = want[l]
        err = float((got.cpu() - w).abs().max() / w.abs().max())
        report.append(f"{name}: {err:.2e}")
        # first layer of each net sees exact bf16 products: only the summation order differs; deeper layers inherit
        # one-ulp bf16 flips of the previous activation
        tol = 2e-5 if l in (0,) else 2e-2
        assert err <= tol, f"{tag} layer {l} ({name}) rel err {err:.3e} > {tol}; all: {report}"
    print(tag, " ".join(report))


@pytest.mark.parametrize("tag,S,A,C,B", [("quadrotor", 12, 2, 2, 5000), ("cartpole", 4, 1, 4, 3000), ("point_robot", 11, 2, 1, 2000)])
def test_bf16_rollout_vs_oracle(tag, S, A, C, B):
    """Step 0 of the bf16 rollout against the fp32 oracle: values within 2e-2; masks may differ only where the
    constraint margin is below that tolerance.  Later steps: the alive counts track the oracle's closely."""
    spec, H = {"quadrotor": O.env_quadrotor(), "cartpole": O.env_cartpole(), "point_robot": O.env_point_robot()}[tag], 10
    wm, ws = O.make_ensemble_weights(31, S, A, diff_scale=0.05), O.make_ssac_weights(32, S, A, C)
    g = torch.Generator().manual_seed(33)
    init = torch.randn(B, S, generator=g) * 0.3
    if tag == "quadrotor":
        init[:, 2] = 0.55 + 0.9 * torch.rand(B, generator=g)
    eps_p, eps_m = torch.randn(H, B, A, generator=g), torch.randn(H, B, S + 1, generator=g)
    members = [int(x) for x in torch.randint(0, 5, (H,), generator=g)]
    ref, counts, _ = O.rollout(ws, wm, spec, init, H, eps_p, eps_m, members)
    alg = _alg(spec, wm, ws, B, S, A)
    alg.horizon = H
    view = alg.rollout(alg.actor, initial_states=to_dev(init), noise=(to_dev(eps_p), to_dev(eps_m)), member_idx=members)
    torch.cuda.synchronize()
    out = view.get(as_dict=True)
    got_counts = view.counts()
    assert got_counts[0] == B
    n0 = B
    for k in ("states", "actions", "next_states", "rewards"):
        assert_close(out[k][:n0], ref[k][:n0], 2e-2, f"{tag}.{k} step 0")
    cv_ref = ref["constraint_values"][:n0].reshape(n0, -1)
    margin = cv_ref.abs().min(dim=1).values > 0.05
    assert torch.equal(out["violations"][:n0].cpu()[margin], ref["violations"][:n0][margin])
    for a, b in zip(got_counts, counts):
        assert abs(a - b) <= max(3, 0.02 * b), (got_counts, counts)
    # the masks stored are exactly the hooks of the stored next_states (bit-exact self-consistency)
    d, v, cv = O.hooks(spec, out["next_states"].cpu().numpy())
    assert np.array_equal(d, out["dones"].cpu().numpy()) and np.array_equal(v, out["violations"].cpu().numpy())
    assert np.array_equal(cv.reshape(-1), out["constraint_values"].cpu().numpy().reshape(-1))
