"""The DRPO_PREC_BF16 critic step (fused tcgen05 forward/loss/dX kernel + split-K dW kernel, csrc/critic_umma.cu) against
the CPU oracle: per-row intermediates, losses, raw gradients, grad norms and the updated parameters, within the 2e-2 the
north star allows for the bf16 GEMM path (BASELINE.json north_star; tolerances are written next to each assert)."""
import ctypes as C

import pytest
import torch

from oracle import drpo_oracle as O
from tests.util import assert_close, dev, make_ssac, to_dev

pytestmark = pytest.mark.gpu


def _octets(x):
    """[rows, F] -> the kernels' slab-octet layout [rows/64][F/8][64][8] (bf16): per 64-row slab one [64, 8] panel per 8 features."""
    rows, F = x.shape
    assert rows % 64 == 0
    return x.view(rows // 64, 64, F // 8, 8).permute(0, 2, 1, 3).contiguous()


@pytest.mark.parametrize("rows,n_in,ksplit", [(128, 256, 1), (1024, 256, 3), (4096, 64, 7), (640, 16, 2)])
def test_dw_kernel_matches_matmul(rows, n_in, ksplit):
    """critic_dw_kernel: dW[256, n_in] = dH^T H with both operands read as MN-major UMMA tiles from the octet layout."""
    from drpo_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(rows + n_in)
    dH = (torch.randn(rows, 256, generator=g) * 0.1).to(torch.bfloat16)
    H = torch.randn(rows, n_in, generator=g).relu().to(torch.bfloat16)
    want = dH.float().t() @ H.float()
    a = to_dev(_octets(dH)); b = to_dev(_octets(H))
    partial = torch.zeros(ksplit * 256 * n_in + 4, device=dev())
    out = torch.zeros(256, n_in, device=dev())
    _lib.check(lib.drpo_debug_critic_dw(a.data_ptr(), b.data_ptr(), n_in // 8, rows, ksplit, partial.data_ptr(), out.data_ptr(),
                                        _lib.stream_ptr()), "drpo_debug_critic_dw")
    _lib.check_kernel_status("dW kernel")
    # bf16 products are exact in fp32; only the summation order differs
    assert_close(out, want, 1e-5, f"dW rows={rows} n_in={n_in} ksplit={ksplit}")


def _inputs(S, A, C, B, seed=62):
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(B, S, generator=g); act = torch.rand(B, A, generator=g) * 2 - 1
    nobs = obs + 0.1 * torch.randn(B, S, generator=g); rew = torch.randn(B, generator=g)
    done = torch.rand(B, generator=g) < 0.1; viol = torch.rand(B, generator=g) < 0.1
    cv = (torch.randn(B, generator=g) - 0.5) if C == 1 else (torch.randn(B, C, generator=g) - 0.5)
    noise = (torch.randn(B, A, generator=g), torch.randn(B, A, generator=g), torch.randn(*cv.shape, generator=g))
    return [obs, act, nobs, rew, done, viol, cv], noise


@pytest.mark.parametrize("S,A,C,B", [(51, 2, 1, 4096), (12, 2, 2, 1000), (4, 1, 4, 333)])
def test_critic_bf16_step_vs_oracle(S, A, C, B):
    import drpo_b200
    from drpo_b200 import _lib
    lib = _lib.load()
    w = O.make_ssac_weights(61, S, A, C)
    solver = make_ssac(w, S, A, C, B)
    solver.precision = drpo_b200.PREC_BF16
    wo = {k: v.clone() for k, v in w.items()}
    batch, noise = _inputs(S, A, C, B)
    # per-row intermediates of the oracle (before its in-place update)
    obs, act, nobs, rew, done, viol, cv = batch
    with torch.no_grad():
        a1, x1, mu1, std1 = O.policy_act(wo, "actor.", nobs, noise[0])
        logp = O.squashed_log_prob(mu1, std1, x1)
        q1t, q2t = O.q_forward(wo, "critic_target.", nobs, a1)
        a2 = O.policy_act(wo, "actor_safe.", nobs, noise[1])[0]
        mt, st = O.qc_forward(wo, "constraint_critic_target.", nobs, a2)
        nqc = mt + torch.clamp(noise[2], -2.0, 2.0) * st
        q1, q2 = O.q_forward(wo, "critic.", obs, act)
        qm, _ = O.qc_forward(wo, "constraint_critic.", obs, act)
    lq, lc, aux = O.critic_update(wo, batch, noise, O.SSACHyper(), 0.0, O.AdamState(), 3e-4)

    rows = torch.zeros(B, 16, device=dev())
    _lib.check(lib.drpo_debug_critic_rows(rows.data_ptr()), "drpo_debug_critic_rows")
    try:
        glq, glc = solver.update_critic(*[to_dev(b) for b in batch], noise=tuple(to_dev(n) for n in noise))
        torch.cuda.synchronize()
    finally:
        lib.drpo_debug_critic_rows(None)
    _lib.check_kernel_status("critic bf16 step")
    r = rows.cpu()
    first = (lambda t: t if t.dim() == 1 else t[:, 0])
    # values: 2e-2 relative to the tensor's scale (bf16 operands through three dense layers)
    assert_close(r[:, 0], a1[:, 0], 2e-2, "a1"); assert_close(r[:, 3], a2[:, 0], 2e-2, "a2")
    assert_close(r[:, 2], logp, 2e-2, "log-prob")
    assert_close(r[:, 5], q1t, 2e-2, "target Q1"); assert_close(r[:, 6], q2t, 2e-2, "target Q2")
    assert_close(r[:, 7], first(nqc), 2e-2, "target Qc sample")
    assert_close(r[:, 8], q1, 2e-2, "Q1"); assert_close(r[:, 9], q2, 2e-2, "Q2")
    assert_close(r[:, 10], first(qm), 2e-2, "Qc mean")
    assert_close(glq, lq, 2e-2, "loss_q"); assert_close(glc, lc, 2e-2, "loss_c")
    assert_close(solver._losses[2], aux["grad_norm_q"], 2e-2, "grad norm Q")
    assert_close(solver._losses[3], aux["grad_norm_c"], 2e-2, "grad norm Qc")
    gviews = solver.critic_arena_views(solver.critic_optimizer.grad)
    for k in [k for k in w if k.startswith(("critic.", "constraint_critic."))]:
        # bf16 activations and activation gradients summed over the batch: the bulk within 2e-2 of the tensor's scale (the
        # per-row rounding errors average out with the batch size, so the small batches here are the hard case: up to 5 % of
        # the entries may exceed it), every entry bounded by 2e-1 of the scale
        assert_close(gviews[k], aux["grads_raw"][k], 2e-2, f"grad {k}", max_outlier_frac=5e-2)
        assert_close(gviews[k], aux["grads_raw"][k], 2e-1, f"grad {k} (outlier bound)")
    # one Adam step moves every weight by about lr: parameters must stay within a step of the oracle's
    sd = solver.state_dict()
    for k in wo:
        if k.startswith(("critic", "constraint_critic")):
            assert float((sd[k].cpu() - wo[k]).abs().max()) <= 2.5 * 3e-4, k


def test_critic_bf16_row_sharding_matches_full_batch():
    """Phase 1 on two row shards (global normaliser, global row ids) sums to the full-batch gradient: what the multi-GPU
    all-reduce relies on."""
    import drpo_b200
    S, A, C, B = 12, 2, 2, 1024
    w = O.make_ssac_weights(5, S, A, C)
    batch, noise = _inputs(S, A, C, B, seed=9)
    full = make_ssac(w, S, A, C, B); full.precision = drpo_b200.PREC_BF16
    full.update_critic(*[to_dev(b) for b in batch], noise=tuple(to_dev(n) for n in noise))
    g_full = full.critic_optimizer.grad.clone()
    total = torch.zeros_like(g_full)
    for lo, hi in ((0, 384), (384, 1024)):
        part = make_ssac(w, S, A, C, hi - lo); part.precision = drpo_b200.PREC_BF16
        part._global_batch_override = B
        part.update_critic(*[to_dev(b[lo:hi]) for b in batch], noise=tuple(to_dev(n[lo:hi]) for n in noise), phases=1)
        total += part.critic_optimizer.grad
    assert_close(total, g_full, 1e-4, "sum of shard gradients")


@pytest.mark.parametrize("B", [1, 129])
def test_critic_bf16_ragged_tiny_batches(B):
    """Batches far below one tile pair (256 rows): padded rows must contribute nothing to losses or gradients."""
    import drpo_b200
    S, A, C = 11, 2, 1
    w = O.make_ssac_weights(3, S, A, C)
    batch, noise = _inputs(S, A, C, B, seed=17)
    wo = {k: v.clone() for k, v in w.items()}
    lq, lc, aux = O.critic_update(wo, batch, noise, O.SSACHyper(), 0.0, O.AdamState(), 3e-4)
    solver = make_ssac(w, S, A, C, B); solver.precision = drpo_b200.PREC_BF16
    glq, glc = solver.update_critic(*[to_dev(b) for b in batch], noise=tuple(to_dev(n) for n in noise), phases=1)
    from drpo_b200 import _lib
    _lib.check_kernel_status("critic bf16 step")
    # a single row has no averaging at all: 5e-2 on the losses, gradients 1e-1 of scale
    assert_close(glq, lq, 5e-2, "loss_q"); assert_close(glc, lc, 5e-2, "loss_c")
    gviews = solver.critic_arena_views(solver.critic_optimizer.grad)
    for k in [k for k in w if k.startswith(("critic.", "constraint_critic."))]:
        assert_close(gviews[k], aux["grads_raw"][k], 1e-1, f"grad {k}", max_outlier_frac=5e-2)


def test_critic_bf16_full_size_matches_fp32_path():
    """BASELINE size (tracking dims, B = 65 536, in-kernel Philox noise): the fused bf16 step against the fp32 CUDA path (itself
    pinned to the oracle at smaller sizes) on identical inputs and noise streams: losses, grad norms, gradients."""
    import drpo_b200
    from drpo_b200 import synthetic
    _, S, A, C = synthetic.WORKLOADS["tracking"]
    B = 65536
    w = synthetic.make_ssac_weights(43567, S, A, C)
    batch = [t.to(dev()) for t in synthetic.make_critic_batch("tracking", B, 49283)]
    res = {}
    for name, prec in (("fp32", drpo_b200.PREC_FP32), ("bf16", drpo_b200.PREC_BF16)):
        cfg = drpo_b200.SSAC.Config(); cfg.batch_size = B; cfg.constraint_critic_cfg.std_ratio = 1.0
        solver = drpo_b200.SSAC(cfg, S, A, C, 10, 100, 1000, 10, 5.0, device=dev())
        solver.load_state_dict(w, strict=False)
        solver.precision = prec
        lq, lc = solver.update_critic(*batch)
        res[name] = (float(lq), float(lc), solver._losses[2:4].clone().cpu(), solver.critic_optimizer.grad.clone().cpu())
    from drpo_b200 import _lib
    _lib.check_kernel_status("critic bf16 step")
    f, b = res["fp32"], res["bf16"]
    assert b[0] == pytest.approx(f[0], rel=2e-2) and b[1] == pytest.approx(f[1], rel=2e-2)
    assert_close(b[2], f[2], 2e-2, "grad norms")
    # at this batch size the per-row rounding errors average out: whole arena within 2e-2 of its scale, <= 0.5 % outliers
    assert_close(b[3], f[3], 2e-2, "gradient arena", max_outlier_frac=5e-3)


def test_multiplier_step_under_bf16_precision():
    """drpo_multiplier_step accepts DRPO_PREC_BF16 (it runs its dense layers as TF32 tensor-op GEMMs)."""
    import drpo_b200
    S, A, C, B = 12, 2, 2, 1000
    w = O.make_ssac_weights(61, S, A, C)
    solver = make_ssac(w, S, A, C, B); solver.precision = drpo_b200.PREC_BF16
    g = torch.Generator().manual_seed(4)
    obs = torch.randn(B, S, generator=g); eps = torch.randn(B, A, generator=g)
    lm, _ = O.multiplier_update({k: v.clone() for k, v in w.items()}, obs, eps, O.SSACHyper(), C, O.AdamState(), 3e-4)
    glm = solver.update_multiplier(to_dev(obs), eps=to_dev(eps))
    assert_close(glm, lm, 2e-2, "multiplier loss")
