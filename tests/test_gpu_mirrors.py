"""The differentiable torch API mirrors that the product keeps for callers who bring their own autograd (``SSAC.actor_loss``,
``ConstraintCritic.forward_torch``, ``SquashedGaussianPolicy.mu_std``, ``BatchedGaussianEnsemble.compute_loss`` / ``_mse_loss``) are not
used by any update step, but they are public: they are pinned to the oracle here (values AND gradients), and the multi-rank path is
driven through torchrun when the box has two GPUs."""
import json
import os
import random
import subprocess
import sys

import pytest
import torch

from oracle import drpo_oracle as O
from tests.util import assert_close, dev, make_ensemble, make_ssac, to_dev

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("S,A,C", [(11, 2, 1), (4, 1, 4)])
def test_torch_mirrors_match_oracle(monkeypatch, S, A, C):
    B = 96
    w = O.make_ssac_weights(81, S, A, C)
    solver = make_ssac(w, S, A, C, B)
    g = torch.Generator().manual_seed(82)
    obs, act = torch.randn(B, S, generator=g), torch.rand(B, A, generator=g) * 2 - 1
    mu, std = solver.actor.mu_std(to_dev(obs))
    mu_o, std_o = O.policy_mu_std(w, "actor.", obs)
    assert_close(mu, mu_o, 1e-5, "mu_std.mu"); assert_close(std, std_o, 1e-5, "mu_std.std")
    qc = solver.constraint_critic
    m_o, s_o = O.qc_forward(w, "constraint_critic.", obs, act)
    assert_close(qc.forward_torch(to_dev(obs), to_dev(act)), m_o, 1e-5, "forward_torch mean")
    assert_close(qc.forward_torch(to_dev(obs), to_dev(act), uncertainty=True), m_o + qc.std_ratio * s_o, 1e-5, "forward_torch uncertainty")
    eps_qc = torch.randn(s_o.shape, generator=g)
    monkeypatch.setattr(torch, "randn_like", lambda x, *a, **k: to_dev(eps_qc).reshape(x.shape))
    m3, s3, smp = qc.forward_torch(to_dev(obs), to_dev(act), sample=True)
    assert_close(smp, m_o + eps_qc.clamp(-2, 2) * s_o, 1e-5, "forward_torch sample")
    # forward() with autograd inputs dispatches to the same differentiable form: gradient w.r.t. the action vs the oracle's
    a_req = to_dev(act).clone().requires_grad_(True)
    qc(to_dev(obs), a_req, uncertainty=True).sum().backward()
    a_o = act.clone().requires_grad_(True)
    mo, so = O.qc_forward(w, "constraint_critic.", obs, a_o)
    (mo + qc.std_ratio * so).sum().backward()
    assert_close(a_req.grad, a_o.grad, 1e-4, "d qc_ub / d action")
    # actor_loss: values and gradients of all three losses w.r.t. the actor / safe-actor / log_alpha
    eps = [torch.randn(B, A, generator=g), torch.randn(B, A, generator=g)]
    tape = [to_dev(e) for e in eps]
    monkeypatch.setattr(torch, "randn_like", lambda x, *a, **k: tape.pop(0))
    monkeypatch.setattr(random, "choice", lambda seq: seq[1])
    losses = solver.actor_loss(to_dev(obs))
    wo = {k: v.clone().requires_grad_(k.startswith(("actor.", "actor_safe."))) for k, v in w.items()}
    la = torch.tensor(0.0, requires_grad=True)
    want, _ = O.actor_losses(wo, obs, eps, O.SSACHyper(), la, 1, C, -float(A))
    for got, wt, name in zip(losses, want, ("actor", "alpha", "safe")):
        assert_close(got, wt, 1e-5, f"actor_loss[{name}]")
    (losses[0] + losses[2]).backward()
    (want[0] + want[2]).backward()
    sd = dict(solver.named_parameters())
    for k in ("actor.net.0.weight", "actor.net.4.bias", "actor_safe.net.2.weight"):
        assert_close(sd[k].grad, wo[k].grad, 2e-4, f"grad {k}", max_outlier_frac=0.01)


def test_ensemble_loss_mirrors_match_oracle():
    S, A = 11, 2
    w = O.make_ensemble_weights(91, S, A)
    ens = make_ensemble(w, S, A)
    g = torch.Generator().manual_seed(92)
    E, B = 7, 40
    s, a = torch.randn(E * B + 3, S, generator=g), torch.rand(E * B + 3, A, generator=g) * 2 - 1         # + a remainder that is dropped
    t = torch.cat([s + 0.05 * torch.randn(E * B + 3, S, generator=g), torch.randn(E * B + 3, 1, generator=g)], -1)
    assert_close(ens.compute_loss(to_dev(s), to_dev(a), to_dev(t)), O.ensemble_compute_loss(w, s, a, t), 1e-5, "compute_loss")
    rb = lambda x: x[:E * B].reshape(E, B, -1)
    assert_close(ens._mse_loss(to_dev(rb(s)), to_dev(rb(a)), to_dev(rb(t))), O.ensemble_mse_loss(w, rb(s), rb(a), rb(t)), 1e-5, "_mse_loss")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.timeout(900)
def test_two_rank_run_equals_single_rank_on_hardware():
    """torchrun x2 over NCCL: bench.py's multi_rank_check - every rank's rollout shard bit-identical to its block of the single-rank
    run; critic replicas started from DIFFERENT weights bit-identical after 3 data-parallel updates and equal to the single-process
    result up to fp32 summation order."""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "1",
           "--batch", "65536", "--skip-critic", "--skip-cpu", "--skip-extra"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=800, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    chk = out["multi_rank_check"]
    assert chk["pass"] and chk["rollout_shard_rows_bit_identical_to_single_rank"] and chk["critic_params_bit_identical_across_ranks"], chk
    assert out["scaling"] == "strong" and out["n_gpus"] == 2
