"""drpo_ensemble_train_step (BatchedGaussianEnsemble.fit's loop, src/dynamics.py:143-189 - SURVEY.md section 8f "next" row 2)
through the C ABI: normaliser fit, three Adam iterations and the holdout ranking against the reference's golden vectors; a
larger batch against the oracle (loss, raw gradients of every tensor incl. the log-var bounds, updated parameters)."""
import numpy as np
import pytest
import torch

from oracle import drpo_oracle as O
from tests.util import assert_close, dev, make_ensemble, to_dev
from tests.test_gpu_parity import _golden_after

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,S,A", [("point_robot", 11, 2), ("quadrotor", 12, 2)])
def test_ensemble_fit_vs_golden(golden, tag, S, A):
    g = golden("ensemble_fit")
    w = O.make_ensemble_weights(int(g[f"{tag}.seed"]), S, A)
    ens = make_ensemble(w, S, A)
    states, actions, targets = to_dev(g[f"{tag}.states"]), to_dev(g[f"{tag}.actions"]), to_dev(g[f"{tag}.targets"])
    ens.state_normalizer.fit(states)
    assert_close(ens.state_normalizer.mean, g[f"{tag}.norm_mean"], 1e-5, "normaliser mean")
    assert_close(ens.state_normalizer.std, g[f"{tag}.norm_std"], 1e-5, "normaliser std")
    keys = [k for k in w if k.startswith(O.ENSEMBLE_TRAINABLE)]
    for it in range(3):
        idx = to_dev(g[f"{tag}.idx{it}"])
        loss = ens.train_step(states[idx], actions[idx], targets[idx])
        assert_close(loss, g[f"{tag}.losses"][it], 1e-5, f"compute_loss it{it}")
        _golden_after(ens.state_dict(), g, f"{tag}.it{it}", keys, 2e-5, f"it{it}")
    hold = to_dev(g[f"{tag}.holdout_idx"])
    hl = ens.holdout_losses(states[hold], actions[hold], targets[hold])
    assert_close(hl, g[f"{tag}.holdout_losses"], 2e-5, "holdout losses")
    assert torch.argsort(hl)[:5].tolist() == [int(x) for x in g[f"{tag}.elites"]]


# up to 512 rows per member all members run in one launch per layer (batched FFMA GEMM); above, member by member with split-K
# weight gradients and, in the tensor mode, TF32 tensor-op GEMMs: both paths, both precisions
@pytest.mark.parametrize("S,A,n,prec,tol", [(12, 2, 7 * 256 + 5, "fp32", 5e-5), (60, 2, 7 * 64, "fp32", 5e-5), (12, 2, 7 * 256, "tf32", 2e-2),
                                            (12, 2, 7 * 640, "fp32", 5e-5), (12, 2, 7 * 640, "tf32", 2e-2)])
def test_ensemble_train_step_vs_oracle(S, A, n, prec, tol):
    import drpo_b200
    w = O.make_ensemble_weights(81, S, A)
    ens = make_ensemble(w, S, A)
    ens.precision = {"fp32": drpo_b200.PREC_FP32, "tf32": drpo_b200.PREC_TF32}[prec]
    g = torch.Generator().manual_seed(82)
    states = torch.randn(n, S, generator=g); actions = torch.rand(n, A, generator=g) * 2 - 1
    targets = torch.cat([states + 0.05 * torch.randn(n, S, generator=g), torch.randn(n, 1, generator=g)], dim=1)
    wo = {k: v.clone() for k, v in w.items()}
    loss, aux = O.ensemble_train_step(wo, states, actions, targets, O.AdamState())
    got = ens.train_step(to_dev(states), to_dev(actions), to_dev(targets))
    assert_close(got, loss, max(tol, 2e-5), "compute_loss")
    views = ens.arena_views(ens.optimizer.grad)
    for k, want in aux["grads_raw"].items():
        # the bound gradients are the 0.01 regulariser plus a batch sum of terms scaled by 1 - sigmoid(.) ~ 1e-5: their fp32
        # rounding differs between the closed form here and autograd's chain; 1e-3 of the (regulariser-sized) scale
        t = max(tol, 1e-3) if k in ("min_log_var", "max_log_var") else tol
        assert_close(views[k], want, t, f"grad {k}", max_outlier_frac=5e-2 if prec == "fp32" else 2e-2)
        assert_close(views[k], want, 10 * t, f"grad {k} (bound)")
    if prec == "fp32":
        sd = ens.state_dict()
        for k in wo:
            if k.startswith(O.ENSEMBLE_TRAINABLE):
                assert_close(sd[k], wo[k], 2e-5, f"param {k}", max_outlier_frac=2e-3)


def test_fit_runs_and_ranks_elites():
    """fit(steps=...) end to end on a replay buffer: losses fall on a learnable target, elites are a permutation prefix."""
    import drpo_b200
    from drpo_b200.sampling import ConstraintSafetySampleBuffer
    S, A, C, n = 12, 2, 2, 4000
    ens = make_ensemble(O.make_ensemble_weights(5, S, A), S, A)
    g = torch.Generator().manual_seed(6)
    s = torch.randn(n, S, generator=g); a = torch.rand(n, A, generator=g) * 2 - 1
    ns = s + 0.1 * a.sum(-1, keepdim=True); r = s[:, 0] * 0.5
    buf = ConstraintSafetySampleBuffer(S, A, 8192, con_dim=C, device=dev())
    buf.extend(states=to_dev(s), actions=to_dev(a), next_states=to_dev(ns), rewards=to_dev(r), dones=to_dev(torch.zeros(n, dtype=torch.bool)),
               violations=to_dev(torch.zeros(n, dtype=torch.bool)), constraint_values=to_dev(torch.zeros(n, C)))
    losses = ens.fit(buf, steps=60)
    assert len(losses) == 60 and all(map(lambda x: x == x, losses)) and losses[-1] < losses[0]
    assert len(ens._elite_inds) == ens.num_elites and len(set(ens._elite_inds)) == ens.num_elites


def test_fit_epochs_form_matches_the_oracle_loop():
    """fit(epochs=...) (src/dynamics.py:186-194 -> epochal_training, src/train.py:58-101): ensemble_size x epochs shuffled passes
    with a ragged last batch (its remainder modulo the ensemble size is dropped); per-epoch mean losses against the oracle's
    train step run over the same permutations."""
    from drpo_b200.sampling import ConstraintSafetySampleBuffer
    S, A, C, n = 4, 1, 4, 2000                     # 2000 rows, batch 7 x 256 = 1792 -> batches of 1792 and 208 (= 7 x 29 + 5)
    w = O.make_ensemble_weights(15, S, A)
    ens = make_ensemble(w, S, A)
    g = torch.Generator().manual_seed(16)
    s = torch.randn(n, S, generator=g); a = torch.rand(n, A, generator=g) * 2 - 1
    ns = s + 0.1 * a.sum(-1, keepdim=True); r = s[:, 0] * 0.5
    buf = ConstraintSafetySampleBuffer(S, A, 4096, con_dim=C, device=dev())
    buf.extend(states=to_dev(s), actions=to_dev(a), next_states=to_dev(ns), rewards=to_dev(r), dones=to_dev(torch.zeros(n, dtype=torch.bool)),
               violations=to_dev(torch.zeros(n, dtype=torch.bool)), constraint_values=to_dev(torch.zeros(n, C)))
    calls = []
    torch.manual_seed(1234)
    losses = ens.fit(buf, epochs=1, post_step_callback=lambda e, b, nb: calls.append((e, b, nb)))
    assert len(losses) == ens.ensemble_size and len(calls) == 2 * ens.ensemble_size and calls[-1] == (ens.ensemble_size - 1, 1, 2)
    # the oracle over the same permutations (torch.randperm on the CPU generator, as the reference draws them)
    wo = {k: v.clone() for k, v in w.items()}
    O.normalizer_fit(wo, s)
    t = torch.cat([ns, r.unsqueeze(1)], dim=1)
    adam = O.AdamState()
    torch.manual_seed(1234)
    want = []
    for _ in range(ens.ensemble_size):
        perm = torch.randperm(n)
        ep = []
        for b0 in (0, 1792):
            idx = perm[b0:b0 + 1792]
            ep.append(float(O.ensemble_train_step(wo, s[idx], a[idx], t[idx], adam)[0]))
        want.append(float(np.mean(ep)))
    assert_close(torch.tensor(losses), torch.tensor(want), 1e-4, "per-epoch mean losses")
    with pytest.raises(ValueError):
        ens.fit(buf)
