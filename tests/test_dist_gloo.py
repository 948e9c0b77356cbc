"""world_size-2 CPU (gloo) tests of the N>1 HOST logic: the critic step's two-phase protocol with the gradient all-reduce,
and the rollout's row sharding.  There is no CPU kernel path in the product, so the C-ABI library is replaced here by a
stand-in that evaluates each phase with the CPU oracle (tests may use the oracle); what is under test is everything the
Python mirror does around the library calls: global-batch normalisation, phases 1/2, which buffers are all-reduced, the
global row-id offsets, and that every rank ends with identical parameters equal to the single-rank result."""
import ctypes
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import drpo_oracle as O

S, A, C, BG = 11, 2, 1, 96
PREFIXES = ("critic.", "constraint_critic.")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _OracleLib:
    """Stand-in for libdrpo_sm100.so: drpo_critic_step evaluated by the CPU oracle, phase by phase."""

    def __init__(self, solver, stash, fail_on_call=None):
        self.solver, self.stash, self.adam, self.calls = solver, stash, O.AdamState(), []
        self.fail_on_call = fail_on_call                  # index of the phase-1 call whose (emulated) watchdog fires on this rank

    def drpo_kernel_status_peek(self):
        return 0

    def drpo_critic_workspace_bytes(self, *a):
        return 1024

    def drpo_last_error(self):
        return b""

    def _names(self):
        return [k for k in self.solver.state_dict() if k.startswith(PREFIXES)]

    def drpo_critic_step(self, args):
        solver = self.solver
        self.calls.append((int(args.phases), int(args.batch_size), int(args.global_batch_size), int(args.row_id_offset)))
        sd = solver.state_dict()
        names = self._names()
        if args.phases & 1:
            w = {k: v.detach().clone() for k, v in sd.items()}
            for k in names:
                w[k].requires_grad_(True)
            lq, lc, _ = O.critic_losses(w, self.stash["batch"], self.stash["noise"], O.SSACHyper(), float(solver.log_alpha))
            (lq + lc).backward()
            scale = args.batch_size / args.global_batch_size          # losses / gradients are normalised by the GLOBAL batch
            gviews = solver.critic_arena_views(solver.critic_optimizer.grad)
            for k in names:
                gviews[k].copy_(w[k].grad * scale)
            solver._losses[0], solver._losses[1] = float(lq) * scale, float(lc) * scale
            n1 = sum(1 for c in self.calls if c[0] & 1)
            solver._losses[15] = 1.0 if self.fail_on_call is not None and n1 - 1 == self.fail_on_call else 0.0   # DRPO_LOSS_ERR_SLOT
        if args.phases & 2:
            if float(solver._losses[15]) != 0.0:             # the library applies no update when any rank's watchdog fired
                return 0
            gviews = solver.critic_arena_views(solver.critic_optimizer.grad)
            grads = {k: gviews[k].clone() for k in names}
            O.clip_grad_norm([grads[k] for k in names if k.startswith("critic.")], 5.0)
            O.clip_grad_norm([grads[k] for k in names if k.startswith("constraint_critic.")], 5.0)
            with torch.no_grad():
                O.adam_step({k: sd[k] for k in names}, grads, self.adam, float(args.adam.lr), 1e-4)      # in place: views of the arena
                O.update_ema(sd, sd, "critic_target.", "critic.", 0.005)
                O.update_ema(sd, sd, "constraint_critic_target.", "constraint_critic.", 0.005)
        return 0


def _make_inputs():
    g = torch.Generator().manual_seed(5)
    obs = torch.randn(BG, S, generator=g); act = torch.rand(BG, A, generator=g) * 2 - 1
    nobs = obs + 0.1 * torch.randn(BG, S, generator=g); rew = torch.randn(BG, generator=g)
    done = torch.rand(BG, generator=g) < 0.1; viol = torch.rand(BG, generator=g) < 0.1
    cv = torch.randn(BG, generator=g) - 0.5
    noise = (torch.randn(BG, A, generator=g), torch.randn(BG, A, generator=g), torch.randn(BG, generator=g))
    return [obs, act, nobs, rew, done, viol, cv], noise


def _worker(rank, world, port, out, veto=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import drpo_b200
        from drpo_b200 import _lib
        cfg = drpo_b200.SSAC.Config()
        cfg.batch_size = BG // world
        solver = drpo_b200.SSAC(cfg, S, A, C, 10, 100, 300, 10, 10.0, device="cpu")
        # every rank starts from DIFFERENT weights and a different host RNG: the first data-parallel step must broadcast rank 0's
        # state (sync_replicas), or the replicas would silently diverge
        solver.load_state_dict(O.make_ssac_weights(7 + 100 * rank, S, A, C), strict=False)
        solver._dp_rng.seed(1234 + rank)
        solver.data_parallel = world > 1
        batch, noise = _make_inputs()
        lo, hi = rank * (BG // world), (rank + 1) * (BG // world)
        stash = dict(batch=[t[lo:hi] for t in batch], noise=tuple(n[lo:hi] for n in noise))
        fake = _OracleLib(solver, stash, fail_on_call=(1 if (veto and rank == 1) else None))
        _lib.load = lambda: fake                                           # the stand-in library
        _lib.ptr = lambda t: None if t is None else t.data_ptr()           # (the real one insists on CUDA tensors)
        _lib.stream_ptr = lambda: None
        losses, snaps = [], []
        for _ in range(2):                                                 # two updates: Adam state and lr schedule carry over
            lq, lc = solver.update_critic(*stash["batch"], noise=stash["noise"])
            losses.append((float(lq), float(lc)))
            snaps.append(solver._critic_arena.clone())
        sd = {k: v.clone() for k, v in solver.state_dict().items() if k.startswith(("critic", "constraint_critic"))}
        picks = [solver._dp_rng.choice(range(2)) for _ in range(8)]        # the actor step's critic pick: same sequence on every rank
        if rank == 0:
            torch.save(dict(sd=sd, losses=losses, calls=fake.calls, second_update_moved=bool((snaps[1] != snaps[0]).any())), out)
        pk = torch.tensor(picks)
        gp = [torch.empty_like(pk) for _ in range(world)]
        dist.all_gather(gp, pk)
        assert all(torch.equal(gp[0], x) for x in gp), "critic picks differ across ranks"
        # every rank must hold bit-identical parameters after the step
        flat = torch.cat([v.reshape(-1) for v in sd.values()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], x) for x in gathered), "ranks diverged"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_critic_two_phase_allreduce_matches_single_rank(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    # single-rank oracle on the full batch, same two updates with the host's cosine schedule
    w = O.make_ssac_weights(7, S, A, C)
    batch, noise = _make_inputs()
    adam, lr, T = O.AdamState(), 3e-4, 100 * 300 * 10
    want_losses = []
    for t in range(2):
        lq, lc, _ = O.critic_update(w, batch, noise, O.SSACHyper(), 0.0, adam, lr)
        want_losses.append((float(lq), float(lc)))
        lr = O.cosine_lr(lr, t + 1, T, 8e-5, 3e-4)
    for (a, b), (c, d) in zip(got["losses"], want_losses):
        assert a == pytest.approx(c, rel=1e-5) and b == pytest.approx(d, rel=1e-5)
    for k, v in got["sd"].items():
        # Adam's first steps move every weight by ~lr * sign(g): summing the gradient over two shards instead of one batch
        # changes g in the last bits, which the update amplifies only where |g| ~ eps; bound the difference by a small
        # fraction of one step
        err = float((v - w[k]).abs().max())
        assert err <= 0.02 * 3e-4, (k, err)
    # protocol: per update phase 1 (local rows, global normaliser, global row offset of the shard) then phase 2
    assert [c[0] for c in got["calls"]] == [1, 2, 1, 2]
    assert all(c[1] == BG // 2 and c[2] == BG and c[3] == 0 for c in got["calls"])
    assert got["second_update_moved"]


@pytest.mark.timeout(300)
def test_watchdog_flag_on_one_rank_vetoes_the_step_on_every_rank(tmp_path):
    """The watchdog slot rides in the gradient all-reduce (tail of the gradient storage): a time-out reported by rank 1 during the
    second update leaves the parameters of EVERY rank untouched by that update, and the replicas stay identical."""
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, True), nprocs=2, join=True)
    got = torch.load(out)
    assert not got["second_update_moved"]


def test_rollout_sharding_offsets():
    """SMBPO.rollout on rank r of N hands the library its own rows with the GLOBAL id of row 0, so Philox noise (keyed by
    global trajectory id) does not depend on N; no collective is involved."""
    import drpo_b200
    from drpo_b200 import _lib
    cfg = drpo_b200.SMBPO.Config()
    cfg.buffer_max, cfg.horizon = 4096, 3
    alg = drpo_b200.SMBPO(cfg, drpo_b200.device_env("quadrotor"), device="cpu")
    alg.model_ensemble._elite_inds = [0, 1, 2, 3, 4]
    seen = {}

    class Fake:
        def drpo_kernel_status_peek(self):
            return 0

        def drpo_rollout_workspace_bytes(self, a):
            return 1024

        def drpo_rollout(self, a):
            seen.update(batch=int(a.batch), off=int(a.traj_id_offset), horizon=int(a.horizon),
                        members=[a.member_idx_host[i] for i in range(a.horizon)])
            return 0

    saved = (_lib.load, _lib.ptr, _lib.stream_ptr)
    fake = Fake()
    try:
        _lib.load = lambda: fake
        _lib.ptr = lambda t: None if t is None else t.data_ptr()
        _lib.stream_ptr = lambda: None
        alg.shard_rank, alg.shard_world = 3, 8
        alg.rollout(alg.actor, initial_states=torch.zeros(500, 12), member_idx=[4, 0, 2])
    finally:
        _lib.load, _lib.ptr, _lib.stream_ptr = saved
    assert seen == dict(batch=500, off=1500, horizon=3, members=[4, 0, 2])
