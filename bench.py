#!/usr/bin/env python
"""bench.py — DRPO hot path on B200: model-rollout transitions/s (headline) and SSAC critic updates/s.

  python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --steps K --warmup W  # the reference's algorithm on the box's host cores (oracle port)

A "step" is one SMBPO.rollout over one batch of synthetic start states (B0 per GPU, horizon 10) through the
probabilistic ensemble; `value` counts the transitions all ranks wrote to their device replay buffers per second with
the start states already resident in HBM; `e2e` is the same metric through the public Python API (SMBPO.rollout) with the
start states in pinned HOST memory (H2D inside the timed region) and a D2H read of the per-step transition counts.
Prints exactly ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

HORIZON = 10
DEFAULT_B0 = {"quadrotor": 1_000_000, "cartpole-move": 100_000, "safetygym-point-synthetic": 400_000, "point-robot": 100_000}
CRITIC_WORKLOAD, CRITIC_B = "tracking", 65536
# dram__bytes_read.sum + dram__bytes_write.sum of one rollout_step_pipe_kernel launch at the bench workload (1 M rows), from the
# `ncu --set full` capture of the shipped round-2 kernel (profiles/r2_ncu_full_rollout_step_1M_raw.csv: 52.6 MB read + 19.7 MB
# written - 48 MB of states in; most of the 60 MB of outputs is still in L2 when the kernel ends); null when no capture exists
TRAFFIC_BYTES_PER_LAUNCH = {("quadrotor", 1_000_000): 52561152 + 19717632,
                            # profiles/r2_ncu_full_rollout_step_safetygym400k_raw.csv / ..._cartpole100k_raw.csv (the 100 k-row launch stays in L2)
                            ("safetygym-point-synthetic", 400_000): 124049408 + 218478080, ("cartpole-move", 100_000): 2475776}
# the same for one critic_fused_kernel launch at B = 65 536, tracking dims (profiles/r2_ncu_full_critic_fused_64k_final_raw.csv:
# 36.0 MB read + 366.6 MB written: the saved bf16 activations the dW kernel consumes)
CRITIC_TRAFFIC_BYTES_PER_LAUNCH = {("tracking", 65536): 36007424 + 366628608}
# dram__bytes_read.sum + dram__bytes_write.sum of one solver_fused_kernel launch (ncu --set full, profiles/r2_ncu_full_solver_actor_64k_raw.csv)
SOLVER_TRAFFIC_BYTES_PER_LAUNCH = {("actor", "tracking", 65536): 81275136 + 396089088}


def flops_per_transition(S, A):
    """2*[MACs_policy + MACs_member]."""
    return 2 * ((256 * S + 65536 + 512 * A) + (120000 + 200 * (S + A) + 400 * (S + 1)))


def flops_per_critic_sample(S, A, C):
    """2*(8Q + 2P + 5T + 9H)."""
    Q = (S + A) * 256 + 65792
    P = 256 * S + 65536 + 512 * A
    T = (S + A) * 256 + 65536
    H = 65536 + 256 * C
    return 2 * (8 * Q + 2 * P + 5 * T + 9 * H)


def flops_per_multiplier_sample(S, A, C):
    """Dense multiply-adds x 2 of SSAC.update_multiplier (src/ssac.py:529-578) per sample: two policy forwards, two constraint-critic
    forwards (trunk + both heads), the lambda net forward and its backward (dX of the hidden layer, both dW, head)."""
    P = 256 * S + 65536 + 512 * A
    QC = (S + A) * 256 + 65536 + 2 * (65536 + 256 * C)
    L = (S + 1) * 256 + 65536 + 256
    return 2 * (2 * P + 2 * QC + 2 * L + 65536)


def flops_per_actor_sample(S, A, C):
    """Dense multiply-adds x 2 of SSAC.update_actor_and_alpha (src/ssac.py:458-527) per sample: two policy forwards, Q_k forward,
    three constraint-critic forwards, the lambda net forward; dX chains through Q_k and (twice) the constraint critic down to the
    action columns; head + hidden dX and both dW of the two policies."""
    P = 256 * S + 65536 + 512 * A
    Q = (S + A) * 256 + 65536 + 256
    QC = (S + A) * 256 + 65536 + 2 * (65536 + 256 * C)
    L = (S + 1) * 256 + 65536 + 256
    dQ = 256 + 65536 + 256 * A
    dQC = 2 * (256 * C + 65536) + 65536 + 256 * A
    dP = 2 * 512 * A + 65536 + 65536 + 256 * S
    return 2 * (2 * P + Q + 3 * QC + L + dQ + 2 * dQC + 2 * dP)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML is polled directly
    (a sample costs ~0.1 ms); nvidia-smi is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        flag = lambda name: "Active" if (r & getattr(n, name, 0)) else "Not Active"
        return [str(sm), str(mx), f"{pw:.2f}", flag("nvmlClocksEventReasonHwSlowdown") if hasattr(n, "nvmlClocksEventReasonHwSlowdown") else flag("nvmlClocksThrottleReasonHwSlowdown"),
                flag("nvmlClocksEventReasonHwThermalSlowdown") if hasattr(n, "nvmlClocksEventReasonHwThermalSlowdown") else flag("nvmlClocksThrottleReasonHwThermalSlowdown"),
                flag("nvmlClocksEventReasonSwThermalSlowdown") if hasattr(n, "nvmlClocksEventReasonSwThermalSlowdown") else flag("nvmlClocksThrottleReasonSwThermalSlowdown"),
                flag("nvmlClocksEventReasonSwPowerCap") if hasattr(n, "nvmlClocksEventReasonSwPowerCap") else flag("nvmlClocksThrottleReasonSwPowerCap")]

    def _run(self):
        self._stop.wait(0.02)              # let the host queue some launches first: one NVML query stalls the launching thread for ms
        while not self._stop.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            # every NVML query perturbs the run it observes (driver locks; a sample landing inside a 0.25 s timed region was
            # measured to cost ~1 ms per rollout step): one sample early in the region, then one every 100 ms (~1 % of the region)
            self._stop.wait(float(os.environ.get("DRPO_BENCH_CLOCK_INTERVAL", "0.1")))

    def sample_now(self):
        """One synchronous sample from the launching thread: called after the timed region's launches are queued and before the
        closing synchronisation, i.e. while the GPU is still executing them - guarantees a sample for regions shorter than the
        sampler thread's first wait (the strong-scaled shards at N = 8 run ~10 ms)."""
        try:
            if self.nvml is not None:
                self.rows.append(self._sample_nvml())
        except Exception:
            pass

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        num = lambda x: x.replace(".", "").isdigit()
        sm = sorted(float(r[0]) for r in self.rows if num(r[0]))
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if num(self.rows[0][1]) else None,
                "power_w_max": max(float(r[2]) for r in self.rows if num(r[2])) if self.rows else None,
                "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi", "reasons": reasons}


def build_alg(workload, B0, device, precision):
    import drpo_b200
    from drpo_b200 import synthetic
    env_name, S, A, C = synthetic.WORKLOADS[workload]
    cfg = drpo_b200.SMBPO.Config()
    cfg.rollout_batch_size, cfg.horizon = B0, HORIZON
    cfg.buffer_max = B0 * HORIZON + 1024           # SampleBuffer.extend needs batch <= capacity (src/sampling.py:131)
    alg = drpo_b200.SMBPO(cfg, drpo_b200.device_env(env_name), device=device)
    alg.model_ensemble.load_state_dict(synthetic.make_ensemble_weights(64578, S, A))
    alg.model_ensemble._elite_inds = [0, 1, 2, 3, 4]
    alg.solver.load_state_dict(synthetic.make_ssac_weights(219803, S, A, C), strict=False)
    alg.rollout_precision = precision
    return alg


def shard_rows(B_global, rank, world):
    """rank r rolls out rows [r*B0/N, (r+1)*B0/N) of the global start-state matrix."""
    B = B_global // world
    return rank * B, B


def measure_rollout(args, lib, workload, B_global, scaling, device, rank, world, precision, with_e2e=True, steps=None, prewarm=0.5):
    """Device-resident value, end-to-end value and the dominant kernel's roofline for one workload.  `scaling` = "strong": the
    B_global start states are split over the ranks (same global matrix, global trajectory ids) ; "weak": B_global per rank."""
    import ctypes
    import torch.distributed as dist
    import drpo_b200
    from drpo_b200 import _lib, synthetic
    _, S, A, C = synthetic.WORKLOADS[workload]
    steps = steps or args.steps
    if scaling == "strong":
        lo, B = shard_rows(B_global, rank, world)
        init_host = synthetic.make_start_states(workload, B_global, 4354)[lo:lo + B].contiguous().pin_memory()
    else:
        lo, B = rank * B_global, B_global
        init_host = synthetic.make_start_states(workload, B, 4354 + rank).pin_memory()
    alg = build_alg(workload, B, device, precision)
    alg.shard_rank, alg.shard_world = rank, world
    alg._traj_id_offset_override = lo                  # global id of this shard's row 0: Philox draws do not depend on N
    init_dev = init_host.to(device)
    members = [i % 5 for i in range(HORIZON)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_device():
        alg.virt_buffer._pointer.zero_()
        return alg.rollout(alg.actor, initial_states=init_dev, member_idx=members)

    def step_e2e():
        alg.virt_buffer._pointer.zero_()
        # H2D of this step's inputs (pinned) happens inside the public call: SMBPO.rollout streams host start states in row blocks
        # on a copy stream and the first step's kernel waits per block
        view = alg.rollout(alg.actor, initial_states=init_host, member_idx=members)
        return view.step_counts.to("cpu", non_blocking=False)                  # D2H of the step's result

    t_pre = time.perf_counter()                       # untimed pre-warm: bring the SM clocks up before the W warm-up steps
    while time.perf_counter() - t_pre < prewarm:
        step_device(); torch.cuda.synchronize()
    total = torch.zeros((), dtype=torch.int64, device=device)
    for _ in range(args.warmup):
        view = step_device()
        total += view.step_counts[-1]                  # every torch op of the timed loop runs once here: on a fresh box the first
    total.zero_()                                      # launch of a torch kernel pages its module in from disk (tens of ms)
    barrier()
    # ---- device-resident timing: K steps between two events ---------------------------------------------------
    launches0 = lib.drpo_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(device.index or 0) as clocks:
        barrier()
        ev0.record()
        for _ in range(steps):
            view = step_device()
            total += view.step_counts[-1]
        ev1.record()
        clocks.sample_now()
        barrier()
    launches = lib.drpo_launch_count() - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=device, dtype=torch.float64)
    tot = total.clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms_total, transitions = float(ms), int(tot)
    res = {"value": transitions / (ms_total * 1e-3), "ms_per_step": ms_total / steps, "transitions_per_step": transitions / steps,
           "gpu_launches": int(launches), "clocks": clocks.summary(), "rows_per_gpu": B, "steps": steps}

    # ---- end-to-end timing through the public API with host inputs ------------------------------------------------
    if with_e2e:
        for _ in range(min(args.warmup, 2)):
            step_e2e()
        barrier()
        tot_e = 0
        t_ev0, t_ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_ev0.record()
        for _ in range(steps):
            tot_e += int(step_e2e()[-1])
        t_ev1.record()
        barrier()
        ms_e = torch.tensor([t_ev0.elapsed_time(t_ev1)], device=device, dtype=torch.float64)
        te = torch.tensor([tot_e], device=device, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
            dist.all_reduce(te)
        res["e2e"] = {"value": int(te) / (float(ms_e) * 1e-3), "unit": "transitions/s", "h2d_bytes_per_step": int(init_host.numel() * 4),
                      "d2h_bytes_per_step": int((HORIZON + 1) * 4)}

    # ---- roofline of the dominant kernel: the fused rollout step kernel, one launch per rollout step.  Its launches are bracketed
    #      by CUDA events on the launching stream (drpo_timing_enable) over K more steps of the same workload ----------------
    pk = peaks()
    step_ms = ms_total / steps
    rows_launched, kt, kn, ksat = 0, ctypes.c_double(0.0), ctypes.c_int64(0), ctypes.c_double(0.0)
    lib.drpo_timing_enable(1)
    for _ in range(steps):
        v = step_device()
        rows_launched += int(v.step_counts[:-1].sum())               # alive rows of every step = rows each launch processed
    torch.cuda.synchronize()
    _lib.check(lib.drpo_timing_read(ctypes.byref(kt), ctypes.byref(kn), ctypes.byref(ksat)), "drpo_timing_read")
    lib.drpo_timing_enable(0)
    if kn.value > 0 and precision == drpo_b200.PREC_BF16:
        avg_ms = kt.value / kn.value
        flops_per_launch = flops_per_transition(S, A) * rows_launched / kn.value
        achieved_tf = flops_per_launch / (avg_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": round(achieved_tf, 3), "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                    "frac": round(achieved_tf / pk["tensor_sustained"], 5), "traffic": TRAFFIC_BYTES_PER_LAUNCH.get((workload, B)),
                    "peak_source": pk["src"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                    "kernel": "rollout_step kernel (policy + ensemble-member tcgen05 GEMM chain + epilogues), one launch per rollout step",
                    "avg_launch_ms": round(avg_ms, 4), "launches_timed": int(kn.value), "kernel_share_of_step": round(kt.value / (step_ms * steps), 3),
                    "algorithmic_flops_per_transition": flops_per_transition(S, A), "rows_per_launch": round(rows_launched / kn.value, 1)}
        # the step's HBM-bound satellites (hooks + ring store, order-preserving compaction): algorithmic bytes per row =
        # read 4(2S+A+1) + write record 4(2S+A+1+C)+2 + done flag 1 (store), read done 1 + 4S + id 4, write survivors 4S+4 (compaction)
        sat_bytes_row = (4 * (2 * S + A + 1) + 4 * (2 * S + A + 1 + C) + 2 + 1) + (1 + 4 * S + 4 + 4 * S + 4)
        if ksat.value > 0:
            sat_gbs = sat_bytes_row * rows_launched / (ksat.value * 1e-3) / 1e9
            roofline["elementwise"] = {"bound": "hbm", "kernels": "hooks_store_kernel (env hooks + ring store + block survivor counts) + compact_scatter_scan_kernel", "achieved": round(sat_gbs, 1),
                                       "peak": pk["hbm"], "unit": "GB/s", "frac": round(sat_gbs / pk["hbm"], 4),
                                       "avg_ms_per_step": round(ksat.value / kn.value, 4), "algorithmic_bytes_per_row": sat_bytes_row}
    else:
        per_gpu_tr_per_step = transitions / world / steps
        achieved_tf = flops_per_transition(S, A) * per_gpu_tr_per_step / (step_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": round(achieved_tf, 3), "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                    "frac": round(achieved_tf / pk["tensor_sustained"], 5), "traffic": None, "peak_source": pk["src"] + " (sustained)",
                    "kernel": "whole rollout step (fp32 path: GEMM + elementwise kernel chain)",
                    "algorithmic_flops_per_transition": flops_per_transition(S, A)}
    res["roofline"] = roofline
    _lib.check_kernel_status("bench rollout")        # a pipeline-protocol time-out inside the fused kernel would invalidate the numbers
    del alg
    return res


def multi_rank_check(device, rank, world, precision):
    """N-rank result == 1-rank result (SURVEY.md §8e "Parity across N"), run inside the scaling bench itself on the real GPUs / NCCL:
    (1) rollout: every rank rolls out its row shard of a 65 536-row global matrix AND the whole matrix; its shard's rows must be
        bit-identical to its block of the full run at every step (Philox draws keyed by global trajectory id, rows independent);
    (2) critic: 3 data-parallel updates from DIFFERENT initial weights per rank (the first step broadcasts rank 0's state) must leave
        bit-identical parameters on every rank, equal to 3 single-process updates on the full batch up to fp32 summation order."""
    import torch.distributed as dist
    import drpo_b200
    from drpo_b200 import synthetic
    out = {}
    # ---- (1) rollout shards ----
    workload, Bg, H = "quadrotor", 65536, 5
    _, S, A, C = synthetic.WORKLOADS[workload]
    init = synthetic.make_start_states(workload, Bg, 977)
    members = [3, 1, 4, 0, 2]

    def roll(rows_lo, rows_n):
        alg = build_alg(workload, rows_n, device, precision)
        alg.horizon = H
        alg._traj_id_offset_override = rows_lo
        view = alg.rollout(alg.actor, initial_states=init[rows_lo:rows_lo + rows_n].to(device), member_idx=members)
        counts = view.counts()
        rows = view.get(as_dict=True)
        return counts, rows
    lo, n = shard_rows(Bg, rank, world)
    c_shard, r_shard = roll(lo, n)
    c_full, r_full = roll(0, Bg)
    allc = [torch.zeros(H, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(allc, torch.tensor(c_shard, dtype=torch.int64, device=device))
    allc = torch.stack(allc).cpu()                                   # [world, H]
    ok = allc.sum(0).tolist() == c_full
    off_s = off_f = 0
    for t in range(H):
        before = int(allc[:rank, t].sum())
        for k, v in r_shard.items():
            ok = ok and torch.equal(v[off_s:off_s + c_shard[t]], r_full[k][off_f + before:off_f + before + c_shard[t]])
        off_s += c_shard[t]; off_f += c_full[t]
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["rollout_shard_rows_bit_identical_to_single_rank"] = bool(int(flag))
    out["rollout_counts_full"] = c_full
    # ---- (2) critic replicas ----
    wl, Bc = "tracking", 8192
    _, S, A, C = synthetic.WORKLOADS[wl]
    full = synthetic.make_critic_batch(wl, Bc, 8101)

    def make(seed, B):
        cfg = drpo_b200.SSAC.Config(); cfg.batch_size = B; cfg.constraint_critic_cfg.std_ratio = 1.0
        sv = drpo_b200.SSAC(cfg, S, A, C, HORIZON, 100, 1000, 10, 5.0, device=device)
        sv.load_state_dict(synthetic.make_ssac_weights(seed, S, A, C), strict=False)
        sv.precision = precision
        return sv
    Bl = Bc // world
    dp = make(500 + rank, Bl)                                        # rank-dependent start: sync_replicas must fix it
    dp.data_parallel = True
    one = make(500, Bc)
    sl = slice(rank * Bl, (rank + 1) * Bl)
    for _ in range(3):
        dp.update_critic(*[t[sl].to(device) for t in full])
        one.update_critic(*[t.to(device) for t in full])
    mine = torch.cat([dp._critic_arena, dp._target_arena, dp.critic_optimizer.m, dp.critic_optimizer.v])
    ref0 = mine.clone()
    dist.broadcast(ref0, src=0)
    same = torch.tensor([1 if torch.equal(ref0, mine) else 0], device=device)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out["critic_params_bit_identical_across_ranks"] = bool(int(same))
    diff = (dp._critic_arena - one._critic_arena).abs()
    lr = 3e-4
    out["critic_vs_single_rank"] = {"max_abs_param_diff": float(diff.max()), "frac_within_1e-5": float((diff <= 1e-5 * one._critic_arena.abs().clamp_min(1e-3)).float().mean()),
                                    "bound": "Adam's first steps move a weight by ~lr*sign(g): fp32 summation order can flip only weights with |g| ~ eps; <= 3 steps * lr",
                                    "updates": 3, "global_batch": Bc}
    ok2 = float(diff.max()) <= 3.5 * lr and out["critic_vs_single_rank"]["frac_within_1e-5"] > 0.98
    out["pass"] = bool(out["rollout_shard_rows_bit_identical_to_single_rank"] and out["critic_params_bit_identical_across_ranks"] and ok2)
    if not out["pass"]:
        raise RuntimeError(f"multi-rank parity check failed on rank {rank}: {out}")
    return out


def run_ours(args):
    # stdout carries exactly one JSON line: libraries that chat on fd 1 (NCCL prints its version banner there) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch.distributed as dist
    import drpo_b200
    from drpo_b200 import _lib, synthetic
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun)"
    lib = _lib.load()
    precision = {"fp32": drpo_b200.PREC_FP32, "bf16": drpo_b200.PREC_BF16}[args.precision]
    workload = args.workload
    _, S, A, C = synthetic.WORKLOADS[workload]
    B0 = args.batch or DEFAULT_B0[workload]
    scaling = args.scaling
    if scaling == "strong" and B0 % world:
        raise SystemExit(f"--scaling strong needs the {B0} start states to divide by {world} ranks")

    # ---- headline: the BASELINE split: B0 GLOBAL start states, B0/N rows per rank, no collective --------------
    head = measure_rollout(args, lib, workload, B0, scaling, device, rank, world, precision)
    rows = head["rows_per_gpu"]
    out = {
        "metric": "model_rollout_transitions_per_s", "value": head["value"], "unit": "transitions/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"{workload} DRPO rollout: {B0 if scaling == 'strong' else B0 * world} start states in total "
                               f"({rows} per GPU) x horizon {HORIZON}, 7x4x200 ensemble + 256x2 actor",
                   "state_dim": S, "action_dim": A, "con_dim": C, "start_states_global": B0 if scaling == "strong" else B0 * world,
                   "start_states_per_gpu": rows, "horizon": HORIZON,
                   "precision": args.precision, "parallelism": f"dp{world} (start-state rows [r*B0/N, (r+1)*B0/N) per rank, global trajectory ids, no collective)",
                   "l2": "per-step working set (states in + records out) exceeds the 126 MB L2" if rows * (8 * S + 4 * A + 12) > 126e6 else
                         "inputs are rewritten every step by the previous step's kernels; the ring (records out) is larger than one step's working set"},
        "transitions_per_step": head["transitions_per_step"],
        "e2e": head["e2e"],
        "gpu_launches": head["gpu_launches"],
        "clocks": head["clocks"],
        "roofline": head["roofline"],
    }
    # ---- second field under N > 1: the weak split (B0 per GPU), short -----------------------------------------------------------
    if world > 1 and scaling == "strong" and not args.skip_extra:
        wk = measure_rollout(args, lib, workload, B0, "weak", device, rank, world, precision, with_e2e=False, steps=min(args.steps, 10), prewarm=0.1)
        out["weak"] = {"scaling": "weak", "value": wk["value"], "unit": "transitions/s", "ms_per_step": wk["ms_per_step"],
                       "start_states_per_gpu": B0, "roofline_frac": wk["roofline"]["frac"]}
    # ---- N > 1: N-rank == 1-rank parity on the real GPUs / NCCL -----------------------------------------------------------------
    if world > 1 and not args.skip_check:
        out["multi_rank_check"] = multi_rank_check(device, rank, world, precision)
    # ---- the other BASELINE rollout configs (cartpole 100 k x 10, safetygym-60 400 k x 10), N = 1 only ---------------------------
    if world == 1 and not args.skip_extra:
        out["other_workloads"] = {}
        for wl in ("cartpole-move", "safetygym-point-synthetic"):
            if wl == workload:
                continue
            r = measure_rollout(args, lib, wl, DEFAULT_B0[wl], "strong", device, rank, world, precision, steps=min(args.steps, 20), prewarm=0.1)
            _, s2, a2, c2 = synthetic.WORKLOADS[wl]
            out["other_workloads"][wl] = {"value": r["value"], "unit": "transitions/s", "ms_per_step": r["ms_per_step"],
                                          "start_states": DEFAULT_B0[wl], "horizon": HORIZON, "state_dim": s2, "action_dim": a2, "con_dim": c2,
                                          "transitions_per_step": r["transitions_per_step"], "e2e": r["e2e"], "roofline": r["roofline"],
                                          "gpu_launches": r["gpu_launches"]}
    # ---- SSAC critic updates/s (second half of the metric) -------------------------------------------------------------
    if not args.skip_critic:
        out["critic"] = bench_critic(args, device, world, rank, peaks())
    alg = None
    if not args.skip_critic and world == 1:
        alg = build_alg(workload, 4096, device, precision)
        # ---- ensemble training iterations/s (BatchedGaussianEnsemble.fit's loop body): replicas only ----------
        out["ensemble_fit"] = bench_ensemble_fit(args, alg, workload, device, not args.skip_cpu)
        # ---- safety shield latency: latency-bound, rank 0's replica only ------------------------------------
        out["shield"] = bench_shield(alg, workload, device, not args.skip_cpu)
    # ---- CPU baseline on rank 0, N=1 only: the reference itself when it is staged, else the oracle port -------------------------
    if world == 1 and not args.skip_cpu:
        out["cpu_baseline"] = cpu_rollout_baseline(workload, args.cpu_batch, reps=3)
        if not args.skip_extra:
            out["reference_on_gpu"] = reference_on_gpu(workload, B0)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(json_fd, 1)
    if rank == 0:
        print(json.dumps(out), flush=True)


def bench_critic(args, device, world, rank, pk):
    import torch.distributed as dist
    import drpo_b200
    from drpo_b200 import _lib, synthetic
    _, S, A, C = synthetic.WORKLOADS[CRITIC_WORKLOAD]
    Bg = CRITIC_B
    B = Bg // world                                     # strong scaling: the 64k minibatch is sharded over ranks
    cfg = drpo_b200.SSAC.Config()
    cfg.batch_size = B
    cfg.constraint_critic_cfg.std_ratio = 1.0
    solver = drpo_b200.SSAC(cfg, S, A, C, HORIZON, 100, 1000, 10, 5.0, device=device)
    solver.load_state_dict(synthetic.make_ssac_weights(43567, S, A, C), strict=False)
    cprec = os.environ.get("DRPO_BENCH_CRITIC_PRECISION", args.precision)       # fp32 | bf16 (fused tcgen05 kernels) | tf32 (cuBLAS)
    solver.precision = {"fp32": drpo_b200.PREC_FP32, "bf16": drpo_b200.PREC_BF16, "tf32": drpo_b200.PREC_TF32}[cprec]
    solver.data_parallel = world > 1
    full = synthetic.make_critic_batch(CRITIC_WORKLOAD, Bg, 49283)
    batch = [t[rank * B:(rank + 1) * B].to(device) for t in full]
    lib = _lib.load()
    for _ in range(3):
        solver.update_critic(*batch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n = args.critic_steps
    l0 = lib.drpo_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        lq, lc = solver.update_critic(*batch)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ups = n / (float(ms) * 1e-3)
    fl = flops_per_critic_sample(S, A, C) * Bg
    ach = fl * ups / 1e12 / world
    res = {"metric": "ssac_critic_updates_per_s", "value": ups, "unit": "updates/s", "samples_per_s": ups * Bg,
           "global_batch": Bg, "per_gpu_batch": B, "scaling": "strong", "ms_per_update": float(ms) / n,
           "dtype": {"bf16": "bf16 (fused tcgen05 forward/loss/dX kernel + tcgen05 dW kernel), fp32 accumulate/optimizer",
                     "tf32": "tf32 tensor-op GEMMs (cuBLAS), fp32 elementwise/optimizer", "fp32": "f32"}[cprec],
           "gpu_launches": int(lib.drpo_launch_count() - l0), "loss_q": float(lq), "loss_c": float(lc),
           "roofline": {"bound": "tensor", "achieved": round(ach, 3), "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                        "frac": round(ach / pk["tensor_sustained"], 5),
                        "traffic": CRITIC_TRAFFIC_BYTES_PER_LAUNCH.get((CRITIC_WORKLOAD, B)) if cprec == "bf16" else None,
                        "kernel": "critic_fused_kernel (traffic) ; achieved = whole update incl. dW, reductions, clip/Adam/EMA",
                        "algorithmic_flops_per_sample": flops_per_critic_sample(S, A, C)}}
    if world == 1 and not args.skip_cpu:
        res["cpu_baseline"] = cpu_critic_baseline(args.cpu_critic_batch)
    # ---- the other two SSAC steps of update_solver, reported separately: actor/alpha/safe-actor and multiplier -----
    obs = batch[0]
    for name, fn in (("actor", lambda: solver.update_actor_and_alpha(obs)), ("multiplier", lambda: solver.update_multiplier(obs))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = lib.drpo_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(max(n // 2, 1)):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms2 = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        k = max(n // 2, 1)
        ups2 = k / (float(ms2) * 1e-3)
        fl2 = (flops_per_actor_sample if name == "actor" else flops_per_multiplier_sample)(S, A, C)
        ach2 = fl2 * Bg * ups2 / 1e12 / world
        res[name] = {"metric": f"ssac_{name}_updates_per_s", "value": ups2, "unit": "updates/s", "global_batch": Bg,
                     "ms_per_update": float(ms2) / k, "gpu_launches": int(lib.drpo_launch_count() - l0),
                     "dtype": {"bf16": "bf16 (fused tcgen05 forward/loss/dX kernel + tcgen05 dW kernel), fp32 accumulate/optimizer",
                               "tf32": "tf32 tensor-op GEMMs (cuBLAS), fp32 elementwise/optimizer", "fp32": "f32"}[cprec],
                     "roofline": {"bound": "tensor", "achieved": round(ach2, 3), "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                                  "frac": round(ach2 / pk["tensor_sustained"], 5),
                                  "traffic": SOLVER_TRAFFIC_BYTES_PER_LAUNCH.get((name, CRITIC_WORKLOAD, B)) if cprec == "bf16" else None,
                                  "kernel": "solver_fused_kernel ; achieved = whole update incl. dW, reductions, clip/Adam",
                                  "algorithmic_flops_per_sample": fl2}}
    if world == 1 and not args.skip_cpu:
        res["actor"]["cpu_baseline"] = cpu_actor_baseline(args.cpu_critic_batch)
    return res


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs: the UNMODIFIED reference (staged under baseline/_ref by oracle/stage_reference.py, or /root/reference) driven by
# oracle/ref_runner.py - kind "reference"; if neither exists, the oracle port (oracle/drpo_oracle.py) - kind "port".
# ---------------------------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(len(os.sched_getaffinity(0)), 1)
    except AttributeError:
        return os.cpu_count() or 1


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _reference_rollout_callable(workload, B0):
    """callable() -> transitions of one rollout of B0 start states x HORIZON on the host cores, and its kind."""
    from drpo_b200 import synthetic
    env_name, S, A, C = synthetic.WORKLOADS[workload]
    wm, ws = synthetic.make_ensemble_weights(64578, S, A), synthetic.make_ssac_weights(219803, S, A, C)
    init = synthetic.make_start_states(workload, B0, 4354)
    from oracle import ref_runner
    if ref_runner.available():
        return ref_runner.rollout_runner(workload, S, A, C, B0, HORIZON, wm, ws, init), "reference"
    from oracle import drpo_oracle as O
    spec = {"quadrotor": O.env_quadrotor, "cartpole-move": O.env_cartpole, "point-robot": O.env_point_robot,
            "safetygym-point-synthetic": O.env_safetygym60}[workload]()
    g = torch.Generator().manual_seed(1)
    eps_p, eps_m = torch.randn(HORIZON, B0, A, generator=g), torch.randn(HORIZON, B0, S + 1, generator=g)
    members = [i % 5 for i in range(HORIZON)]
    return (lambda: sum(O.rollout(ws, wm, spec, init, HORIZON, eps_p, eps_m, members)[1])), "port"


def _time_rollouts(run, reps, threads):
    torch.set_num_threads(threads)
    run()                                                   # full-size warm-up: first-touch of the 1e6-slot buffers, MKL thread pool
    best, n = None, 0
    t_all = time.perf_counter()
    for _ in range(reps):
        t0 = time.perf_counter()
        n = run()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n / best, time.perf_counter() - t_all


def cpu_rollout_baseline(workload, B0, reps):
    """SMBPO.rollout (src/smbpo.py:229-249) of the reference on the box's host cores: all threads (`value`) and the reference's own
    setting of 4 threads (src/cli.py:108).  Best of `reps` after a full-size warm-up (the first run on a fresh box is 2x slower)."""
    run, kind = _reference_rollout_callable(workload, B0)
    threads = host_threads()
    v_all, t1 = _time_rollouts(run, reps, threads)
    v_4, t2 = _time_rollouts(run, reps, min(4, threads))
    torch.set_num_threads(threads)
    what = "the reference's own SMBPO.rollout (unmodified src/, eager torch CPU fp32 + numpy hooks)" if kind == "reference" else "oracle port rollout (torch CPU fp32)"
    return {"value": v_all, "unit": "transitions/s", "cores": threads, "kind": kind, "cpu": _cpu_model(), "host_cpus": os.cpu_count(),
            "threads_4": {"value": v_4, "cores": min(4, threads), "note": "torch.set_num_threads(4): the reference's own setting (src/cli.py:108)"},
            "sample": f"best of {reps} x {what} of {B0} start states x horizon {HORIZON} ({workload} dims) per thread setting, {t1 + t2:.1f}s"}


def reference_on_gpu(workload, B0):
    """Second, clearly labelled baseline: the UNMODIFIED reference's SMBPO.rollout in eager PyTorch ON THE B200 (cuBLAS GEMMs, one
    ATen launch per op, 3 device->host->device hook round trips per step) at the bench's own config.  Runs in a child process
    because the reference binds its device at import (src/torch_util.py:9)."""
    code = r"""
import json, sys, time, torch
sys.path.insert(0, %r)
from oracle import ref_shim
if not ref_shim.reference_available():
    print(json.dumps({"unavailable": "reference not staged"})); sys.exit(0)
ref_shim.import_reference(device="cuda")
from oracle import ref_runner
from drpo_b200 import synthetic
wl, B0, H = %r, %d, %d
_, S, A, C = synthetic.WORKLOADS[wl]
wm, ws = synthetic.make_ensemble_weights(64578, S, A), synthetic.make_ssac_weights(219803, S, A, C)
init = synthetic.make_start_states(wl, B0, 4354).cuda()
run = ref_runner.rollout_runner(wl, S, A, C, B0, H, wm, ws, init, device="cuda")
run(); torch.cuda.synchronize()
best = None
for _ in range(3):
    t0 = time.perf_counter(); n = run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
print(json.dumps({"value": n / best, "unit": "transitions/s", "ms_per_step": best * 1e3, "transitions_per_step": n,
                  "kind": "reference, eager PyTorch on the B200 (not the graded CPU baseline)", "start_states": B0, "horizon": H}))
""" % (ROOT, workload, B0, HORIZON)
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
        lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else {"unavailable": (r.stderr or "no output")[-300:]}
    except Exception as e:                                   # a reported baseline: its failure must not void the measurement
        return {"unavailable": repr(e)[:300]}


def _cpu_update_baseline(which, B):
    """One SSAC update step of the reference at minibatch B (tracking dims) on the host cores: which = 0 critic, 1 actor, 2 multiplier."""
    from drpo_b200 import synthetic
    _, S, A, C = synthetic.WORKLOADS[CRITIC_WORKLOAD]
    threads = host_threads()
    w = synthetic.make_ssac_weights(43567, S, A, C)
    batch = synthetic.make_critic_batch(CRITIC_WORKLOAD, B, 49283)
    from oracle import ref_runner
    name = ("update_critic", "update_actor_and_alpha", "update_multiplier")[which]
    if ref_runner.available():
        fn, kind = ref_runner.critic_runner(S, A, C, B, w, batch, std_ratio=1.0)[which], "reference"
        what = f"the reference's own SSAC.{name} (unmodified src/ssac.py, torch CPU fp32 autograd)"
    else:
        from oracle import drpo_oracle as O
        g = torch.Generator().manual_seed(2)
        hp = O.SSACHyper(std_ratio=1.0)
        if which == 0:
            noise = (torch.randn(B, A, generator=g), torch.randn(B, A, generator=g), torch.randn(B, generator=g))
            adam = O.AdamState()
            fn = lambda: O.critic_update(w, batch, noise, hp, 0.0, adam, 3e-4)
        elif which == 1:
            noise = (torch.randn(B, A, generator=g), torch.randn(B, A, generator=g))
            la, adams = torch.tensor(0.0), {k: O.AdamState() for k in ("actor", "alpha", "safe")}
            fn = lambda: O.actor_update(w, batch[0], noise, hp, la, 0, C, -float(A), adams, dict(actor=8e-5, alpha=8e-5, safe=8e-5))
        else:
            eps, adam = torch.randn(B, A, generator=g), O.AdamState()
            fn = lambda: O.multiplier_update(w, batch[0], eps, hp, C, adam, 3e-4)
        kind, what = "port", f"oracle {name} (torch CPU fp32 autograd)"
    res = {}
    for th in (threads, min(4, threads)):
        torch.set_num_threads(th)
        fn()
        best = None
        for _ in range(3):
            t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        res[th] = best
    torch.set_num_threads(threads)
    dt = res[threads]
    return {"value": (B / dt) / CRITIC_B, "unit": "updates/s (64k-sample equivalents)", "samples_per_s": B / dt, "cores": threads, "kind": kind,
            "threads_4": {"value": (B / res[min(4, threads)]) / CRITIC_B, "cores": min(4, threads)},
            "sample": f"best of 3 x {what} at B={B} (tracking dims) per thread setting"}


def cpu_critic_baseline(B):
    return _cpu_update_baseline(0, B)


def bench_ensemble_fit(args, alg, workload, device, with_cpu):
    """One training iteration of the dynamics ensemble at the reference's batch (7 members x 256 rows, src/dynamics.py:66)."""
    from drpo_b200 import _lib, synthetic
    lib = _lib.load()
    ens = alg.model_ensemble
    _, S, A, C = synthetic.WORKLOADS[workload]
    n = ens.ensemble_size * ens.batch_size
    g = torch.Generator().manual_seed(11)
    s = torch.randn(n, S, generator=g); a = torch.rand(n, A, generator=g) * 2 - 1
    t = torch.cat([s + 0.05 * torch.randn(n, S, generator=g), torch.randn(n, 1, generator=g)], dim=1)
    sd, ad, td = s.to(device), a.to(device), t.to(device)
    for _ in range(5):
        ens.train_step(sd, ad, td)
    torch.cuda.synchronize()
    k, l0 = 100, lib.drpo_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        ens.train_step(sd, ad, td)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    res = {"metric": "ensemble_train_iterations_per_s", "value": 1e3 / ms, "unit": "iterations/s", "ms_per_iteration": ms,
           "rows": n, "gpu_launches": int(lib.drpo_launch_count() - l0), "dtype": "f32", "scaling": "replicas only"}
    if with_cpu:
        from oracle import drpo_oracle as O
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        w = synthetic.make_ensemble_weights(64578, S, A)
        adam = O.AdamState()
        O.ensemble_train_step(w, s, a, t, adam)
        t0, reps = time.perf_counter(), 10
        for _ in range(reps):
            O.ensemble_train_step(w, s, a, t, adam)
        dt = (time.perf_counter() - t0) / reps
        res["cpu_baseline"] = {"value": 1.0 / dt, "unit": "iterations/s", "cores": threads, "kind": "port",
                               "sample": f"{reps} x oracle ensemble_train_step at {n} rows ({workload} dims), torch CPU fp32 autograd, {threads} threads"}
    return res


def bench_shield(alg, workload, device, with_cpu):
    """Latency of one shielded action selection: the training step's switch on one state (src/smbpo.py:124-136) and the
    evaluation sampler's linear shield on the 10 evaluation envs (src/sampling.py:420-439, N_EVAL_TRAJ = 10)."""
    from drpo_b200 import _lib, synthetic
    lib = _lib.load()
    _, S, A, C = synthetic.WORKLOADS[workload]
    solver = alg.solver
    g = torch.Generator().manual_seed(12)
    s1, s10 = torch.randn(1, S, generator=g), torch.randn(10, S, generator=g)
    d1, d10 = s1.to(device), s10.to(device)
    cases = {"train_step_1_state": lambda: solver.shield_act(d1, eval=False, shield_type="safe", safe_shield_threshold=-0.1, uncertainty=True),
             "eval_linear_10_envs": lambda: solver.shield_act(d10, eval=True, shield_type="linear", safe_shield_threshold=-0.05)}
    res = {"metric": "shielded_action_latency", "unit": "us/call", "higher_is_better": False, "dtype": "f32", "scaling": "replicas only"}
    for name, fn in cases.items():
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        k, l0 = 200, lib.drpo_launch_count()
        t0 = time.perf_counter()
        for _ in range(k):
            a = fn()
        a.cpu()                                        # the env needs the action on the host: the read-back is part of the latency
        dt = (time.perf_counter() - t0) / k
        res[name] = {"value": dt * 1e6, "gpu_launches_per_call": (lib.drpo_launch_count() - l0) / k}
    if with_cpu:
        from oracle import drpo_oracle as O
        w = synthetic.make_ssac_weights(64578, S, A, C)
        torch.set_num_threads(4)                       # the reference's own setting for this host-side path (src/cli.py:108)
        eps = torch.randn(1, A, generator=g)
        for name, fn in (("train_step_1_state", lambda: O.shield_actions(w, s1, C, "safe", -0.1, eps_perf=eps, uncertainty=True)),
                         ("eval_linear_10_envs", lambda: O.shield_actions(w, s10, C, "linear", -0.05))):
            with torch.no_grad():
                fn()
                t0, reps = time.perf_counter(), 50
                for _ in range(reps):
                    fn()
            res[name]["cpu_baseline"] = {"value": (time.perf_counter() - t0) / reps * 1e6, "unit": "us/call", "cores": 4, "kind": "port",
                                         "sample": f"{reps} x oracle shield_actions, torch CPU fp32, 4 threads"}
        torch.set_num_threads(os.cpu_count() or 1)
    return res


def cpu_actor_baseline(B):
    return _cpu_update_baseline(1, B)


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path (SMBPO.rollout, unmodified, from baseline/_ref or
    /root/reference; the oracle port only if neither is present) on the box's host cores with every host thread, on our arm's
    metric / unit; each step is a bounded sample of the workload (--cpu-batch start states x horizon 10), shrunk if K steps would
    not finish within a few minutes.  Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from drpo_b200 import synthetic
    workload = args.workload
    _, S, A, C = synthetic.WORKLOADS[workload]
    threads = host_threads()
    torch.set_num_threads(threads)
    B0 = args.cpu_batch
    run, kind = _reference_rollout_callable(workload, B0)
    t0 = time.perf_counter(); run(); first = time.perf_counter() - t0        # warm-up 1 (also sizes the sample)
    budget = 150.0
    if first * (args.steps + max(args.warmup - 1, 0)) > budget and B0 > 2000:
        B0 = max(2000, int(B0 * budget / (first * (args.steps + args.warmup))))
        run, kind = _reference_rollout_callable(workload, B0)
        run()
    for _ in range(max(args.warmup - 1, 0)):
        run()
    n, t0 = 0, time.perf_counter()
    for _ in range(max(args.steps, 1)):
        n += run()
    dt = time.perf_counter() - t0
    value = n / dt
    what = "the reference's own SMBPO.rollout (unmodified src/, eager torch CPU fp32 + numpy hooks)" if kind == "reference" else "oracle port rollout"
    res = {"value": value, "unit": "transitions/s", "cores": threads, "kind": kind, "cpu": _cpu_model(), "host_cpus": os.cpu_count(),
           "sample": f"{max(args.steps, 1)} x {what} of {B0} start states x horizon {HORIZON} ({workload} dims), {threads} threads, {dt:.1f}s"}
    out = {"impl": "reference", "metric": "model_rollout_transitions_per_s", "value": value, "unit": "transitions/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3,
           "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{workload} DRPO rollout: bounded CPU sample of {B0} start states x horizon {HORIZON} per step",
                      "state_dim": S, "action_dim": A, "con_dim": C, "horizon": HORIZON},
           "cpu_baseline": res,
           "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="quadrotor", choices=list(DEFAULT_B0))
    ap.add_argument("--batch", type=int, default=0, help="start states per GPU (default: the workload's BASELINE size)")
    ap.add_argument("--precision", default=os.environ.get("DRPO_BENCH_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--critic-steps", type=int, default=20)
    ap.add_argument("--cpu-batch", type=int, default=100_000)
    ap.add_argument("--cpu-critic-batch", type=int, default=16384)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default, SURVEY.md 8d/e): the workload's start states are split over the ranks; weak: that many per rank")
    ap.add_argument("--skip-check", action="store_true", help="skip the N-rank == 1-rank parity check of the multi-GPU run")
    ap.add_argument("--skip-extra", action="store_true", help="skip the secondary measurements (other workloads, weak split, reference on GPU)")
    ap.add_argument("--skip-critic", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
