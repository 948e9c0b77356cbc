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
# dram__bytes_read.sum + dram__bytes_write.sum of one rollout_step_umma_kernel launch at the bench workload (1 M rows), from the
# `ncu --set full` capture of the final round-1 kernel (profiles/r1_ncu_full_rollout_step_1M_final_raw.csv: 52.9 MB read + 17.2 MB
# written - 48 MB of states in; most of the 60 MB of outputs is still in L2 when the kernel ends); null when no capture exists
TRAFFIC_BYTES_PER_LAUNCH = {"quadrotor": 70104064}
# the same for one critic_fused_kernel launch at B = 65 536, tracking dims (profiles/r1_ncu_full_critic_fused_64k_raw.csv:
# 44.2 MB read + 378.2 MB written: the saved bf16 activations the dW kernel consumes)
CRITIC_TRAFFIC_BYTES_PER_LAUNCH = {("tracking", 65536): 422400000}


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
            # measured to cost ~1 ms per rollout step): one sample early in the region, then one per second
            self._stop.wait(float(os.environ.get("DRPO_BENCH_CLOCK_INTERVAL", "1.0")))

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
    alg = build_alg(workload, B0, device, precision)
    alg.shard_rank, alg.shard_world = rank, world          # weak scaling: every rank rolls out its own B0 start states
    init_host = synthetic.make_start_states(workload, B0, 4354 + rank).pin_memory()
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
    while time.perf_counter() - t_pre < 0.5:
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
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            view = step_device()
            total += view.step_counts[-1]
        ev1.record()
        barrier()
    launches = lib.drpo_launch_count() - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=device, dtype=torch.float64)
    tot = total.clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms_total = float(ms)
    transitions = int(tot)
    value = transitions / (ms_total * 1e-3)

    # ---- end-to-end timing through the public API with host inputs ------------------------------------------------
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    tot_e = 0
    t_ev0, t_ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_ev0.record()
    for _ in range(args.steps):
        tot_e += int(step_e2e()[-1])
    t_ev1.record()
    barrier()
    ms_e = torch.tensor([t_ev0.elapsed_time(t_ev1)], device=device, dtype=torch.float64)
    te = torch.tensor([tot_e], device=device, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
        dist.all_reduce(te)
    e2e_value = int(te) / (float(ms_e) * 1e-3)

    # ---- roofline of the dominant kernel: rollout_step_umma_kernel, one launch per rollout step.  Its launches are bracketed
    #      by CUDA events on the launching stream (drpo_timing_enable) over K more steps of the same workload ----------------
    pk = peaks()
    step_ms = ms_total / args.steps
    import ctypes
    rows_launched, kt, kn, ksat = 0, ctypes.c_double(0.0), ctypes.c_int64(0), ctypes.c_double(0.0)
    lib.drpo_timing_enable(1)
    for _ in range(args.steps):
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
                    "frac": round(achieved_tf / pk["tensor_sustained"], 5), "traffic": TRAFFIC_BYTES_PER_LAUNCH.get(workload),
                    "peak_source": pk["src"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                    "kernel": "rollout_step_umma_kernel (policy + ensemble-member GEMM chain + epilogues), one launch per rollout step",
                    "avg_launch_ms": round(avg_ms, 4), "launches_timed": int(kn.value), "kernel_share_of_step": round(kt.value / (step_ms * args.steps), 3),
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
        per_gpu_tr_per_step = transitions / world / args.steps
        achieved_tf = flops_per_transition(S, A) * per_gpu_tr_per_step / (step_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": round(achieved_tf, 3), "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                    "frac": round(achieved_tf / pk["tensor_sustained"], 5), "traffic": None, "peak_source": pk["src"] + " (sustained)",
                    "kernel": "whole rollout step (fp32 path: GEMM + elementwise kernel chain)",
                    "algorithmic_flops_per_transition": flops_per_transition(S, A)}

    out = {
        "metric": "model_rollout_transitions_per_s", "value": value, "unit": "transitions/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"{workload} DRPO rollout: {B0} start states/GPU x horizon {HORIZON}, 7x4x200 ensemble + 256x2 actor",
                   "state_dim": S, "action_dim": A, "con_dim": C, "start_states_per_gpu": B0, "horizon": HORIZON,
                   "precision": args.precision, "parallelism": f"dp{world} (start states sharded, no collective)",
                   "l2": "per-step working set (states in + records out) exceeds the 126 MB L2"},
        "transitions_per_step": transitions / args.steps,
        "e2e": {"value": e2e_value, "unit": "transitions/s", "h2d_bytes_per_step": int(init_host.numel() * 4),
                "d2h_bytes_per_step": int((HORIZON + 1) * 4)},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": roofline,
    }

    _lib.check_kernel_status("bench rollout")        # a pipeline-protocol time-out inside the fused kernel would invalidate the numbers
    # ---- SSAC critic updates/s (second half of the metric) -------------------------------------------------------------
    if not args.skip_critic:
        out["critic"] = bench_critic(args, device, world, rank, pk)
    # ---- ensemble training iterations/s (BatchedGaussianEnsemble.fit's loop body): replicas only ----------
    if not args.skip_critic:
        out["ensemble_fit"] = bench_ensemble_fit(args, alg, workload, device, world == 1 and not args.skip_cpu)
    # ---- safety shield latency: latency-bound, rank 0's replica only ------------------------------------
    if not args.skip_critic:
        out["shield"] = bench_shield(alg, workload, device, world == 1 and not args.skip_cpu)
    # ---- CPU baseline (oracle port) on rank 0, N=1 only --------------------------------------------------------------
    if world == 1 and not args.skip_cpu:
        out["cpu_baseline"] = cpu_rollout_baseline(workload, args.cpu_batch, reps=5)
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
        res[name] = {"metric": f"ssac_{name}_updates_per_s", "value": k / (float(ms2) * 1e-3), "unit": "updates/s", "global_batch": Bg,
                     "ms_per_update": float(ms2) / k, "gpu_launches": int(lib.drpo_launch_count() - l0),
                     "dtype": "f32" if cprec == "fp32" else "tf32 tensor-op GEMMs (cuBLAS), fp32 elementwise/optimizer"}
    if world == 1 and not args.skip_cpu:
        res["actor"]["cpu_baseline"] = cpu_actor_baseline(args.cpu_critic_batch)
    return res


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs: the oracle port of the reference's algorithm (the reference itself is Python and cannot travel to the box)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_rollout_baseline(workload, B0, reps):
    from oracle import drpo_oracle as O
    from drpo_b200 import synthetic
    env_name, S, A, C = synthetic.WORKLOADS[workload]
    spec = {"quadrotor": O.env_quadrotor, "cartpole-move": O.env_cartpole, "point-robot": O.env_point_robot,
            "safetygym-point-synthetic": O.env_safetygym60}[workload]()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    wm, ws = synthetic.make_ensemble_weights(64578, S, A), synthetic.make_ssac_weights(219803, S, A, C)
    init = synthetic.make_start_states(workload, B0, 4354)
    g = torch.Generator().manual_seed(1)
    eps_p, eps_m = torch.randn(HORIZON, B0, A, generator=g), torch.randn(HORIZON, B0, S + 1, generator=g)
    members = [i % 5 for i in range(HORIZON)]
    O.rollout(ws, wm, spec, init[:2000], HORIZON, eps_p[:, :2000], eps_m[:, :2000], members)        # warm-up
    t0, n = time.perf_counter(), 0
    for _ in range(reps):
        res, counts, _ = O.rollout(ws, wm, spec, init, HORIZON, eps_p, eps_m, members)
        n += sum(counts)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "transitions/s", "cores": threads, "kind": "port",
            "sample": f"{reps} x oracle rollout of {B0} start states x horizon {HORIZON} ({workload} dims), torch CPU fp32, {threads} threads, {dt:.1f}s"}


def cpu_critic_baseline(B):
    from oracle import drpo_oracle as O
    from drpo_b200 import synthetic
    _, S, A, C = synthetic.WORKLOADS[CRITIC_WORKLOAD]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    w = synthetic.make_ssac_weights(43567, S, A, C)
    batch = synthetic.make_critic_batch(CRITIC_WORKLOAD, B, 49283)
    g = torch.Generator().manual_seed(2)
    noise = (torch.randn(B, A, generator=g), torch.randn(B, A, generator=g), torch.randn(B, generator=g))
    adam, hp = O.AdamState(), O.SSACHyper(std_ratio=1.0)
    O.critic_update(w, batch, noise, hp, 0.0, adam, 3e-4)
    t0, reps = time.perf_counter(), 3
    for _ in range(reps):
        O.critic_update(w, batch, noise, hp, 0.0, adam, 3e-4)
    dt = (time.perf_counter() - t0) / reps
    return {"value": (B / dt) / CRITIC_B, "unit": "updates/s (64k-sample equivalents)", "samples_per_s": B / dt, "cores": threads,
            "kind": "port", "sample": f"{reps} x oracle critic_update at B={B} (tracking dims), torch CPU fp32 autograd, {threads} threads"}


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
    from oracle import drpo_oracle as O
    from drpo_b200 import synthetic
    _, S, A, C = synthetic.WORKLOADS[CRITIC_WORKLOAD]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    w = synthetic.make_ssac_weights(43567, S, A, C)
    obs = synthetic.make_critic_batch(CRITIC_WORKLOAD, B, 49283)[0]
    g = torch.Generator().manual_seed(3)
    noise = (torch.randn(B, A, generator=g), torch.randn(B, A, generator=g))
    la, hp = torch.tensor(0.0), O.SSACHyper(std_ratio=1.0)
    adams = {k: O.AdamState() for k in ("actor", "alpha", "safe")}
    lrs = dict(actor=8e-5, alpha=8e-5, safe=8e-5)
    O.actor_update(w, obs, noise, hp, la, 0, C, -float(A), adams, lrs)
    t0, reps = time.perf_counter(), 3
    for _ in range(reps):
        O.actor_update(w, obs, noise, hp, la, 0, C, -float(A), adams, lrs)
    dt = (time.perf_counter() - t0) / reps
    return {"value": (B / dt) / CRITIC_B, "unit": "updates/s (64k-sample equivalents)", "samples_per_s": B / dt, "cores": threads,
            "kind": "port", "sample": f"{reps} x oracle actor_update at B={B} (tracking dims), torch CPU fp32 autograd, {threads} threads"}


def run_reference(args):
    """The reference's algorithm on the host cores: the oracle port (the reference is pure Python + torch; its files cannot
    travel to the GPU box, oracle/drpo_oracle.py restates it and is pinned to it by tests/golden)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from drpo_b200 import synthetic
    workload = args.workload
    _, S, A, C = synthetic.WORKLOADS[workload]
    B0 = args.cpu_batch
    for _ in range(max(args.warmup - 1, 0)):
        cpu_rollout_baseline(workload, min(B0, 5000), 1)
    t0 = time.perf_counter()
    res = cpu_rollout_baseline(workload, B0, max(args.steps, 1))
    dt = time.perf_counter() - t0
    out = {"impl": "reference", "metric": "model_rollout_transitions_per_s", "value": res["value"], "unit": "transitions/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{workload} DRPO rollout: bounded CPU sample of {B0} start states x horizon {HORIZON} per step",
                      "state_dim": S, "action_dim": A, "con_dim": C, "horizon": HORIZON},
           "cpu_baseline": res,
           "e2e": {"value": res["value"], "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    ap.add_argument("--skip-critic", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
