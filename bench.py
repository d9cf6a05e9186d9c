#!/usr/bin/env python
"""bench.py -- KS env control-periods/s on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...      # the reference's CPU path

A "step" is one control period (250 RK4 sub-steps) of every env of the workload:
``configs[1]`` of BASELINE.json at N=1 -- 4096 batched KS envs, default grid (N=64, L=22, 4 jets),
fp64, random actions -- and the same 4096 envs PER GPU at N>1 (weak scaling; each rank owns a
contiguous env shard, and the per-period NCCL all-gather of obs/reward/flags is inside the timed
region).  One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.

Timed regions
* ``value``: K x ``ks_step`` with actions already resident in HBM, CUDA events around every
  step on the launching stream, L2 flushed between timed steps, max over ranks.
* ``e2e``: K x ``KSVecEnv.step(numpy actions)`` -- the call a user of the gym API makes -- host
  buffers in, host buffers out, synchronised every step (wall clock around synchronous calls); the
  host->device and device->host traffic happens inside the kernel (zero-copy over PCIe).
* ``cpu_baseline`` / ``--impl reference``: the oracle's NumPy/SciPy port of the reference's
  ``step`` (same third-party calls the reference makes) on every host core.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ks_control_periods_per_s"
UNIT = "control-periods/s"
ENVS_PER_GPU = 4096
FLOPS_PER_POINT_SUBSTEP = 191          # un-merged stencils, mul/add = 1 flop, FMA = 2
FP64_NOMINAL_TFLOPS = 37.2             # 148 SM x 64 lanes x 2 x 1.965 GHz
# Spectral ETDRK4 mode (extra leg, not the headline): algorithmic flops per env per ETDRK4 step at
# N = 64 -- 8 complex 64-point FFTs per PAIR of envs at the textbook 5 N log2 N (15360) + nonlinear
# term (1536) + stage combinations (2176) + reward (128) = 19200 per pair = 9600 per env.
ETD_FLOPS_PER_ENV_STEP = 9600
ETD_DT, ETD_STEPS = 0.025, 10          # 10 x 0.025 = the reference's 0.25 time units per control period


# -------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (oracle port; the ONLY place bench.py touches oracle/)
# -------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process = one env stepped with the reference-style NumPy/SciPy port, single-threaded."""
    seed, periods, warm = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import numpy as np
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass
    from oracle import ks_numpy as ko

    cfg = ko.KSConfig()
    F = ko.forcing_matrix(cfg)
    rng = np.random.default_rng(seed)
    u = rng.uniform(-0.4, 0.4, (1, cfg.N))
    acts = rng.uniform(-1, 1, (warm + periods, 1, cfg.J)).astype(np.float32)
    for k in range(warm):
        u, _ = ko.step(cfg, u, ko.forcing(acts[k], F), rhs_fn=ko.rhs_scipy)
    t0 = time.perf_counter()
    for k in range(warm, warm + periods):
        u, _ = ko.step(cfg, u, ko.forcing(acts[k], F), rhs_fn=ko.rhs_scipy)
    return time.perf_counter() - t0, float(np.abs(u).max())


def cpu_port_throughput(periods_per_proc: int, procs: int | None = None, warm: int = 1):
    """Aggregate control-periods/s of ``procs`` independent single-env processes (the reference's
    own parallelism is one process per env, mbrl.py:81-86)."""
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(1000 + i, periods_per_proc, warm) for i in range(procs)])
    wall = time.perf_counter() - t0
    slowest = max(r[0] for r in res)
    return procs * periods_per_proc / slowest, procs, wall


def cpu_c_port_throughput(envs: int = 64, periods: int = 4):
    """The plain-C oracle on all cores (context: how fast a compiled CPU port is)."""
    import numpy as np
    from oracle import ks_c, ks_numpy as ko

    cfg = ko.KSConfig()
    rng = np.random.default_rng(0)
    u = rng.uniform(-0.4, 0.4, (envs, cfg.N))
    phi = np.zeros((envs, cfg.N), np.float32)
    u, _ = ks_c.step(cfg, u, phi)
    t0 = time.perf_counter()
    for _ in range(periods):
        u, _ = ks_c.step(cfg, u, phi)
    dt = time.perf_counter() - t0
    return envs * periods / dt, ks_c.num_threads()


def workload_name(B, world, N, L, J, S, dt, precision):
    return (f"{B} KS envs per GPU x {world} GPU(s) = {B * world} envs, N={N} L={L} J={J}, cfg_steps={S} RK4 sub-steps "
            f"per control period, dt={dt}, {precision}, random actions (BASELINE.json configs[1] per GPU)")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warm = max(1, args.steps), max(0, args.warmup)
    procs = os.cpu_count() or 1
    # bounded sample: each "step" = one control period of one env per host core
    per_proc = min(steps, 40)
    thr, procs, wall = cpu_port_throughput(per_proc, procs, warm=min(warm, 2))
    line = {
        "impl": "reference", "metric": METRIC, "value": thr, "unit": UNIT, "n_gpus": args.gpus,
        "steps": per_proc, "warmup": min(warm, 2), "ms_per_step": 1e3 * procs / thr,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.envs_per_gpu, max(1, args.gpus), 64, 22.0, 4, 250, 0.001, "f64"),
                   "sample": f"bounded sample of that workload: {procs} of its envs, one per host core (the reference's "
                             f"process-per-env parallelism, mbrl.py:81-86), {per_proc} control periods each"},
        "cpu_baseline": {"value": thr, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": f"{procs} processes x {per_proc} control periods of 1 env each, NumPy/SciPy port "
                                   "of the reference step (scipy.ndimage.convolve1d stencils, per-sub-step reward)"},
        "e2e": {"value": thr, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smmax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smmax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the observed power range
        thr = (max(power) + min(power)) / 2 if len(power) > 1 else 0.0
        loaded = [c for c, p in zip(sm, power) if p >= thr] or sm
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(smmax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def spectral_leg(B, K, W, device, fp64_peak, flush):
    """K control periods of the ETDRK4 solver (solver="etdrk4", dt=0.025 x 10 steps), inputs resident
    in HBM, one CUDA-event pair per step, L2 flushed between steps.  NOT the reference's scheme --
    reported next to the headline, never instead of it."""
    import numpy as np
    import torch

    from model_based_pde_control_b200 import KSVecEnv

    dev = torch.device("cuda", device)
    env = KSVecEnv(B, dict(dt=ETD_DT, cfg_steps=ETD_STEPS), device=device, solver="etdrk4")
    rng = np.random.default_rng(77)
    env.set_state(rng.uniform(-0.4, 0.4, (B, env.N)), 0)
    env.rollout_device(None, K=40, outputs=False)
    env.set_state(None, 0)
    actions = torch.as_tensor(rng.uniform(-1, 1, (W + K, B, env.J)).astype(np.float32)).to(dev)
    stream = torch.cuda.current_stream(dev)
    for k in range(W):
        env.step_device(actions[k])
    torch.cuda.synchronize(dev)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    n0 = env.launch_count
    for k in range(K):
        flush.zero_()
        starts[k].record(stream)
        env.step_device(actions[W + k])
        stops[k].record(stream)
    torch.cuda.synchronize(dev)
    ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops)) / K
    launches = env.launch_count - n0
    bad = bool(env.nonfinite().any())
    info = env.launch_info()
    env.close()
    tf = ETD_FLOPS_PER_ENV_STEP * ETD_STEPS * B / (ms * 1e-3) / 1e12
    return {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "gpu_launches": int(launches), "nonfinite": bad,
            "config": {"workload": f"{B} KS envs, N=64 L=22 J=4, solver=etdrk4 (pseudo-spectral ETDRK4, 2/3 dealiasing), "
                                   f"dt={ETD_DT} x {ETD_STEPS} steps per control period, f64, random actions",
                       "note": "different discretisation from the reference (FD-RK4): validated against oracle/ks_etdrk4.py "
                               "at 1e-10 and against the reference only statistically / by convergence",
                       "layout": info},
            "roofline": {"bound": "fp64", "kernel": "ks_etd_kernel", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": tf / fp64_peak, "flops_per_launch": ETD_FLOPS_PER_ENV_STEP * ETD_STEPS * B,
                         "flops_model": "9600 per env per ETDRK4 step (8 FFTs per env pair at 5 N log2 N + pointwise)"}}


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from model_based_pde_control_b200 import KSVecEnv, _lib
    from model_based_pde_control_b200.sharding import connect_fused_gather, gather_packed

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    json_out = sys.stdout
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly ONE JSON line.  NCCL's C code prints its version banner straight to
        # file descriptor 1 when NCCL_DEBUG is set, so fd 1 is pointed at stderr for the whole run and
        # the JSON line goes to a private duplicate of the original stdout.
        sys.stdout.flush()
        json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    B = args.envs_per_gpu
    K, W = args.steps, max(3, args.warmup)
    spectral = args.solver == "etdrk4"          # non-default: the spectral solver as the timed workload
    env_cfg = dict(dt=ETD_DT, cfg_steps=ETD_STEPS) if spectral else {}
    env = KSVecEnv(B, env_cfg, device=local_rank, precision=args.precision, points_per_lane=args.points_per_lane,
                   solver=args.solver)
    N, J, S = env.N, env.J, env.cfg_steps
    total_envs = B * world

    # synthetic workload: seeded ICs, short device burn-in onto the attractor,
    # random actions for every timed period, all resident in HBM before timing starts
    rng = np.random.default_rng(1000 + rank)
    env.set_state(rng.uniform(-0.4, 0.4, (B, N)), 0)
    env.rollout_device(None, K=args.burnin, outputs=False)
    env.set_state(None, 0)
    actions = torch.as_tensor(rng.uniform(-1, 1, (W + K, B, J)).astype(np.float32)).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    fields = env.packed_fields()
    fused = world > 1 and args.gather == "fused"
    gather_note = None
    if fused:
        # CUDA-IPC handles exchanged once; afterwards no collective call per period.  If peer mapping is
        # not possible on this box (ranks in different IPC namespaces, no P2P), every rank switches to
        # the NCCL all-gather together and the JSON line says so -- the exchange is never skipped.
        try:
            connect_fused_gather(env)         # raises on every rank if any rank fails
        except Exception as exc:              # noqa: BLE001 - reported in the JSON line
            fused = False
            args.gather = "nccl"
            gather_note = f"fused exchange unavailable ({type(exc).__name__}: {exc}); NCCL all-gather used"
            print(gather_note, file=sys.stderr)

    def one_step(k):
        if fused:
            return env.step_gather(actions[k])   # kernel epilogue stores into every peer's buffer + handshake
        out = env.step_device(actions[k])
        if world > 1 and args.gather == "nccl":
            out = gather_packed(out["packed"], fields, B)
        return out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)            # before the warm-up, so that the GPU does not idle between warm-up and timing
    if world > 1:
        dist.barrier()             # nobody enters the first exchange while rank 0 is still sleeping
    for k in range(W):
        one_step(k)
    torch.cuda.synchronize(dev)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    launches0 = env.launch_count
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    wall0 = time.perf_counter()
    align = torch.zeros(1, device=dev)
    for k in range(K):
        flush.zero_()                      # L2 flush between timed steps (outside the event pair)
        if world > 1:
            # the 256 MiB memsets do not take equally long on every GPU; re-align the ranks on the
            # device (stream-ordered 4-byte all-reduce, outside the event pair) so that a timed step
            # is the period + exchange, not the previous flush's skew
            dist.all_reduce(align)
        starts[k].record(stream)
        one_step(W + k)
        stops[k].record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = env.launch_count - launches0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    total_ms = sum(step_ms)
    if os.environ.get("KS_BENCH_DEBUG"):
        ss = sorted(step_ms)
        print(f"rank {rank} timed steps: median {ss[len(ss) // 2]:.4f} ms, min {ss[0]:.4f}, max {ss[-1]:.4f}, "
              f"first 3 {[round(x, 4) for x in step_ms[:3]]}", file=sys.stderr)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    flags_bad = bool(env.nonfinite().any())
    if fused and env.gather_timed_out():
        raise SystemExit("fused gather: a peer never signalled (handshake timed out)")

    # ---- e2e: the gym-facing host API, host buffers, copies inside the timed region ----
    acts_host = rng.uniform(-1, 1, (W + K, B, 1, J)).astype(np.float32)
    env.set_state(None, 0)
    for k in range(W):
        obs, rew, term, trunc, info = env.step(acts_host[k])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    step_s = []
    t0 = time.perf_counter()
    for k in range(K):
        t1 = time.perf_counter()
        obs, rew, term, trunc, info = env.step(acts_host[W + k])
        step_s.append(time.perf_counter() - t1)
    e2e_s = time.perf_counter() - t0
    if os.environ.get("KS_BENCH_DEBUG"):
        ss = sorted(step_s)
        print(f"e2e steps: median {1e3 * ss[len(ss) // 2]:.4f} ms, min {1e3 * ss[0]:.4f}, max {1e3 * ss[-1]:.4f}, "
              f"5 slowest {[round(1e3 * x, 3) for x in ss[-5:]]}, first 5 {[round(1e3 * x, 3) for x in step_s[:5]]}",
              file=sys.stderr)
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_envs * K / e2e_s

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel ----
    kernel_ms = total_ms / K            # N=1: the event pair brackets exactly one kernel launch
    value = total_envs * K / (total_ms * 1e-3)
    flops_per_launch = (ETD_FLOPS_PER_ENV_STEP * S * B) if spectral else FLOPS_PER_POINT_SUBSTEP * N * S * B
    achieved_tf = flops_per_launch / (kernel_ms * 1e-3) / 1e12
    lib = _lib.load()
    best, mean = ctypes.c_double(), ctypes.c_double()
    rc = lib.ks_bench_fp64_peak(local_rank, 20000, 5, ctypes.byref(best), ctypes.byref(mean))
    fp64_peak = best.value if rc == 0 and best.value > 0 else FP64_NOMINAL_TFLOPS
    peak_src = "self-measured DFMA micro-kernel (ks_bench_fp64_peak, best of 5)" if rc == 0 else "nominal"
    bound, kernel_peak, nominal = "fp64", fp64_peak, FP64_NOMINAL_TFLOPS
    if args.precision == "f32":        # optional fp32 mode: the FP32 FMA pipes bound it (no self-measured figure)
        bound, kernel_peak, nominal = "fp32", 2 * FP64_NOMINAL_TFLOPS, 2 * FP64_NOMINAL_TFLOPS
        peak_src = "nominal 148 SM x 128 lanes x 2 x 1.965 GHz"
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    esz = 8 if args.precision == "f64" else 4
    bytes_per_launch = (2 * esz * N + 4 * N + 4 * J + 16) * B     # 20N+4J+16 per env (fp64)
    hbm_achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9

    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f).get(f"{args.precision}/envs{B}/P{env.launch_info()['points_per_lane']}")
        if t:
            traffic, traffic_src = t["dram_bytes_read"] + t["dram_bytes_write"], t["source"]
    except Exception:
        pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {
            "workload": workload_name(B, world, N, env.L, J, S, env.dt, args.precision)
                        + (" -- NON-DEFAULT solver=etdrk4 (pseudo-spectral ETDRK4, not the reference's scheme)" if spectral else ""),
            "envs_per_gpu": B, "total_envs": total_envs, "N": N, "J": J, "cfg_steps": S,
            "l2": "flushed (256 MiB memset) between timed steps, outside the per-step event pairs"
                  + ("; ranks re-aligned after each flush by a 4-byte all-reduce, also outside the pairs" if world > 1 else ""),
            "collective": "none (N=1)" if world == 1 else (
                "fused: the period kernel's epilogue stores the packed obs/reward/step/truncated/flags block into every "
                "peer's gather buffer over NVLink (CUDA-IPC peer stores) + one-warp epoch handshake, timed"
                if fused else ("one NCCL all-gather of the packed obs/reward/step/truncated/flags block per period, timed"
                               if args.gather == "nccl" else "NONE (diagnostic run: every rank keeps its shard to itself)")),
            "layout": env.launch_info(),
            **({"collective_note": gather_note} if gather_note else {}),
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": env.h2d_bytes_per_step,
                "d2h_bytes_per_step": env.d2h_bytes_per_step, "ms_per_step": 1e3 * e2e_s / K,
                "api": "KSVecEnv.step(numpy actions) -> ks_step_host: one launch + sync; the kernel reads the pinned actions over PCIe "
                       "and mirrors the packed outputs into the pinned host block (KS_HOST_IO=copy: H2D copy, kernel, D2H copy)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": bound, "kernel": "ks_etd_kernel" if spectral else "ks_period_kernel", "achieved": achieved_tf,
            "peak": kernel_peak,
            "unit": "TFLOP/s", "frac": achieved_tf / kernel_peak, "traffic": traffic,
            "traffic_unit": "bytes per launch (dram read+write, ncu --set full)", "traffic_source": traffic_src,
            "peak_source": peak_src, "peak_nominal": nominal, "frac_of_nominal": achieved_tf / nominal,
            "flops_per_launch": flops_per_launch,
            "flops_model": ("9600 per env per ETDRK4 step (8 FFTs per env pair at 5 N log2 N + pointwise)" if spectral
                            else "191*N*cfg_steps per env-period (SURVEY.md 8d)"),
            "kernel_ms": kernel_ms,
            "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                    "bytes_per_launch": bytes_per_launch, "peak_source": hbm_src},
        },
        "wall_s_timed_region": wall,
        "nonfinite": flags_bad,
    }

    # ---- extra leg: the spectral ETDRK4 solver on the same batch (device-resident, same timing rules) ----
    if world == 1 and not args.no_spectral and not spectral:
        try:
            line["spectral_mode"] = spectral_leg(B, K, W, local_rank, fp64_peak, flush)
        except Exception as exc:
            line["spectral_mode"] = {"error": str(exc)}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only; bounded sample) ----
    if world == 1 and not args.no_cpu_baseline:
        thr, procs, cpu_wall = cpu_port_throughput(args.cpu_periods)
        line["cpu_baseline"] = {
            "value": thr, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": f"{procs} processes x {args.cpu_periods} control periods of 1 env each (default grid), NumPy/SciPy "
                      f"port of the reference step; {cpu_wall:.1f} s wall"}
        try:
            cthr, cthreads = cpu_c_port_throughput()
            line["cpu_baseline_c"] = {"value": cthr, "unit": UNIT, "cores": cthreads, "kind": "port",
                                      "sample": "plain-C oracle (oracle/ks_oracle.c, -O2, pthreads), 64 envs x 4 periods"}
        except Exception as exc:   # the C oracle is optional context
            line["cpu_baseline_c"] = {"error": str(exc)}
    print(json.dumps(line), file=json_out, flush=True)
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--points-per-lane", type=int, default=0)
    ap.add_argument("--burnin", type=int, default=40, help="device burn-in periods before timing")
    ap.add_argument("--cpu-periods", type=int, default=20, help="control periods per host core for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl", "none"],
                    help="N>1: how every rank gets the full batch each period (default: fused peer stores)")
    ap.add_argument("--solver", default="fd_rk4", choices=["fd_rk4", "etdrk4"],
                    help="timed solver; the default is the reference's scheme (the headline), etdrk4 is the spectral mode")
    ap.add_argument("--no-spectral", action="store_true", help="skip the extra spectral-ETDRK4 leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
