#!/usr/bin/env python
"""bench.py -- KS env control-periods/s on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...      # the reference's CPU path

A "step" is one control period (250 RK4 sub-steps) of every env of the workload.  Headline
(``value``): ``configs[1]`` of BASELINE.json at N=1 -- 4096 batched KS envs, default grid (N=64, L=22,
4 jets), fp64, random actions -- and the same 4096 envs PER GPU at N>1 (weak scaling; each rank owns
a contiguous env shard, the per-period exchange of obs/reward/flags is inside the timed region).
One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.

Timed regions of the GPU arm
* ``value``            K x ``ks_step`` (N>1: ``ks_step_gather``), actions resident in HBM, one CUDA-event
                       pair per step on the launching stream, L2 flushed between steps, max over ranks.
* ``sustained``        >= 2 s of the same step back to back (no flush, one event pair around all), with
                       its own nvidia-smi clock / power record -- the burst ``value`` next to a figure
                       the clocks had time to react to.
* ``e2e``              K x ``KSVecEnv.step(numpy actions)`` -- the call a user of the gym API makes --
                       host buffers in / out, synchronised every step.
* ``e2e_episode_amortised``  one whole 400-step episode through ``KSVecEnv.step`` INCLUDING the auto-reset
                       (800-period burn-in launch) that the last step triggers: the rate a collection
                       loop with the reference's reset semantics sustains.
* ``config_65536``     BASELINE configs[2]: 65 536 envs over the N GPUs (strong scaling: 65536/N per GPU).
* ``large_domain``     BASELINE configs[3] (N=1 only): N=256, L=88, 8 jets, 4096 envs, fp64 and fp32.
* ``spectral_mode``    (N=1 only) the ETDRK4 solver -- NOT the reference's scheme, never the headline.
* ``gather_verified``  (N>1) before anything is timed: the fused exchange's full-batch block is compared
                       bit for bit with an NCCL all-gather of a twin shard and with the single-GPU run
                       of this rank's shard; any difference aborts the run.
* ``cpu_baseline`` / ``--impl reference``: the UNMODIFIED reference ``KuramotoSivashinskyEnv.step``
  (``pdegym/kuramoto/kuramoto.py:78-98``, run from ``baseline/_ref`` through ``oracle/ref_loader.py``), one
  single-threaded process per host core -- ``kind: "reference"``; the oracle's NumPy/SciPy port
  (``kind: "port"``) only if the reference tree is not available.
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
TOTAL_ENVS_CONFIG2 = 65536
FLOPS_PER_POINT_SUBSTEP = 191          # un-merged stencils, mul/add = 1 flop, FMA = 2
FP64_NOMINAL_TFLOPS = 37.2             # 148 SM x 64 lanes x 2 x 1.965 GHz
FP32_NOMINAL_TFLOPS = 74.4             # 148 SM x 128 lanes x 2 x 1.965 GHz
# Spectral ETDRK4 mode (extra leg, not the headline): algorithmic flops per env per ETDRK4 step at
# N = 64 -- 8 complex 64-point FFTs per PAIR of envs at the textbook 5 N log2 N (15360) + nonlinear
# term (1536) + stage combinations (2176) + reward (128) = 19200 per pair = 9600 per env.
ETD_FLOPS_PER_ENV_STEP = 9600
ETD_DT, ETD_STEPS = 0.025, 10          # 10 x 0.025 = the reference's 0.25 time units per control period
LARGE = dict(N=256, L=88.0, J=8)       # BASELINE configs[3]: dx unchanged, 8 jets at k/8


# -------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (the ONLY place bench.py touches oracle/)
# -------------------------------------------------------------------------------------------------
def cpu_kind() -> str:
    """"reference" when the unmodified reference tree can be executed here (baseline/_ref travels to
    the GPU box), else "port" (the oracle's NumPy/SciPy restatement)."""
    try:
        from oracle import ref_loader
        return "reference" if ref_loader.reference_available() else "port"
    except Exception:
        return "port"


def _cpu_worker(args):
    """One process = one env, single-threaded: the reference's own ``env.step`` (kuramoto.py:78-98) or,
    if the reference tree is absent, the oracle port of it."""
    seed, periods, warm, kind = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import numpy as np
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass
    rng = np.random.default_rng(seed)
    u0 = rng.uniform(-0.4, 0.4, 64)
    acts = rng.uniform(-1, 1, (warm + periods, 1, 4)).astype(np.float32)
    if kind == "reference":
        from oracle import ref_loader

        env = ref_loader.make_reference_env()            # KuramotoSivashinskyEnv(), defaults = configs[0]
        env.u, env.timestep = u0, 0
        for k in range(warm):
            env.step(acts[k])
        t0 = time.perf_counter()
        for k in range(warm, warm + periods):
            env.step(acts[k])
        return time.perf_counter() - t0, float(np.abs(env.u).max())
    from oracle import ks_numpy as ko

    cfg = ko.KSConfig()
    F = ko.forcing_matrix(cfg)
    u = u0[None]
    for k in range(warm):
        u, _ = ko.step(cfg, u, ko.forcing(acts[k], F), rhs_fn=ko.rhs_scipy)
    t0 = time.perf_counter()
    for k in range(warm, warm + periods):
        u, _ = ko.step(cfg, u, ko.forcing(acts[k], F), rhs_fn=ko.rhs_scipy)
    return time.perf_counter() - t0, float(np.abs(u).max())


def cpu_throughput(periods_per_proc: int, procs: int | None = None, warm: int = 1, kind: str | None = None):
    """Aggregate control-periods/s of ``procs`` independent single-env processes (the reference's
    own parallelism is one process per env, mbrl.py:81-86)."""
    procs = procs or os.cpu_count() or 1
    kind = kind or cpu_kind()
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(1000 + i, periods_per_proc, warm, kind) for i in range(procs)])
    wall = time.perf_counter() - t0
    slowest = max(r[0] for r in res)
    return procs * periods_per_proc / slowest, procs, wall, kind


def cpu_sample_text(kind, procs, periods, warm):
    what = ("the UNMODIFIED reference KuramotoSivashinskyEnv.step (pdegym/kuramoto/kuramoto.py:78-98, executed from "
            "baseline/_ref under oracle/ref_loader.py's gym / pytorch_lightning stubs)" if kind == "reference" else
            "NumPy/SciPy port of the reference step (oracle/ks_numpy.py: scipy.ndimage.convolve1d stencils, per-sub-step "
            "reward); the reference tree was not found")
    return (f"{procs} single-threaded processes (one per host core; the reference's process-per-env parallelism, "
            f"mbrl.py:81-86) x {periods} timed control periods of 1 default-grid env each after {warm} warm-up periods: {what}")


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


def shared_config(B, world, N, L, J, S, dt, precision, gather):
    """The ``config`` object both arms print (identical for the same flags, so the two lines can be
    matched key by key); statements that concern one arm only say so."""
    return {
        "workload": workload_name(B, world, N, L, J, S, dt, precision),
        "envs_per_gpu": B, "total_envs": B * world, "N": N, "J": J, "cfg_steps": S,
        "l2": "GPU arm: L2 flushed (256 MiB memset) between timed steps, outside the per-step CUDA-event pairs; at N>1 the "
              "ranks are re-aligned after each flush by a 4-byte all-reduce + (fused exchange) the library's device-side flag "
              "rendezvous, also outside the pairs.  CPU arm: not applicable",
        "collective": "GPU arm at N>1 (--gather %s): %s; N=1 and CPU arm: none" % (gather, {
            "fused": "the period kernel's epilogue stores the packed obs/reward/step/truncated/flags block into every rank's "
                     "gather buffer over NVLink + one-warp epoch handshake, inside the timed region; `gather_mode` says which "
                     "transport ran: fused_mc (symmetric memory, observation rows sent once through an NVLS multicast address, "
                     "multimem.st) or fused_ipc (CUDA-IPC mapped peer buffers, one store per peer)",
            "fused_mc": "forced fused_mc (see fused)", "fused_ipc": "forced fused_ipc (see fused)",
            "nccl": "one NCCL all-gather of the packed obs/reward/step/truncated/flags block per period, inside the timed region",
            "none": "NONE (diagnostic run: every rank keeps its shard to itself)"}[gather]),
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # bounded sample of the workload: each "step" = one control period of one env on every host core
    thr, procs, wall, kind = cpu_throughput(steps, os.cpu_count() or 1, warm=warm)
    world = max(1, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": thr, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * procs / thr,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(args.envs_per_gpu, world, 64, 22.0, 4, 250, 0.001, "f64", args.gather),
        "cpu_baseline": {"value": thr, "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": cpu_sample_text(kind, procs, steps, warm) + f"; {wall:.1f} s wall incl. process start-up"},
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
        return {"sm_mhz": statistics.median(loaded), "sm_min_mhz_under_load": min(loaded), "sm_max_mhz": max(smmax),
                "power_w_max": max(power), "samples": len(sm), "samples_under_load": len(loaded), "reasons": sorted(reasons)}


def load_json(*parts):
    try:
        with open(os.path.join(ROOT, *parts)) as f:
            return json.load(f)
    except Exception:
        return {}


class Peaks:
    """Roofline denominators: FP64 / FP32 FMA peaks self-measured in this run with the library's
    micro-kernels (MEASURED_PEAKS.json carries HBM and bf16 only), HBM from MEASURED_PEAKS.json."""

    def __init__(self, lib, device):
        best, mean = ctypes.c_double(), ctypes.c_double()
        rc = lib.ks_bench_fp64_peak(device, 20000, 5, ctypes.byref(best), ctypes.byref(mean))
        self.fp64 = best.value if rc == 0 and best.value > 0 else FP64_NOMINAL_TFLOPS
        self.fp64_src = "self-measured DFMA micro-kernel (ks_bench_fp64_peak, best of 5)" if rc == 0 else "nominal"
        rc = lib.ks_bench_fp32_peak(device, 20000, 5, ctypes.byref(best), ctypes.byref(mean))
        self.fp32 = best.value if rc == 0 and best.value > 0 else FP32_NOMINAL_TFLOPS
        self.fp32_src = "self-measured FFMA micro-kernel (ks_bench_fp32_peak, best of 5)" if rc == 0 else "nominal"
        peaks = load_json("MEASURED_PEAKS.json")
        self.hbm = float(peaks.get("hbm_gbs", 6650.0))
        self.hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"


def fd_roofline(env, B, kernel_ms, precision, peaks):
    """Roofline object of one ks_period_kernel launch over B envs (algorithmic flops of SURVEY.md 8d, the
    executed-instruction fraction from the SASS counts, HBM side, ncu traffic where a capture exists)."""
    N, J, S = env.N, env.J, env.cfg_steps
    P = env.launch_info()["points_per_lane"]
    flops = FLOPS_PER_POINT_SUBSTEP * N * S * B
    tf = flops / (kernel_ms * 1e-3) / 1e12
    f64 = precision == "f64"
    peak, src, nominal = (peaks.fp64, peaks.fp64_src, FP64_NOMINAL_TFLOPS) if f64 else (peaks.fp32, peaks.fp32_src, FP32_NOMINAL_TFLOPS)
    esz = 8 if f64 else 4
    nbytes = (2 * esz * N + 4 * N + 4 * J + 16) * B              # 20N+4J+16 per env (fp64)
    hbm = nbytes / (kernel_ms * 1e-3) / 1e9
    out = {"bound": "fp64" if f64 else "fp32", "kernel": f"ks_period_kernel<{'double' if f64 else 'float'},{P},0>",
           "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "peak_source": src, "peak_nominal": nominal,
           "frac_of_nominal": tf / nominal, "flops_per_launch": flops,
           "flops_model": "191*N*cfg_steps per env-period (SURVEY.md 8d)", "kernel_ms": kernel_ms,
           "hbm": {"achieved": hbm, "peak": peaks.hbm, "unit": "GB/s", "frac": hbm / peaks.hbm, "bytes_per_launch": nbytes,
                   "peak_source": peaks.hbm_src}}
    sass = load_json("profiles", "sass_counts.json").get(str(P)) if f64 else None
    if sass:
        # one FP64 instruction = one pipe slot = 2 flop-equivalents per lane; a lane holds P points
        per_point = 2.0 * sass["fp64"] / P
        out["frac_executed"] = out["frac"] * per_point / FLOPS_PER_POINT_SUBSTEP
        out["executed"] = {"fp64_instr_per_warp_substep": sass["fp64"], "issued_instr_per_warp_substep": sass["issue"],
                           "flop_equiv_per_point_substep": per_point, "algorithmic_per_point_substep": FLOPS_PER_POINT_SUBSTEP,
                           "source": "profiles/sass_counts.json (cuobjdump -sass of the shipped kernel, tools/sass_stats.py --table)",
                           "meaning": "fraction of the FP64 pipe's issue slots this launch used = what ncu reports as "
                                      "sm__pipe_fp64_cycles_active averaged over all SMs"}
    t = load_json("profiles", "ncu_traffic.json").get(f"{precision}/envs{B}/N{N}/P{P}") or \
        (load_json("profiles", "ncu_traffic.json").get(f"{precision}/envs{B}/P{P}") if N == 64 else None)
    out["traffic"] = (t["dram_bytes_read"] + t["dram_bytes_write"]) if t else None
    out["traffic_unit"] = "bytes per launch (dram read+write, ncu --set full capture of this configuration; static, see source)"
    out["traffic_source"] = t["source"] if t else None
    if t and "pipe_active_pct" in t:      # what ncu measured for the same configuration (compare with frac_executed)
        out["ncu_pipe_active_pct"] = {"pipe": t["pipe"], "pct": t["pipe_active_pct"], "issue_active_pct": t.get("issue_active_pct"),
                                      "source": t["source"]}
    return out


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
                       "note": "NOT the reference's scheme (FD-RK4): long-horizon statistics differ from the reference's by "
                               "design (dissipation -3.4 %, spectrum bins up to 8.5 %, DESIGN.md section 9) -- the north star's "
                               "1 % statistics gate holds for the FD-RK4 headline only; validated against oracle/ks_etdrk4.py at "
                               "1e-10 (parity unpinned by the reference, which has no spectral code)",
                       "layout": info},
            "roofline": {"bound": "fp64", "kernel": "ks_etd_kernel", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": tf / fp64_peak, "flops_per_launch": ETD_FLOPS_PER_ENV_STEP * ETD_STEPS * B,
                         "flops_model": "9600 per env per ETDRK4 step (8 FFTs per env pair at 5 N log2 N + pointwise)"}}


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from model_based_pde_control_b200 import KSVecEnv, _lib
    from model_based_pde_control_b200.sharding import connect_fused_gather, connect_fused_gather_symm, gather_packed

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
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    align = torch.zeros(1, device=dev)
    notes = []

    def allmax(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make_env(n, cfg=None, **kw):
        return KSVecEnv(n, dict(cfg or env_cfg), device=local_rank, solver=args.solver, **kw)

    def prepare(env, seed, Ksteps):
        """seeded ICs, short device burn-in onto the attractor, random actions for every
        period resident in HBM before anything is timed."""
        rng = np.random.default_rng(seed + rank)
        env.set_state(rng.uniform(-0.4, 0.4, (env.num_envs, env.N)), 0)
        env.rollout_device(None, K=args.burnin, outputs=False)
        env.set_state(None, 0)
        return torch.as_tensor(rng.uniform(-1, 1, (Ksteps, env.num_envs, env.J)).astype(np.float32)).to(dev)

    def fused(env):
        return getattr(env, "_bench_mode", "none") in ("fused_mc", "fused_ipc")

    def verify_gather(env, precision, ppl):
        """The fused exchange against two independent routes, bit for bit, both buffer parities:
        (i) an NCCL all-gather of a twin shard's packed block, (ii) the twin's own (single-GPU) outputs
        for this rank's slot.  All ranks agree on the verdict."""
        if os.environ.get("KS_GATHER_DEBUG"):      # measurement knobs that break the completion guarantee
            return "skipped (KS_GATHER_DEBUG=%s: cost-breakdown run, results not guaranteed complete)" % os.environ["KS_GATHER_DEBUG"]
        twin = make_env(env.num_envs, precision=precision, points_per_lane=ppl)
        rng = np.random.default_rng(4242 + rank)
        u0 = rng.uniform(-1.0, 1.0, (env.num_envs, env.N))
        acts = torch.as_tensor(rng.uniform(-1, 1, (3, env.num_envs, env.J)).astype(np.float32)).to(dev)
        env.set_state(u0, 0)
        twin.set_state(u0, 0)
        ok = True
        for k in range(3):
            got = env.step_gather(acts[k])
            mine = twin.step_device(acts[k])
            ref = gather_packed(mine["packed"], twin.packed_fields(), env.num_envs)
            torch.cuda.synchronize(dev)
            for name in twin.packed_fields():
                ok = ok and torch.equal(got[name], ref[name]) and torch.equal(got[name][rank], mine[name])
        twin.close()
        t = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item()) == 1.0

    def connected_env(n, precision, ppl):
        """A local shard of n envs with the per-period exchange set up, and whether it was verified.
        --gather fused (default) tries, in this order, and every rank takes the same branch:
          fused_mc   torch symmetric-memory buffers + NVLS multicast address: the period kernel sends its observation
                     rows once (multimem.st), the NVSwitch replicates them; small outputs + handshake unicast
          fused_ipc  CUDA-IPC mapped peer buffers, every store once per peer (round 1's path)
          nccl       one NCCL all-gather of the packed block per period
        A mode is used only if its set-up succeeded on EVERY rank and its results were verified bit for bit; the
        JSON line names the mode that ran (`gather_mode`) and why earlier ones were passed over (`collective_note`)."""
        order = {"fused": ["fused_mc", "fused_ipc", "nccl"], "fused_mc": ["fused_mc"], "fused_ipc": ["fused_ipc"],
                 "nccl": ["nccl"], "none": ["none"]}[args.gather if world > 1 else "none"]
        for i, mode in enumerate(order):
            env = make_env(n, precision=precision, points_per_lane=ppl)
            env._bench_mode = mode
            if mode in ("nccl", "none"):
                return env, None
            try:
                if mode == "fused_mc":
                    if args.solver != "fd_rk4":
                        raise RuntimeError("multicast stores are built into the FD-RK4 kernel only")
                    info = connect_fused_gather_symm(env)         # raises on every rank if any rank fails
                    if not info["multicast"]:
                        raise RuntimeError("symmetric memory gave no multicast address on some rank")
                else:
                    connect_fused_gather(env)
                verdict = verify_gather(env, precision, ppl)
                if verdict is not False:
                    return env, verdict
                why = "results differ from the NCCL all-gather / single-GPU run"
            except Exception as exc:              # noqa: BLE001 - reported in the JSON line
                why = f"{type(exc).__name__}: {exc}"
            env.close()
            if i + 1 == len(order):
                raise SystemExit(f"--gather {args.gather}: {mode} failed ({why})")
            notes.append(f"{mode} passed over ({why})")
            print(notes[-1], file=sys.stderr)

    def stepper(env):
        fields, n, mode = env.packed_fields(), env.num_envs, getattr(env, "_bench_mode", "none")

        def one_step(a):
            if mode in ("fused_mc", "fused_ipc"):
                return env.step_gather(a)     # kernel epilogue stores into every peer's buffer + handshake
            out = env.step_device(a)
            if mode == "nccl":
                out = gather_packed(out["packed"], fields, n)
            return out
        return one_step

    def timed(env, actions, Ksteps, Wsteps):
        """W warm-up + K timed steps: one CUDA-event pair per step on the launching stream, L2 flush
        (and, at N>1, a rank re-alignment) between steps outside the pairs; max over ranks of the sum."""
        step = stepper(env)
        for k in range(Wsteps):
            step(actions[k])
        torch.cuda.synchronize(dev)
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(Ksteps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(Ksteps)]
        n0 = env.launch_count
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        wall0 = time.perf_counter()
        for k in range(Ksteps):
            flush.zero_()                      # L2 flush between timed steps (outside the event pair)
            if world > 1:
                # the 256 MiB memsets do not take equally long on every GPU; re-align the ranks on the
                # device (stream-ordered, outside the event pair) so that a timed step is the period +
                # exchange, not the previous flush's skew: a 4-byte all-reduce, followed -- when the fused
                # exchange is connected -- by the library's own flag rendezvous (ks_gather_barrier), which
                # every rank leaves within about one NVLink flag flight (NCCL kernels exit a few us apart)
                dist.all_reduce(align)
                if fused(env) and not args.no_flag_barrier:
                    env.gather_barrier()
            starts[k].record(stream)
            step(actions[Wsteps + k])
            stops[k].record(stream)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - wall0
        step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
        if os.environ.get("KS_BENCH_DEBUG"):
            ss = sorted(step_ms)
            print(f"rank {rank} timed steps: median {ss[len(ss) // 2]:.4f} ms, min {ss[0]:.4f}, max {ss[-1]:.4f}, "
                  f"first 3 {[round(x, 4) for x in step_ms[:3]]}", file=sys.stderr)
        return allmax(sum(step_ms)), env.launch_count - n0, wall

    def check_health(env):
        bad = bool(env.nonfinite().any())
        if fused(env) and env.gather_timed_out():
            raise SystemExit("fused gather: a peer never signalled (handshake timed out)")
        return bad

    # =========================================== headline: 4096 envs per GPU =======================
    env, gather_verified = connected_env(B, args.precision, args.points_per_lane)
    N, J, S = env.N, env.J, env.cfg_steps
    total_envs = B * world
    actions = prepare(env, 1000, W + K)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)            # before the warm-up, so that the GPU does not idle between warm-up and timing
    if world > 1:
        dist.barrier()             # nobody enters the first exchange while rank 0 is still sleeping
    total_ms, launches, wall = timed(env, actions, K, W)
    clocks = sampler.stop() if rank == 0 else None
    flags_bad = check_health(env)

    # ---- sustained: >= 2 s of the same step back to back, own clock record ----
    sustained = None
    if not args.no_sustained:
        step = stepper(env)
        n_sus = max(K, int(args.sustained_s * 1e3 / (total_ms / K)) + 1)
        s2 = ClockSampler(local_rank)
        if rank == 0:
            s2.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = env.launch_count
        e0.record(stream)
        for k in range(n_sus):
            step(actions[k % (W + K)])
        e1.record(stream)
        torch.cuda.synchronize(dev)
        sus_ms = allmax(e0.elapsed_time(e1))
        sus_clocks = s2.stop() if rank == 0 else None
        flags_bad = check_health(env) or flags_bad
        sustained = {"value": total_envs * n_sus / (sus_ms * 1e-3), "unit": UNIT, "steps": n_sus, "seconds": sus_ms * 1e-3,
                     "ms_per_step": sus_ms / n_sus, "gpu_launches": int(env.launch_count - n0), "clocks": sus_clocks,
                     "note": "the headline step repeated back to back (no L2 flush; the 2 MiB working set is L2-resident either "
                             "way), one CUDA-event pair around the whole run, max over ranks; `value` above is the burst figure"}

    # ---- e2e: the gym-facing host API, host buffers, copies inside the timed region ----
    rng = np.random.default_rng(5000 + rank)
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
    e2e_s = allmax(time.perf_counter() - t0)
    if os.environ.get("KS_BENCH_DEBUG"):
        ss = sorted(step_s)
        print(f"e2e steps: median {1e3 * ss[len(ss) // 2]:.4f} ms, min {1e3 * ss[0]:.4f}, max {1e3 * ss[-1]:.4f}, "
              f"5 slowest {[round(1e3 * x, 3) for x in ss[-5:]]}, first 5 {[round(1e3 * x, 3) for x in step_s[:5]]}",
              file=sys.stderr)
    e2e_value = total_envs * K / e2e_s

    # ---- e2e over a whole episode, auto-reset included (reference semantics: 800 burn-in periods per 400 steps) ----
    episode = None
    if not args.no_episode and not spectral:
        E = env.max_episode_steps
        env.set_state(None, 0)
        n0 = env.launch_count
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        ntrunc = 0
        for k in range(E):
            obs, rew, term, trunc, info = env.step(acts_host[k % (W + K)])
            ntrunc += int(trunc.sum())
        ep_s = allmax(time.perf_counter() - t0)
        assert ntrunc == B and "final_observation" in info, "the episode's last step must truncate and auto-reset every env"
        episode = {"value": total_envs * E / ep_s, "unit": UNIT, "steps": E, "seconds": ep_s,
                   "burnin_periods_per_reset": env.burnin_periods, "gpu_launches": int(env.launch_count - n0),
                   "api": f"{E} x KSVecEnv.step(numpy actions); the last step truncates every env and runs the auto-reset "
                          f"(fresh initial conditions + {env.burnin_periods} no-op control periods in one launch, kuramoto.py:100-116) "
                          "before it returns -- solver work per episode = 400 + 800 periods, as in the reference"}

    # =========================================== BASELINE configs[2]: 65 536 envs over N GPUs ======
    config2 = None
    if not args.no_config_65536 and not spectral and args.precision == "f64":
        B2 = TOTAL_ENVS_CONFIG2 // world
        K2 = min(K, 20)
        env2, v2 = connected_env(B2, "f64", 0)
        a2 = prepare(env2, 3000, W + K2)
        ms2, l2, _ = timed(env2, a2, K2, W)
        bad2 = check_health(env2)
        config2 = dict(B=B2, K=K2, ms=ms2, launches=l2, bad=bad2, verified=v2, layout=env2.launch_info(), env=env2)

    if rank != 0:
        if config2:
            config2["env"].close()
        env.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- rooflines (rank 0) ----
    lib = _lib.load()
    peaks = Peaks(lib, local_rank)
    kernel_ms = total_ms / K            # N=1: the event pair brackets exactly one kernel launch
    value = total_envs * K / (total_ms * 1e-3)
    if spectral:
        tf = ETD_FLOPS_PER_ENV_STEP * S * B / (kernel_ms * 1e-3) / 1e12
        roofline = {"bound": "fp64", "kernel": "ks_etd_kernel", "achieved": tf, "peak": peaks.fp64, "unit": "TFLOP/s",
                    "frac": tf / peaks.fp64, "traffic": None, "peak_source": peaks.fp64_src,
                    "flops_per_launch": ETD_FLOPS_PER_ENV_STEP * S * B, "kernel_ms": kernel_ms,
                    "flops_model": "9600 per env per ETDRK4 step (8 FFTs per env pair at 5 N log2 N + pointwise)"}
    else:
        roofline = fd_roofline(env, B, kernel_ms, args.precision, peaks)

    cfg = shared_config(B, world, N, env.L, J, S, env.dt, args.precision, args.gather)
    if spectral:
        cfg["workload"] += " -- NON-DEFAULT solver=etdrk4 (pseudo-spectral ETDRK4, not the reference's scheme)"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": cfg,
        "layout": env.launch_info(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": env.h2d_bytes_per_step,
                "d2h_bytes_per_step": env.d2h_bytes_per_step, "ms_per_step": 1e3 * e2e_s / K,
                "api": "KSVecEnv.step(numpy actions) -> ks_step_host: one launch + sync; the kernel reads the pinned actions over PCIe "
                       "and mirrors the packed outputs into the pinned host block (KS_HOST_IO=copy: H2D copy, kernel, D2H copy)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "wall_s_timed_region": wall,
        "nonfinite": flags_bad,
        "value_kind": "burst (K short steps); see `sustained` for >= 2 s of the same step",
    }
    if notes:
        line["collective_note"] = "; ".join(notes)
    if world > 1:
        line["gather_verified"] = gather_verified
        line["gather_mode"] = env._bench_mode
    if sustained:
        if not spectral:
            sustained["roofline"] = {k: v for k, v in fd_roofline(env, B, sustained["ms_per_step"], args.precision, peaks).items()
                                     if k in ("bound", "achieved", "peak", "unit", "frac", "frac_executed", "kernel_ms")}
        line["sustained"] = sustained
    if episode:
        line["e2e_episode_amortised"] = episode
    if config2:
        env2 = config2.pop("env")
        B2, K2, ms2 = config2["B"], config2["K"], config2["ms"]
        line["config_65536"] = {
            "value": TOTAL_ENVS_CONFIG2 * K2 / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / K2, "steps": K2, "warmup": W,
            "scaling": "strong", "n_gpus": world, "gpu_launches": int(config2["launches"]), "nonfinite": config2["bad"],
            "config": {"workload": f"BASELINE.json configs[2]: {TOTAL_ENVS_CONFIG2} KS envs over {world} GPU(s) = {B2} per GPU, N={N} "
                                   f"L={env.L} J={J}, cfg_steps={S}, dt={env.dt}, f64, random actions; exchange "
                                   f"{env2._bench_mode} inside the timed region at N>1",
                       "layout": config2["layout"]},
            "roofline": fd_roofline(env2, B2, ms2 / K2, "f64", peaks),
            **({"gather_verified": config2["verified"]} if world > 1 else {}),
        }
        env2.close()

    # =========================================== BASELINE configs[3]: large domain (N=1 only) ======
    if world == 1 and not args.no_large_domain and not spectral:
        large = {}
        Xi = [k / LARGE["J"] for k in range(LARGE["J"])]
        for prec in ("f64", "f32"):
            try:
                e3 = KSVecEnv(B, dict(N=LARGE["N"], L=LARGE["L"]), Xi=Xi, device=local_rank, precision=prec)
                a3 = prepare(e3, 7000, W + K)
                ms3, l3, _ = timed(e3, a3, K, W)
                large[prec] = {"value": B * K / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3 / K, "steps": K, "warmup": W,
                               "gpu_launches": int(l3), "nonfinite": bool(e3.nonfinite().any()), "layout": e3.launch_info(),
                               "roofline": fd_roofline(e3, B, ms3 / K, prec, peaks)}
                e3.close()
            except Exception as exc:      # noqa: BLE001
                large[prec] = {"error": f"{type(exc).__name__}: {exc}"}
        line["large_domain"] = {
            "config": {"workload": f"BASELINE.json configs[3]: {B} KS envs, N={LARGE['N']} L={LARGE['L']} J={LARGE['J']} (jets at k/8), "
                                   f"cfg_steps={S}, dt={env.dt}, random actions, fp64 and fp32; the statistics check "
                                   "(spectrum / dissipation within 1 % of the reference) is tests/test_gpu_statistics.py"},
            **large}

    # ---- extra leg: the spectral ETDRK4 solver on the same batch (device-resident, same timing rules) ----
    if world == 1 and not args.no_spectral and not spectral:
        try:
            line["spectral_mode"] = spectral_leg(B, K, W, local_rank, peaks.fp64, flush)
        except Exception as exc:
            line["spectral_mode"] = {"error": str(exc)}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only; bounded sample) ----
    if world == 1 and not args.no_cpu_baseline:
        thr, procs, cpu_wall, kind = cpu_throughput(args.cpu_periods, warm=1)
        line["cpu_baseline"] = {"value": thr, "unit": UNIT, "cores": procs, "kind": kind,
                                "sample": cpu_sample_text(kind, procs, args.cpu_periods, 1) + f"; {cpu_wall:.1f} s wall"}
        try:
            cthr, cthreads = cpu_c_port_throughput()
            line["cpu_baseline_c"] = {"value": cthr, "unit": UNIT, "cores": cthreads, "kind": "port",
                                      "sample": "context only: plain-C oracle (oracle/ks_oracle.c, -O2, pthreads), 64 envs x 4 periods"}
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
    ap.add_argument("--gather", default="fused", choices=["fused", "fused_mc", "fused_ipc", "nccl", "none"],
                    help="N>1: how every rank gets the full batch each period.  fused (default): stores fused into the period "
                         "kernel's epilogue -- fused_mc (symmetric memory + NVLS multicast stores) if it sets up and verifies on "
                         "every rank, else fused_ipc (CUDA-IPC unicast peer stores), else nccl; the other values force one mode")
    ap.add_argument("--solver", default="fd_rk4", choices=["fd_rk4", "etdrk4"],
                    help="timed solver; the default is the reference's scheme (the headline), etdrk4 is the spectral mode")
    ap.add_argument("--no-spectral", action="store_true", help="skip the extra spectral-ETDRK4 leg")
    ap.add_argument("--sustained-s", type=float, default=2.0, help="length of the sustained leg in seconds")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-episode", action="store_true", help="skip e2e_episode_amortised")
    ap.add_argument("--no-config-65536", action="store_true")
    ap.add_argument("--no-large-domain", action="store_true")
    ap.add_argument("--no-flag-barrier", action="store_true", help="N>1: re-align the ranks between timed steps with the all-reduce only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
