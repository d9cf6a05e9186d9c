#!/usr/bin/env python
"""Device-resident throughput sweep over batch size and lane layout (development tool).

    python tools/sweep.py --envs 4096,4736,8192 --ppl 4,8,16 [--precision f64] [--steps 20]
Prints one JSON line per configuration: control-periods/s, ms per period, FP64 fraction.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from model_based_pde_control_b200 import KSVecEnv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", default="4096")
    ap.add_argument("--ppl", default="0")
    ap.add_argument("--N", type=int, default=64)
    ap.add_argument("--L", type=float, default=22.0)
    ap.add_argument("--J", type=int, default=4)
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--reward-mode", default="l2")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--solver", default="fd_rk4")
    ap.add_argument("--dt", type=float, default=None)
    ap.add_argument("--cfg-steps", type=int, default=None)
    ap.add_argument("--rollout", type=int, default=0, help="K periods per launch (0 = one launch per period)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    for B in [int(x) for x in args.envs.split(",")]:
        for P in [int(x) for x in args.ppl.split(",")]:
            Xi = [k / args.J for k in range(args.J)]
            try:
                cfg = dict(N=args.N, L=args.L)
                if args.dt is not None:
                    cfg["dt"] = args.dt
                if args.cfg_steps is not None:
                    cfg["cfg_steps"] = args.cfg_steps
                env = KSVecEnv(B, cfg, Xi=Xi, precision=args.precision, solver=args.solver,
                               reward_mode=args.reward_mode, points_per_lane=P)
            except Exception as exc:
                print(json.dumps({"envs": B, "ppl": P, "error": str(exc)}))
                continue
            rng = np.random.default_rng(0)
            env.set_state(rng.uniform(-0.4, 0.4, (B, args.N)), 0)
            env.rollout_device(None, K=20, outputs=False)
            K = args.steps
            acts = torch.as_tensor(rng.uniform(-1, 1, (K + 3, B, args.J)).astype(np.float32)).to(dev)
            for k in range(3):
                env.step_device(acts[k])
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            if args.rollout:
                env.rollout_device(acts[3:3 + K])
            else:
                for k in range(K):
                    env.step_device(acts[3 + k])
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / K
            flops = 191.0 * args.N * env.cfg_steps * B
            info = env.launch_info()
            print(json.dumps({"envs": B, "N": args.N, "solver": args.solver, "cfg_steps": env.cfg_steps, "dt": env.dt,
                              "precision": args.precision, "ppl": info["points_per_lane"],
                              "lanes": info["lanes_per_env"], "regs": info["regs_per_thread"],
                              "grid": info["grid_blocks"], "ms_per_period": round(ms, 4),
                              "periods_per_s": round(B / ms * 1e3),
                              "us_per_substep": round(ms * 1e3 / env.cfg_steps, 4), "tflops_alg": round(flops / ms / 1e9, 2),
                              "nonfinite": bool(env.nonfinite().any())}), flush=True)
            env.close()


if __name__ == "__main__":
    main()
