#!/usr/bin/env python
"""Data-collection loop timing (BASELINE.json configs[4], functional: "env seconds per iteration").

The reference's MBRL loop collects transitions with ``Worker.rollout`` (pdecontrol/mbrl/worker.py:39-93)
over ``cpus`` (default 10) gym sub-processes: policy forward (SAC tanh-Gaussian MLP on ``[B,1,N]``
observations, pdecontrol/sac/policies.py:73-130), ``envs.step``, wrapper bookkeeping.  This tool runs
the same loop with the GPU env:

  device : ``DeviceEnvPipeline.rollout`` -- policy, env kernel and wrapper plumbing all on the GPU,
           no host synchronisation inside the loop;
  graphed: ``DeviceEnvPipeline.rollout_graphed`` -- the same step captured once in a CUDA graph and
           replayed (one graph launch per step instead of ~40 kernel launches);
  host   : ``KSVecEnv.step`` with NumPy actions (what the reference's wrappers would call), policy on
           the GPU with a host round trip per step.

The MLP is a random-init stand-in of the reference policy's shape (there is no checkpoint; the env
cost does not depend on the weights).  Prints one JSON line per batch size.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from model_based_pde_control_b200 import DeviceEnvPipeline, KSVecEnv


class TanhGaussianPolicy(torch.nn.Module):
    """Shape of GaussianPolicy (policies.py:73-130): flatten (B,1,N) -> 256 -> 256 -> mean / log_std (J)."""

    def __init__(self, N, J, hidden=256):
        super().__init__()
        self.l1, self.l2 = torch.nn.Linear(N, hidden), torch.nn.Linear(hidden, hidden)
        self.mean, self.log_std = torch.nn.Linear(hidden, J), torch.nn.Linear(hidden, J)
        self.J = J

    def forward(self, obs):
        x = torch.relu(self.l2(torch.relu(self.l1(obs.reshape(obs.shape[0], -1)))))
        mean, log_std = self.mean(x), self.log_std(x).clamp(-20, 2)
        return torch.tanh(mean + log_std.exp() * torch.randn_like(mean)).reshape(-1, 1, self.J)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", default="10,1024,4096")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--solver", default="fd_rk4")
    args = ap.parse_args()
    for B in [int(x) for x in args.envs.split(",")]:
        cfg = dict(dt=0.025, cfg_steps=10) if args.solver == "etdrk4" else {}
        cfg["Tmax"] = 1000.0        # 4000-step episodes: no auto-reset (burn-in launch) inside the timed loops
        env = KSVecEnv(B, cfg, ic="device", solver=args.solver, burnin_periods=40)     # short burn-in: timing tool
        policy = TanhGaussianPolicy(env.N, env.J).cuda()
        pipe = DeviceEnvPipeline(env)
        last = pipe.reset(seed=0)
        _, last = pipe.rollout(policy, args.steps, last_obs=last, reuse_buffers=True)   # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        batch, last = pipe.rollout(policy, args.steps, last_obs=last, reuse_buffers=True)
        torch.cuda.synchronize()
        dev_s = time.perf_counter() - t0
        # the same loop with the step captured in a CUDA graph
        _, last = pipe.rollout_graphed(policy, args.steps, last_obs=last)          # warm-up + capture
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        batch_g, last = pipe.rollout_graphed(policy, args.steps, last_obs=last)
        torch.cuda.synchronize()
        graph_s = time.perf_counter() - t0
        # host API path: NumPy observations out, NumPy actions in, every step
        obs = env.reset(seed=0, burnin_periods=40)
        for _ in range(5):
            with torch.no_grad():
                a = policy(torch.as_tensor(obs).cuda()).cpu().numpy()
            obs, *_ = env.step(a)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            with torch.no_grad():
                a = policy(torch.as_tensor(obs).cuda()).cpu().numpy()
            obs, rew, term, trunc, info = env.step(a)
        host_s = time.perf_counter() - t0
        print(json.dumps({"envs": B, "steps": args.steps, "solver": args.solver,
                          "device_pipeline_env_steps_per_s": round(B * args.steps / dev_s),
                          "device_pipeline_ms_per_step": round(1e3 * dev_s / args.steps, 4),
                          "graphed_pipeline_env_steps_per_s": round(B * args.steps / graph_s),
                          "graphed_pipeline_ms_per_step": round(1e3 * graph_s / args.steps, 4),
                          "host_api_env_steps_per_s": round(B * args.steps / host_s),
                          "host_api_ms_per_step": round(1e3 * host_s / args.steps, 4),
                          "transitions_shape": list(batch.obs.shape)}), flush=True)
        env.close()


if __name__ == "__main__":
    main()
