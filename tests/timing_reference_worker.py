#!/usr/bin/env python
"""BASELINE configs[4], measured ("env seconds per iteration before / after"): the reference's OWN
data-collection code -- wrapper stack of ``mbrl.py:257-291`` + ``Worker.rollout`` (``worker.py:39-93``), unmodified,
executed from ``baseline/_ref`` -- over (a) a real ``KSVecEnv`` on the GPU and (b) the reference's own env stepped in
this process (what ONE worker process of gym's AsyncVectorEnv does; the reference runs ``cpus`` of them in
parallel, ``mbrl.py:81-86``).  Test-side tool (it executes the reference through ``oracle/ref_loader.py``); not part
of the product path.

    python tests/timing_reference_worker.py [--envs 10,512,4096] [--steps 50]

Prints one JSON line per batch size: seconds per ``Worker.rollout`` iteration of ``steps`` vector steps, split into
env time (inside ``envs.step``) and the reference's wrapper / replay bookkeeping around it.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class TimedEnv:
    """Pass-through that accumulates the wall time spent inside the wrapped vector env."""

    def __init__(self, env):
        self.__dict__["env"] = env
        self.__dict__["seconds"] = 0.0

    def __getattr__(self, name):
        return getattr(self.__dict__["env"], name)

    def step_async(self, actions):
        t0 = time.perf_counter()
        self.env.step_async(actions)
        self.__dict__["seconds"] += time.perf_counter() - t0

    def step_wait(self, **kw):
        t0 = time.perf_counter()
        out = self.env.step_wait(**kw)
        self.__dict__["seconds"] += time.perf_counter() - t0
        return out

    def reset(self, **kw):
        return self.env.reset(**kw)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", default="10,512,4096")
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()

    import test_gpu_reference_stack as trs          # build_reference_stack, SeededAgent (the test's own helpers)
    from model_based_pde_control_b200 import KSVecEnv
    from oracle import ref_loader

    # (b) the reference env itself: seconds per control period of ONE env in one process
    ref = ref_loader.make_reference_env()
    rng = np.random.default_rng(0)
    ref.u, ref.timestep = rng.uniform(-0.4, 0.4, 64), 0
    ref.step(np.zeros((1, 4), np.float32))
    t0 = time.perf_counter()
    for _ in range(10):
        ref.step(rng.uniform(-1, 1, (1, 4)).astype(np.float32))
    ref_step_s = (time.perf_counter() - t0) / 10

    for B in [int(x) for x in args.envs.split(",")]:
        trs.B = B
        envs = KSVecEnv(B, burnin_periods=8)          # (short burn-in: the reset is not what is being timed)
        timed = TimedEnv(envs)
        worker, _ = trs.build_reference_stack(timed)
        agent = trs.SeededAgent(envs.J)
        worker._last_obs = worker.stack.envs.reset(seed=0)
        worker._last_stored_obs = worker.stack.ostore.obs.copy()[worker.stack.ostore.mask]
        worker.rollout(agent, stop=lambda ts, eps: ts >= B * 3)          # warm-up
        timed.__dict__["seconds"] = 0.0
        t0 = time.perf_counter()
        replay = worker.rollout(agent, stop=lambda ts, eps: ts >= B * args.steps)
        total = time.perf_counter() - t0
        env_s = timed.seconds
        print(json.dumps({
            "num_envs": B, "vector_steps": args.steps, "transitions": int(replay.ntimesteps),
            "rollout_seconds": total, "env_seconds": env_s, "reference_plumbing_seconds": total - env_s,
            "ms_per_vector_step": 1e3 * total / args.steps, "env_ms_per_vector_step": 1e3 * env_s / args.steps,
            "reference_env_ms_per_step_one_process": 1e3 * ref_step_s,
            "reference_env_seconds_same_iteration": ref_step_s * args.steps,
            "note": "reference: every AsyncVectorEnv worker process needs reference_env_ms_per_step per vector step (cpus processes "
                    "in parallel); the per-env Python loop of Sample.split / ExperienceReplay.add is the reference's own"}), flush=True)
        envs.close()


if __name__ == "__main__":
    main()
