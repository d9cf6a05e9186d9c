#!/usr/bin/env python
"""Where does the end-to-end (host-buffer) step spend its time?  Development tool.

Times, per control period of 4096 default envs: (a) ks_step on device tensors + stream sync,
(b) ks_step_host (pinned H2D + kernel + packed D2H + sync), (c) KSVecEnv.step (adds the NumPy
wrapping).  Wall clock over `--steps` iterations after warm-up.
"""
import argparse
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from model_based_pde_control_b200 import KSVecEnv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    B, K = args.envs, args.steps
    env = KSVecEnv(B)
    rng = np.random.default_rng(0)
    env.set_state(rng.uniform(-0.4, 0.4, (B, env.N)), 0)
    env.rollout_device(None, K=20, outputs=False)
    a_dev = torch.as_tensor(rng.uniform(-1, 1, (B, env.J)).astype(np.float32)).cuda()
    a_host = rng.uniform(-1, 1, (B, 1, env.J)).astype(np.float32)
    res = {}

    def timeit(name, fn):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            fn()
        torch.cuda.synchronize()
        res[name] = round(1e3 * (time.perf_counter() - t0) / K, 4)
        env.set_state(None, 0)

    def dev_step():
        env.step_device(a_dev)
        torch.cuda.current_stream().synchronize()

    timeit("ks_step_device_plus_sync_ms", dev_step)
    np.copyto(env._h_act, a_host.reshape(B, env.J))

    def host_call():
        env._lib.ks_step_host(env._h, env._h_act.ctypes.data, env._blocks[0]["pinned"].data_ptr(), env._stream())

    timeit("ks_step_host_ms", host_call)
    timeit("KSVecEnv_step_ms", lambda: env.step(a_host))
    keep = {}

    def step_keep():                      # like a training loop: the previous results stay referenced
        keep["r"] = env.step(a_host)

    timeit("KSVecEnv_step_results_kept_ms", step_keep)
    many = rng.uniform(-1, 1, (64, B, 1, env.J)).astype(np.float32)
    it = [0]

    def step_fresh_actions():
        it[0] += 1
        keep["r"] = env.step(many[it[0] % 64])

    timeit("KSVecEnv_step_fresh_actions_ms", step_fresh_actions)
    keep.clear()
    env.copy = False
    timeit("KSVecEnv_step_nocopy_ms", lambda: env.step(a_host))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(K):
        env.step_device(a_dev)
    t1.record()
    torch.cuda.synchronize()
    res["kernel_back_to_back_ms"] = round(t0.elapsed_time(t1) / K, 4)
    res["layout"] = env.launch_info()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
