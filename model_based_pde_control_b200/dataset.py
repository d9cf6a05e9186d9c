"""Random-action episode datasets in the reference's on-disk format, generated on the GPU.

Mirror of ``pdecontrol/surrogates/evaluation/generate.py:21-63`` (the script behind
``KSattractor.pl`` in ``runscripts/offline.sh:5``): for every episode ``env.reset()`` (initial
condition + 800 no-op control periods), then ``action_space.sample()`` / ``env.step`` until the
``TimeLimit`` truncates at 400 steps; the result is ``torch.save(TensorDataset(obs, actions, nxt,
rewards, terminated, truncated, steps))`` with

    obs, nxt   float32 [E, T, 1, N]        actions  float32 [E, T, 1, J]
    rewards    float32 [E, T]              terminated / truncated  bool [E, T]
    steps      int64   [E, T] = arange(T)  (generate.py:49-50)

The reference runs the episodes one after another on one core (33 s of burn-in + ~20 s per
episode); here ``num_envs`` episodes run side by side and a whole episode is ONE persistent
launch (``ks_rollout`` with K = 400 periods), burn-in another one.

    python -m model_based_pde_control_b200.dataset --output KSattractor.pl --episodes 100 --config '{}'
"""
from __future__ import annotations

import argparse
import json
from typing import Optional

import torch
from torch.utils.data import TensorDataset

from .env import KSVecEnv


def generate_episodes(env: KSVecEnv, episodes: int, seed: Optional[int] = None) -> TensorDataset:
    """``episodes`` full random-action episodes (``ceil(episodes / num_envs)`` batches of
    reset + one rollout launch).  Actions are i.i.d. U(-1, 1) like ``Box(-1,1).sample()``
    (generate.py:32), drawn on the device from ``seed``."""
    B, T, dev = env.num_envs, env.max_episode_steps, env.device
    gen = torch.Generator(device=dev)
    if seed is not None:
        gen.manual_seed(int(seed))
    else:
        gen.seed()
    parts = []
    done = 0
    batch = 0
    while done < episodes:
        env.reset_device(seed=None if seed is None else int(seed) + 7919 * batch)     # IC + burn-in, timestep = 0
        u0, _ = env.get_state_device()
        first = u0[:, env.sensor_stride // 2::env.sensor_stride].to(torch.float32)      # obs = env.reset()
        actions = torch.rand((T, B, env.J), generator=gen, device=dev, dtype=torch.float32) * 2 - 1
        out = env.rollout_device(actions)                                               # one launch, T periods
        nxt = out["obs"]                                                                # [T, B, No]
        obs = torch.cat([first[None], nxt[:-1]], dim=0)
        n = min(B, episodes - done)
        parts.append(tuple(t[:, :n].transpose(0, 1).contiguous().cpu() for t in (
            obs, actions, nxt, out["reward"].to(torch.float32), out["truncated"].bool(), out["step"])))
        if bool(out["nonfinite"].any()):
            raise FloatingPointError("overflow encountered in KS state while generating episodes")
        done += n
        batch += 1
    obs, actions, nxt, rewards, truncated, step = (torch.cat(x, dim=0) for x in zip(*parts))
    E = obs.shape[0]
    assert bool((step[:, -1] == T).all()) and bool(truncated[:, -1].all()) and not bool(truncated[:, :-1].any())
    steps = torch.arange(T, dtype=torch.int64).reshape(1, -1).repeat(E, 1)              # generate.py:49-50
    return TensorDataset(obs.unsqueeze(2), actions.unsqueeze(2), nxt.unsqueeze(2), rewards,
                         torch.zeros_like(truncated), truncated, steps)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--env", type=str, default="KuramotoSivashinskyEnv-v0")
    ap.add_argument("--output", type=str, required=True)
    ap.add_argument("--episodes", type=int, default=100)
    ap.add_argument("--config", type=str, default="{}")
    ap.add_argument("--num-envs", type=int, default=0, help="episodes generated side by side (default: all)")
    ap.add_argument("--seed", type=int, default=None)
    args = ap.parse_args(argv)
    if args.env != "KuramotoSivashinskyEnv-v0":
        raise SystemExit(f"only KuramotoSivashinskyEnv-v0 is provided, got {args.env}")
    env = KSVecEnv(args.num_envs or args.episodes, json.loads(args.config), ic="device")
    data = generate_episodes(env, args.episodes, args.seed)
    torch.save(data, args.output)
    env.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
