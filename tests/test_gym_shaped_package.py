"""The code paths that run only when ``import gym`` succeeds (``spaces.py``: real ``gym.vector.VectorEnv`` /
``gym.Env`` / ``Box`` as base classes; ``registration.py``: ``gym.envs.register`` on import, gym's ``TimeLimit``).
gym 0.25.2 is not installable here, so a test double with the same surface (``tests/fake_gym``, read its
docstring) is put on ``sys.path`` -- in a SUBPROCESS, so that the rest of the suite keeps the gym-less classes.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAKE = os.path.join(ROOT, "tests", "fake_gym")

CPU_SCRIPT = r'''
import gym, numpy as np
import model_based_pde_control_b200 as ksb
from model_based_pde_control_b200 import spaces, registration
assert spaces.HAVE_GYM and spaces.Box is gym.spaces.Box
assert issubclass(ksb.KSVecEnv, gym.vector.VectorEnv) and issubclass(ksb.KSEnv, gym.Env)
# registered on import, with the reference's id / keywords (pdegym/kuramoto/__init__.py:26-31)
spec = gym.envs.registry["KuramotoSivashinskyEnv-v0"]
assert spec["entry_point"] == "model_based_pde_control_b200.registration:make"
assert spec["kwargs"] == {"order_enforce": False, "new_step_api": True}
assert registration.register() is False          # already taken -> untouched
print("GYM_CPU_OK")
'''

GPU_SCRIPT = r'''
import gym, numpy as np
import model_based_pde_control_b200 as ksb
# gym.make resolves the id to TimeLimit(KSEnv(**config)) -- gym's own TimeLimit (generate.py:23)
env = gym.make("KuramotoSivashinskyEnv-v0", config=dict(cfg_steps=5, Tmax=0.02), burnin_periods=2)
assert type(env) is gym.wrappers.TimeLimit and isinstance(env.unwrapped, ksb.KSEnv) and isinstance(env.unwrapped, gym.Env)
obs = env.reset(seed=3)
assert obs.shape == (1, 64) and obs.dtype == np.float64
flags = [env.step(env.action_space.sample())[3] for _ in range(4)]
assert flags == [False, False, False, True]
env.close()
# a real gym.vector.VectorEnvWrapper (isinstance assertion) accepts the vector env; step / reset go through the base class
envs = ksb.vector_make("KuramotoSivashinskyEnv-v0", num_envs=6, config=dict(cfg_steps=5, Tmax=0.02), burnin_periods=2)
assert isinstance(envs, gym.vector.VectorEnv) and envs.is_vector_env and envs.observation_space.shape == (6, 1, 64)
w = gym.vector.VectorEnvWrapper(envs)
o = w.reset(seed=1)
assert o.shape == (6, 1, 64) and o.dtype == np.float32
for k in range(4):
    o, r, term, trunc, info = w.step(envs.action_space.sample())
assert trunc.all() and info["_final_observation"].all() and (info["step"] == 4).all() and r.shape == (6,)
o, r, term, trunc, info = w.step(envs.action_space.sample())
assert (info["step"] == 1).all() and not trunc.any()          # auto-reset happened inside the truncating step
w.close()
assert envs.closed
print("GYM_GPU_OK")
'''


def _run(script):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([FAKE, ROOT, os.environ.get("PYTHONPATH", "")]))
    res = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    return res.stdout


def test_gym_branch_class_hierarchy_and_registration():
    assert "GYM_CPU_OK" in _run(CPU_SCRIPT)


@pytest.mark.gpu
def test_gym_make_and_vector_wrapper_over_the_gpu_env():
    assert "GYM_GPU_OK" in _run(GPU_SCRIPT)
