"""B200-native batched Kuramoto-Sivashinsky control environment.

Drop-in for the hot path of ``stwerner97/model-based-pde-control``'s ``KuramotoSivashinskyEnv-v0``
(``pdegym/kuramoto``): the same finite-difference + RK4 scheme, float32 jet forcing and reward,
advanced for thousands of independent environments by one persistent sm_100a kernel per control
period, behind the C ABI of ``libks_b200.so`` (``include/ks_b200.h``).

The package holds only what that path needs: ``csrc/`` (CUDA kernels + C ABI), the ctypes
binding, the host mirror of the gym-facing interface, and the env-sharding helper for
multi-GPU runs.  Importing it does not require a GPU; constructing an env does.
"""
from .device_pipeline import DeviceEnvPipeline, RolloutBatch, ScaleTransformDevice, SensorTransformDevice
from .env import KSVecEnv
from .forcing import GaussianForcing
from .registration import ENV_ID, make, register, vector_make
from .sharding import ShardedKSVecEnv, shard_range
from .single_env import KSEnv

__all__ = ["KSVecEnv", "KSEnv", "GaussianForcing", "ENV_ID", "make", "vector_make", "register", "ShardedKSVecEnv", "shard_range",
           "DeviceEnvPipeline", "RolloutBatch", "ScaleTransformDevice", "SensorTransformDevice"]
__version__ = "0.2.0"

register()      # no-op without gym; with gym the reference's env id resolves to the GPU env
