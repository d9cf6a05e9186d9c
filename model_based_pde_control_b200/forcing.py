"""Host-side mirror of the reference's Gaussian-jet actuation (``pdegym/common/transforms.py:250-279``).

The forcing matrix is a one-off host precomputation; it is built with the same sequence of
torch float32 CPU operations the reference uses so that the matrix uploaded to the GPU is
bit-identical to ``GaussianForcing.forcing`` (a 1-ulp change of ``phi`` moves the one-period
state by 2.5e-8 relative L2).  The per-step product ``phi = a @ F`` on the
hot path is NOT computed here: the control-period kernel evaluates it as the same sequential
float32 FMA chain.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch


class GaussianForcing:
    """``forcing[j, i] = exp(-(x_i - L*Xi_j)^2 / (2 sigma^2)) / sqrt(2 pi sigma)``, float32.

    Same call surface as the reference transform: ``forcing(values)`` maps actions ``[..., J]``
    to spatial forcing ``[..., N]`` (NumPy in -> NumPy out, tensor in -> tensor out) and
    ``forcing.Inverse(values)`` maps a forcing pattern back to jet amplitudes by sampling it at
    the jet positions (``transforms.py:267-279``).
    """

    def __init__(self, x: Sequence, Xi: Sequence, sigma: float, L: float, N: int):
        self.sigma, self.L, self.N = sigma, L, N
        self.x = torch.as_tensor(np.asarray(x)).to(dtype=torch.float32)
        self.Xi = torch.as_tensor(np.asarray(Xi, dtype=np.float64)).to(dtype=torch.float32)
        self.xi = (self.L * self.Xi).reshape(-1, 1)
        # identical op order to transforms.py:258-260 (float32 throughout; note sqrt(2*pi*sigma))
        gauss = torch.exp((-((self.x - self.xi) ** 2.0) / (2.0 * sigma ** 2)))
        self.forcing = gauss / np.sqrt(2.0 * np.pi * self.sigma)
        self._inverse = None

    @property
    def J(self) -> int:
        return int(self.forcing.shape[0])

    def matrix(self) -> np.ndarray:
        """Contiguous float32 ``[J, N]`` copy for ``ks_config.forcing``."""
        return np.ascontiguousarray(self.forcing.numpy(), dtype=np.float32)

    def __call__(self, values):
        is_np = isinstance(values, np.ndarray)
        t = torch.from_numpy(values) if is_np else values
        out = t @ self.forcing.to(t.device)
        return out.detach().cpu().numpy() if is_np else out

    class _Inverse:
        def __init__(self, transf: "GaussianForcing"):
            self.transf = transf
            xpos = transf.Xi.reshape(-1, 1)
            self.xpos = (transf.N * xpos).to(dtype=torch.long).reshape(-1)
            self.inv_forcing = torch.inverse(transf.forcing[:, self.xpos])

        def __call__(self, values):
            is_np = isinstance(values, np.ndarray)
            t = torch.from_numpy(values) if is_np else values
            out = t[..., self.xpos.to(t.device)] @ self.inv_forcing.to(t.device)
            return out.detach().cpu().numpy() if is_np else out

        def update(self, values):
            pass

        @property
        def Inverse(self):
            return self.transf

    def update(self, values):
        pass

    @property
    def Inverse(self):
        if self._inverse is None:
            self._inverse = GaussianForcing._Inverse(self)
        return self._inverse
