"""Temporal kernels for time-variable surfaces (reference: starry_process/temporal.py:8-16).

``StarryProcess(tau=..., temporal_kernel=Matern32Kernel)`` multiplies the flux covariance
elementwise by ``temporal_kernel(t, t, tau)`` (sp.py:697-698).  On the CUDA path the two kernels of
the reference are evaluated inside the assembly kernels (``spb_noise_model.temporal_kind``); the
callables below carry that code as ``spb_kind`` and, called directly, return the kernel matrix as a
torch tensor (same signature as the reference's functions).
"""
import math

import torch

__all__ = ["ExpSquaredKernel", "Matern32Kernel"]


def _dt(t1, t2):
    t1 = torch.as_tensor(t1, dtype=torch.float64)
    t2 = torch.as_tensor(t2, dtype=torch.float64).to(t1.device)
    return (t1.reshape(-1, 1) - t2.reshape(1, -1)).abs()


def ExpSquaredKernel(t1, t2, tau):
    """temporal.py:8-10."""
    dt = _dt(t1, t2)
    return torch.exp(-(dt ** 2) / (2 * tau))


def Matern32Kernel(t1, t2, tau):
    """temporal.py:13-16."""
    x = math.sqrt(3) * _dt(t1, t2) / tau
    return (1 + x) * torch.exp(-x)


Matern32Kernel.spb_kind = 1
ExpSquaredKernel.spb_kind = 2
