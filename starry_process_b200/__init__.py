"""starry_process_b200 -- B200-native (sm_100a) batched log-likelihood path of starry_process.

Public surface mirrors the reference for this path: ``StarryProcess``, ``gauss2beta``,
``beta2gauss``.  Importing the package does not touch the GPU; constructing a ``StarryProcess``
loads ``libspb200.so`` (build it with ``python -m starry_process_b200.build``) and fails loudly if it
or a CUDA device is missing.
"""
from .sp import (StarryProcess, StarryProcessSum, beta2gauss, defaults, gauss2beta,  # noqa: F401
                 get_context)
from .distributed import (design_matrix_sharded, ensemble_log_likelihood_sharded,  # noqa: F401
                          gather_lnlike, log_likelihood_sharded, shard_range)
from .temporal import ExpSquaredKernel, Matern32Kernel  # noqa: F401

__version__ = "0.1.0"
