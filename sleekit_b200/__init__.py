"""sleekit_b200 -- B200 (sm_100a) implementation of sleekit's layer-wise quantization hot path.

Same public functions as ``sleekit.codebook / obq / scaling / statistics`` of the
reference (Coloquinte/sleekit); the arithmetic runs in hand-written CUDA kernels
behind the C ABI of ``include/sleekit_b200.h``.  There is no CPU fallback.
"""

from . import codebook, obq, scaling, statistics  # noqa: F401
from .statistics import Sleekit  # noqa: F401

__version__ = "0.1.0"
