"""Drop-in shim: ``import sleekit`` resolves to the B200 implementation, so the reference's
``experiments/*.py`` (``from sleekit.codebook import *`` ...) run unchanged against it."""

import sys

import sleekit_b200
from sleekit_b200 import codebook, obq, scaling, statistics  # noqa: F401
from sleekit_b200.statistics import Sleekit  # noqa: F401

for _name in ("codebook", "obq", "scaling", "statistics"):
    sys.modules[__name__ + "." + _name] = getattr(sleekit_b200, _name)
