"""foragax_b200 -- B200-native (sm_100a) implementation of Foragax's per-tick agent hot path
behind the reference's ``foragax.base`` API.

Layout: ``csrc/`` hand-written CUDA kernels + the C ABI (``include/foragax_b200.h``),
built in-tree as ``libforagax_b200.so``; ``base/`` and ``policy/`` the host-side mirror of
``foragax/base`` and ``foragax/policy``; ``examples/`` the reference's example agents with
native ops.  Importing the package loads the shared library and fails if it is missing:
there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (loads libforagax_b200.so; raises if absent)
from . import random, struct  # noqa: F401

__version__ = "0.1.0"
