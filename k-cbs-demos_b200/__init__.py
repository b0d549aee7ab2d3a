"""kcbs_b200 -- B200-native K-CBS low-level hot path.

The product is the C-ABI shared library ``libkcbs_b200.so`` (``include/kcbs_b200.h``) and the C++
host-side mirror of the reference's OMPL plugin headers (``includes/``).  This Python package is
only the thin ctypes binding that ``tests/`` and ``bench.py`` drive the C ABI through, plus the
scene constants of the reference demos and the synthetic workload generators.

The directory name contains hyphens (it is prescribed by the repo layout); import it through
``__graft_entry__.load_package()``, which registers it as ``kcbs_b200``.
"""
from . import scenes  # noqa: F401
from .capi import (  # noqa: F401
    Conflict,
    Context,
    KcbsError,
    KcbsParams,
    RrtParams,
    World,
    lib_path,
    load_library,
)
