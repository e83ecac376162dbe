"""gorder_b200 — B200-native per-frame order-parameter engine behind gorder's frame loop.

The product is the CUDA library ``libgorder_b200.so`` (C ABI in ``include/gorder_b200.h``); this
package is its thin Python host mirror.  It never falls back to a CPU implementation.
"""
from . import abi
from .abi import EngineSetup, GorderError, MolType, RawResults

__all__ = ["abi", "EngineSetup", "MolType", "RawResults", "GorderError", "SystemTopology"]


def __getattr__(name):
    if name == "SystemTopology":
        from .topology import SystemTopology
        return SystemTopology
    raise AttributeError(name)
