"""mindrec_b200 — B200-native (sm_100a) implementation of MindRec's embedding-and-interaction hot path.

The compute lives in ``libmindrec_b200.so`` behind the MindSpore ``ops.Custom(func_type="aot")`` C-ABI
(``include/mindrec_b200.h``); this package is the host-side mirror of the reference's operator surface.
"""
from . import _lib  # noqa: F401

__all__ = ["EmbeddingLookup", "HashEmbeddingLookup", "MapParameter", "Adam", "LazyAdam", "FTRL",
           "WideDeepModel", "DeepFMModel", "DeepCrossModel", "CrossLayer"]


def __getattr__(name):
    # torch-backed modules are imported lazily so `import mindrec_b200` stays cheap
    import importlib
    for mod in ("nn", "cells", "hash", "interaction"):
        try:
            m = importlib.import_module("." + mod, __name__)
        except ImportError:
            continue
        if hasattr(m, name):
            return getattr(m, name)
    raise AttributeError("module 'mindrec_b200' has no attribute %r" % name)
