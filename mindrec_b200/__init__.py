"""mindrec_b200 — B200-native (sm_100a) implementation of MindRec's embedding-and-interaction hot path.

The compute lives in ``libmindrec_b200.so`` behind the MindSpore ``ops.Custom(func_type="aot")`` C-ABI
(``include/mindrec_b200.h``); this package is the host-side mirror of the reference's operator surface.
"""
from . import _lib  # noqa: F401

__all__ = ["EmbeddingLookup", "HashEmbeddingLookup", "MapParameter", "Adam", "LazyAdam", "FTRL",
           "WideDeepModel", "DeepFMModel", "DeepCrossModel", "CrossLayer"]


def __getattr__(name):
    # the torch-backed cells are imported lazily, and only for the names they export: `from mindrec_b200 import ops,
    # runtime` (the torch-free layer) must not drag them in
    if name not in __all__:
        raise AttributeError("module 'mindrec_b200' has no attribute %r" % name)
    import importlib
    for mod in ("nn", "cells", "hash", "interaction"):
        try:
            m = importlib.import_module("." + mod, __name__)
        except ImportError:
            continue
        if hasattr(m, name):
            return getattr(m, name)
    raise AttributeError("module 'mindrec_b200' has no attribute %r" % name)
