"""Synchronous stand-ins for the torch.cuda objects mindrec_b200.sharded uses (streams, events, pinned buffers) — TEST
ONLY.  The gloo CPU tests of the exchange logic assign this module to `sharded._cu`; on the CPU every "stream" is the
host thread, so ordering calls are no-ops.  The product code never imports this module."""
import contextlib


class Stream:
    def __init__(self, device=None):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass


class Event:
    def record(self, stream=None):
        pass

    def synchronize(self):
        pass


_MAIN = Stream()


def current_stream(device=None):
    return _MAIN


@contextlib.contextmanager
def stream(s):
    yield


def synchronize():
    pass


class CUDAGraph:
    def __init__(self):
        raise RuntimeError("no graph capture on the CPU stand-in")
