"""ctypes loader for libmindrec_b200.so and the aot call marshaller.

Every kernel of the product path is reached through the MindSpore
``ops.Custom(func_type="aot")`` C signature declared in ``include/mindrec_b200.h``::

    int f(int nparam, void **params, int *ndims, int64_t **shapes,
          const char **dtypes, void *stream, void *extra)

Under MindSpore the framework builds that argument pack; in this repository the same symbols are driven
through ctypes with device buffers owned either by torch or by mindrec_b200.runtime (plain CUDA allocations: no
torch in the process) — plumbing only: memory + streams.  There is no CPU fallback: a missing library raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MREC_LIB_PATH: an experimental build of the same library (kernel tuning A/B runs); the product path is the default
LIB_PATH = os.environ.get("MREC_LIB_PATH") or os.path.join(_HERE, "libmindrec_b200.so")

ERROR_NAMES = {
    0: "OK", 1: "ERR_NPARAM", 2: "ERR_DTYPE", 3: "ERR_SHAPE", 4: "ERR_ALIGN",
    5: "ERR_DIM", 6: "ERR_CUDA", 7: "ERR_WORKSPACE", 8: "ERR_NULL",
}

_DTYPE_NAMES = {}
_lib = None


class MindrecKernelError(RuntimeError):
    """Raised when an aot entry point returns non-zero (MindSpore would raise RuntimeError too)."""

    def __init__(self, symbol, code, message):
        super().__init__("%s failed with %s (%d): %s" % (symbol, ERROR_NAMES.get(code, "?"), code, message))
        self.symbol = symbol
        self.code = code


def lib():
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libmindrec_b200.so is missing at %s: run `python -m mindrec_b200.build` "
                "(there is no CPU / PyTorch fallback for the product path)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mrec_version.restype = ctypes.c_char_p
        _lib.mrec_last_error.restype = ctypes.c_char_p
        _lib.mrec_launch_count.restype = ctypes.c_ulonglong
    return _lib


def version():
    return lib().mrec_version().decode()


def last_error():
    return lib().mrec_last_error().decode()


def launch_count():
    return int(lib().mrec_launch_count())


def _dtype_name(t):
    """MindSpore dtype string of a buffer: torch tensors ('torch.float32') and mindrec_b200.runtime buffers
    ('float32') both spell it in their dtype."""
    d = t.dtype
    name = _DTYPE_NAMES.get(d)
    if name is None:
        name = (d if isinstance(d, str) else str(d).replace("torch.", "")).encode()
        _DTYPE_NAMES[d] = name
    return name


def _current_stream(tensors):
    """The stream the kernels go to when the caller names none: the current stream of whoever owns the buffers."""
    for t in tensors:
        dev = getattr(t, "device", None)
        if getattr(dev, "__mrec_rt__", False):
            return dev.current_stream_handle()
    import torch
    return torch.cuda.current_stream().cuda_stream


_AOT_SIG = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
            ctypes.POINTER(ctypes.POINTER(ctypes.c_int64)), ctypes.POINTER(ctypes.c_char_p),
            ctypes.c_void_p, ctypes.c_void_p]
_fn_cache = {}


def _fn(symbol):
    f = _fn_cache.get(symbol)
    if f is None:
        f = getattr(lib(), symbol)
        f.argtypes = _AOT_SIG
        f.restype = ctypes.c_int
        _fn_cache[symbol] = f
    return f


_pack_cache = {}
_PACK_CACHE_MAX = 4096


_I64 = ctypes.sizeof(ctypes.c_int64)
_P_I64 = ctypes.POINTER(ctypes.c_int64)


def _build_pack(symbol, tensors, check_device):
    n = len(tensors)
    params = (ctypes.c_void_p * n)()
    ndims = (ctypes.c_int * n)()
    shapes = (_P_I64 * n)()
    dtypes = (ctypes.c_char_p * n)()
    flat = []
    for t in tensors:
        flat.extend(t.shape if t.dim() else (1,))
    shape_store = (ctypes.c_int64 * max(1, len(flat)))(*flat)   # one array for every shape
    base = ctypes.addressof(shape_store)
    off = 0
    for i, t in enumerate(tensors):
        if check_device and not t.is_cuda:
            raise RuntimeError("%s: param %d is not a CUDA tensor (no CPU path exists)" % (symbol, i))
        if not t.is_contiguous():
            raise RuntimeError("%s: param %d is not contiguous" % (symbol, i))
        params[i] = t.data_ptr() if t.numel() > 0 else None
        d = t.dim()
        ndims[i] = d
        shapes[i] = ctypes.cast(base + off * _I64, _P_I64)
        off += d if d else 1
        dtypes[i] = _dtype_name(t)
    return (n, params, ndims, shapes, dtypes, shape_store)


def aot_call(symbol, tensors, stream=None, check_device=True):
    """Call aot entry point `symbol` with `tensors` (inputs then outputs, torch CUDA tensors).

    The marshalled argument pack is cached per (symbol, buffer addresses, shapes, dtypes): a training step
    calls the same entry points on the same preallocated buffers, so the steady-state host cost of a call
    is one dictionary lookup (this matters for the eager, launch-bound multi-GPU step)."""
    # key: buffers, dtypes and ranks; extents are refreshed in place (they vary per step on the sharded path)
    key = (symbol, tuple((t.data_ptr(), t.dtype, t.dim()) for t in tensors))
    pack = _pack_cache.get(key)
    if pack is None:
        pack = _build_pack(symbol, tensors, check_device)
        if len(_pack_cache) >= _PACK_CACHE_MAX:
            _pack_cache.clear()
        _pack_cache[key] = pack
    else:
        store = pack[5]
        off = 0
        for t in tensors:
            if not t.is_contiguous():
                raise RuntimeError("%s: a param is not contiguous" % symbol)
            for d in t.shape:
                store[off] = d
                off += 1
            if t.dim() == 0:
                off += 1
    n, params, ndims, shapes, dtypes, _ = pack
    if stream is None:
        stream = _current_stream(tensors)
    rc = _fn(symbol)(n, params, ndims, shapes, dtypes, ctypes.c_void_p(stream), None)
    if rc != 0:
        raise MindrecKernelError(symbol, rc, last_error())
    return rc


def aot_call_raw(symbol, params, shapes, dtypes, stream=0):
    """Call an aot entry point with raw integer pointers (used by ABI-level tests without a GPU)."""
    n = len(params)
    c_params = (ctypes.c_void_p * n)(*params)
    c_ndims = (ctypes.c_int * n)(*[len(s) for s in shapes])
    keep = [(ctypes.c_int64 * max(1, len(s)))(*s) for s in shapes]
    c_shapes = (ctypes.POINTER(ctypes.c_int64) * n)(
        *[ctypes.cast(k, ctypes.POINTER(ctypes.c_int64)) for k in keep])
    c_dtypes = (ctypes.c_char_p * n)(*[d.encode() for d in dtypes])
    return _fn(symbol)(n, c_params, c_ndims, c_shapes, c_dtypes, ctypes.c_void_p(stream), None)
