"""Torch-free host runtime: device buffers, streams, events and CUDA graphs over the `mrec_rt_*` entry points of
libmindrec_b200.so (ctypes only).  BASELINE north_star: "Python host code calls a .so through ... the aot C-ABI, with
no PyTorch".  `mindrec_b200.ops` allocates its outputs and workspaces through the device of its inputs, so handing it
`DeviceBuffer`s runs the whole embedding path — gather, dedup, fused sparse optimizers, FM / cross kernels, the hash
table's probe kernels — without torch in the process (tests/test_runtime_gpu.py asserts it on `sys.modules`).

    dev = runtime.Device(0)
    table = dev.from_numpy(w)                         # numpy is the host-side array type
    ids = dev.from_numpy(batch_ids)
    rows = ops.gather(table, ids)                     # a DeviceBuffer
    uq = ops.unique(ids, table_like=table)
    ops.sparse_lazy_adam(table, m, v, hyper, grads, mask, uq)
    with dev.capture() as g: ...                      # the same calls recorded once
    g.launch()

What stays on torch in this repository: the DenseLayer GEMMs (cuBLAS), torch.distributed (NCCL bootstrap) and the
model cells built on them (cells / nn / hash / interaction); see DESIGN.md.
"""
import contextlib
import ctypes

import numpy as np

from . import _lib

_NP = {"float32": np.float32, "float16": np.float16, "int32": np.int32, "int64": np.int64, "uint8": np.uint8,
       "int8": np.int8, "float64": np.float64, "bool": np.bool_}
_bound = False


def _c():
    global _bound
    L = _lib.lib()
    if not _bound:
        vp, sz, i64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int64
        for name, res, args in (
                ("mrec_rt_malloc", vp, [sz]), ("mrec_rt_free", ctypes.c_int, [vp]),
                ("mrec_rt_malloc_host", vp, [sz]), ("mrec_rt_free_host", ctypes.c_int, [vp]),
                ("mrec_rt_memcpy", ctypes.c_int, [vp, vp, sz, ctypes.c_int, vp]),
                ("mrec_rt_memset", ctypes.c_int, [vp, ctypes.c_int, sz, vp]),
                ("mrec_rt_fill32", ctypes.c_int, [vp, ctypes.c_uint32, i64, vp]),
                ("mrec_rt_stream_create", vp, []), ("mrec_rt_stream_destroy", ctypes.c_int, [vp]),
                ("mrec_rt_stream_sync", ctypes.c_int, [vp]), ("mrec_rt_device_sync", ctypes.c_int, []),
                ("mrec_rt_event_create", vp, [ctypes.c_int]), ("mrec_rt_event_record", ctypes.c_int, [vp, vp]),
                ("mrec_rt_event_sync", ctypes.c_int, [vp]), ("mrec_rt_stream_wait_event", ctypes.c_int, [vp, vp]),
                ("mrec_rt_event_elapsed_ms", ctypes.c_float, [vp, vp]), ("mrec_rt_event_destroy", ctypes.c_int, [vp]),
                ("mrec_rt_graph_begin", ctypes.c_int, [vp]), ("mrec_rt_graph_end", vp, [vp]),
                ("mrec_rt_graph_launch", ctypes.c_int, [vp, vp]), ("mrec_rt_graph_destroy", ctypes.c_int, [vp]),
                ("mrec_rt_gemm", ctypes.c_int, [vp, vp, vp, vp, i64, i64, i64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int, vp]),
                ("mrec_rt_device_count", ctypes.c_int, []), ("mrec_rt_set_device", ctypes.c_int, [ctypes.c_int]),
                ("mrec_rt_mem_info", ctypes.c_int, [ctypes.POINTER(sz), ctypes.POINTER(sz)])):
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _bound = True
    return L


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, _lib.last_error()))


def device_count():
    return int(_c().mrec_rt_device_count())


class Stream:
    def __init__(self, handle=None):
        self._own = handle is None
        self.handle = _c().mrec_rt_stream_create() if handle is None else handle
        if self._own and not self.handle:
            raise RuntimeError("cudaStreamCreate failed: " + _lib.last_error())

    def synchronize(self):
        _check(_c().mrec_rt_stream_sync(ctypes.c_void_p(self.handle)), "stream sync")

    def wait_event(self, ev):
        _check(_c().mrec_rt_stream_wait_event(ctypes.c_void_p(self.handle), ctypes.c_void_p(ev.handle)), "stream wait")

    def wait_stream(self, other):
        ev = Event()
        ev.record(other)
        self.wait_event(ev)


class Event:
    def __init__(self, timing=False):
        self.handle = _c().mrec_rt_event_create(1 if timing else 0)
        if not self.handle:
            raise RuntimeError("cudaEventCreate failed: " + _lib.last_error())

    def record(self, stream):
        _check(_c().mrec_rt_event_record(ctypes.c_void_p(self.handle), ctypes.c_void_p(stream.handle)), "event record")

    def synchronize(self):
        _check(_c().mrec_rt_event_sync(ctypes.c_void_p(self.handle)), "event sync")

    def elapsed_time(self, end):
        return float(_c().mrec_rt_event_elapsed_ms(ctypes.c_void_p(self.handle), ctypes.c_void_p(end.handle)))

    def __del__(self):
        try:
            _c().mrec_rt_event_destroy(ctypes.c_void_p(self.handle))
        except Exception:                      # interpreter shutdown
            pass


class Graph:
    def __init__(self, handle, device):
        self.handle, self.device = handle, device

    def launch(self, stream=None):
        s = stream or self.device.stream
        _check(_c().mrec_rt_graph_launch(ctypes.c_void_p(self.handle), ctypes.c_void_p(s.handle)), "graph launch")

    replay = launch


def gemm(a, b, out, bias=None, trans_a=False, trans_b=False, relu=False, alpha=1.0, beta=0.0):
    """Row-major out[M,N] = act(alpha * op(a) op(b) + beta * out + bias) on the device's current stream (cuBLASLt through
    mrec_rt_gemm: fp16 or fp32 storage, fp32 accumulation).  op(a) is a or a^T, op(b) likewise; relu needs a bias."""
    if a.dtype != b.dtype or a.dtype not in ("float16", "float32") or out.dtype not in ("float16", "float32"):
        raise TypeError("gemm: a, b fp16 or fp32 (same type), out fp16 or fp32")
    m, k = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    k2, n = (b.shape[1], b.shape[0]) if trans_b else (b.shape[0], b.shape[1])
    if k != k2 or tuple(out.shape) != (m, n):
        raise ValueError("gemm: shapes %r x %r -> %r do not match" % (a.shape, b.shape, out.shape))
    if relu and bias is None:
        raise ValueError("gemm: the relu epilogue comes with a bias")
    if bias is not None and (bias.numel() != n or bias.dtype != out.dtype):
        raise ValueError("gemm: bias must be [N] of out's type")
    _check(_c().mrec_rt_gemm(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                             ctypes.c_void_p(bias.data_ptr() if bias is not None else 0), m, n, k, int(trans_a), int(trans_b),
                             int(a.dtype == "float16"), int(out.dtype == "float16"), float(alpha), float(beta),
                             (2 if relu else 1) if bias is not None else 0,
                             ctypes.c_void_p(a.device.current_stream_handle())), "gemm")
    return out


class _Allocation:
    """Owner of one cudaMalloc: freed when the last buffer viewing it goes away."""

    def __init__(self, nbytes):
        self.ptr = _c().mrec_rt_malloc(nbytes)
        if not self.ptr:
            raise MemoryError("cudaMalloc(%d) failed: %s" % (nbytes, _lib.last_error()))

    def __del__(self):
        try:
            _c().mrec_rt_free(ctypes.c_void_p(self.ptr))
        except Exception:
            pass


class DeviceBuffer:
    """A contiguous device array: the subset of the tensor interface that ops.py and _lib.aot_call use."""
    __mrec_rt__ = True
    is_cuda = True

    def __init__(self, device, shape, dtype, alloc=None, offset=0):
        self.device, self.shape, self.dtype = device, tuple(int(x) for x in shape), dtype
        self.itemsize = np.dtype(_NP[dtype]).itemsize
        n = self.numel()
        self._alloc = alloc if alloc is not None else _Allocation(max(n * self.itemsize, 16))
        self._ptr = self._alloc.ptr + offset

    # ---- what aot_call reads --------------------------------------------------------------------
    def data_ptr(self):
        return self._ptr

    def dim(self):
        return len(self.shape)

    def numel(self):
        n = 1
        for d in self.shape:
            n *= d
        return n

    def element_size(self):
        return self.itemsize

    def is_contiguous(self):
        return True

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": np.dtype(_NP[self.dtype]).str, "data": (int(self._ptr), False), "version": 2}

    # ---- views ----------------------------------------------------------------------------------
    def view(self, *shape):
        shape = tuple(shape[0]) if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else tuple(shape)
        n = self.numel()
        if -1 in shape:
            known = 1
            for d in shape:
                known *= d if d != -1 else 1
            shape = tuple(n // max(known, 1) if d == -1 else d for d in shape)
        m = 1
        for d in shape:
            m *= d
        if m != n:
            raise ValueError("view%r of a buffer with %d elements" % (shape, n))
        return DeviceBuffer(self.device, shape, self.dtype, self._alloc, self._ptr - self._alloc.ptr)

    reshape = view

    def __getitem__(self, key):
        """Rows [a:b] of dim 0 (a contiguous view)."""
        if isinstance(key, int):
            key = slice(key, key + 1)
        if not isinstance(key, slice) or key.step not in (None, 1):
            raise IndexError("DeviceBuffer supports contiguous slices of dim 0 only")
        a, b, _ = key.indices(self.shape[0])
        row = self.numel() // max(self.shape[0], 1)
        return DeviceBuffer(self.device, (max(b - a, 0),) + self.shape[1:], self.dtype, self._alloc,
                            self._ptr - self._alloc.ptr + a * row * self.itemsize)

    # ---- data movement (on the device's current stream) -----------------------------------------
    def copy_(self, src):
        nbytes = self.numel() * self.itemsize
        s = ctypes.c_void_p(self.device.current_stream_handle())
        if isinstance(src, DeviceBuffer):
            if src.numel() * src.itemsize != nbytes:
                raise ValueError("copy_ between buffers of different size")
            _check(_c().mrec_rt_memcpy(ctypes.c_void_p(self._ptr), ctypes.c_void_p(src._ptr), nbytes, 3, s), "memcpy d2d")
        else:
            a = np.ascontiguousarray(src, dtype=_NP[self.dtype])
            if a.size != self.numel():
                raise ValueError("copy_ from an array of another size")
            _check(_c().mrec_rt_memcpy(ctypes.c_void_p(self._ptr), ctypes.c_void_p(a.ctypes.data), nbytes, 1, s), "memcpy h2d")
            if not self.device.async_host_copies:          # `a` may be a temporary: it must outlive the copy
                Stream(self.device.current_stream_handle()).synchronize()
        return self

    def zero_(self):
        _check(_c().mrec_rt_memset(ctypes.c_void_p(self._ptr), 0, self.numel() * self.itemsize,
                                   ctypes.c_void_p(self.device.current_stream_handle())), "memset")
        return self

    def fill_(self, value):
        if self.itemsize != 4:
            raise TypeError("fill_ handles 4-byte element types")
        pattern = int(np.asarray(value, dtype=_NP[self.dtype]).view(np.uint32))
        _check(_c().mrec_rt_fill32(ctypes.c_void_p(self._ptr), pattern, self.numel(),
                                   ctypes.c_void_p(self.device.current_stream_handle())), "fill")
        return self

    def numpy(self):
        """Synchronous copy to a new host array."""
        out = np.empty(self.shape, dtype=_NP[self.dtype])
        s = ctypes.c_void_p(self.device.current_stream_handle())
        if out.size:
            _check(_c().mrec_rt_memcpy(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(self._ptr), out.nbytes, 2, s), "memcpy d2h")
        Stream(self.device.current_stream_handle()).synchronize()
        return out

    def item(self):
        return self.numpy().reshape(-1)[0].item()

    def clone(self):
        return self.device.empty(self.shape, self.dtype).copy_(self)

    def __repr__(self):
        return "DeviceBuffer(shape=%s, dtype=%s, device=%s)" % (self.shape, self.dtype, self.device)


class Device:
    """One GPU: allocator + current stream.  ops.py treats it as the `device` of the buffers it allocates."""
    __mrec_rt__ = True

    def __init__(self, index=0):
        self.index = int(index)
        _check(_c().mrec_rt_set_device(self.index), "set_device")
        self.stream = Stream()
        self._current = self.stream
        # True: copy_(numpy array) returns at once — the caller keeps the (pinned) array alive and unchanged until the
        # stream has passed the copy (double-buffered input pipelines)
        self.async_host_copies = False

    def __hash__(self):
        return hash(("mrec_rt", self.index))

    def __eq__(self, other):
        return isinstance(other, Device) and other.index == self.index

    def __repr__(self):
        return "mrec_rt:%d" % self.index

    def current_stream_handle(self):
        return self._current.handle

    @contextlib.contextmanager
    def use_stream(self, stream):
        prev, self._current = self._current, stream
        try:
            yield stream
        finally:
            self._current = prev

    def synchronize(self):
        _check(_c().mrec_rt_device_sync(), "device sync")

    def mem_info(self):
        f, t = ctypes.c_size_t(0), ctypes.c_size_t(0)
        _check(_c().mrec_rt_mem_info(ctypes.byref(f), ctypes.byref(t)), "mem_info")
        return f.value, t.value

    # ---- allocation -------------------------------------------------------------------------------
    def empty(self, shape, dtype="float32"):
        shape = tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)
        return DeviceBuffer(self, shape, dtype)

    def zeros(self, shape, dtype="float32"):
        return self.empty(shape, dtype).zero_()

    def full(self, shape, value, dtype="float32"):
        return self.empty(shape, dtype).fill_(value)

    def tensor(self, values, dtype="float32"):
        return self.from_numpy(np.asarray(values, dtype=_NP[dtype]))

    def from_numpy(self, a):
        a = np.ascontiguousarray(a)
        name = {v: k for k, v in _NP.items()}[a.dtype.type]
        return self.empty(a.shape, name).copy_(a)

    def pinned(self, shape, dtype="float32"):
        """A page-locked host array (numpy view of cudaMallocHost memory) for asynchronous H2D / D2H copies."""
        shape = tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)
        n = int(np.prod(shape)) * np.dtype(_NP[dtype]).itemsize
        p = _c().mrec_rt_malloc_host(max(n, 16))
        if not p:
            raise MemoryError("cudaMallocHost failed: " + _lib.last_error())
        buf = (ctypes.c_char * max(n, 16)).from_address(p)
        return np.frombuffer(buf, dtype=_NP[dtype], count=int(np.prod(shape))).reshape(shape)

    # ---- graphs -----------------------------------------------------------------------------------
    @contextlib.contextmanager
    def capture(self, stream=None):
        """Record every aot call issued inside the block (on the device's current stream) into one CUDA graph."""
        s = stream or self._current
        holder = Graph(None, self)
        _check(_c().mrec_rt_graph_begin(ctypes.c_void_p(s.handle)), "graph begin")
        try:
            yield holder
        finally:
            h = _c().mrec_rt_graph_end(ctypes.c_void_p(s.handle))
            if not h:
                raise RuntimeError("graph capture failed: " + _lib.last_error())
            holder.handle = h
