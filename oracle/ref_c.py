"""ctypes binding of oracle/ref_c.c (TEST / BASELINE INFRASTRUCTURE ONLY, see that file's header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmrec_ref.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int32)
_i64, _int, _flt = ctypes.c_int64, ctypes.c_int, ctypes.c_float


def build(force=False):
    src = os.path.join(_HERE, "ref_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        # -march=native is decided on the machine that runs it: rebuild there if the ISA differs
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", src,
                               "-o", _SO, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        try:
            _lib = ctypes.CDLL(build())
            _lib.mrec_ref_threads()
        except OSError:
            _lib = ctypes.CDLL(build(force=True))
        L = _lib
        L.mrec_ref_threads.restype = _int
        L.mrec_ref_gather_masked.argtypes = [_f, _i64, _int, _i, _f, _i64, _f]
        L.mrec_ref_gather_reduce.argtypes = [_f, _i64, _i, _f, _i64, _int, _flt, _f]
        L.mrec_ref_unique.argtypes = [_i, _i64, _i64, _i, _i, _i, _i]
        L.mrec_ref_unique.restype = _i64
        L.mrec_ref_segment_sum.argtypes = [_f, _int, _int, _f, _i, _i, _i64, _f]
        L.mrec_ref_lazy_adam.argtypes = [_f, _f, _f, _i64, _int, _i, _i64, _f, _flt, _flt, _flt, _flt, _flt]
        L.mrec_ref_adam_dense.argtypes = [_f, _f, _f, _f, _i64, _flt, _flt, _flt, _flt, _flt]
        L.mrec_ref_ftrl.argtypes = [_f, _f, _f, _i64, _int, _i, _i64, _f, _flt, _flt, _flt, _flt]
        L.mrec_ref_fm_fwd.argtypes = [_f, _i64, _int, _int, _f]
        L.mrec_ref_fm_bwd.argtypes = [_f, _f, _i64, _int, _int, _f]
        L.mrec_ref_cross_fwd.argtypes = [_f, _f, _f, _i64, _int, _int, _f, _f]
    return _lib


def _p(a, t=_f):
    return a.ctypes.data_as(t) if a is not None else None


def threads():
    return lib().mrec_ref_threads()


def use_all_cores():
    """Run the port (OpenMP loops AND numpy's BLAS) on every host core whatever OMP_NUM_THREADS says (torchrun sets
    it to 1).  Returns the thread count in force."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    L = lib()
    L.mrec_ref_set_threads.argtypes = [_int]
    L.mrec_ref_set_threads.restype = _int
    got = L.mrec_ref_set_threads(n)
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
    except Exception:
        pass
    return got


def gather_masked(table, ids, mask):
    n = ids.size
    dim = table.shape[1]
    out = np.empty((ids.shape[0], n // ids.shape[0] * dim), dtype=np.float32)
    lib().mrec_ref_gather_masked(_p(table), table.shape[0], dim, _p(ids, _i), _p(mask), n, _p(out))
    return out


def gather_reduce(table, ids, mask, bias):
    out = np.empty((ids.shape[0], 1), dtype=np.float32)
    lib().mrec_ref_gather_reduce(_p(table), table.shape[0], _p(ids, _i), _p(mask), ids.shape[0],
                                 ids.shape[1], float(bias), _p(out))
    return out


def unique(ids, bound):
    flat = np.ascontiguousarray(ids.reshape(-1), dtype=np.int32)
    n = flat.size
    uniq = np.empty(n, np.int32); inverse = np.empty(n, np.int32)
    perm = np.empty(n, np.int32); seg_start = np.empty(n + 1, np.int32)
    u = lib().mrec_ref_unique(_p(flat, _i), n, bound, _p(uniq, _i), _p(inverse, _i), _p(perm, _i),
                              _p(seg_start, _i))
    return uniq[:u], inverse, perm, seg_start[:u + 1]


def segment_sum(g, dim, div, mask, perm, seg_start):
    u = seg_start.size - 1
    out = np.empty((u, dim), dtype=np.float32)
    lib().mrec_ref_segment_sum(_p(g), dim, div, _p(mask), _p(perm, _i), _p(seg_start, _i), u, _p(out))
    return out


def lazy_adam(w, m, v, uniq, gsum, lr_t, b1, b2, eps, scale):
    lib().mrec_ref_lazy_adam(_p(w), _p(m), _p(v), w.shape[0], w.shape[1], _p(uniq, _i), uniq.size,
                             _p(gsum), lr_t, b1, b2, eps, scale)


def adam_dense(w, m, v, g, lr_t, b1, b2, eps, scale):
    lib().mrec_ref_adam_dense(_p(w), _p(m), _p(v), _p(g), w.size, lr_t, b1, b2, eps, scale)


def ftrl(w, acc, lin, uniq, gsum, lr, l1, l2, scale):
    dim = w.shape[1] if w.ndim == 2 else 1
    lib().mrec_ref_ftrl(_p(w), _p(acc), _p(lin), w.shape[0], dim, _p(uniq, _i), uniq.size, _p(gsum), lr,
                        l1, l2, scale)


def fm_fwd(vx):
    b, f, d = vx.shape
    out = np.empty((b, 1), np.float32)
    lib().mrec_ref_fm_fwd(_p(vx), b, f, d, _p(out))
    return out


def fm_bwd(vx, gout):
    b, f, d = vx.shape
    out = np.empty_like(vx)
    lib().mrec_ref_fm_bwd(_p(vx), _p(np.ascontiguousarray(gout, np.float32)), b, f, d, _p(out))
    return out


def cross_fwd(x0, w, b):
    n, dp = x0.shape
    y = np.empty_like(x0)
    s = np.empty((n, w.shape[0]), np.float32)
    lib().mrec_ref_cross_fwd(_p(x0), _p(w), _p(b), n, dp, w.shape[0], _p(y), _p(s))
    return y, s


class WideDeepCpu:
    """The reference's Wide&Deep step on host cores: embedding path in C/OpenMP, DenseLayers through
    numpy's BLAS (fp32).  Same math as ref_numpy.WideDeepOracle(mode='lazy'), fp32 throughout."""

    def __init__(self, vocab, dim, hidden=(1024, 512, 256, 128), fields=39, sens=1024.0, seed=0):
        rng = np.random.default_rng(seed)
        self.ww = (rng.standard_normal((vocab, 1), dtype=np.float32) * 0.01)
        self.wd = (rng.standard_normal((vocab, dim), dtype=np.float32) * 0.01)
        self.acc = np.ones_like(self.ww); self.lin = np.zeros_like(self.ww)
        self.md = np.zeros_like(self.wd); self.vd = np.zeros_like(self.wd)
        dims = [fields * dim] + list(hidden) + [1]
        self.w = [(rng.standard_normal((dims[i], dims[i + 1]), dtype=np.float32) * 0.01) for i in range(len(dims) - 1)]
        self.b = [np.zeros(dims[i + 1], np.float32) for i in range(len(dims) - 1)]
        self.mw = [np.zeros_like(x) for x in self.w]; self.vw = [np.zeros_like(x) for x in self.w]
        self.mb = [np.zeros_like(x) for x in self.b]; self.vb = [np.zeros_like(x) for x in self.b]
        self.wide_b = np.zeros(1, np.float32); self.mwb = np.zeros(1, np.float32); self.vwb = np.zeros(1, np.float32)
        self.sens = np.float32(sens)
        self.b1p = np.float32(1); self.b2p = np.float32(1)
        self.dim, self.fields = dim, fields

    def step(self, ids, wts, label):
        b, f = ids.shape
        wide = gather_reduce(self.ww, ids, wts, self.wide_b[0])
        x = gather_masked(self.wd, ids, wts)
        acts = [x]
        h = x
        for i in range(len(self.w)):
            a = h @ self.w[i] + self.b[i]
            h = np.maximum(a, 0, out=a) if i + 1 < len(self.w) else a
            acts.append(h)
        logit = wide + h
        loss = np.mean(np.maximum(logit, 0) - logit * label + np.log1p(np.exp(-np.abs(logit))))
        delta = ((1.0 / (1.0 + np.exp(-logit)) - label) * (self.sens / b)).astype(np.float32)
        g = delta
        gw, gb = [None] * len(self.w), [None] * len(self.w)
        for i in range(len(self.w) - 1, -1, -1):
            if i + 1 < len(self.w):
                g = g * (acts[i + 1] > 0)
            gw[i] = acts[i].T @ g
            gb[i] = g.sum(axis=0)
            g = g @ self.w[i].T
        gx = np.ascontiguousarray(g.reshape(b * f, self.dim), dtype=np.float32)
        mask = np.ascontiguousarray(wts.reshape(-1))
        uniq, _, perm, seg_start = unique(ids, self.wd.shape[0])
        gs_w = segment_sum(np.ascontiguousarray(delta), 1, f, mask, perm, seg_start)
        gs_d = segment_sum(gx, self.dim, 1, mask, perm, seg_start)
        scale = np.float32(1) / self.sens
        ftrl(self.ww, self.acc, self.lin, uniq, gs_w, 5e-2, 1e-8, 1e-8, scale)
        self.b1p *= np.float32(0.9); self.b2p *= np.float32(0.999)
        lr_t = np.float32(3.5e-4) * np.sqrt(np.float32(1) - self.b2p) / (np.float32(1) - self.b1p)
        lazy_adam(self.wd, self.md, self.vd, uniq, gs_d, lr_t, 0.9, 0.999, 1e-8, scale)
        for i in range(len(self.w)):
            adam_dense(self.w[i], self.mw[i], self.vw[i], np.ascontiguousarray(gw[i], np.float32), lr_t, 0.9, 0.999, 1e-8, scale)
            adam_dense(self.b[i], self.mb[i], self.vb[i], np.ascontiguousarray(gb[i], np.float32), lr_t, 0.9, 0.999, 1e-8, scale)
        adam_dense(self.wide_b, self.mwb, self.vwb, np.array([delta.sum()], np.float32), lr_t, 0.9, 0.999, 1e-8, scale)
        return float(loss)
