"""CPU oracle for the MindRec embedding-and-interaction hot path (numpy).

TEST INFRASTRUCTURE ONLY.  Nothing under ``mindrec_b200/`` imports this module; it may be imported by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` and nowhere else.

PARITY UNPINNED.  The reference repository (mindspore-lab/mindrec) contains no kernels and no numerical
tests for this path: every arithmetic op is delegated to the unpinned third-party dependency
``mindspore`` (requirements/cpu_requirements.txt:3; >= 2.0 is needed for
mindspore_rec/ops/embedding.py:22,27), which is neither vendored under /root/reference nor installable
here.  This file therefore restates (a) the reference's own model code line by line and (b) the published
semantics of the MindSpore ops that code calls (listed in oracle/ASSUMPTIONS.md, one testable item
each).  The pins we can offer are algebraic identities and finite-difference checks in
tests/test_oracle.py plus frozen seeded vectors in tests/golden/.

Every function cites the reference file:line it follows.  Accumulations that the tolerance contract
(1e-5 relative, fp32) is judged against are carried in float64 and rounded once at the end; index /
key / gathered-row outputs are exact.
"""
import numpy as np

F32 = np.float32


# ------------------------------------------------------------------------------------------------
# a1/a2  gather, mask multiply, wide reduce
# ------------------------------------------------------------------------------------------------
def gather(table, ids):
    """P.Gather(table, ids, 0) — models/wide_deep/src/wide_and_deep.py:300-302 via nn.EmbeddingLookup,
    models/deepfm/src/deepfm.py:217,221, models/deep_and_cross/src/deep_and_cross.py:199.
    Out-of-range ids give a zero row (GPU Gather semantics, ASSUMPTIONS B2)."""
    table = np.asarray(table)
    t2 = table.reshape(table.shape[0], -1)
    ids = np.asarray(ids)
    flat = ids.reshape(-1).astype(np.int64)
    ok = (flat >= 0) & (flat < t2.shape[0])
    out = np.zeros((flat.size, t2.shape[1]), dtype=table.dtype)
    out[ok] = t2[flat[ok]]
    return out.reshape(ids.shape + (t2.shape[1],))


def gather_masked(table, ids, mask):
    """deep_in = reshape(table[ids] * mask[..., None], (B, F*D)) — wide_and_deep.py:303,308-309;
    deepfm.py:215,222,230; deep_and_cross.py:295-298.  One fp32 multiply per element: exact."""
    e = gather(table, ids)
    out = e * np.asarray(mask, dtype=F32)[..., None]
    return out.reshape(ids.shape[0], -1).astype(F32)


def gather_reduce(table, ids, mask, bias=None):
    """wide_out = sum_f table[ids] * mask + Wide_b — wide_and_deep.py:300,305-306; deepfm.py:217-219."""
    e = gather(np.asarray(table).reshape(-1, 1), ids)[..., 0].astype(np.float64)
    s = (e * np.asarray(mask, dtype=np.float64)).sum(axis=1)
    if bias is not None:
        s = s + np.float64(np.asarray(bias).reshape(-1)[0])
    return s.astype(F32).reshape(-1, 1)


# ------------------------------------------------------------------------------------------------
# a4  unique (both orders) and deterministic segment-sum
# ------------------------------------------------------------------------------------------------
def unique_sorted(ids, bound=None):
    """P.Unique, GPU order (ascending) — mindspore_rec/ops/embedding.py:191-192; ASSUMPTIONS B3.
    Returns (uniq, inverse int32, perm int32 = stable argsort, seg_start int32[U+1]).
    bound: ids outside [0, bound) collapse onto the value `bound` (mrec_unique_bounded)."""
    flat = np.asarray(ids).reshape(-1)
    if bound is not None:
        flat = np.where((flat >= 0) & (flat < bound), flat, bound).astype(flat.dtype)
    perm = np.argsort(flat, kind="stable").astype(np.int32)
    s = flat[perm]
    head = np.ones(s.size, dtype=bool)
    head[1:] = s[1:] != s[:-1]
    uniq = s[head]
    seg_of = np.cumsum(head) - 1
    inverse = np.empty(flat.size, dtype=np.int32)
    inverse[perm] = seg_of.astype(np.int32)
    seg_start = np.concatenate([np.nonzero(head)[0], [flat.size]]).astype(np.int32)
    return uniq, inverse, perm, seg_start


def unique_first(ids):
    """P.Unique, CPU order (first occurrence) — ASSUMPTIONS B3; docs example [1,2,5,2] -> ([1,2,5],[0,1,2,1])."""
    flat = np.asarray(ids).reshape(-1)
    seen = {}
    uniq = []
    inverse = np.empty(flat.size, dtype=np.int32)
    for i, k in enumerate(flat.tolist()):
        j = seen.get(k)
        if j is None:
            j = len(uniq)
            seen[k] = j
            uniq.append(k)
        inverse[i] = j
    return np.asarray(uniq, dtype=flat.dtype), inverse


def segment_sum(values, inverse, num_segments, mask=None, div=1):
    """UnsortedSegmentSum(values, inverse, U) — the RowTensor dedup of ASSUMPTIONS B4 (float64 accumulate).
    values[n // div] is the row of lookup position n; mask[n] scales it (the bprop of the mask Mul,
    wide_and_deep.py:307-308)."""
    values = np.asarray(values)
    v2 = values.reshape(values.shape[0], -1) if values.ndim > 1 else values.reshape(-1, 1)
    n = np.asarray(inverse).size
    rows = v2[np.arange(n) // div].astype(np.float64)
    if mask is not None:
        rows = rows * np.asarray(mask, dtype=np.float64).reshape(-1, 1)
    out = np.zeros((num_segments, v2.shape[1]), dtype=np.float64)
    np.add.at(out, np.asarray(inverse, dtype=np.int64), rows)
    return out


# ------------------------------------------------------------------------------------------------
# a5/a6  optimizers (ASSUMPTIONS B4-B7)
# ------------------------------------------------------------------------------------------------
class AdamState:
    """nn.Adam / nn.LazyAdam state — wide_and_deep.py:420-422,435-437 (lr 3.5e-4, eps 1e-8, loss_scale sens)."""

    def __init__(self, lr, beta1=0.9, beta2=0.999, eps=1e-8, loss_scale=1.0):
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.grad_scale = F32(1.0) / F32(loss_scale)
        self.beta1_power = F32(1.0)
        self.beta2_power = F32(1.0)
        self.lr_t = F32(0.0)

    def begin_step(self):
        """beta powers are multiplied before use (step t uses beta^t); lr_t = lr*sqrt(1-b2^t)/(1-b1^t)."""
        self.beta1_power = F32(self.beta1_power * F32(self.beta1))
        self.beta2_power = F32(self.beta2_power * F32(self.beta2))
        self.lr_t = F32(F32(self.lr) * np.sqrt(F32(1) - self.beta2_power) / (F32(1) - self.beta1_power))


def adam_rows(w, m, v, g, st):
    """m = b1*m + (1-b1)g; v = b2*v + (1-b2)g^2; w -= lr_t*m/(sqrt(v)+eps)  (float64 math, fp32 state)."""
    g = np.asarray(g, dtype=np.float64) * np.float64(st.grad_scale)
    # beta and (1 - beta) are fp32 scalars in the upstream kernels (1 - 0.999f != 0.001 to 1.3e-5)
    b1, b2 = np.float64(F32(st.beta1)), np.float64(F32(st.beta2))
    omb1, omb2 = np.float64(F32(1) - F32(st.beta1)), np.float64(F32(1) - F32(st.beta2))
    m64 = b1 * m.astype(np.float64) + omb1 * g
    v64 = b2 * v.astype(np.float64) + omb2 * g * g
    w64 = w.astype(np.float64) - np.float64(st.lr_t) * m64 / (np.sqrt(v64) + st.eps)
    return w64.astype(F32), m64.astype(F32), v64.astype(F32)


def lazy_adam_sparse(w, m, v, uniq, gsum, st):
    """nn.LazyAdam with a deduplicated RowTensor gradient: only rows in `uniq` move (ASSUMPTIONS B6)."""
    uniq = np.asarray(uniq, dtype=np.int64)
    ok = (uniq >= 0) & (uniq < w.shape[0])
    r = uniq[ok]
    w[r], m[r], v[r] = adam_rows(w[r], m[r], v[r], np.asarray(gsum)[ok], st)


def adam_dense(w, m, v, g, st):
    """nn.Adam dense kernel (ASSUMPTIONS B5) — in place."""
    w[...], m[...], v[...] = adam_rows(w, m, v, g, st)


class FtrlState:
    """nn.FTRL hyper-parameters — wide_and_deep.py:423-430 (lr 5e-2, l1=l2=1e-8, initial_accum 1.0)."""

    def __init__(self, lr, l1=0.0, l2=0.0, lr_power=-0.5, loss_scale=1.0):
        self.lr, self.l1, self.l2, self.lr_power = lr, l1, l2, lr_power
        self.grad_scale = F32(1.0) / F32(loss_scale)


def ftrl_rows(w, acc, lin, g, st):
    """ApplyFtrl (ASSUMPTIONS B7): a'=a+g^2; sigma=(a'^-p - a^-p)/lr; lin+=g-sigma*w;
    w = |lin|>l1 ? (sign(lin)*l1-lin)/(a'^-p/lr+2*l2) : 0."""
    g = np.asarray(g, dtype=np.float64) * np.float64(st.grad_scale)
    a = acc.astype(np.float64)
    a_new = a + g * g
    p = -st.lr_power
    pa_new, pa_old = np.power(a_new, p), np.power(a, p)
    sigma = (pa_new - pa_old) / st.lr
    lin64 = lin.astype(np.float64) + g - sigma * w.astype(np.float64)
    q = pa_new / st.lr + 2.0 * st.l2
    w64 = np.where(np.abs(lin64) > st.l1, (np.sign(lin64) * st.l1 - lin64) / q, 0.0)
    return w64.astype(F32), a_new.astype(F32), lin64.astype(F32)


def ftrl_sparse(w, acc, lin, uniq, gsum, st):
    """SparseApplyFtrl / FusedSparseFtrl: identical math on the touched (deduplicated) rows only."""
    uniq = np.asarray(uniq, dtype=np.int64)
    ok = (uniq >= 0) & (uniq < w.shape[0])
    r = uniq[ok]
    w[r], acc[r], lin[r] = ftrl_rows(w[r], acc[r], lin[r], np.asarray(gsum)[ok].reshape(w[r].shape), st)


def ftrl_dense(w, acc, lin, g, st):
    w[...], acc[...], lin[...] = ftrl_rows(w, acc, lin, g, st)


# ------------------------------------------------------------------------------------------------
# a8  loss
# ------------------------------------------------------------------------------------------------
def sigmoid_xent(logit, label):
    """SigmoidCrossEntropyWithLogits = max(x,0) - x*z + log1p(exp(-|x|)) — wide_and_deep.py:354 (B11)."""
    x = np.asarray(logit, dtype=np.float64)
    z = np.asarray(label, dtype=np.float64)
    return np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))


def sigmoid(x):
    x = np.asarray(x, dtype=np.float64)
    return 1.0 / (1.0 + np.exp(-x))


# ------------------------------------------------------------------------------------------------
# a11  DeepFM second-order FM  — models/deepfm/src/deepfm.py:222-228
# ------------------------------------------------------------------------------------------------
def fm_forward(vx):
    """fm = 0.5 * sum_d[(sum_f vx)^2 - sum_f vx^2]; vx is [B,F,D] (already multiplied by the mask)."""
    vx = np.asarray(vx, dtype=np.float64)
    v1 = np.square(vx.sum(axis=1))
    v2 = np.square(vx).sum(axis=1)
    return (0.5 * (v1 - v2).sum(axis=1)).reshape(-1, 1)


def fm_backward(vx, gout):
    """d fm / d vx[b,f,d] = g[b] * (S[b,d] - vx[b,f,d]),  S = sum_f vx."""
    vx = np.asarray(vx, dtype=np.float64)
    s = vx.sum(axis=1, keepdims=True)
    return np.asarray(gout, dtype=np.float64).reshape(-1, 1, 1) * (s - vx)


def fm_pairwise(vx):
    """Independent statement of the same quantity: sum_{i<j} <v_i, v_j> (used only to pin fm_forward)."""
    vx = np.asarray(vx, dtype=np.float64)
    b, f, _ = vx.shape
    out = np.zeros((b, 1))
    for i in range(f):
        for j in range(i + 1, f):
            out[:, 0] += (vx[:, i] * vx[:, j]).sum(axis=1)
    return out


# ------------------------------------------------------------------------------------------------
# a13  DCN cross stack — models/deep_and_cross/src/deep_and_cross.py:139-149 (x6: 301-306)
# ------------------------------------------------------------------------------------------------
def cross_forward(x0, w, b):
    """x_{l+1} = x_0 * (x_l . w_l) + b_l + x_l, layer by layer exactly as CrossLayer.construct.
    w, b: [L, D'].  Returns (x_L, [x_0..x_{L-1}], s[B,L])."""
    x0 = np.asarray(x0, dtype=np.float64)
    xl = x0
    xs, ss = [], []
    for l in range(w.shape[0]):
        s = xl @ np.asarray(w[l], dtype=np.float64)
        xs.append(xl)
        ss.append(s)
        xl = x0 * s[:, None] + np.asarray(b[l], dtype=np.float64)[None, :] + xl
    return xl, xs, np.stack(ss, axis=1) if ss else np.zeros((x0.shape[0], 0))


def cross_backward(x0, w, b, gy):
    """Reverse-mode through the L layers: returns (dx0_total, dw[L,D'], db[L,D'])."""
    x0 = np.asarray(x0, dtype=np.float64)
    _, xs, ss = cross_forward(x0, w, b)
    g = np.asarray(gy, dtype=np.float64)
    dx0 = np.zeros_like(x0)
    dw = np.zeros(w.shape, dtype=np.float64)
    db = np.zeros(b.shape, dtype=np.float64)
    for l in range(w.shape[0] - 1, -1, -1):
        db[l] = g.sum(axis=0)
        ds = (g * x0).sum(axis=1)
        dx0 += g * ss[:, l][:, None]
        dw[l] = (ds[:, None] * xs[l]).sum(axis=0)
        g = g + ds[:, None] * np.asarray(w[l], dtype=np.float64)[None, :]
    return g + dx0, dw, db


# ------------------------------------------------------------------------------------------------
# a9  MapParameter (hash) model — README.md:155-205, mindspore_rec/ops/embedding.py:136-149 (B9)
# ------------------------------------------------------------------------------------------------
class MapParameterModel:
    """dict-backed model of mindspore.experimental.MapParameter with permit / evict filters."""

    def __init__(self, dim, default_value=0.0, permit_filter_value=1, evict_filter_value=None):
        self.dim = dim
        self.default = np.full((dim,), default_value, dtype=F32) if np.isscalar(default_value) \
            else np.asarray(default_value, dtype=F32)
        self.permit = permit_filter_value
        self.evict_after = evict_filter_value
        self.rows = {}       # key -> row (resident)
        self.seen = {}       # key -> number of steps in which it was looked up
        self.last = {}       # key -> last step it was looked up
        self.step = 0
        self._touched = set()   # keys looked up / put since the last export (incremental export, ASSUMPTIONS)
        self._erased = set()    # keys erased / evicted since the last export

    def get(self, keys, insert_default_value=True):
        """MapTensorGet: one sighting per distinct key per call; a key becomes resident on its
        `permit`-th sighting; before that the default row is returned and nothing is stored."""
        keys = np.asarray(keys).reshape(-1)
        self.step += 1
        out = np.empty((keys.size, self.dim), dtype=F32)
        touched = set()
        for i, k in enumerate(keys.tolist()):
            if k not in touched:
                touched.add(k)
                self.seen[k] = self.seen.get(k, 0) + 1
                self.last[k] = self.step
                self._touched.add(k)
                if k not in self.rows and insert_default_value and self.seen[k] >= self.permit:
                    self.rows[k] = self.default.copy()
            out[i] = self.rows[k] if k in self.rows else self.default
        return out

    def put(self, keys, values):
        for k, v in zip(np.asarray(keys).reshape(-1).tolist(), np.asarray(values, dtype=F32)):
            self.rows[k] = v.copy()
            self.last[k] = self.step
            self.seen[k] = max(self.seen.get(k, 0), self.permit)
            self._touched.add(k)

    def erase(self, keys):
        for k in np.asarray(keys).reshape(-1).tolist():
            if k in self.rows or k in self.seen:
                self._erased.add(k)
            self.rows.pop(k, None)
            self.seen.pop(k, None)
            self.last.pop(k, None)

    def evict(self):
        """Drop every key not looked up for more than `evict_filter_value` steps."""
        if self.evict_after is None:
            return
        for k in [k for k, s in self.last.items() if self.step - s > self.evict_after]:
            self.erase([k])

    def keys(self):
        return np.asarray(sorted(self.rows.keys()), dtype=np.int64)

    def export_data(self, incremental=False):
        """(keys, values, statuses): full = every resident key, status 0; incremental = resident keys looked up or put
        since the previous export (status 1) + keys erased since then and not resident again (status 2, zero rows)."""
        if incremental:
            mod = sorted(k for k in self._touched if k in self.rows)
            gone = sorted(k for k in self._erased if k not in self.rows)
            keys = np.asarray(mod + gone, dtype=np.int64)
            vals = np.stack([self.rows[k] for k in mod] + [np.zeros(self.dim, F32)] * len(gone)) if keys.size \
                else np.zeros((0, self.dim), F32)
            status = np.asarray([1] * len(mod) + [2] * len(gone), dtype=np.int32)
        else:
            keys = self.keys()
            vals = np.stack([self.rows[k] for k in keys.tolist()]) if keys.size else np.zeros((0, self.dim), F32)
            status = np.zeros(keys.size, dtype=np.int32)
        self._touched.clear()
        self._erased.clear()
        return keys, vals, status

    def import_data(self, data):
        keys, vals, status = data
        gone = status == 2
        self.erase(keys[gone])
        self.put(keys[~gone], vals[~gone])


# ------------------------------------------------------------------------------------------------
# Appendix D of SURVEY.md: one full Wide&Deep training step
# ------------------------------------------------------------------------------------------------
class WideDeepOracle:
    """models/wide_deep/src/wide_and_deep.py:293-316 (forward), :349-362 (losses), :472-492 (step), fp32
    DenseLayers (use_mixed_precision=False).  State arrays are float32, arithmetic is float64.

    mode: "lazy"  LazyAdam on the deep table (sparse + auto-parallel / PS / dynamic embedding, :415-422)
          "adam"  nn.Adam with a RowTensor gradient = dense-equivalent update (sparse=True, one device)
          "dense" sparse=False: dense gradient + l2_coef * sum(Wd^2)/2 in deep_loss (:359-360)."""

    def __init__(self, wide_w, deep_w, mlp_w, mlp_b, wide_b, sens=1024.0, mode="lazy", l2_coef=8e-5):
        self.ww = wide_w.astype(F32).copy()
        self.wd = deep_w.astype(F32).copy()
        self.mlp_w = [w.astype(F32).copy() for w in mlp_w]
        self.mlp_b = [b.astype(F32).copy() for b in mlp_b]
        self.wide_b = np.asarray(wide_b, dtype=F32).reshape(1).copy()
        self.sens = sens
        self.mode = mode
        self.l2_coef = l2_coef
        self.adam = AdamState(3.5e-4, eps=1e-8, loss_scale=sens)
        self.ftrl = FtrlState(5e-2, l1=1e-8, l2=1e-8, loss_scale=sens)
        self.acc = np.full_like(self.ww, 1.0)
        self.lin = np.zeros_like(self.ww)
        self.md, self.vd = np.zeros_like(self.wd), np.zeros_like(self.wd)
        self.m_w = [np.zeros_like(w) for w in self.mlp_w]
        self.v_w = [np.zeros_like(w) for w in self.mlp_w]
        self.m_b = [np.zeros_like(b) for b in self.mlp_b]
        self.v_b = [np.zeros_like(b) for b in self.mlp_b]
        self.m_wb, self.v_wb = np.zeros(1, F32), np.zeros(1, F32)

    def forward(self, ids, wts):
        b, f = ids.shape
        wide = gather_reduce(self.ww, ids, wts, self.wide_b).astype(np.float64)
        x = gather_masked(self.wd, ids, wts).astype(np.float64)
        acts = [x]
        h = x
        for i, (w, bb) in enumerate(zip(self.mlp_w, self.mlp_b)):
            a = h @ w.astype(np.float64) + bb.astype(np.float64)
            h = np.maximum(a, 0) if i + 1 < len(self.mlp_w) else a
            acts.append(h)
        return wide + h, acts

    def step(self, ids, wts, label):
        b, f = ids.shape
        d = self.wd.shape[1]
        logit, acts = self.forward(ids, wts)
        loss = sigmoid_xent(logit, label).mean()
        loss_d = loss
        if self.mode == "dense":
            loss_d = loss + self.l2_coef * np.square(self.wd.astype(np.float64)).sum() / 2
        delta = self.sens * (sigmoid(logit) - label) / b                      # [B,1]
        g = delta
        gw, gb = [None] * len(self.mlp_w), [None] * len(self.mlp_w)
        for i in range(len(self.mlp_w) - 1, -1, -1):
            if i + 1 < len(self.mlp_w):
                g = g * (acts[i + 1] > 0)
            gw[i] = acts[i].T @ g
            gb[i] = g.sum(axis=0)
            g = g @ self.mlp_w[i].astype(np.float64).T
        gx = g.reshape(b * f, d)
        mask = np.asarray(wts, dtype=np.float64).reshape(-1)
        uniq, inverse, _, _ = unique_sorted(ids, bound=self.wd.shape[0])
        gsum_w = segment_sum(delta, inverse, uniq.size, mask, div=f)
        gsum_d = segment_sum(gx, inverse, uniq.size, mask)
        # optimizer_w = FTRL on the wide table only (Wide_b is in the Adam group, :405-413)
        ftrl_sparse(self.ww, self.acc, self.lin, uniq, gsum_w, self.ftrl)
        self.adam.begin_step()
        if self.mode == "lazy":
            lazy_adam_sparse(self.wd, self.md, self.vd, uniq, gsum_d, self.adam)
        else:
            gd = np.zeros(self.wd.shape, dtype=np.float64)
            ok = uniq < self.wd.shape[0]
            gd[uniq[ok]] = gsum_d[ok]
            if self.mode == "dense":
                gd += self.sens * self.l2_coef * self.wd.astype(np.float64)
            adam_dense(self.wd, self.md, self.vd, gd, self.adam)
        for i in range(len(self.mlp_w)):
            adam_dense(self.mlp_w[i], self.m_w[i], self.v_w[i], gw[i], self.adam)
            adam_dense(self.mlp_b[i], self.m_b[i], self.v_b[i], gb[i], self.adam)
        adam_dense(self.wide_b, self.m_wb, self.v_wb, np.array([delta.sum()]), self.adam)
        return F32(loss), F32(loss_d)


def _mlp_forward(x, ws, bs, last_activation=False):
    acts = [x]
    h = x
    for i, (w, b) in enumerate(zip(ws, bs)):
        a = h @ w.astype(np.float64) + b.astype(np.float64)
        h = np.maximum(a, 0) if (i + 1 < len(ws) or last_activation) else a
        acts.append(h)
    return h, acts


def _mlp_backward(g, acts, ws, last_activation=False):
    gw, gb = [None] * len(ws), [None] * len(ws)
    for i in range(len(ws) - 1, -1, -1):
        if i + 1 < len(ws) or last_activation:
            g = g * (acts[i + 1] > 0)
        gw[i] = acts[i].T @ g
        gb[i] = g.sum(axis=0)
        g = g @ ws[i].astype(np.float64).T
    return g, gw, gb


def _scatter_rows(shape, ids, rows, mask, div=1):
    """Dense gradient of P.Gather: UnsortedSegmentSum(mask * rows, ids, V) (sparse=False)."""
    out = np.zeros(shape, dtype=np.float64)
    flat = np.asarray(ids).reshape(-1).astype(np.int64)
    r = np.asarray(rows, dtype=np.float64).reshape(-1, shape[1])[np.arange(flat.size) // div]
    np.add.at(out, flat, r * np.asarray(mask, dtype=np.float64).reshape(-1, 1))
    return out


class DeepFMOracle:
    """models/deepfm/src/deepfm.py:208-237 (forward), :252-260 (loss with full-table l2), :263-297 (one
    nn.Adam over everything, dense gradients).  fp32 DenseLayers (convert_dtype=False)."""

    def __init__(self, fm_w, fm_v, mlp_w, mlp_b, lr=5e-4, eps=5e-8, l2_coef=8e-5, sens=1024.0):
        self.w = fm_w.astype(F32).copy()
        self.v = fm_v.astype(F32).copy()
        self.mlp_w = [x.astype(F32).copy() for x in mlp_w]
        self.mlp_b = [x.astype(F32).copy() for x in mlp_b]
        self.l2, self.sens = l2_coef, sens
        self.adam = AdamState(lr, eps=eps, loss_scale=sens)
        z = np.zeros_like
        self.m = {k: z(a) for k, a in (("w", self.w), ("v", self.v))}
        self.s = {k: z(a) for k, a in (("w", self.w), ("v", self.v))}
        self.mm = [(z(a), z(a)) for a in self.mlp_w]
        self.mb = [(z(a), z(a)) for a in self.mlp_b]

    def step(self, ids, wts, label):
        b, f = ids.shape
        d = self.v.shape[1]
        linear = gather_reduce(self.w, ids, wts).astype(np.float64)
        vx = gather_masked(self.v, ids, wts).astype(np.float64)
        fm = fm_forward(vx.reshape(b, f, d))
        deep, acts = _mlp_forward(vx, self.mlp_w, self.mlp_b)
        logit = linear + fm + deep
        loss = sigmoid_xent(logit, label).mean() + self.l2 * 0.5 * (
            np.square(self.v.astype(np.float64)).sum() + np.square(self.w.astype(np.float64)).sum())
        delta = self.sens * (sigmoid(logit) - label) / b
        gx, gw, gb = _mlp_backward(delta, acts, self.mlp_w)
        dvx = fm_backward(vx.reshape(b, f, d), delta).reshape(b * f, d) + gx.reshape(b * f, d)
        g_w = _scatter_rows(self.w.shape, ids, delta, wts, div=f) + self.sens * self.l2 * self.w.astype(np.float64)
        g_v = _scatter_rows(self.v.shape, ids, dvx, wts) + self.sens * self.l2 * self.v.astype(np.float64)
        self.adam.begin_step()
        adam_dense(self.w, self.m["w"], self.s["w"], g_w, self.adam)
        adam_dense(self.v, self.m["v"], self.s["v"], g_v, self.adam)
        for i in range(len(self.mlp_w)):
            adam_dense(self.mlp_w[i], self.mm[i][0], self.mm[i][1], gw[i], self.adam)
            adam_dense(self.mlp_b[i], self.mb[i][0], self.mb[i][1], gb[i], self.adam)
        return F32(loss)


class DeepCrossOracle:
    """models/deep_and_cross/src/deep_and_cross.py:293-309 (forward), :312-328 (loss), :331-357 (nn.Adam,
    dense gradients), all fp32."""

    def __init__(self, table, tower_w, tower_b, head_w, head_b, cross_w, cross_b, lr=1e-4, eps=1e-8, sens=1000.0):
        c = lambda a: np.asarray(a, dtype=F32).copy()
        self.table = c(table)
        self.tw, self.tb = [c(x) for x in tower_w], [c(x) for x in tower_b]
        self.hw, self.hb = [c(x) for x in head_w], [c(x) for x in head_b]
        self.cw, self.cb = c(cross_w), c(cross_b)
        self.sens = sens
        self.adam = AdamState(lr, eps=eps, loss_scale=sens)
        self.state = {}

    def _adam(self, name, param, grad):
        if name not in self.state:
            self.state[name] = (np.zeros_like(param), np.zeros_like(param))
        m, v = self.state[name]
        adam_dense(param, m, v, grad, self.adam)

    def step(self, ids, wts, label):
        b, f = ids.shape
        d = self.table.shape[1]
        x = gather_masked(self.table, ids, wts).astype(np.float64)
        d2, t_acts = _mlp_forward(x, self.tw, self.tb, last_activation=True)
        c6, _, _ = cross_forward(x, self.cw, self.cb)
        cat = np.concatenate([d2, c6], axis=1)
        logit, h_acts = _mlp_forward(cat, self.hw, self.hb)
        loss = sigmoid_xent(logit, label).mean()
        delta = self.sens * (sigmoid(logit) - label) / b
        g_cat, ghw, ghb = _mlp_backward(delta, h_acts, self.hw)
        k = d2.shape[1]
        gx_t, gtw, gtb = _mlp_backward(g_cat[:, :k], t_acts, self.tw, last_activation=True)
        gx_c, gcw, gcb = cross_backward(x, self.cw, self.cb, g_cat[:, k:])
        g_table = _scatter_rows(self.table.shape, ids, (gx_t + gx_c).reshape(b * f, d), wts)
        self.adam.begin_step()
        self._adam("table", self.table, g_table)
        for i in range(len(self.tw)):
            self._adam("tw%d" % i, self.tw[i], gtw[i]); self._adam("tb%d" % i, self.tb[i], gtb[i])
        for i in range(len(self.hw)):
            self._adam("hw%d" % i, self.hw[i], ghw[i]); self._adam("hb%d" % i, self.hb[i], ghb[i])
        self._adam("cw", self.cw, gcw); self._adam("cb", self.cb, gcb)
        return F32(loss)


def gather_pool(table, ids, mask):
    """Multi-hot field of the multitable model: ReduceMean(table[ids] * mask[..., None], axis 1) over ALL slots
    — models/wide_and_deep_multitable/src/wide_and_deep.py:301-307."""
    e = gather(table, ids).astype(np.float64) * np.asarray(mask, dtype=np.float64)[..., None]
    return e.mean(axis=1)


# ------------------------------------------------------------------------------------------------
# a15  Wide&Deep multitable (models/wide_and_deep_multitable/src/wide_and_deep.py)
# ------------------------------------------------------------------------------------------------
class MultitableOracle:
    """models/wide_and_deep_multitable/src/wide_and_deep.py:271-427 (forward), :431-480 (loss), :495-600 (step),
    DenseLayers in fp32 arithmetic.  `deep` / `wide` are dicts of float32 arrays under the reference's parameter
    names (wide vectors [V,1]); every `P.Gather` has a dense gradient, so every parameter takes a dense optimizer
    step: FTRL(lr, l1=l2=5e-4, initial_accum=0.1) on the "wide" names (wide_bias included), Adam(lr, eps=1e-6) on
    the rest.  multi: list of six (ids [B,S], mask [B,S])."""

    def __init__(self, deep, wide, mlp_w, mlp_b, adam_lr=3e-3, ftrl_lr=0.1, sens=1000.0):
        self.deep = {k: v.astype(F32).copy() for k, v in deep.items()}
        self.wide = {k: v.astype(F32).copy() for k, v in wide.items()}
        self.mlp_w = [w.astype(F32).copy() for w in mlp_w]
        self.mlp_b = [b.astype(F32).copy() for b in mlp_b]
        self.sens = sens
        self.adam = AdamState(adam_lr, eps=1e-6, loss_scale=sens)
        self.ftrl = FtrlState(ftrl_lr, l1=5e-4, l2=5e-4, loss_scale=sens)
        z = np.zeros_like
        self.mv = {k: (z(v), z(v)) for k, v in self.deep.items()}
        self.mv_w = [(z(w), z(w)) for w in self.mlp_w]
        self.mv_b = [(z(b), z(b)) for b in self.mlp_b]
        self.al = {k: (np.full_like(v, 0.1), z(v)) for k, v in self.wide.items()}

    def forward(self, continue_val, indicator_id, emb_128_id, emb_64_single_id, multi):
        b = continue_val.shape[0]
        d, w = self.deep, self.wide
        cv = continue_val.astype(np.float64)
        parts = [cv, gather(d["emb64_indicator"], indicator_id).reshape(b, -1),
                 gather(d["emb128_embedding"], emb_128_id).reshape(b, -1),
                 gather(d["emb64_single"], emb_64_single_id).reshape(b, -1)]
        parts += [gather_pool(d["emb64_multi"], ids, m) for ids, m in multi]
        x = np.concatenate([p.astype(np.float64) for p in parts], axis=1)
        deep_out, acts = _mlp_forward(x, self.mlp_w, self.mlp_b)
        ones = lambda ids: np.ones(ids.shape)
        wide = (cv * w["wide_continue_w"].astype(np.float64)[None, :]).sum(1, keepdims=True)
        wide = wide + gather_reduce(w["wide_indicator_w"], indicator_id, ones(indicator_id))
        wide = wide + gather_reduce(w["wide_emb128_w"], emb_128_id, ones(emb_128_id))
        wide = wide + gather_reduce(w["wide_emb64_single_w"], emb_64_single_id, ones(emb_64_single_id))
        for ids, m in multi:
            wide = wide + gather_reduce(w["wide_emb64_multi_w"], ids, m)
        wide = wide + w["wide_bias"].astype(np.float64)
        return wide + deep_out, acts

    def step(self, label, continue_val, indicator_id, emb_128_id, emb_64_single_id, multi):
        b = continue_val.shape[0]
        d, w = self.deep, self.wide
        logit, acts = self.forward(continue_val, indicator_id, emb_128_id, emb_64_single_id, multi)
        label = np.asarray(label, dtype=np.float64).reshape(-1, 1)
        loss = sigmoid_xent(logit, label).mean()
        delta = self.sens * (sigmoid(logit) - label) / b
        gx, gw, gb = _mlp_backward(delta, acts, self.mlp_w)
        nf = continue_val.shape[1]
        gd = {}
        o = nf
        n = indicator_id.shape[1] * 64
        gd["emb64_indicator"] = _scatter_rows(d["emb64_indicator"].shape, indicator_id, gx[:, o:o + n].reshape(-1, 64),
                                              np.ones(indicator_id.size)); o += n
        n = emb_128_id.shape[1] * 128
        gd["emb128_embedding"] = _scatter_rows(d["emb128_embedding"].shape, emb_128_id, gx[:, o:o + n].reshape(-1, 128),
                                               np.ones(emb_128_id.size)); o += n
        n = emb_64_single_id.shape[1] * 64
        gd["emb64_single"] = _scatter_rows(d["emb64_single"].shape, emb_64_single_id, gx[:, o:o + n].reshape(-1, 64),
                                           np.ones(emb_64_single_id.size)); o += n
        gd["emb64_multi"] = np.zeros(d["emb64_multi"].shape)
        for ids, m in multi:
            s = ids.shape[1]
            gd["emb64_multi"] += _scatter_rows(d["emb64_multi"].shape, ids, gx[:, o:o + 64],
                                               np.asarray(m, dtype=np.float64) / s, div=s)
            o += 64
        gwd = {"wide_continue_w": (delta * continue_val.astype(np.float64)).sum(0),
               "wide_bias": np.array([delta.sum()]),
               "wide_indicator_w": _scatter_rows(w["wide_indicator_w"].shape, indicator_id, delta,
                                                 np.ones(indicator_id.size), div=indicator_id.shape[1]),
               "wide_emb128_w": _scatter_rows(w["wide_emb128_w"].shape, emb_128_id, delta, np.ones(emb_128_id.size),
                                              div=emb_128_id.shape[1]),
               "wide_emb64_single_w": _scatter_rows(w["wide_emb64_single_w"].shape, emb_64_single_id, delta,
                                                    np.ones(emb_64_single_id.size), div=emb_64_single_id.shape[1]),
               "wide_emb64_multi_w": np.zeros(w["wide_emb64_multi_w"].shape)}
        for ids, m in multi:
            gwd["wide_emb64_multi_w"] += _scatter_rows(w["wide_emb64_multi_w"].shape, ids, delta, m, div=ids.shape[1])
        for k in w:
            acc, lin = self.al[k]
            ftrl_dense(w[k], acc, lin, gwd[k].reshape(w[k].shape), self.ftrl)
        self.adam.begin_step()
        for k in d:
            mm, vv = self.mv[k]
            adam_dense(d[k], mm, vv, gd[k], self.adam)
        for i in range(len(self.mlp_w)):
            adam_dense(self.mlp_w[i], self.mv_w[i][0], self.mv_w[i][1], gw[i], self.adam)
            adam_dense(self.mlp_b[i], self.mv_b[i][0], self.mv_b[i][1], gb[i], self.adam)
        return F32(loss)


# ------------------------------------------------------------------------------------------------
# 8e  row-sharded exchange: index arithmetic of the device-driven protocol (mindrec_b200/peer_sharded.py)
# ------------------------------------------------------------------------------------------------
def shard_exchange_offsets(bounds_all, me):
    """bounds_all[s, :] = bucket bounds of rank s's sorted unique keys (bucket o = keys owned by rank o).
    Returns what mrec_shard_offsets computes on rank `me`:
      dst_off[s]   where rank s wants my rows in ITS landing buffer   = bounds_all[s, me]
      src_off[s]   where rank s's requests start in MY key inbox      = sum_{t<s} |bucket(t -> me)|   (src_off[G] = n_r)
      inbox_off[o] where MY bucket for owner o starts in o's inbox    = sum_{t<me} |bucket(t -> o)|
    The reference reaches the same layout with AllGather / ReduceScatter over a contiguous row split
    (models/wide_deep/src/wide_and_deep.py:232-249); this is its all-to-all form."""
    b = np.asarray(bounds_all, dtype=np.int64)
    g = b.shape[0]
    size = b[:, 1:] - b[:, :-1]                        # size[s, o]
    dst_off = b[:, me].copy()
    src_off = np.concatenate([[0], np.cumsum(size[:, me])])
    inbox_off = np.array([size[:me, o].sum() for o in range(g)], dtype=np.int64)
    return dst_off, src_off, inbox_off, int(src_off[-1])


def shard_exchange_simulate(keys_per_rank, world, rows_per_rank):
    """Full forward exchange on the host for G ranks: per-rank sorted unique owner-major keys -> key inboxes ->
    (owner, local row) served back into landing buffers.  Returns (inboxes, landings) where landing[s][u] is the
    (owner, local_row) pair that arrived for rank s's u-th unique key — it must name that very key."""
    uniq = [np.unique(np.asarray(k, dtype=np.int64)) for k in keys_per_rank]                 # owner-major, sorted
    edges = np.arange(world + 1) * rows_per_rank
    bounds = np.stack([np.searchsorted(u, edges, side="left") for u in uniq])
    inbox = [np.full(sum(int(bounds[s, o + 1] - bounds[s, o]) for s in range(world)), -1, dtype=np.int64)
             for o in range(world)]
    for me in range(world):
        _, _, inbox_off, _ = shard_exchange_offsets(bounds, me)
        for o in range(world):
            seg = uniq[me][bounds[me, o]:bounds[me, o + 1]] % rows_per_rank                    # local rows
            inbox[o][inbox_off[o]:inbox_off[o] + seg.size] = seg
    landing = [np.full((u.size, 2), -1, dtype=np.int64) for u in uniq]
    for me in range(world):
        dst_off, src_off, _, n_r = shard_exchange_offsets(bounds, me)
        assert n_r == inbox[me].size
        for s in range(world):
            rows = inbox[me][src_off[s]:src_off[s + 1]]
            landing[s][dst_off[s]:dst_off[s] + rows.size] = np.stack([np.full(rows.size, me), rows], 1)
    return uniq, inbox, landing
