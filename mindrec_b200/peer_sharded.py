"""Device-driven exchange for row-sharded tables: no host-side sizes, so the whole step is CUDA-graph capturable.

Same sharding as mindrec_b200.sharded (owner = key mod G, owner-major remap, one dedup per batch), but every
transfer is a peer store over NVLink into CUDA-IPC inboxes at offsets computed ON THE DEVICE from the all-ranks
bucket-bounds matrix, and every phase boundary is a signal / wait pair of flag kernels (bounded spin):

    plan     shard_remap -> unique -> shard_bounds                      (local)
    publish  bounds row -> every peer's matrix                 signal 0 | wait 0
    keys     shard_offsets; uniq % R -> owners' key inboxes    signal 1 | wait 1
    serve    gather fused with peer stores into requesters' landing buffers (deep + wide)
                                                               signal 2 | wait 2
    expand   gather_masked / gather_reduce from the landing buffers by the inverse index
    ...      DenseLayers forward / loss / backward ...
    grads    segment_sum per unique key -> owners' gradient inboxes (deep + wide)
                                                               signal 3 | wait 3
    update   owner: unique over the key inbox (static capacity, device-side valid count) -> fused FTRL / LazyAdam

`PeerRank` holds one rank's state and phase methods.  `EmulatedPeerGroup` runs G ranks inside ONE process on ONE
GPU in phase-major order (tests: all of the offset / inbox logic without a multi-GPU box);
`PeerShardedTables` is the multi-process form (one rank per GPU, buffers shared with CUDA IPC) and plugs into
sharded.ShardedWideDeepStep.
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib, ops
from .sharded import ShardPlan, ShardedWideDeepStep, _RawCuda

N_PHASES = 4
# bisection switches (diagnostics only): keep the dim-1 twins / the dense update in line on the main stream
_NO_WIDE_FORK = os.environ.get("MREC_PEER_NO_WIDE_FORK", "0") == "1"
_NO_DENSE_FORK = os.environ.get("MREC_PEER_NO_DENSE_FORK", "0") == "1"
_NO_EARLY_SERVE = os.environ.get("MREC_PEER_NO_EARLY_SERVE", "0") == "1"    # every row is served after the update
_PEER_BUFFERS = ("ball", "keys_in", "land_deep", "land_wide", "grad_in", "gwide_in", "flags")


class _Fork:
    """`with _Fork(side):` runs the body on `side` after everything queued on the current stream so far; `_join(side)`
    makes the current stream wait for it.  side=None: the body runs in line."""

    def __init__(self, side):
        self.side = side
        self.ctx = None

    def __enter__(self):
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream())
            self.ctx = torch.cuda.stream(self.side)
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _join(side):
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)


class PeerRank:
    """State and phases of one rank.  `alloc(name, shape, dtype)` returns the rank's peer-writable buffers."""

    def __init__(self, rank, world, vocab_size, emb_dim, n_lookups, device, alloc, seed=1, sens=1024.0,
                 init_std=0.01, cap_rows=None, adam=(3.5e-4, 1e-8), ftrl=(5e-2, 1e-8, 1e-8, 1.0)):
        """adam = (lr, eps) of the LazyAdam on the deep shard, ftrl = (lr, l1, l2, initial_accum) of the FTRL on the
        dim-1 shard: wide_and_deep.py:420-430 by default; the multitable model passes its own (:525-535)."""
        self.rank, self.world = rank, world
        self.plan = ShardPlan(vocab_size, world)
        self.dim, self.n = emb_dim, n_lookups
        self.device = torch.device(device)
        g, r, dev = world, self.plan.rows_per_rank, self.device
        # Rows an inbox can hold.  A requester never has more than N unique keys (the landing buffers are N rows), but
        # an OWNER receives the keys of all G ranks: up to min(G * N, R) rows when no two lookups repeat a key and every
        # rank asks for this owner's rows only.  Default: 2 N rows — the whole worst case for G <= 2; for G > 2 it
        # covers twice the load of G ranks with N distinct keys each spread evenly over the owners (N rows), which a
        # key-mod-G owner function only exceeds under an adversarial key set.  An overflow is never silent: the push
        # kernel drops the row and raises err bit 1, which poisons the step's loss with NaN
        # (PeerShardedWideDeepStep._a2).  cap_rows = G * N removes the possibility altogether.
        self.cap = (int(cap_rows or min(g * n_lookups, max(r, n_lookups), 2 * n_lookups)) + 3) // 4 * 4
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed * 1000 + rank)
        self.wide = torch.empty((r, 1), dtype=torch.float32, device=dev).normal_(0, init_std, generator=gen)
        self.deep = torch.empty((r, emb_dim), dtype=torch.float32, device=dev).normal_(0, init_std, generator=gen)
        self.acc, self.lin = torch.full_like(self.wide, float(ftrl[3])), torch.zeros_like(self.wide)
        self.m, self.v = torch.zeros_like(self.deep), torch.zeros_like(self.deep)
        self.adam_hyper = ops.adam_hyper(adam[0], eps=adam[1], loss_scale=sens * world, device=dev)
        self.ftrl_hyper = ops.ftrl_hyper(ftrl[0], l1=ftrl[1], l2=ftrl[2], loss_scale=sens * world, device=dev)
        # peer-writable buffers.  The KEY phase (bounds matrix, key inbox) exists twice: while step t runs on set
        # t & 1, the key phase of batch t + 1 (plan, publish, key push, owner dedup) runs one step ahead on the other
        # set underneath step t's DenseLayers, so that only serve -> expand stays in front of them.
        i32, f32 = torch.int32, torch.float32
        self.buf = {
            "ball": alloc("ball", (2 * g * (g + 1),), i32), "keys_in": alloc("keys_in", (2 * self.cap,), i32),
            "land_deep": alloc("land_deep", (n_lookups, emb_dim), f32), "land_wide": alloc("land_wide", (n_lookups, 1), f32),
            "grad_in": alloc("grad_in", (self.cap, emb_dim), f32), "gwide_in": alloc("gwide_in", (self.cap, 1), f32),
            "flags": alloc("flags", (N_PHASES * g,), i32),
        }
        self.buf["ball"].zero_()
        self.buf["flags"].zero_()
        self.buf["keys_in"].zero_()
        # local state
        self.ctrl = torch.tensor([rank, world], dtype=i32, device=dev)
        self.epoch = [torch.zeros(1, dtype=i32, device=dev) for _ in range(N_PHASES)]
        self.err = torch.zeros(1, dtype=i32, device=dev)
        self.edges = (torch.arange(g + 1, device=dev, dtype=i32) * r)
        self._sets = []
        for si in range(2):
            st = {
                "ball": self.buf["ball"][si * g * (g + 1):(si + 1) * g * (g + 1)],
                "keys_in": self.buf["keys_in"][si * self.cap:(si + 1) * self.cap],
                "bounds": torch.zeros(g + 1, dtype=i32, device=dev), "dst_off": torch.zeros(g, dtype=i32, device=dev),
                "src_off": torch.zeros(g + 1, dtype=i32, device=dev), "inbox_off": torch.zeros(g, dtype=i32, device=dev),
                "n_r": torch.zeros(1, dtype=i32, device=dev), "key": torch.empty(n_lookups, dtype=i32, device=dev),
                "uq": ops.UniqueResult(n_lookups, i32, dev, packed=True), "uq_owner": ops.UniqueResult(self.cap, i32, dev),
                "tag": "unique_peer_plan%d" % si, "owner_tag": "unique_peer_owner%d" % si,
            }
            self._sets.append(st)
        self.cur = 0
        # one bit per owned row: set for the rows the CURRENT step's update rewrites (its owner-side dedup names them).
        # Rows of the NEXT batch whose bit is clear are served one step early, underneath the DenseLayers (p_serve_early);
        # only the others — the Zipf head both batches share — wait for the update (p_serve(late=True)).
        self.dirty = torch.zeros((r + 31) // 32, dtype=i32, device=dev)
        self.gs_wide = torch.empty((n_lookups, 1), dtype=f32, device=dev)
        self._bound_like = torch.empty((g * r, 0), device=dev)
        self._vocab_like = torch.empty((vocab_size, 0), device=dev)
        self._owners_like = torch.empty((g, r, 0), device=dev)
        self._cap_like = torch.empty((self.cap, 0), device=dev)
        self._mod_rows = torch.empty((r, 0), device=dev)
        self._mod_none = torch.empty((0, 0), device=dev)
        self._no_payload = torch.empty(0, dtype=i32, device=dev)
        self.flag_views = [self.buf["flags"][ph * g:(ph + 1) * g] for ph in range(N_PHASES)]
        self.ptrs = None
        self._wts = None

    # the CURRENT set's state under the names the phases (and the tests) use
    def _s(self, nxt=False):
        return self._sets[self.cur ^ 1 if nxt else self.cur]

    key = property(lambda self: self._s()["key"])
    uq = property(lambda self: self._s()["uq"])
    bounds = property(lambda self: self._s()["bounds"])
    dst_off = property(lambda self: self._s()["dst_off"])
    src_off = property(lambda self: self._s()["src_off"])
    inbox_off = property(lambda self: self._s()["inbox_off"])
    n_r = property(lambda self: self._s()["n_r"])
    uq_owner = property(lambda self: self._s()["uq_owner"])
    keys_in = property(lambda self: self._s()["keys_in"])

    def use_set(self, i):
        self.cur = int(i) & 1

    def connect(self, base_ptrs):
        """base_ptrs[name][s] = address of rank s's buffer `name` as mapped in THIS process."""
        g, me, dev = self.world, self.rank, self.device
        t = lambda lst: torch.tensor(lst, dtype=torch.int64, device=dev)
        for si, st in enumerate(self._sets):
            st["p_ball_row"] = t([base_ptrs["ball"][s] + (si * g * (g + 1) + me * (g + 1)) * 4 for s in range(g)])
            st["p_keys_in"] = t([base_ptrs["keys_in"][s] + si * self.cap * 4 for s in range(g)])
        self.ptrs = {
            "land_deep": t(base_ptrs["land_deep"]),
            "land_wide": t(base_ptrs["land_wide"]), "grad_in": t(base_ptrs["grad_in"]),
            "gwide_in": t(base_ptrs["gwide_in"]),
            "flag": [t([base_ptrs["flags"][s] + (ph * g + me) * 4 for s in range(g)]) for ph in range(N_PHASES)],
            "none": t([0] * g),
        }

    # ---- phase boundaries ----------------------------------------------------------------------------
    def signal(self, phase, payload=None, payload_ptrs=None):
        ops.peer_signal(self._no_payload if payload is None else payload,
                        self.ptrs["none"] if payload_ptrs is None else payload_ptrs,
                        self.ptrs["flag"][phase], self.epoch[phase])

    def wait(self, phase):
        ops.peer_wait(self.flag_views[phase], self.epoch[phase], self.err)

    # ---- phases --------------------------------------------------------------------------------------
    def p_plan_local(self, ids, nxt=False):
        """The rank-local part of the plan: owner-major remap, dedup, bucket bounds.  nxt=True works on the OTHER set
        (the next batch, one step ahead); p_adopt makes that set the current one."""
        st = self._s(nxt)
        ops.shard_remap(ids.reshape(-1), self._vocab_like, self._owners_like, out=st["key"])
        ops.unique(st["key"], table_like=self._bound_like, result=st["uq"], ws_tag=st["tag"])
        ops.shard_bounds(st["uq"].uniq, st["uq"].count, self.edges, out=st["bounds"])

    def p_adopt(self):
        """The set prepared one step ahead becomes the current one (no copy: the two sets swap roles)."""
        self.cur ^= 1

    def p_publish(self, nxt=False):
        st = self._s(nxt)
        self.signal(0, st["bounds"], st["p_ball_row"])

    def p_plan_publish(self, ids):
        self.p_plan_local(ids)
        self.p_publish()

    def p_keys(self, nxt=False):
        st = self._s(nxt)
        ops.shard_offsets(st["ball"], self.ctrl, st["dst_off"], st["src_off"], st["inbox_off"], st["n_r"])
        ops.push_rows_to_peers(st["uq"].uniq, st["bounds"], st["inbox_off"], st["p_keys_in"], self._cap_like,
                               self._mod_rows, self.err)
        self.signal(1)

    def p_key_phase(self, ids, nxt=False, early_serve=False):
        """Everything of a batch that does not depend on the table's values: plan, publish, key push, owner dedup.
        With nxt=True it runs for the NEXT batch on the other set (callers put it on a forked branch underneath the
        current step's DenseLayers); its waits only depend on the peers' key phases.  early_serve: additionally serve
        the requested rows that the current step's update will not touch (p_mark_updates must have run)."""
        self.p_plan_local(ids, nxt)
        self.p_publish(nxt)
        self.wait(0)
        self.p_keys(nxt)
        self.wait(1)
        self.p_owner_dedup(nxt)
        if early_serve:
            self.p_serve_early(nxt)

    # `side` (a stream, optional) in the next ones: the dim-1 work of the wide vector runs there, beside the
    # bandwidth-bound deep rows (forked from / joined back into the current stream — a parallel branch when captured)
    def _serve(self, st, mode, side):
        with _Fork(side):
            ops.gather_to_peers(self.wide, st["keys_in"], self.ptrs["land_wide"], st["dst_off"], st["src_off"],
                                dirty=self.dirty, mode=mode)
        ops.gather_to_peers(self.deep, st["keys_in"], self.ptrs["land_deep"], st["dst_off"], st["src_off"],
                            dirty=self.dirty, mode=mode)
        _join(side)

    def p_serve(self, side=None, late=False):
        """Serve the current batch's rows (late=True: only those the previous update rewrote — the rest went out with
        p_serve_early one step ahead), then signal 2."""
        self._serve(self._s(), 2 if late else 0, side)
        self.signal(2)

    def p_serve_early(self, nxt=True, side=None):
        """Rows of the batch on the other set that the CURRENT step's update does not touch: their values are final, so
        they cross NVLink now.  No signal: the late part of the same serve sends it.  (Every requester has finished
        expanding its current batch: this runs after wait 1 of the next batch's key phase, and a peer signals 1 after
        its expand.)"""
        self._serve(self._s(nxt), 1, side)

    def p_mark_updates(self):
        """dirty := the rows this step's update rewrites (the current set's owner-side dedup)."""
        self.dirty.zero_()
        ops.bitmap_set(self.uq_owner.uniq, self.uq_owner.count, self.dirty)

    def p_expand(self, ids_shape, wts, wide_bias, deep_out, wide_out, side=None):
        inverse = self.uq.inverse.view(ids_shape)
        # ids outside [0, V) collapse onto ONE extra unique entry behind the valid ones (index bounds[G]), which no
        # owner serves: make that landing row the zero row, what the unsharded gather returns for such ids
        n_valid_keys = self.bounds[self.world:self.world + 1]
        with _Fork(side):
            ops.zero_row(self.buf["land_wide"], n_valid_keys)
            ops.gather_reduce(self.buf["land_wide"], inverse, wts, wide_bias, out=wide_out)
        ops.zero_row(self.buf["land_deep"], n_valid_keys)
        ops.gather_masked(self.buf["land_deep"], inverse, wts, out=deep_out)
        _join(side)
        self._wts = wts

    def p_grads(self, delta, gx, side=None):
        mask = self._wts.reshape(-1)
        with _Fork(side):
            ops.segment_sum(delta, mask, self.uq, dim=1, out=self.gs_wide)
            ops.push_rows_to_peers(self.gs_wide, self.bounds, self.inbox_off, self.ptrs["gwide_in"], self._cap_like,
                                   self._mod_none, self.err)
        # deep rows: the segment sums are stored straight into the owners' gradient inboxes (no local gsum pass)
        ops.segment_sum_to_peers(gx.view(self.n, self.dim), mask, self.uq, self.bounds, self.inbox_off,
                                 self.ptrs["grad_in"], self._cap_like, self.err, dim=self.dim)
        _join(side)
        self.signal(3)

    def p_owner_dedup(self, nxt=False):
        """The same row can be asked for by several ranks: dedup the key inbox (needs only the keys, so callers
        run it on a side stream underneath the DenseLayer segment).  Work follows n_r, not the capacity."""
        st = self._s(nxt)
        ops.unique(st["keys_in"], table_like=self.deep, result=st["uq_owner"], ws_tag=st["owner_tag"], n_valid=st["n_r"])

    def p_update_wide(self):
        ops.sparse_ftrl(self.wide, self.acc, self.lin, self.ftrl_hyper, self.buf["gwide_in"], None, self.uq_owner,
                        n_valid=self.n_r)

    def p_update_deep(self):
        ops.adam_begin_step(self.adam_hyper)
        ops.sparse_lazy_adam(self.deep, self.m, self.v, self.adam_hyper, self.buf["grad_in"], None, self.uq_owner,
                             n_valid=self.n_r)

    def p_update(self):
        self.p_owner_dedup()
        self.p_update_wide()
        self.p_update_deep()


class EmulatedPeerGroup:
    """G ranks in one process on one GPU, phases run in phase-major order (every signal of a phase is issued
    before any wait of that phase, so the wait kernels never spin).  Test vehicle for the protocol's data flow."""

    def __init__(self, world, vocab_size, emb_dim, n_lookups, device, seed=1, sens=1024.0):
        def alloc(name, shape, dtype):
            return torch.empty(shape, dtype=dtype, device=device)
        self.ranks = [PeerRank(r, world, vocab_size, emb_dim, n_lookups, device, alloc, seed=seed, sens=sens)
                      for r in range(world)]
        base = {name: [rk.buf[name].data_ptr() for rk in self.ranks] for name in _PEER_BUFFERS}
        for rk in self.ranks:
            rk.connect(base)

    def key_phase_next(self, ids_list, early_serve=True):
        """The NEXT batch's key phase on the other buffer set, phase-major (what the step graph's forked branch does
        per rank), followed by the early part of its serve."""
        for rk, ids in zip(self.ranks, ids_list):
            rk.p_plan_local(ids, nxt=True)
            rk.p_publish(nxt=True)
        for rk in self.ranks:
            rk.wait(0)
            rk.p_keys(nxt=True)
        for rk in self.ranks:
            rk.wait(1)
            rk.p_owner_dedup(nxt=True)
        if early_serve:
            for rk in self.ranks:
                rk.p_serve_early(nxt=True)
        self._early = bool(early_serve)

    def forward(self, ids_list, wts_list, bias, deep_outs, wide_outs, planned=False):
        """planned=True: the batch went through key_phase_next during the previous step; adopt that set (and serve only
        the rows the previous update rewrote, if the rest went out early)."""
        late = False
        if planned:
            for rk in self.ranks:
                rk.p_adopt()
            late = getattr(self, "_early", False)
        else:
            for rk, ids in zip(self.ranks, ids_list):
                rk.p_plan_publish(ids)
            for rk in self.ranks:
                rk.wait(0)
                rk.p_keys()
        self._early = False
        for rk in self.ranks:
            if not planned:
                rk.wait(1)
                rk.p_owner_dedup()
            rk.p_serve(late=late)
        for rk, ids, wts, do, wo in zip(self.ranks, ids_list, wts_list, deep_outs, wide_outs):
            rk.wait(2)
            rk.p_expand(ids.shape, wts, bias, do, wo)
            rk.p_mark_updates()

    def backward(self, deltas, gxs):
        for rk, d, g in zip(self.ranks, deltas, gxs):
            rk.p_grads(d, g)
        for rk in self.ranks:
            rk.wait(3)
            rk.p_update()

    def full_tables(self):
        g = len(self.ranks)
        v = self.ranks[0].plan.vocab_size
        r = self.ranks[0].plan.rows_per_rank
        wide = torch.stack([rk.wide for rk in self.ranks], 1).reshape(r * g, 1)[:v]
        deep = torch.stack([rk.deep for rk in self.ranks], 1).reshape(r * g, -1)[:v]
        return wide, deep


_HASH_BUFFERS = ("ball", "keys_in", "land", "grad_in", "flags")


class PeerHashRank:
    """One rank of a MapParameter sharded by owner = hash(key) mod G: every rank owns an independent
    open-addressing table (mindrec_b200.hash.MapParameter) plus LazyAdam moments addressed by slot.

    Same four-phase exchange as PeerRank; what differs is the key handling: keys are int64 in [0, 2^key_bits),
    made owner-major as owner << key_bits | key (mrec_shard_remap_hash) so one bounded sort dedups and buckets
    them; the owner gets the ORIGINAL key back (key' % 2^key_bits), runs find-or-insert on its table (admission /
    eviction counters live there), initialises rows of new keys (Philox keyed by the key, so a key's first row does
    not depend on G) and serves rows by slot.  The stale tail of the static key inbox is blanked with the
    reserved key -1 so that it cannot touch the table."""

    def __init__(self, rank, world, emb_dim, n_lookups, device, alloc, key_bits=40, capacity=1 << 16, seed=0,
                 learning_rate=1e-3, loss_scale=1.0, permit_filter_value=1, evict_filter_value=None, cap_rows=None,
                 eps=1e-8):
        from . import hash as _hash
        self.rank, self.world, self.dim, self.n = rank, world, emb_dim, n_lookups
        self.bits = int(key_bits)
        if world << self.bits >= 1 << 62:
            raise ValueError("key_bits + log2(world) must stay below 62")
        self.device = torch.device(device)
        dev, g = self.device, world
        self.cap = int(cap_rows or min(g * n_lookups, 2 * n_lookups))     # see PeerRank: inbox rows, overflow -> err bit 1
        self.table = _hash.MapParameter(key_dtype=torch.int64, value_shape=emb_dim, default_value="normal",
                                        permit_filter_value=permit_filter_value,
                                        evict_filter_value=evict_filter_value or _hash.MAX_SIZE, capacity=capacity,
                                        device=dev, seed=seed)
        self.m, self.v = self.table.add_arena(0.0), self.table.add_arena(0.0)
        self.hyper = ops.adam_hyper(learning_rate, eps=eps, loss_scale=loss_scale, device=dev)
        i32, i64, f32 = torch.int32, torch.int64, torch.float32
        self.buf = {"ball": alloc("ball", (g * (g + 1),), i32), "keys_in": alloc("keys_in", (self.cap,), i64),
                    "land": alloc("land", (n_lookups, emb_dim), f32), "grad_in": alloc("grad_in", (self.cap, emb_dim), f32),
                    "flags": alloc("flags", (N_PHASES * g,), i32)}
        self.buf["ball"].zero_()
        self.buf["flags"].zero_()
        self.buf["keys_in"].fill_(-1)
        self.ctrl = torch.tensor([rank, world], dtype=i32, device=dev)
        self.epoch = [torch.zeros(1, dtype=i32, device=dev) for _ in range(N_PHASES)]
        self.err = torch.zeros(1, dtype=i32, device=dev)
        self.bounds = torch.zeros(g + 1, dtype=i32, device=dev)
        self.dst_off = torch.zeros(g, dtype=i32, device=dev)
        self.src_off = torch.zeros(g + 1, dtype=i32, device=dev)
        self.inbox_off = torch.zeros(g, dtype=i32, device=dev)
        self.n_r = torch.zeros(1, dtype=i32, device=dev)
        self.edges = torch.arange(g + 1, device=dev, dtype=i64) << self.bits
        self.key = torch.empty(n_lookups, dtype=i64, device=dev)
        self.uq = ops.UniqueResult(n_lookups, i64, dev)
        self.uq_owner = ops.UniqueResult(self.cap, i32, dev)
        self.gs = torch.empty((n_lookups, emb_dim), dtype=f32, device=dev)
        self._bound_like = torch.empty((g << self.bits, 0), device=dev)
        self._owners_like = torch.empty((g, 0), device=dev)
        self._bits_like = torch.empty((self.bits, 0), device=dev)
        self._mod_keys = torch.empty((1 << self.bits, 0), device=dev)
        self._mod_none = torch.empty((0, 0), device=dev)
        self._cap_like = torch.empty((self.cap, 0), device=dev)
        self._slot_like = torch.empty((self.table.capacity, 0), device=dev)      # slot C (default row) is out of range
        self._minus_one = torch.tensor([-1], dtype=i64, device=dev)
        self._no_payload = torch.empty(0, dtype=i32, device=dev)
        self.flag_views = [self.buf["flags"][ph * g:(ph + 1) * g] for ph in range(N_PHASES)]
        self.slots = None
        self.ptrs = None

    def connect(self, base_ptrs):
        g, me, dev = self.world, self.rank, self.device
        t = lambda lst: torch.tensor(lst, dtype=torch.int64, device=dev)
        self.ptrs = {"ball_row": t([base_ptrs["ball"][s] + me * (g + 1) * 4 for s in range(g)]),
                     "keys_in": t(base_ptrs["keys_in"]), "land": t(base_ptrs["land"]), "grad_in": t(base_ptrs["grad_in"]),
                     "flag": [t([base_ptrs["flags"][s] + (ph * g + me) * 4 for s in range(g)]) for ph in range(N_PHASES)],
                     "none": t([0] * g)}

    signal = PeerRank.signal
    wait = PeerRank.wait

    def p_plan_publish(self, keys):
        ops.shard_remap_hash(keys.reshape(-1), self._owners_like, self._bits_like, out=self.key)
        ops.unique(self.key, table_like=self._bound_like, result=self.uq, ws_tag="unique_peerhash_plan")
        ops.shard_bounds(self.uq.uniq, self.uq.count, self.edges, out=self.bounds)
        self.signal(0, self.bounds, self.ptrs["ball_row"])

    def p_keys(self):
        ops.shard_offsets(self.buf["ball"], self.ctrl, self.dst_off, self.src_off, self.inbox_off, self.n_r)
        ops.push_rows_to_peers(self.uq.uniq, self.bounds, self.inbox_off, self.ptrs["keys_in"], self._cap_like,
                               self._mod_keys, self.err)
        self.signal(1)

    def p_serve(self):
        ops.fill_tail(self.buf["keys_in"], self.n_r, self._minus_one)
        self.slots = self.table.lookup_slots(self.buf["keys_in"], insert_default_value=True)
        ops.gather_to_peers(self.table.values, self.slots, self.ptrs["land"], self.dst_off, self.src_off)
        self.signal(2)

    def p_expand(self, out):
        """out[..., D] = rows of the looked-up keys (keys.shape + (D,))."""
        ops.gather(self.buf["land"], self.uq.inverse, out=out.view(self.n, self.dim))

    def p_grads(self, g_out):
        if self.dim % 4 == 0:      # fused: segment sums go straight into the owners' gradient inboxes
            ops.segment_sum_to_peers(g_out.reshape(self.n, self.dim), None, self.uq, self.bounds, self.inbox_off,
                                     self.ptrs["grad_in"], self._cap_like, self.err, dim=self.dim)
        else:
            ops.segment_sum(g_out.reshape(self.n, self.dim), None, self.uq, dim=self.dim, out=self.gs)
            ops.push_rows_to_peers(self.gs, self.bounds, self.inbox_off, self.ptrs["grad_in"], self._cap_like,
                                   self._mod_none, self.err)
        self.signal(3)

    def p_update(self):
        c = self.table.capacity
        uq2 = ops.unique(self.slots, table_like=self._slot_like, result=self.uq_owner, ws_tag="unique_peerhash_owner",
                         n_valid=self.n_r)
        ops.adam_begin_step(self.hyper)
        ops.sparse_lazy_adam(self.table.values[:c], self.m[:c], self.v[:c], self.hyper, self.buf["grad_in"], None, uq2,
                             n_valid=self.n_r)


class EmulatedPeerHashGroup:
    """G PeerHashRank on one GPU in phase-major order (tests)."""

    def __init__(self, world, emb_dim, n_lookups, device, **kw):
        def alloc(name, shape, dtype):
            return torch.empty(shape, dtype=dtype, device=device)
        self.ranks = [PeerHashRank(r, world, emb_dim, n_lookups, device, alloc, **kw) for r in range(world)]
        base = {name: [rk.buf[name].data_ptr() for rk in self.ranks] for name in _HASH_BUFFERS}
        for rk in self.ranks:
            rk.connect(base)

    def forward(self, keys_list, outs):
        for rk, keys in zip(self.ranks, keys_list):
            rk.p_plan_publish(keys)
        for rk in self.ranks:
            rk.wait(0)
            rk.p_keys()
        for rk in self.ranks:
            rk.wait(1)
            rk.p_serve()
        for rk, out in zip(self.ranks, outs):
            rk.wait(2)
            rk.p_expand(out)

    def backward(self, grads):
        for rk, g in zip(self.ranks, grads):
            rk.p_grads(g)
        for rk in self.ranks:
            rk.wait(3)
            rk.p_update()

    def get_data(self):
        """All (key, row) pairs of the G tables, sorted by key."""
        ks, vs = zip(*[rk.table.get_data() for rk in self.ranks])
        k, v = torch.cat(ks), torch.cat(vs)
        order = torch.argsort(k)
        return k[order], v[order]


class PeerShardedHashEmbedding:
    """Multi-process form of PeerHashRank (one rank per GPU, CUDA-IPC inboxes): `lookup(keys)` returns the rows,
    `update(g_out)` applies LazyAdam to the owners' rows.  Mirrors HashEmbeddingLookup(sparse=True) + LazyAdam on a
    MapParameter (mindspore_rec/ops/embedding.py:85-206, wide_and_deep.py:415-422) under row sharding."""

    def __init__(self, emb_dim, n_lookups, device, group=None, **kw):
        self.group = group
        arena = self._arena = _IpcArena(group)
        self.rk = _connect_collectively(
            lambda: PeerHashRank(dist.get_rank(group), dist.get_world_size(group), emb_dim, n_lookups, device,
                                 arena.alloc(device), **kw), arena, group, torch.device(device))
        self.dim = emb_dim

    def close(self):
        """Unmap the peers' buffers and free this rank's (collective: every rank of the group calls it)."""
        self._arena.close()

    def lookup(self, keys, out=None):
        rk = self.rk
        if out is None:
            out = torch.empty(tuple(keys.shape) + (self.dim,), dtype=torch.float32, device=keys.device)
        rk.p_plan_publish(keys)
        rk.wait(0)
        rk.p_keys()
        rk.wait(1)
        rk.p_serve()
        rk.wait(2)
        rk.p_expand(out)
        return out

    def update(self, g_out):
        rk = self.rk
        rk.p_grads(g_out)
        rk.wait(3)
        rk.p_update()

    def error_flags(self):
        return int(self.rk.err.item()) | (4 if self.rk.table.overflowed else 0)


class PeerMemoryUnavailable(RuntimeError):
    """CUDA-IPC peer memory could not be set up on SOME rank; raised on EVERY rank (collectively agreed), so callers
    can fall back to the NCCL exchange without dead-locking."""


def _all_ok(ok, group, device):
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(t.item())


def _connect_collectively(make_rank, arena, group, device):
    """Allocate + export on every rank, agree, open the peers' handles, agree again."""
    rk, err = None, None
    try:
        rk = make_rank()
        torch.cuda.synchronize()
    except Exception as exc:                                   # noqa: BLE001 - reported through the collective
        err = exc
    if not _all_ok(err is None, group, device):
        raise PeerMemoryUnavailable("peer buffer allocation / export failed on a rank: %s" % (err,))
    base = None
    try:
        base = arena.exchange()
    except Exception as exc:                                   # noqa: BLE001
        err = exc
    if not _all_ok(err is None, group, device):
        raise PeerMemoryUnavailable("opening the peers' IPC handles failed on a rank: %s" % (err,))
    rk.connect(base)
    dist.barrier(group=group)
    return rk


class _IpcArena:
    """cudaMalloc'ed buffers exported to the other ranks of the node with CUDA IPC."""

    def __init__(self, group):
        self.group = group
        self.lib = _lib.lib()
        self.lib.mrec_peer_alloc.restype = ctypes.c_void_p
        self.lib.mrec_peer_alloc.argtypes = [ctypes.c_size_t]
        self.lib.mrec_ipc_open_handle.restype = ctypes.c_void_p
        self.lib.mrec_ipc_open_handle.argtypes = [ctypes.c_char_p]
        self.lib.mrec_ipc_get_handle.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        self.lib.mrec_ipc_close_handle.argtypes = [ctypes.c_void_p]
        self.lib.mrec_peer_free.argtypes = [ctypes.c_void_p]
        self.local = {}
        self.opened = []

    def alloc(self, device):
        def _alloc(name, shape, dtype):
            nbytes = 1
            for d in shape:
                nbytes *= d
            nbytes = max(nbytes * (8 if dtype == torch.int64 else 4), 256)
            ptr = self.lib.mrec_peer_alloc(nbytes)
            if not ptr:
                raise RuntimeError("mrec_peer_alloc failed: " + _lib.last_error())
            h = ctypes.create_string_buffer(64)
            if self.lib.mrec_ipc_get_handle(ctypes.c_void_p(ptr), h) != 0:
                raise RuntimeError("mrec_ipc_get_handle failed: " + _lib.last_error())
            self.local[name] = (ptr, h.raw)
            typestr = {torch.int32: "<i4", torch.int64: "<i8"}.get(dtype, "<f4")
            return torch.as_tensor(_RawCuda(ptr, shape, typestr), device=device)
        return _alloc

    def exchange(self):
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        gathered = [None] * world
        dist.all_gather_object(gathered, {k: h for k, (_, h) in self.local.items()}, group=self.group)
        base = {}
        for name, (ptr, _) in self.local.items():
            lst = []
            for r in range(world):
                if r == rank:
                    lst.append(ptr)
                else:
                    p = self.lib.mrec_ipc_open_handle(gathered[r][name])
                    if not p:
                        raise RuntimeError("mrec_ipc_open_handle failed: " + _lib.last_error())
                    self.opened.append(p)
                    lst.append(p)
            base[name] = lst
        return base

    def close(self):
        """Every rank first unmaps what it opened, then (after a barrier) frees what it owns."""
        torch.cuda.synchronize()
        for p in self.opened:
            self.lib.mrec_ipc_close_handle(ctypes.c_void_p(p))
        self.opened = []
        dist.barrier(group=self.group)
        for ptr, _ in self.local.values():
            self.lib.mrec_peer_free(ctypes.c_void_p(ptr))
        self.local = {}


class PeerAllReduce:
    """Sum of a flat fp32 buffer over the ranks through CUDA-IPC peer memory, capturable in a CUDA graph (NCCL inside a
    captured graph hung on this stack, and an eager collective cuts the step into several graphs): `src` is this
    rank's contribution (write it in place — e.g. make it the DenseLayers' flat gradient buffer), `run()` leaves the
    sum of all ranks in `dst` on every rank.  signal 0 / wait 0: every rank's src is complete; mrec_peer_allreduce:
    rank r adds the r-th slice of the G sources in rank order and stores it into all G dst buffers; signal 1 / wait 1:
    every slice has arrived.  The replicas' sums are bit-identical.  Re-use is safe without a third barrier: a peer
    reads my src only before its signal 1, which I wait for before anything can overwrite src; a peer overwrites my
    dst only after wait 0 of the NEXT round, which needs my next signal 0 — issued after I consumed dst."""

    def __init__(self, n, device, group=None):
        self.group = group
        g, me = dist.get_world_size(group), dist.get_rank(group)
        self.device = dev = torch.device(device)
        arena = self._arena = _IpcArena(group)
        holder = {}

        def make():
            alloc = arena.alloc(dev)
            holder["src"] = alloc("ar_src", (n,), torch.float32)
            holder["dst"] = alloc("ar_dst", (n,), torch.float32)
            holder["flags"] = alloc("ar_flags", (2 * g,), torch.int32)
            for t in holder.values():
                t.zero_()
            return self

        self._base = None
        _connect_collectively(make, arena, group, dev)
        self.src, self.dst, self.flags = holder["src"], holder["dst"], holder["flags"]
        base = self._base
        t = lambda lst: torch.tensor(lst, dtype=torch.int64, device=dev)
        self.p_src, self.p_dst = t(base["ar_src"]), t(base["ar_dst"])
        self.p_flag = [t([base["ar_flags"][s] + (ph * g + me) * 4 for s in range(g)]) for ph in range(2)]
        self.flag_views = [self.flags[ph * g:(ph + 1) * g] for ph in range(2)]
        self.epoch = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(2)]
        self.ctrl = torch.tensor([me, g], dtype=torch.int32, device=dev)
        self._none = t([0] * g)
        self._no_payload = torch.empty(0, dtype=torch.int32, device=dev)

    def connect(self, base):                                 # called by _connect_collectively
        self._base = base

    def run(self, err):
        """On the current stream.  err[1] int32: bit 0 is raised when a wait times out."""
        for ph in range(2):
            ops.peer_signal(self._no_payload, self._none, self.p_flag[ph], self.epoch[ph])
            ops.peer_wait(self.flag_views[ph], self.epoch[ph], err)
            if ph == 0:
                ops.peer_allreduce(self.p_src, self.p_dst, self.ctrl, self.dst)

    def close(self):
        self._arena.close()


class PeerShardedTables:
    """Multi-process form: drop-in for sharded.ShardedWideDeepTables inside sharded.ShardedWideDeepStep
    (`plan_batch` / `lookup` / `update` / `gather_full`), with nothing read back to the host."""

    def __init__(self, vocab_size, emb_dim, n_lookups, device, group=None, seed=1, sens=1024.0, cap_rows=None, **rank_kw):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.plan_stream = None
        self.owner_stream = torch.cuda.Stream(device=self.device, priority=-1)       # exchange kernels: see main_stream
        arena = self._arena = _IpcArena(group)
        self.rk = _connect_collectively(
            lambda: PeerRank(self.rank, self.world, vocab_size, emb_dim, n_lookups, device, arena.alloc(device),
                             seed=seed, sens=sens, cap_rows=cap_rows, **rank_kw), arena, group, self.device)
        self.plan = self.rk.plan
        self.dim = emb_dim
        self._ids = None

    # sharded.ShardedWideDeepStep drives these three; the "plan" is implicit in device state
    def plan_batch(self, ids, ahead=False, adopt=False):
        """The key phase of the batch.  adopt=True: it already ran one step ahead (key_phase_next) on the other set,
        which now becomes the current one."""
        rk = self.rk
        if adopt:
            rk.p_adopt()
        else:
            rk.p_key_phase(ids)
        return _DevicePlan(ids)

    def key_phase_next(self, ids, early_serve=True):
        """Key phase of the NEXT batch on the other set, then the early part of its serve: the requested rows this
        step's update does not touch (call it on a forked branch underneath the DenseLayers, after `lookup`: the bitmap
        of the rows this step rewrites replaces the previous step's, which `lookup(late=True)` has just used)."""
        early = early_serve and not _NO_EARLY_SERVE
        if early:
            self.rk.p_mark_updates()
        self.rk.p_key_phase(ids, nxt=True, early_serve=early)

    def lookup(self, plan, wts, wide_bias, deep_out, wide_out, late=False):
        """late=True: the batch's key phase ran one step ahead WITH its early serve, so only the rows the previous update
        rewrote still have to cross NVLink."""
        rk = self.rk
        side = None if _NO_WIDE_FORK else self.owner_stream
        rk.p_serve(side=side, late=late and not _NO_EARLY_SERVE)
        rk.wait(2)
        rk.p_expand(plan.ids.shape, wts, wide_bias, deep_out, wide_out, side=side)
        return wide_out, deep_out

    def update(self, delta, gx):
        self.push_grads(delta, gx)
        self.apply_grads()

    def push_grads(self, delta, gx):
        """Segment sums of this rank's gradients into the owners' inboxes (NVLink stores), then signal 3."""
        self.rk.p_grads(delta, gx, side=None if _NO_WIDE_FORK else self.owner_stream)

    def apply_grads(self):
        """wait 3, then the fused row updates of the rows this rank owns."""
        rk = self.rk
        rk.wait(3)
        main = torch.cuda.current_stream()
        self.owner_stream.wait_stream(main)
        with torch.cuda.stream(self.owner_stream):       # latency-bound FTRL beside the LazyAdam rows
            rk.p_update_wide()
        rk.p_update_deep()
        main.wait_stream(self.owner_stream)

    @property
    def wide(self):
        return self.rk.wide

    @property
    def deep(self):
        return self.rk.deep

    def error_flags(self):
        """bit 0: a peer wait timed out; bit 1: an inbox overflowed (host read — call outside the hot loop)."""
        return int(self.rk.err.item())

    def close(self):
        """Unmap the peers' buffers and free this rank's (collective: every rank of the group calls it)."""
        self._arena.close()

    def gather_full(self):
        g, r, v = self.world, self.plan.rows_per_rank, self.plan.vocab_size
        wl = [torch.empty_like(self.rk.wide) for _ in range(g)]
        dl = [torch.empty_like(self.rk.deep) for _ in range(g)]
        dist.all_gather(wl, self.rk.wide, group=self.group)
        dist.all_gather(dl, self.rk.deep, group=self.group)
        return (torch.stack(wl, 1).reshape(r * g, 1)[:v], torch.stack(dl, 1).reshape(r * g, self.dim)[:v])


class _DevicePlan:
    __slots__ = ("ids",)

    def __init__(self, ids):
        self.ids = ids


class PeerShardedWideDeepStep(ShardedWideDeepStep):
    """Wide&Deep step over PeerShardedTables.  Nothing in the step depends on a host-side size and the DenseLayer
    gradient all-reduce is a kernel over peer memory (PeerAllReduce), so `capture` records the WHOLE step — forward
    exchange, DenseLayers, gradient exchange, fused row updates, all-reduce, dense Adam, and the next batch's key phase
    underneath the DenseLayers — as ONE CUDA graph per key-phase buffer set: a step costs one graph launch of host
    time.  Branches of the graph (forked streams while capturing):

        main    serve -> signal/wait 2 -> expand -> DenseLayers fwd / loss / bwd -> segment sums into the owners'
                inboxes -> signal/wait 3 -> LazyAdam on the owned rows
        wide    the dim-1 twin of every exchange kernel (serve, reduce, gradient push, FTRL) beside the deep one
        plan    after expand: wait for the staged copy of the NEXT batch (an external event, so the copy itself stays
                outside the graph and overlaps the forward exchange), then its whole key phase on the other buffer set
        dense   after the DenseLayer backward: all-reduce over peer memory -> dense Adam -> loss poison, beside the
                gradient exchange

    The input batch is double-buffered too: set s reads slots[s]; the next batch is copied (H2D from pinned memory, or
    D2D) straight into slots[s ^ 1] on a copy stream, so no input copy sits on the step's critical path."""

    def __init__(self, batch_size, vocab_size, emb_dim, hidden, device, seed=1, sens=1024.0, fields=39,
                 use_mixed_precision=True, group=None, graph=True, cap_rows=None):
        self._ar = None

        def storage(n):
            self._ar = PeerAllReduce(n, device, group)
            return torch.zeros(n, dtype=torch.float32, device=device), self._ar.src

        super().__init__(batch_size, vocab_size, emb_dim, hidden, device, seed=seed, sens=sens, fields=fields,
                         use_mixed_precision=use_mixed_precision, group=group, graph_dense=False,
                         tables_factory=lambda: PeerShardedTables(vocab_size, emb_dim, batch_size * fields, device,
                                                                  group=group, seed=seed, sens=sens, cap_rows=cap_rows),
                         dense_storage=storage)
        self._nan = torch.tensor(float("nan"), dtype=torch.float32, device=self.device)
        self._zero = torch.zeros((), dtype=torch.float32, device=self.device)
        self._graph_step = graph
        self._graphs = None
        self._loss = torch.zeros((), dtype=torch.float32, device=self.device)
        # Priorities (captured into the graph's kernel nodes): the step's critical path — exchange kernels, DenseLayers,
        # row updates — is recorded on high-priority streams; the all-reduce + dense Adam branch and the next batch's
        # key phase only have to finish somewhere under it, so their CTAs queue behind the critical ones.
        self._main_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self._dense_stream = torch.cuda.Stream(device=self.device)
        self._stage_stream = torch.cuda.Stream(device=self.device)
        self._plan_stream = torch.cuda.Stream(device=self.device)
        self._staged_ev = torch.cuda.Event(external=True)    # a wait NODE in the graph: the copy stays outside
        self._staged_for = None

    def _dense_update(self):
        """All-reduce over peer memory + dense Adam (capturable; replaces the NCCL collective of the base class)."""
        self._ar.run(self.tables.rk.err)
        self._dense_adam()

    def _dense_adam(self):
        ops.adam_begin_step(self.dense_hyper)
        ops.adam_dense(self.dense.flat, self.dense_m, self.dense_v, self.dense_hyper, self._ar.dst)

    def _pre(self):
        """Key phase of the current batch in line (first step, or no look-ahead was given)."""
        self.tables.plan_batch(self._slots[self.tables.rk.cur][0])

    def _body(self, ahead=True, late=False):
        """ahead: run the next batch's key phase + early serve on the forked branch.  late: this batch went through that
        one step ago (it is being adopted), so the serve only moves the rows the previous update rewrote."""
        main = torch.cuda.current_stream()
        rk = self.tables.rk
        ids, wts, label = self._slots[rk.cur]
        self.tables.lookup(_DevicePlan(ids), wts, self.wide_b, self._io["deep_in"], self._io["wide_out"], late=late)
        if ahead:
            self._plan_stream.wait_stream(main)
            with torch.cuda.stream(self._plan_stream):
                self._plan_stream.wait_event(self._staged_ev)            # the next batch is on the device
                self.tables.key_phase_next(self._slots[rk.cur ^ 1][0])
        # The all-reduce branch forks as soon as the last weight gradient has been issued, underneath the
        # input-gradient GEMM of layer 0 (compute-bound, no NVLink traffic), so that most of it is over when the
        # gradient push — which shares the NVLink ports and IS on the critical path — starts.  The dense Adam follows
        # on the same branch once that GEMM, the last reader of the weights, has been issued.
        dense_stream = main if _NO_DENSE_FORK else self._dense_stream

        def fork_allreduce():
            dense_stream.wait_stream(main)
            with torch.cuda.stream(dense_stream):
                self._ar.run(rk.err)

        loss, delta, gx = self._dense_segment(label, on_weight_grads=fork_allreduce)
        dense_stream.wait_stream(main)
        with torch.cuda.stream(dense_stream):
            self._dense_adam()
            # an exchange error (a wait that timed out, an inbox that overflowed) must not train on silently: the loss
            # the caller reads turns NaN from the step after the one that raised the bit
            torch.add(loss, torch.where(rk.err[0] != 0, self._nan, self._zero), out=self._loss)
        self.tables.push_grads(delta, gx)
        self.tables.apply_grads()
        main.wait_stream(self._dense_stream)
        if ahead:
            main.wait_stream(self._plan_stream)
        return self._loss

    def _one_step(self, adopt=False, ahead=True):
        rk = self.tables.rk
        if adopt:
            rk.p_adopt()                                     # the set whose key phase ran under the previous step
        g = self._graphs[rk.cur] if self._graphs is not None else None
        if not adopt:
            g["pre"].replay() if g is not None else self._pre()
        # an adopted batch had the clean part of its rows served one step early (late = adopt)
        if g is not None:
            g[("step" if ahead else "step_plain") + ("" if adopt else "_full")].replay()
        else:
            self._body(ahead, late=adopt)
        return self._loss

    def capture(self, ids, wts, label, warmup=3):
        self._slots = [tuple(t.clone() for t in (ids, wts, label)) for _ in range(2)]
        self._ensure_io(ids)
        with torch.cuda.stream(self._stage_stream):
            self._staged_ev.record()
        for _ in range(warmup):
            self._one_step(ahead=False)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        if self._graph_step:
            # capture executes nothing: the pieces are recorded back to back, real steps are replayed afterwards
            rk = self.tables.rk
            cur0 = rk.cur
            sets, pool = [], None
            self.launches_per_step = 0
            for si in range(2):
                rk.use_set(si)
                graphs = {}
                for name, fn in (("pre", self._pre), ("step", lambda: self._body(True, late=True)),
                                 ("step_full", lambda: self._body(True, late=False)),
                                 ("step_plain", lambda: self._body(False, late=True)),
                                 ("step_plain_full", lambda: self._body(False, late=False))):
                    gr = torch.cuda.CUDAGraph()
                    n0 = _lib.launch_count()
                    with torch.cuda.graph(gr, pool=pool, stream=self._main_stream):
                        fn()
                    if si == 0 and name == "step":                   # steady state replays `step` only
                        self.launches_per_step = _lib.launch_count() - n0
                    pool = pool or gr.pool()
                    graphs[name] = gr
                sets.append(graphs)
            rk.use_set(cur0)
            self._graphs = sets
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
        self._staged_for = None
        return self._slots[self.tables.rk.cur]

    def replay(self, ids=None, wts=None, label=None, next_batch=None):
        """One step on (ids, wts, label).  next_batch (host-pinned or device tensors): copied into the other input
        slot on a copy stream during this step AND taken through its key phase (dedup, bucket bounds, key exchange,
        owner dedup) underneath this step's DenseLayers; pass the same tensors as the next call's batch to use both.
        Every rank must make the same choice (the key phase is a collective of the group)."""
        main = torch.cuda.current_stream()
        rk = self.tables.rk
        adopt = ids is not None and self._staged_for is not None and self._staged_for is ids
        if ids is not None and not adopt:
            for d, s in zip(self._slots[rk.cur], (ids, wts, label)):
                d.copy_(s, non_blocking=True)
        self._staged_for = None
        ahead = next_batch is not None
        if ahead:
            nxt = rk.cur if adopt else rk.cur ^ 1            # the slot of the set the next step will run on
            ev = torch.cuda.Event()
            ev.record(main)                                  # its last reader (the step before this one) is enqueued
            self._stage_stream.wait_event(ev)
            with torch.cuda.stream(self._stage_stream):
                for d, s in zip(self._slots[nxt], next_batch):
                    d.copy_(s, non_blocking=True)
                self._staged_ev.record()
            self._staged_for = next_batch[0]
        loss = self._one_step(adopt, ahead)
        return loss, loss

    def close(self):
        self.tables.close()
        self._ar.close()
