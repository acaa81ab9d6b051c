"""Wide&Deep multitable (models/wide_and_deep_multitable/src/wide_and_deep.py:146-427, 431-480, 495-600).

Four embedding tables and five 1-D wide weight vectors looked up by single-hot inputs (Gather + Flatten) and six
multi-hot inputs over ONE shared table (Gather * mask -> ReduceMean over all slots, masked ones included,
:291-346); a six-layer DenseLayer stack in mixed precision; the sigmoid cross-entropy loss; FTRL on every
parameter whose name contains "wide" (here that includes `wide_bias`, lower case, :510-520) and nn.Adam on the
rest.  The reference gathers with `P.Gather` (not sparse), so every table gets a DENSE gradient
(UnsortedSegmentSum into [V, D]) and a dense optimizer step over all rows; that is what this cell does:

    forward   mrec_gather (single-hot), mrec_gather_pool (multi-hot mean), mrec_gather_reduce (wide vectors)
    backward  mrec_unique_bounded per input -> mrec_segment_sum_scatter_add into the table's dense gradient
              (the six multi-hot inputs accumulate into one [V,64] gradient), then mrec_adam_dense / mrec_ftrl_dense

Input names follow the reference's construct signature; the four trailing inputs it ignores (display_id, ad_id,
display_ad_and_is_leak, is_leak) are accepted and unused.
"""
import torch

from . import ops
from .nn import DenseStack

MULTI_NAMES = ("multi_doc_ad_category_id", "multi_doc_event_entity_id", "multi_doc_ad_entity_id",
               "multi_doc_event_topic_id", "multi_doc_event_category_id", "multi_doc_ad_topic_id")


class MultitableConfig:
    """Model constants of wide_and_deep.py:154-163 and src/config.py:24-35; the per-input field counts come from the
    dataset's input_shape_dict (datasets.py:296-313) and are constructor arguments here."""

    def __init__(self, batch_size=131072, n_indicator=4, n_emb128=6, n_emb64_single=8, multi_slots=(4, 4, 4, 4, 4, 4),
                 continue_field_size=32, emb_128_size=650000, emb64_single_size=17300, emb64_multi_size=20900,
                 indicator_size=16, deep_dim_list=(1024, 1024, 1024, 1024, 1024), adam_lr=3e-3, ftrl_lr=0.1,
                 use_mixed_precision=True, seed=1):
        if len(multi_slots) != len(MULTI_NAMES):
            raise ValueError("multi_slots needs one slot count per multi-hot input (%d)" % len(MULTI_NAMES))
        self.batch_size = batch_size
        self.n_indicator, self.n_emb128, self.n_emb64_single = n_indicator, n_emb128, n_emb64_single
        self.multi_slots = tuple(multi_slots)
        self.continue_field_size = continue_field_size
        self.emb_128_size, self.emb64_single_size = emb_128_size, emb64_single_size
        self.emb64_multi_size, self.indicator_size = emb64_multi_size, indicator_size
        self.deep_dim_list = tuple(deep_dim_list)
        self.adam_lr, self.ftrl_lr = adam_lr, ftrl_lr
        self.use_mixed_precision = use_mixed_precision
        self.seed = seed
        # datasets.py:310-315
        self.input_emb_dim = (continue_field_size + n_indicator * 64 + n_emb128 * 128 + n_emb64_single * 64 +
                              64 * len(MULTI_NAMES))


class MultitableWideDeepModel:
    """wide_and_deep.py:146-427.  construct(...) -> logit [B,1]."""

    def __init__(self, config, device="cuda"):
        self.config = c = config
        self.device = dev = torch.device(device)
        gen = torch.Generator(device=dev)
        gen.manual_seed(c.seed)
        normal = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev).normal_(0.0, 0.01, generator=gen)
        # deep tables (Adam group)
        self.emb128_embedding = normal(c.emb_128_size, 128)
        self.emb64_single = normal(c.emb64_single_size, 64)
        self.emb64_multi = normal(c.emb64_multi_size, 64)
        self.emb64_indicator = normal(c.indicator_size, 64)
        # wide vectors (FTRL group: every name contains "wide"), kept [V,1] for the dim-1 kernels
        self.wide_continue_w = normal(c.continue_field_size)
        self.wide_emb128_w = normal(c.emb_128_size, 1)
        self.wide_emb64_single_w = normal(c.emb64_single_size, 1)
        self.wide_emb64_multi_w = normal(c.emb64_multi_size, 1)
        self.wide_indicator_w = normal(c.indicator_size, 1)
        self.wide_bias = normal(1)
        dims = [c.input_emb_dim] + list(c.deep_dim_list) + [1]
        self.dense = DenseStack(dims, c.use_mixed_precision, dev, generator=gen, weight_init="normal", bias_init="normal")
        self._zero_bias = torch.zeros(1, dtype=torch.float32, device=dev)
        self._ctx = None

    def deep_tables(self):
        return {"emb128_embedding": self.emb128_embedding, "emb64_single": self.emb64_single,
                "emb64_multi": self.emb64_multi, "emb64_indicator": self.emb64_indicator}

    def wide_tables(self):
        return {"wide_continue_w": self.wide_continue_w, "wide_emb128_w": self.wide_emb128_w,
                "wide_emb64_single_w": self.wide_emb64_single_w, "wide_emb64_multi_w": self.wide_emb64_multi_w,
                "wide_indicator_w": self.wide_indicator_w, "wide_bias": self.wide_bias}

    def __call__(self, *a, **kw):
        return self.construct(*a, **kw)

    def construct(self, continue_val, indicator_id, emb_128_id, emb_64_single_id, multi_ids, multi_masks,
                  display_id=None, ad_id=None, display_ad_and_is_leak=None, is_leak=None):
        """multi_ids / multi_masks: the six (ids [B,S_k], mask [B,S_k]) pairs in MULTI_NAMES order."""
        b = continue_val.shape[0]
        dev = self.device
        ones = lambda t: torch.ones(t.shape, dtype=torch.float32, device=dev)
        # ---- deep input (:291-349): [continue | indicator | emb128 | emb64 single | six pooled] ----
        parts = [continue_val.to(torch.float32),
                 ops.gather(self.emb64_indicator, indicator_id).view(b, -1),
                 ops.gather(self.emb128_embedding, emb_128_id).view(b, -1),
                 ops.gather(self.emb64_single, emb_64_single_id).view(b, -1)]
        masks = [m.to(torch.float32).contiguous() for m in multi_masks]
        for ids, m in zip(multi_ids, masks):
            parts.append(ops.gather_pool(self.emb64_multi, ids, m))          # mean over ALL slots (:303-307)
        deep_in = torch.cat(parts, 1)
        deep_out = self.dense.forward(deep_in.half() if self.config.use_mixed_precision else deep_in)
        # ---- wide (:357-420) ----
        wide = (continue_val.to(torch.float32) * self.wide_continue_w[None, :]).sum(1, keepdim=True)
        wide = wide + ops.gather_reduce(self.wide_indicator_w, indicator_id, ones(indicator_id), self._zero_bias)
        wide = wide + ops.gather_reduce(self.wide_emb128_w, emb_128_id, ones(emb_128_id), self._zero_bias)
        wide = wide + ops.gather_reduce(self.wide_emb64_single_w, emb_64_single_id, ones(emb_64_single_id), self._zero_bias)
        for ids, m in zip(multi_ids, masks):
            wide = wide + ops.gather_reduce(self.wide_emb64_multi_w, ids, m, self._zero_bias)
        wide = wide + self.wide_bias
        self._ctx = (continue_val, indicator_id, emb_128_id, emb_64_single_id, list(multi_ids), masks)
        self.wide_out, self.deep_out = wide, deep_out.float()
        return wide + self.deep_out


class NetWithLossClass:
    """wide_and_deep.py:431-480: wide_loss = deep_loss = mean SigmoidCrossEntropyWithLogits."""

    def __init__(self, network, config):
        self.network = network
        self.sens_t = torch.ones(1, dtype=torch.float32, device=network.device)

    def __call__(self, label, *inputs, **kw):
        net = self.network
        net(*inputs, **kw)
        _, loss, self.delta, self.delta16, self.delta_sum = ops.sigmoid_xent(
            net.wide_out, net.deep_out, label.reshape(-1, 1).to(torch.float32), self.sens_t,
            half=net.config.use_mixed_precision)
        return loss[0], loss[0]


class TrainStepWrap:
    """wide_and_deep.py:495-600: FTRL(lr=ftrl_lr, l1=l2=5e-4, initial_accum=0.1) on the "wide" names,
    Adam(lr=adam_lr, eps=1e-6) on the rest, loss scale sens (default 1000)."""

    def __init__(self, network, config, sens=1000.0):
        self.network = network
        self.model = m = network.network
        self.sens = float(sens)
        network.sens_t.fill_(self.sens)
        dev = m.device
        self.adam_hyper = ops.adam_hyper(config.adam_lr, eps=1e-6, loss_scale=sens, device=dev)
        self.ftrl_hyper = ops.ftrl_hyper(config.ftrl_lr, l1=5e-4, l2=5e-4, loss_scale=sens, device=dev)
        self.deep = m.deep_tables()
        self.wide = m.wide_tables()
        z = torch.zeros_like
        self.adam_state = {k: (z(w), z(w)) for k, w in self.deep.items()}
        self.adam_state["dense"] = (z(m.dense.flat), z(m.dense.flat))
        self.ftrl_state = {k: (torch.full_like(w, 0.1), z(w)) for k, w in self.wide.items()}
        self.grads = {k: z(w) for k, w in list(self.deep.items()) + list(self.wide.items())}

    def __call__(self, label, *inputs, **kw):
        return self.construct(label, *inputs, **kw)

    def _scatter(self, name, table, ids, values, mask):
        """Dense gradient of a Gather: grads[name][ids[n]] += mask[n] * values[n // div]."""
        uq = ops.unique(ids.reshape(-1), table_like=table)
        ops.segment_sum_scatter_add(self.grads[name], values, mask, uq)

    def construct(self, label, *inputs, **kw):
        m = self.model
        loss_w, loss_d = self.network(label, *inputs, **kw)
        net = self.network
        continue_val, indicator_id, emb_128_id, emb_64_single_id, multi_ids, masks = m._ctx
        b = continue_val.shape[0]
        c = m.config
        delta = net.delta                                    # sens * (sigmoid - y) / B, [B,1]
        seed = net.delta16 if net.delta16.numel() else delta
        gx = m.dense.backward(seed).float()                  # [B, input_emb_dim]
        for g in self.grads.values():
            g.zero_()
        # ---- deep tables: slices of gx in concat order (:347-349) ----
        o = c.continue_field_size
        w = c.n_indicator * 64
        self._scatter("emb64_indicator", m.emb64_indicator, indicator_id, gx[:, o:o + w].reshape(-1, 64).contiguous(), None)
        o += w
        w = c.n_emb128 * 128
        self._scatter("emb128_embedding", m.emb128_embedding, emb_128_id, gx[:, o:o + w].reshape(-1, 128).contiguous(), None)
        o += w
        w = c.n_emb64_single * 64
        self._scatter("emb64_single", m.emb64_single, emb_64_single_id, gx[:, o:o + w].reshape(-1, 64).contiguous(), None)
        o += w
        for ids, mask in zip(multi_ids, masks):
            s = ids.shape[1]
            # ReduceMean over the S slots: every slot gets g_pooled * mask / S (values broadcast over the slots)
            self._scatter("emb64_multi", m.emb64_multi, ids, gx[:, o:o + 64].contiguous(), (mask / s).contiguous())
            o += 64
        # ---- wide vectors: d logit = delta on every term ----
        self.grads["wide_continue_w"].copy_((delta * continue_val.to(torch.float32)).sum(0))
        self.grads["wide_bias"].copy_(net.delta_sum)
        self._scatter("wide_indicator_w", m.wide_indicator_w, indicator_id, delta, None)
        self._scatter("wide_emb128_w", m.wide_emb128_w, emb_128_id, delta, None)
        self._scatter("wide_emb64_single_w", m.wide_emb64_single_w, emb_64_single_id, delta, None)
        for ids, mask in zip(multi_ids, masks):
            self._scatter("wide_emb64_multi_w", m.wide_emb64_multi_w, ids, delta, mask)
        # ---- optimizers: dense steps over every parameter (P.Gather gradients are dense) ----
        for k, wt in self.wide.items():
            acc, lin = self.ftrl_state[k]
            ops.ftrl_dense(wt, acc, lin, self.ftrl_hyper, self.grads[k])
        ops.adam_begin_step(self.adam_hyper)
        for k, wt in self.deep.items():
            mm, vv = self.adam_state[k]
            ops.adam_dense(wt, mm, vv, self.adam_hyper, self.grads[k])
        mm, vv = self.adam_state["dense"]
        ops.adam_dense(m.dense.flat, mm, vv, self.adam_hyper, m.dense.flat_grad)
        return loss_w, loss_d
