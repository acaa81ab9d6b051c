"""DeepFM and Deep&Cross cells of the reference, restated over the aot kernels.

    DeepFMModel / DeepFMNetWithLoss / DeepFMTrainStep     models/deepfm/src/deepfm.py:152-297
    CrossLayer / DeepCrossModel / DeepCrossTrainStep      models/deep_and_cross/src/deep_and_cross.py:117-370

Lookup, mask multiply, linear reduce, FM second order, the cross stack, the sparse gradient dedup and the
optimizer updates run in libmindrec_b200.so; DenseLayers use library GEMMs.
"""
import torch

from . import ops
from .nn import Adam, DenseStack, Parameter, RowTensor, _init_table


# --------------------------------------------------------------------------------------------------
# DeepFM
# --------------------------------------------------------------------------------------------------
class DeepFMConfig:
    """models/deepfm/default_config.yaml:13-33 (same names)."""

    def __init__(self, batch_size=16000, data_field_size=39, data_vocab_size=184965, data_emb_dim=80,
                 deep_layer_args=((1024, 512, 256, 128), "relu"), init_args=(-0.01, 0.01),
                 weight_bias_init=("normal", "normal"), keep_prob=0.9, convert_dtype=True,
                 learning_rate=5e-4, epsilon=5e-8, l2_coef=8e-5, loss_scale=1024.0, seed=1):
        self.batch_size = batch_size
        self.data_field_size = data_field_size
        self.data_vocab_size = data_vocab_size
        self.data_emb_dim = data_emb_dim
        self.deep_layer_args = deep_layer_args
        self.init_args = init_args
        self.weight_bias_init = weight_bias_init
        self.keep_prob = keep_prob
        self.convert_dtype = convert_dtype
        self.learning_rate = learning_rate
        self.epsilon = epsilon
        self.l2_coef = l2_coef
        self.loss_scale = loss_scale
        self.seed = seed


class DeepFMModel:
    """deepfm.py:152-237: out = linear + fm + deep; returns (out, fm_w, embedding_table)."""

    def __init__(self, config, device="cuda"):
        self.config = config
        self.batch_size = config.batch_size
        self.field_size = config.data_field_size
        self.vocab_size = config.data_vocab_size
        self.emb_dim = config.data_emb_dim
        dims_hidden, act = config.deep_layer_args
        if act != "relu":
            raise ValueError("only deep_layer_act='relu' is implemented")
        self.device = torch.device(device)
        gen = torch.Generator(device=self.device)
        gen.manual_seed(config.seed)
        self.fm_w = Parameter(_init_table((self.vocab_size, 1), "normal", self.device, gen), name="W_l2")
        self.embedding_table = Parameter(_init_table((self.vocab_size, self.emb_dim), "normal", self.device, gen),
                                         name="V_l2")
        dims = [self.field_size * self.emb_dim] + list(dims_hidden) + [1]
        w_init, b_init = config.weight_bias_init
        self.dense = DenseStack(dims, config.convert_dtype, self.device, generator=gen, weight_init=w_init,
                                bias_init=b_init)
        self._zero_bias = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._buf = None

    def __call__(self, id_hldr, wt_hldr):
        return self.construct(id_hldr, wt_hldr)

    def construct(self, id_hldr, wt_hldr):
        b, f, d = id_hldr.shape[0], self.field_size, self.emb_dim
        if self._buf is None or self._buf[0].shape[0] != b:
            self._buf = (torch.empty((b, 1), dtype=torch.float32, device=self.device),
                         torch.empty((b, f * d), dtype=torch.float32, device=self.device),
                         torch.empty((b, 1), dtype=torch.float32, device=self.device))
        linear_out, vx, fm_out = self._buf
        ops.gather_reduce(self.fm_w.data, id_hldr, wt_hldr, self._zero_bias, out=linear_out)   # :217-219
        ops.gather_masked(self.embedding_table.data, id_hldr, wt_hldr, out=vx)                # :221-222
        ops.fm_fwd(vx.view(b, f, d), out=fm_out)                                               # :223-228
        deep_out = self.dense.forward(vx)                                                      # :230-235
        self.vx = vx
        return linear_out + fm_out + deep_out, self.fm_w, self.embedding_table


class DeepFMNetWithLoss:
    """deepfm.py:240-260: mean sigmoid xent + l2_coef * (sum V^2 + sum W^2) * 0.5 over the FULL tables."""

    def __init__(self, network, l2_coef=1e-6):
        self.network = network
        self.l2_coef = l2_coef
        self.logit = None

    def __call__(self, batch_ids, batch_wts, label):
        predict, fm_w, fm_v = self.network(batch_ids, batch_wts)
        self.logit = predict
        log_loss = torch.clamp(predict, min=0) - predict * label + torch.log1p(torch.exp(-predict.abs()))
        l2 = self.l2_coef * (fm_v.data.square().sum() + fm_w.data.square().sum()) * 0.5
        return log_loss.mean() + l2


class DeepFMTrainStep:
    """deepfm.py:263-297 (TrainStepWrap): one nn.Adam(lr, eps, loss_scale) over every weight.  The l2 term
    makes both table gradients dense, so the tables take the dense-equivalent Adam update with the
    l2 * w term on every row (mrec_adam_rowsparse, hyper[8] = l2_coef)."""

    def __init__(self, network, lr=5e-4, eps=5e-8, loss_scale=1000.0):
        self.network = network
        self.model = network.network
        m = self.model
        self.sens = float(loss_scale)
        self.weights = [m.fm_w, m.embedding_table, Parameter(m.dense.flat, name="dense_layers")]
        self.optimizer = Adam(self.weights, learning_rate=lr, eps=eps, loss_scale=loss_scale)
        self.optimizer.hyper[8] = network.l2_coef
        self._uq = None
        self._dvx = None

    def __call__(self, batch_ids, batch_wts, label):
        m = self.model
        b, f, d = batch_ids.shape[0], m.field_size, m.emb_dim
        loss = self.network(batch_ids, batch_wts, label)
        delta = (torch.sigmoid(self.network.logit) - label) * (self.sens / b)
        g_deep = m.dense.backward(delta)                                  # [B, F*D] (fp16 when mixed)
        if self._dvx is None or self._dvx.shape[0] != b:
            self._dvx = torch.empty((b, f, d), dtype=torch.float32, device=batch_ids.device)
        ops.fm_bwd(m.vx.view(b, f, d), delta, out=self._dvx, addend=g_deep.view(b, f, d))
        n = batch_ids.numel()
        if self._uq is None or self._uq.n != n:
            self._uq = ops.UniqueResult(n, batch_ids.dtype, batch_ids.device)
        uq = ops.unique(batch_ids, table_like=m.embedding_table.data, result=self._uq)
        mask = batch_wts.reshape(-1)
        self.optimizer([RowTensor(batch_ids, delta, mask, uq),
                        RowTensor(batch_ids, self._dvx.view(n, d), mask, uq), m.dense.flat_grad])
        return loss


# --------------------------------------------------------------------------------------------------
# Deep & Cross
# --------------------------------------------------------------------------------------------------
class DeepCrossConfig:
    """models/deep_and_cross/src/config.py:61-88."""

    def __init__(self, batch_size=16000, field_size=39, vocab_size=200000, emb_dim=27,
                 deep_layer_dim=(1024, 1024), cross_layer_num=6, keep_prob=1.0, learning_rate=1e-4,
                 epsilon=1e-8, loss_scale=1000.0, seed=1):
        self.batch_size = batch_size
        self.field_size = field_size
        self.vocab_size = vocab_size
        self.emb_dim = emb_dim
        self.deep_layer_dim = list(deep_layer_dim)
        self.cross_layer_num = cross_layer_num
        self.keep_prob = keep_prob
        self.learning_rate = learning_rate
        self.epsilon = epsilon
        self.loss_scale = loss_scale
        self.seed = seed


class CrossLayer:
    """deep_and_cross.py:117-149: y = x_0 * (x_l . w) + b + x_l.  A single layer is the L = 1 stack."""

    def __init__(self, cross_raw_dim, cross_col_dim, weight_bias_init=("normal", "normal"), device="cuda",
                 generator=None):
        self.cross_weight = Parameter(_init_table((1, cross_col_dim), weight_bias_init[0], device, generator), "weight")
        self.cross_bias = Parameter(_init_table((1, cross_col_dim), weight_bias_init[1], device, generator), "bias")

    def __call__(self, inputs, x_0):
        if inputs.data_ptr() == x_0.data_ptr():
            return ops.cross_fwd(x_0, self.cross_weight.data, self.cross_bias.data)[0]
        # general x_l: x_0 * (x_l . w) + b + x_l  (only reached when a caller chains layers by hand)
        s = inputs @ self.cross_weight.data.view(-1, 1)
        return x_0 * s + self.cross_bias.data + inputs


class DeepCrossModel:
    """deep_and_cross.py:206-309: embedding -> (2-layer deep tower || 6-layer cross stack) -> concat -> logit."""

    def __init__(self, config, device="cuda"):
        self.config = config
        self.batch_size = config.batch_size
        self.field_size = config.field_size
        self.emb_dim = config.emb_dim
        self.input_size = self.field_size * self.emb_dim
        self.layers = config.cross_layer_num
        self.device = torch.device(device)
        gen = torch.Generator(device=self.device)
        gen.manual_seed(config.seed)
        # normal_weight(shape, emb_dim) = N(0, emb_dim^-0.5)  (:49-51,175-177)
        table = torch.empty((config.vocab_size, self.emb_dim), dtype=torch.float32, device=self.device)
        table.normal_(0.0, self.emb_dim ** -0.5, generator=gen)
        self.embedding_table = Parameter(table, name="deep_embeddinglookup.embedding_table")
        tower_dims = [self.input_size] + list(config.deep_layer_dim)
        head_dims = [self.input_size + config.deep_layer_dim[-1], 1]
        al = lambda n: (n + 3) // 4 * 4          # keep every block 16-byte aligned inside the flat buffer
        n_tower = DenseStack.numel(tower_dims)
        n_head = DenseStack.numel(head_dims)
        lw = al(self.layers * self.input_size)
        self.flat = torch.zeros(al(n_tower) + al(n_head) + 2 * lw, dtype=torch.float32, device=self.device)
        self.flat_grad = torch.zeros_like(self.flat)
        o = 0
        self.tower = DenseStack(tower_dims, False, self.device, generator=gen, weight_init="normal",
                                bias_init="normal", last_activation=True,
                                storage=(self.flat[o:o + n_tower], self.flat_grad[o:o + n_tower]))
        o += al(n_tower)
        self.head = DenseStack(head_dims, False, self.device, generator=gen, weight_init="normal",
                               bias_init="normal", storage=(self.flat[o:o + n_head], self.flat_grad[o:o + n_head]))
        o += al(n_head)
        lw = self.layers * self.input_size
        self.cross_weight = self.flat[o:o + lw].view(self.layers, self.input_size)
        self.cross_weight_grad = self.flat_grad[o:o + lw].view(self.layers, self.input_size)
        o += al(lw)
        self.cross_bias = self.flat[o:o + lw].view(self.layers, self.input_size)
        self.cross_bias_grad = self.flat_grad[o:o + lw].view(self.layers, self.input_size)
        self.cross_weight.normal_(0.0, 0.01, generator=gen)
        self.cross_bias.normal_(0.0, 0.01, generator=gen)
        self._buf = None

    def __call__(self, id_hldr, wt_hldr):
        return self.construct(id_hldr, wt_hldr)

    def construct(self, id_hldr, wt_hldr):
        b = id_hldr.shape[0]
        if self._buf is None or self._buf[0].shape[0] != b:
            self._buf = (torch.empty((b, self.input_size), dtype=torch.float32, device=self.device),
                         torch.empty((b, self.input_size + self.tower.dims[-1]), dtype=torch.float32, device=self.device),
                         torch.empty((b, self.layers), dtype=torch.float32, device=self.device))
        input_x, cat, p = self._buf
        ops.gather_masked(self.embedding_table.data, id_hldr, wt_hldr, out=input_x)    # :295-298
        d_2 = self.tower.forward(input_x)                                              # :299-300
        k = self.tower.dims[-1]
        cat[:, :k].copy_(d_2)
        c_6 = torch.empty_like(input_x)
        ops.cross_fwd(input_x, self.cross_weight, self.cross_bias, y=c_6, p=p)         # :301-306
        cat[:, k:].copy_(c_6)                                                          # :307 concat((d_2, c_6))
        self.input_x, self.p = input_x, p
        return self.head.forward(cat)                                                  # :308


class DeepCrossNetWithLoss:
    """deep_and_cross.py:312-328: mean sigmoid cross entropy."""

    def __init__(self, network):
        self.network = network
        self.logit = None

    def __call__(self, batch_ids, batch_wts, label):
        predict = self.network(batch_ids, batch_wts)
        self.logit = predict
        log_loss = torch.clamp(predict, min=0) - predict * label + torch.log1p(torch.exp(-predict.abs()))
        return log_loss.mean()


class DeepCrossTrainStep:
    """deep_and_cross.py:331-357 (TrainStepWrap): nn.Adam(lr=1e-4, eps=1e-8, loss_scale=1000) on every weight;
    the table gradient of P.Gather is dense, so the table takes the dense-equivalent Adam update."""

    def __init__(self, network, lr=1e-4, eps=1e-8, loss_scale=1000.0):
        self.network = network
        self.model = network.network
        m = self.model
        self.sens = float(loss_scale)
        self.weights = [m.embedding_table, Parameter(m.flat, name="dense+cross")]
        self.optimizer = Adam(self.weights, learning_rate=lr, eps=eps, loss_scale=loss_scale)
        self._uq = None

    def __call__(self, batch_ids, batch_wts, label):
        m = self.model
        b = batch_ids.shape[0]
        loss = self.network(batch_ids, batch_wts, label)
        delta = (torch.sigmoid(self.network.logit) - label) * (self.sens / b)
        g_cat = m.head.backward(delta)                                   # [B, 1024 + D']
        k = m.tower.dims[-1]
        g_x = m.tower.backward(g_cat[:, :k].contiguous())               # deep-tower path to input_x
        g_c6 = g_cat[:, k:].contiguous()
        dx, _, _ = ops.cross_bwd(m.input_x, g_c6, m.cross_weight, m.cross_bias, m.p,
                                 dw=m.cross_weight_grad, db=m.cross_bias_grad)
        g_x = g_x + dx
        n = batch_ids.numel()
        if self._uq is None or self._uq.n != n:
            self._uq = ops.UniqueResult(n, batch_ids.dtype, batch_ids.device)
        uq = ops.unique(batch_ids, table_like=m.embedding_table.data, result=self._uq)
        self.optimizer([RowTensor(batch_ids, g_x.view(n, m.emb_dim), batch_wts.reshape(-1), uq), m.flat_grad])
        return loss
