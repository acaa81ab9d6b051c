"""DenseLayer backward glue (mrec_relu_bwd_bias) and the fused loss kernel vs float64 numpy restatements."""
import numpy as np
import pytest
import torch

from mindrec_b200 import nn, ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("rows,cols", [(16000, 1024), (16000, 128), (1000, 1), (37, 24), (5, 7), (1, 512), (4099, 260)])
@pytest.mark.parametrize("masked", [True, False])
def test_relu_bwd_bias_matches_numpy(cuda, dtype, rows, cols, masked):
    rng = np.random.default_rng(rows * 31 + cols)
    npdt = np.float16 if dtype == torch.float16 else np.float32
    g = rng.standard_normal((rows, cols)).astype(npdt)
    y = np.maximum(rng.standard_normal((rows, cols)), 0).astype(npdt)      # post-ReLU activations (many exact zeros)
    g_t, y_t = torch.from_numpy(g).to(cuda), torch.from_numpy(y).to(cuda)
    gb = torch.full((cols,), 7.0, dtype=torch.float32, device=cuda)
    for _ in range(2):                                                        # twice: the ticket counters self-reset
        gz = ops.relu_bwd_bias(g_t.clone(), y_t if masked else None, gb)
        want = np.where(y > 0, g, 0).astype(npdt) if masked else g
        assert np.array_equal(gz.cpu().numpy(), want)                         # the mask is a bit-exact select
        ref = want.astype(np.float64).sum(0)
        np.testing.assert_allclose(gb.cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * np.sqrt(rows))


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("rows,k,masked", [(16000, 128, True), (777, 128, False), (33, 5, True), (1, 300, True)])
def test_dense_head_matches_numpy(cuda, dtype, rows, k, masked):
    rng = np.random.default_rng(rows + k)
    npdt = np.float16 if dtype == torch.float16 else np.float32
    h = np.maximum(rng.standard_normal((rows, k)), 0).astype(npdt)
    w = (rng.standard_normal(k) * 0.1).astype(npdt)
    bias = np.array([0.3], dtype=npdt)
    delta = (rng.standard_normal((rows, 1)) * 0.01).astype(npdt)
    t = lambda a: torch.from_numpy(a).to(cuda)
    out = ops.dense_head_fwd(t(h), t(w), t(bias))
    want = h.astype(np.float64) @ w.astype(np.float64) + float(bias[0])
    np.testing.assert_allclose(out.cpu().numpy()[:, 0], want, rtol=1e-5, atol=1e-5)
    gw = torch.empty(k, device=cuda)
    gb_head = torch.empty(1, device=cuda)
    gb_prev = torch.empty(k, device=cuda)
    for _ in range(2):
        gh = ops.dense_head_bwd(t(delta), t(h), t(w), masked, gw, gb_head, gb_prev)
        prod = (delta.astype(np.float32) * w.astype(np.float32)[None, :]).astype(npdt)     # one rounding, like the GEMM
        want_gh = np.where(h > 0, prod, 0).astype(npdt) if masked else prod
        assert np.array_equal(gh.cpu().numpy(), want_gh)
        np.testing.assert_allclose(gw.cpu().numpy(), delta[:, 0].astype(np.float64) @ h.astype(np.float64),
                                   rtol=1e-5, atol=1e-6 * np.sqrt(rows))
        np.testing.assert_allclose(gb_head.cpu().numpy(), [delta.astype(np.float64).sum()], rtol=1e-5, atol=1e-7 * rows)
        np.testing.assert_allclose(gb_prev.cpu().numpy(), want_gh.astype(np.float64).sum(0), rtol=1e-5,
                                   atol=1e-6 * np.sqrt(rows))


def test_relu_bwd_bias_is_run_to_run_deterministic(cuda):
    gen = torch.Generator(device=cuda)
    gen.manual_seed(11)
    g = torch.randn((16000, 512), device=cuda, generator=gen).half()
    y = torch.relu(torch.randn((16000, 512), device=cuda, generator=gen)).half()
    outs = []
    for _ in range(3):
        gb = torch.empty(512, device=cuda)
        ops.relu_bwd_bias(g.clone(), y, gb)
        outs.append(gb.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("mixed", [True, False])
def test_dense_stack_backward_matches_autograd(cuda, mixed):
    """Explicit backward of the DenseLayer chain vs torch autograd on the same (fp32 master) weights."""
    dims, b = [48, 64, 32, 1], 300
    gen = torch.Generator(device=cuda)
    gen.manual_seed(5)
    stack = nn.DenseStack(dims, mixed, cuda, generator=gen, weight_init="normal", bias_init="normal")
    for w in stack.weights:
        w.mul_(20.0)                                     # activations of order 1 so the ReLU mask is exercised
    x = torch.randn((b, dims[0]), device=cuda, generator=gen)
    g_out = torch.randn((b, 1), device=cuda, generator=gen)
    out = stack.forward(x.half() if mixed else x)
    gx = stack.backward(g_out)
    ws = [w.detach().clone().requires_grad_(True) for w in stack.weights]
    bs = [v.detach().clone().requires_grad_(True) for v in stack.biases]
    xr = x.clone().requires_grad_(True)
    h = xr
    for i, (w, v) in enumerate(zip(ws, bs)):
        h = h @ w + v
        if i + 1 < len(ws):
            h = torch.relu(h)
    h.backward(g_out)
    errs = {}

    def close(name, a, b):
        if mixed:      # fp16 activations: a ReLU mask can flip next to zero, so compare in norm, not element-wise
            errs[name] = float((a.float() - b).norm() / b.norm())
        else:
            torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
    close("out", out.float(), h.detach())
    close("gx", gx.float(), xr.grad)
    for i in range(len(ws)):
        close("gw%d" % i, stack.gw[i], ws[i].grad)
        close("gb%d" % i, stack.gb[i], bs[i].grad)
    assert all(e < 5e-2 for e in errs.values()), errs


@pytest.mark.parametrize("mixed", [False, True])
def test_dense_stack_input_in_column_blocks_equals_concatenated_input(cuda, mixed):
    """DenseStack.forward((x_a, x_b)) — two lookups feeding one tower without a concatenation copy (config 5) — against
    the same stack on cat(x_a, x_b): outputs, every weight / bias gradient and the per-block input gradients; and
    on_weight_grads fires once, after the last weight gradient and before the layer-0 input gradient."""
    dims = [96 + 160, 64, 32, 1]
    gen = torch.Generator(device=cuda)
    gen.manual_seed(3)
    a = nn.DenseStack(dims, mixed, cuda, generator=gen, bias_init="normal")
    b = nn.DenseStack(dims, mixed, cuda)
    b.flat.copy_(a.flat)                                  # (the constructor initialises its own weights)
    x = torch.randn((300, dims[0]), device=cuda, generator=gen)
    if mixed:
        x = x.half()
    xa, xb = x[:, :96].contiguous(), x[:, 96:].contiguous()
    ya, yb = a.forward(x), b.forward((xa, xb))
    tol = dict(rtol=2e-2, atol=2e-3) if mixed else dict(rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(yb, ya, **tol)
    go = torch.randn((300, 1), device=cuda, generator=gen) * 0.01
    calls = []
    ga = a.backward(go)
    gb = b.backward(go, on_weight_grads=lambda: calls.append(float(b.gw[0].abs().sum())))
    assert len(calls) == 1 and calls[0] > 0          # issued after layer 0's weight gradient
    assert isinstance(gb, tuple) and [g.shape[1] for g in gb] == [96, 160]
    torch.testing.assert_close(torch.cat(gb, 1).float(), ga.float(), **tol)
    torch.testing.assert_close(b.flat_grad, a.flat_grad, **(dict(rtol=2e-2, atol=2e-3) if mixed else dict(rtol=1e-4, atol=1e-6)))
