"""Design validation for the training row: the backward pass of every distinct conv-block geometry of the three
networks, written with the forward kernels' primitives + wgrad GEMMs + the BatchNorm/ReLU/residual pass
(oracle/backward_recipe.py), against torch autograd (float64, so the comparison checks the recipe, not rounding)."""
import pytest
import torch

from oracle import backward_recipe as R
from oracle import w2l_oracle as O


def distinct_rows():
    rows, seen = [], set()
    allr = [r for _, r in O.generator_layers()] + [r for _, r in O.syncnet_layers()] + [r for _, r in O.disc_layers()]
    # input sizes as they occur in the nets (one representative per geometry, shrunk channels keep the test fast)
    for r in allr:
        key = (r[0], r[3] if isinstance(r[3], int) else tuple(r[3]), r[4] if isinstance(r[4], int) else tuple(r[4]),
               r[5] if isinstance(r[5], int) else tuple(r[5]), r[6], r[7])
        if key not in seen:
            seen.add(key)
            rows.append(r)
    return rows


def shrink(row, c=6):
    kind, cin, cout, k, s, p, op, res = row
    cin2 = c if cin > c else cin
    cout2 = cin2 if res else (c + 2 if cout > c else cout)
    return (kind, cin2, cout2, k, s, p, op, res)


@pytest.mark.parametrize("row", distinct_rows(), ids=lambda r: f"{r[0]}-k{r[3]}-s{r[4]}-p{r[5]}-op{r[6]}-res{int(r[7])}")
@pytest.mark.parametrize("hw", [(12, 13), (9, 16)])
def test_block_backward_recipe_matches_autograd(row, hw):
    row = shrink(row)
    kind, cin, cout, k, s, p, op, res = row
    (kh, kw), (ph, pw) = O._pair(k), O._pair(p)
    H, W = hw
    if kind != "t" and (H + 2 * ph < kh or W + 2 * pw < kw):
        pytest.skip("input smaller than the filter")
    g = torch.Generator().manual_seed(hash((kind, cin, cout, H, W)) % 1000)
    dt = torch.float64
    x = torch.randn((3, cin, H, W), generator=g, dtype=dt, requires_grad=True)
    wshape = (cin, cout, kh, kw) if kind == "t" else (cout, cin, kh, kw)
    w = (0.3 * torch.randn(wshape, generator=g, dtype=dt)).requires_grad_(True)
    b = (0.1 * torch.randn(cout, generator=g, dtype=dt)).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(cout, generator=g, dtype=dt)).requires_grad_(True)
    beta = (0.1 * torch.randn(cout, generator=g, dtype=dt)).requires_grad_(True)
    y, saved = R.block_forward_train(x, w, b, gamma, beta, row)
    dy = torch.randn(y.shape, generator=g, dtype=dt)
    params = [x, w, b] + ([] if kind == "n" else [gamma, beta])
    ref = torch.autograd.grad(y, params, dy)
    with torch.no_grad():
        got = R.block_backward(dy, x, w, gamma, row, {k_: v.detach() for k_, v in saved.items()})
    assert torch.allclose(got["dx"], ref[0], rtol=1e-9, atol=1e-10)
    assert torch.allclose(got["dw"], ref[1], rtol=1e-9, atol=1e-10)
    assert torch.allclose(got["db"], ref[2], rtol=1e-9, atol=1e-8)
    if kind != "n":
        assert torch.allclose(got["dgamma"], ref[3], rtol=1e-9, atol=1e-10)
        assert torch.allclose(got["dbeta"], ref[4], rtol=1e-9, atol=1e-10)
        # the train-mode forward itself equals torch's batch_norm path
        y2 = O.block_forward(x.detach().float(), {"b.conv_block.0.weight": w.detach().float(), "b.conv_block.0.bias": b.detach().float(),
                                                  "b.conv_block.1.weight": gamma.detach().float(), "b.conv_block.1.bias": beta.detach().float(),
                                                  "b.conv_block.1.running_mean": torch.zeros(cout), "b.conv_block.1.running_var": torch.ones(cout)},
                             "b", row, training=True)
        assert torch.allclose(y2, y.detach().float(), rtol=1e-4, atol=1e-5)


def test_generator_backward_schedule_matches_autograd():
    """The whole generator: train-mode forward + backward as an explicit schedule of block calls (skip-concat gradient
    split, residual accumulation) == autograd through oracle.generator_forward(training=True), every parameter."""
    torch.manual_seed(0)
    sd32 = O.make_state_dict("generator", 0, init="default")
    sd = {k: v.double() for k, v in sd32.items() if v.dtype.is_floating_point}
    mel, face = O.make_generator_inputs(2, seed=1)
    mel, face = mel.double(), face.double()
    dout = torch.randn((2, 3, 96, 96), dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        out, grads = R.generator_forward_backward(sd, mel, face, dout)
    leaves = {k: v.clone().requires_grad_(k.endswith(".weight") or k.endswith(".bias")) for k, v in sd.items()}
    ref_out = O.generator_forward(leaves, mel, face, training=True)
    assert torch.allclose(out, ref_out.detach(), rtol=1e-10, atol=1e-12)
    names = [k for k, v in leaves.items() if v.requires_grad]
    ref = dict(zip(names, torch.autograd.grad(ref_out, [leaves[k] for k in names], dout)))
    assert set(grads) == set(names)
    worst = 0.0
    for k in names:
        scale = ref[k].abs().max().item() + 1e-30
        err = (grads[k] - ref[k]).abs().max().item() / scale
        if k.endswith("conv_block.0.bias"):
            # conv bias before BatchNorm: the true gradient is 0; both sides hold rounding noise
            assert grads[k].abs().max().item() <= 1e-9 and ref[k].abs().max().item() <= 1e-9, k
            continue
        worst = max(worst, err)
        assert err <= 1e-8, (k, err)
    assert worst > 0.0
