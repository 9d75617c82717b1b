"""Training row, operator level: one conv.py block in TRAIN mode — forward on batch statistics
and the full backward (dx via the dgrad launches, dW via the tcgen05 wgrad kernel + split-K reduction, dgamma / dbeta /
dbias, running-average update) — through the C-ABI entry `w2l_conv_block_train`, for every distinct block geometry of
the three networks, against oracle/backward_recipe.py (float64; itself equal to torch autograd, tests/
test_backward_recipe.py).

Tolerances: operands (activations, weights, incoming gradients) are rounded to bf16 (8-bit mantissa, 2^-9 relative)
and every stored intermediate (z, y, dz) once more; accumulation, statistics and reductions are fp32/fp64.  Per block
that is a few 1e-3 relative in the L2 sense (tests/test_precision_model.py: "a single block's backward in bf16 is within
3 %"); asserted: relative L2 error <= 2e-2 for y / dx / dW, <= 2e-2 for dgamma / dbeta."""
import ctypes as C

import pytest
import torch

from oracle import backward_recipe as R
from oracle import w2l_oracle as O

pytestmark = pytest.mark.gpu

REL = 2e-2


def rel_l2(got, ref):
    return ((got.double() - ref.double()).norm() / (ref.double().norm() + 1e-30)).item()


def _cases():
    rows, seen = [], set()
    allr = [r for _, r in O.generator_layers()] + [r for _, r in O.syncnet_layers()] + [r for _, r in O.disc_layers()]
    for r in allr:
        key = (r[0], O._pair(r[3]), O._pair(r[4]), O._pair(r[5]), r[6], r[7])
        if key not in seen:
            seen.add(key)
            rows.append(r)
    out = []
    for r in rows:                      # every distinct geometry with moderate channel counts
        kind, cin, cout, k, s, p, op, res = r
        cin2 = min(cin, 48)
        cout2 = cin2 if res else min(cout, 32)
        if res:
            cin2 = cout2 = 32
        out.append(((kind, cin2, cout2, k, s, p, op, res), 2, (12, 13)))
    # real channel counts that exercise every wgrad tile shape (N tiles of 16/32/64/128/256, several M tiles, tap groups)
    out += [
        (O._c(6, 16, 7, 1, 3), 2, (24, 24)),            # first block: Cin 6 (padded to 16), 49 taps in two groups
        (O._c(64, 64, 3, 1, 1, True), 2, (24, 24)),      # the 96x96 residual blocks' shape
        (O._c(80, 32, 3, 1, 1), 2, (16, 16)),            # output block: Cin 80 = 64 + 16
        (O._t(160, 64, 3, 2, 1, 1), 2, (8, 8)),          # last transposed conv: two M tiles, the second 32 channels wide
        (O._c(384, 384, 3, 1, 1, True), 2, (6, 6)),      # three N tiles of 128, three M tiles
        (O._c(256, 256, 3, 1, 1, True), 2, (6, 6)),      # N tile 256
        (O._t(1024, 512, 3, 1, 0), 4, (1, 1)),           # 1x1 -> 3x3 transposed conv (GEMM form in the forward)
        (O._c(512, 512, 1, 1, 0), 4, (1, 1)),            # 1x1 conv on a 1x1 map: K = batch only
        (O._c(32, 64, 3, (3, 1), 1), 2, (20, 16)),       # audio encoder's (3,1) stride
        (O._n(32, 64, 5, (1, 2), 2), 2, (12, 24)),       # discriminator, k5 s(1,2), LeakyReLU, real bias gradient
    ]
    return out


def _id(c):
    r, n, hw = c
    return f"{r[0]}-{r[1]}to{r[2]}-k{r[3]}-s{r[4]}-p{r[5]}-op{r[6]}-res{int(r[7])}-n{n}-{hw[0]}x{hw[1]}"


@pytest.fixture(scope="module")
def ctx():
    from wav2lip_b200 import _lib
    return _lib.Context(0, _lib.PREC_BF16)


@pytest.mark.parametrize("case", _cases(), ids=_id)
def test_block_train_forward_backward(case, ctx):
    from wav2lip_b200 import _lib
    row, n, (H, W) = case
    kind, cin, cout, k, s, p, op, res = row
    (kh, kw), (sh, sw), (ph, pw) = O._pair(k), O._pair(s), O._pair(p)
    if kind != "t" and (H + 2 * ph < kh or W + 2 * pw < kw):
        pytest.skip("input smaller than the filter")
    g = torch.Generator().manual_seed((cin * 131 + cout * 17 + kh * 7 + kw * 3 + sh * 5 + sw + H * 11 + W + ord(kind)) % 100000)
    x = torch.randn((n, cin, H, W), generator=g)
    wshape = (cin, cout, kh, kw) if kind == "t" else (cout, cin, kh, kw)
    fan = cin * kh * kw
    w = torch.randn(wshape, generator=g) / fan ** 0.5
    b = 0.1 * torch.randn(cout, generator=g)
    gamma = 1 + 0.2 * torch.randn(cout, generator=g)
    beta = 0.1 * torch.randn(cout, generator=g)
    rmean, rvar = 0.1 * torch.randn(cout, generator=g), 0.5 + torch.rand(cout, generator=g)
    y_ref, saved = R.block_forward_train(x.double(), w.double(), b.double(), gamma.double(), beta.double(), row)
    dy = torch.randn(y_ref.shape, generator=g)
    ref = R.block_backward(dy.double(), x.double(), w.double(), gamma.double(), row, saved)

    li = _lib.LayerInfo()
    li.name = b"block"
    li.kind = {"c": _lib.BLOCK_CONV_BN_RELU, "t": _lib.BLOCK_CONVT_BN_RELU, "n": _lib.BLOCK_CONV_LRELU}[kind]
    li.cin, li.cout, li.kh, li.kw, li.sh, li.sw, li.ph, li.pw = cin, cout, kh, kw, sh, sw, ph, pw
    li.out_pad, li.residual = op, int(res)
    dev = "cuda:0"
    t = lambda a: a.float().contiguous().to(dev)
    xd, wd, bd, gd, bed, rmd, rvd, dyd = map(t, (x, w, b, gamma, beta, rmean, rvar, dy))
    yd = torch.empty(y_ref.shape, device=dev)
    dxd, dwd = torch.empty_like(xd), torch.full_like(wd, float("nan"))
    dbd, dgd, dbed = (torch.full((cout,), float("nan"), device=dev) for _ in range(3))
    P = lambda a: C.c_void_p(a.data_ptr())
    bn = kind != "n"
    _lib.check(ctx.lib.w2l_conv_block_train(ctx.h, C.byref(li), P(xd), n, H, W, P(wd), P(bd), P(gd) if bn else None,
                                            P(bed) if bn else None, P(rmd) if bn else None, P(rvd) if bn else None, P(dyd), P(yd),
                                            P(dxd), P(dwd), P(dbd), P(dgd) if bn else None, P(dbed) if bn else None, None))
    torch.cuda.synchronize()
    assert rel_l2(yd.cpu(), y_ref) <= REL, ("y", rel_l2(yd.cpu(), y_ref))
    assert rel_l2(dxd.cpu(), ref["dx"]) <= REL, ("dx", rel_l2(dxd.cpu(), ref["dx"]))
    assert rel_l2(dwd.cpu(), ref["dw"]) <= REL, ("dw", rel_l2(dwd.cpu(), ref["dw"]))
    if bn:
        assert rel_l2(dgd.cpu(), ref["dgamma"]) <= REL, ("dgamma", rel_l2(dgd.cpu(), ref["dgamma"]))
        assert rel_l2(dbed.cpu(), ref["dbeta"]) <= REL, ("dbeta", rel_l2(dbed.cpu(), ref["dbeta"]))
        assert dbd.abs().max().item() == 0.0          # conv bias under a BatchNorm: exactly zero gradient
        # running averages as nn.BatchNorm2d updates them (momentum 0.1, unbiased variance, conv bias included in the mean)
        if kind == "t":
            z = torch.nn.functional.conv_transpose2d(x.double(), w.double(), b.double(), stride=(sh, sw), padding=(ph, pw), output_padding=op)
        else:
            z = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=(sh, sw), padding=(ph, pw))
        m_ref = 0.9 * rmean.double() + 0.1 * z.mean(dim=(0, 2, 3))
        v_ref = 0.9 * rvar.double() + 0.1 * z.var(dim=(0, 2, 3), unbiased=True)
        assert (rmd.cpu().double() - m_ref).abs().max().item() <= 5e-3
        assert rel_l2(rvd.cpu(), v_ref) <= 1e-2
    else:
        assert rel_l2(dbd.cpu(), ref["db"]) <= REL, ("db", rel_l2(dbd.cpu(), ref["db"]))
