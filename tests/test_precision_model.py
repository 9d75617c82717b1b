"""Why the stress-weight tolerance is 8e-3 and not 1e-3: a CPU model of TF32-class arithmetic.

The reference run on any Ampere-or-later GPU executes its convolutions through cuDNN with
`torch.backends.cudnn.allow_tf32 = True` (PyTorch's default): operands rounded to a 10-bit mantissa, fp32
accumulation.  Our kernels use fp16 operands (the same 10-bit mantissa) and additionally store activations in
fp16.  This test emulates both roundings inside the fp32 CPU oracle and shows that (a) plain TF32 operand
rounding alone already moves the stress-weight output by more than 1e-3, (b) the fp16-storage model lands at
the error the GPU tests measure (~4e-3), and (c) on the reference's own initialisation statistics everything is
two orders of magnitude below the 1e-3 bar."""
import torch

from oracle import w2l_oracle as O


def round_mantissa10(t: torch.Tensor) -> torch.Tensor:
    """Round fp32 to a 10-bit mantissa (TF32 / fp16 significand), round-to-nearest-even, range untouched."""
    i = t.contiguous().view(torch.int32)
    lsb = (i >> 13) & 1
    r = (i + 0x0FFF + lsb) & ~0x1FFF
    return r.view(torch.float32)


def forward_with_rounding(sd, mel, face, round_ops: bool, fp16_store: bool):
    orig = O.block_forward

    def patched(x, sd_, prefix, row, training=False):
        kind = row[0]
        w = sd_[f"{prefix}.conv_block.0.weight"]
        sd2 = dict(sd_)
        xin = x
        if round_ops:
            sd2[f"{prefix}.conv_block.0.weight"] = round_mantissa10(w)
            xin = round_mantissa10(x)
        y = orig(xin, sd2, prefix, row)
        if fp16_store:
            y = y.half().float()
        return y

    O.block_forward = patched
    try:
        with torch.no_grad():
            return O.generator_forward(sd, mel, face)
    finally:
        O.block_forward = orig


def test_tf32_class_arithmetic_explains_the_stress_error():
    mel, face = O.make_generator_inputs(1, 0)
    sd = O.make_state_dict("generator", 0)  # stress weights
    with torch.no_grad():
        ref = O.generator_forward(sd, mel, face)
    tf32 = forward_with_rounding(sd, mel, face, round_ops=True, fp16_store=False)
    ours = forward_with_rounding(sd, mel, face, round_ops=True, fp16_store=True)
    e_tf32 = (tf32 - ref).abs().max().item()
    e_ours = (ours - ref).abs().max().item()
    assert e_tf32 > 1e-3, e_tf32          # the reference's own GPU arithmetic misses 1e-3 on these weights
    assert 1e-3 < e_ours < 8e-3, e_ours   # the model of our kernels: what tests/test_gpu_nets.py measures
    assert e_ours < 4 * e_tf32            # and it is the same class of error, not a different regime


def test_default_init_is_far_below_the_bar_in_the_same_model():
    mel, face = O.make_generator_inputs(1, 0)
    sd = O.make_state_dict("generator", 0, init="default")
    with torch.no_grad():
        ref = O.generator_forward(sd, mel, face)
    ours = forward_with_rounding(sd, mel, face, round_ops=True, fp16_store=True)
    assert (ours - ref).abs().max().item() < 2e-4


def _tf32(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _bf16(t):
    return t.to(torch.bfloat16).to(t.dtype)


def test_single_block_backward_is_well_conditioned_in_bf16():
    """One block's backward with bf16 GEMM operands and bf16-stored activations (fp32 accumulate) stays within a few
    per cent of the exact result (measured: dx 3 %, dw 1 % — ReLU-mask flips of near-zero pre-activations dominate): the
    per-block parity bar the training kernels can be held to."""
    from oracle import backward_recipe as R
    g = torch.Generator().manual_seed(0)
    for row in (O._c(64, 64, 3, 1, 1, True), O._c(64, 128, 3, 2, 1), O._t(160, 64, 3, 2, 1, 1)):
        kind, cin, cout = row[0], row[1], row[2]
        x = torch.randn((4, cin, 24, 24), generator=g)
        wshape = (cin, cout, 3, 3) if kind == "t" else (cout, cin, 3, 3)
        w = 0.05 * torch.randn(wshape, generator=g)
        b = torch.zeros(cout); gamma = torch.ones(cout); beta = torch.zeros(cout)
        with torch.no_grad():
            y, saved = R.block_forward_train(x.double(), w.double(), b.double(), gamma.double(), beta.double(), row)
            dy = torch.randn(y.shape, generator=g)
            ref = R.block_backward(dy.double(), x.double(), w.double(), gamma.double(), row, saved)
            R.set_precision_model(_bf16, _bf16)
            try:
                y16, saved16 = R.block_forward_train(x, w, b, gamma, beta, row)
                got = R.block_backward(dy, x, w, gamma, row, saved16)
            finally:
                R.set_precision_model(None, None)
        for k in ("dx", "dw", "dgamma", "dbeta"):
            rel = ((got[k].double() - ref[k]).norm() / ref[k].norm()).item()
            assert rel <= 6e-2, (row, k, rel)


def test_end_to_end_gradients_are_ill_conditioned_at_initialisation():
    """Why end-to-end gradient parity against fp32 is NOT the bar for the training kernels: on the
    reference's initialisation the generator's gradients amplify relative perturbations by ~1e5 — fp32 and fp64 already
    disagree by ~0.5 %, and TF32-class operands (cuDNN's default for the reference on any Ampere+ GPU) by ~25 %."""
    from oracle import backward_recipe as R
    sd = O.make_state_dict("generator", 0, init="default")
    N = 8
    mel, face = O.make_generator_inputs(N, seed=1)
    gt = torch.rand((N, 3, 96, 96), generator=torch.Generator().manual_seed(4))

    def med(ga, gb):
        rel = sorted(((ga[k].double() - v.double()).norm() / (v.double().norm() + 1e-30)).item()
                     for k, v in gb.items() if not k.endswith("conv_block.0.bias"))
        return rel[len(rel) // 2]

    with torch.no_grad():
        out, _ = R.generator_forward_backward(sd, mel, face, torch.zeros((N, 3, 96, 96)))
        dout = torch.sign(out - gt) / out.numel()                       # dL1/dout, wav2lip_train.py:227
        _, g32 = R.generator_forward_backward(sd, mel, face, dout)
        sd64 = {k: v.double() for k, v in sd.items() if v.dtype.is_floating_point}
        _, g64 = R.generator_forward_backward(sd64, mel.double(), face.double(), dout.double())
        R.set_precision_model(_tf32, None)
        try:
            _, gtf = R.generator_forward_backward(sd, mel, face, dout)
        finally:
            R.set_precision_model(None, None)
    m32, mtf = med(g32, g64), med(gtf, g64)
    assert 1e-4 <= m32 <= 5e-2, m32          # fp32 eps is 6e-8: five orders of magnitude of amplification
    assert 5e-2 <= mtf <= 1.0, mtf           # measured 0.24
    # the head's gradient (no amplification yet) is accurate in both
    hk = "output_block.1.weight"
    assert ((gtf[hk].double() - g64[hk]).norm() / g64[hk].norm()).item() <= 5e-3
