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
