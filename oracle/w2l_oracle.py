"""CPU oracle for the conv hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.  The product (wav2lip_b200/) never does.

It restates, as plain functional fp32 PyTorch on the CPU, the algorithm of the
reference's three networks:

  * blocks           /root/reference/models/conv.py:5-19   (Conv2d: conv -> BN -> [+x] -> ReLU)
                     /root/reference/models/conv.py:21-31  (nonorm_Conv2d: conv -> LeakyReLU(0.01))
                     /root/reference/models/conv.py:33-44  (Conv2dTranspose: convT -> BN -> ReLU)
  * generator        /root/reference/models/wav2lip.py:9-125  (Wav2Lip ctor + forward)
  * quality disc     /root/reference/models/wav2lip.py:128-184 (Wav2Lip_disc_qual)
  * sync expert      /root/reference/models/syncnet.py:8-66   (SyncNet_color)

The architecture is restated as data (the *_SPEC tables below) and evaluated by one
interpreter, so it is an independent statement from both the reference's nn.Module
trees and the product's C++ layer tables.  Pinning: tests/golden/make_golden.py runs
the REAL reference modules (imported from /root/reference in the build container)
on seeded weights/inputs and commits their outputs under tests/golden/; the CPU test
tests/test_oracle_golden.py checks this restatement against those vectors, and
tests/test_oracle_golden.py checks it against the live reference whenever
/root/reference is present.

State-dict keys are the reference's own (352 tensors for the generator), so a
dict produced by make_state_dict() loads with strict=True into the reference modules.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Architecture tables.  Row = (kind, cin, cout, k, stride, pad, out_pad, residual)
# kind: "c" = Conv2d block (BN+ReLU), "t" = Conv2dTranspose block, "n" = nonorm_Conv2d.
# Strides/kernels may be ints or (h, w) pairs.
# --------------------------------------------------------------------------------------
Row = Tuple[str, int, int, object, object, object, int, bool]


def _c(cin, cout, k, s, p, res=False) -> Row:
    return ("c", cin, cout, k, s, p, 0, res)


def _t(cin, cout, k, s, p, op=0) -> Row:
    return ("t", cin, cout, k, s, p, op, False)


def _n(cin, cout, k, s, p) -> Row:
    return ("n", cin, cout, k, s, p, 0, False)


# wav2lip.py:12-36
GEN_FACE_ENCODER: List[List[Row]] = [
    [_c(6, 16, 7, 1, 3)],
    [_c(16, 32, 3, 2, 1), _c(32, 32, 3, 1, 1, True), _c(32, 32, 3, 1, 1, True)],
    [_c(32, 64, 3, 2, 1)] + [_c(64, 64, 3, 1, 1, True)] * 3,
    [_c(64, 128, 3, 2, 1)] + [_c(128, 128, 3, 1, 1, True)] * 2,
    [_c(128, 256, 3, 2, 1)] + [_c(256, 256, 3, 1, 1, True)] * 2,
    [_c(256, 512, 3, 2, 1), _c(512, 512, 3, 1, 1, True)],
    [_c(512, 512, 3, 1, 0), _c(512, 512, 1, 1, 0)],
]
# wav2lip.py:38-55
GEN_AUDIO_ENCODER: List[Row] = [
    _c(1, 32, 3, 1, 1), _c(32, 32, 3, 1, 1, True), _c(32, 32, 3, 1, 1, True),
    _c(32, 64, 3, (3, 1), 1), _c(64, 64, 3, 1, 1, True), _c(64, 64, 3, 1, 1, True),
    _c(64, 128, 3, 3, 1), _c(128, 128, 3, 1, 1, True), _c(128, 128, 3, 1, 1, True),
    _c(128, 256, 3, (3, 2), 1), _c(256, 256, 3, 1, 1, True),
    _c(256, 512, 3, 1, 0), _c(512, 512, 1, 1, 0),
]
# wav2lip.py:57-81
GEN_FACE_DECODER: List[List[Row]] = [
    [_c(512, 512, 1, 1, 0)],
    [_t(1024, 512, 3, 1, 0), _c(512, 512, 3, 1, 1, True)],
    [_t(1024, 512, 3, 2, 1, 1), _c(512, 512, 3, 1, 1, True), _c(512, 512, 3, 1, 1, True)],
    [_t(768, 384, 3, 2, 1, 1), _c(384, 384, 3, 1, 1, True), _c(384, 384, 3, 1, 1, True)],
    [_t(512, 256, 3, 2, 1, 1), _c(256, 256, 3, 1, 1, True), _c(256, 256, 3, 1, 1, True)],
    [_t(320, 128, 3, 2, 1, 1), _c(128, 128, 3, 1, 1, True), _c(128, 128, 3, 1, 1, True)],
    [_t(160, 64, 3, 2, 1, 1), _c(64, 64, 3, 1, 1, True), _c(64, 64, 3, 1, 1, True)],
]
# wav2lip.py:83-85 : Conv2d(80,32,3,1,1) ; plain nn.Conv2d(32,3,1) ; Sigmoid
GEN_OUTPUT_BLOCK0: Row = _c(80, 32, 3, 1, 1)

# syncnet.py:11-33
SYNC_FACE_ENCODER: List[Row] = [
    _c(15, 32, 7, 1, 3),
    _c(32, 64, 5, (1, 2), 1), _c(64, 64, 3, 1, 1, True), _c(64, 64, 3, 1, 1, True),
    _c(64, 128, 3, 2, 1), _c(128, 128, 3, 1, 1, True), _c(128, 128, 3, 1, 1, True), _c(128, 128, 3, 1, 1, True),
    _c(128, 256, 3, 2, 1), _c(256, 256, 3, 1, 1, True), _c(256, 256, 3, 1, 1, True),
    _c(256, 512, 3, 2, 1), _c(512, 512, 3, 1, 1, True), _c(512, 512, 3, 1, 1, True),
    _c(512, 512, 3, 2, 1), _c(512, 512, 3, 1, 0), _c(512, 512, 1, 1, 0),
]
# syncnet.py:35-53 (one more 256-ch residual block than the generator's audio encoder)
SYNC_AUDIO_ENCODER: List[Row] = [
    _c(1, 32, 3, 1, 1), _c(32, 32, 3, 1, 1, True), _c(32, 32, 3, 1, 1, True),
    _c(32, 64, 3, (3, 1), 1), _c(64, 64, 3, 1, 1, True), _c(64, 64, 3, 1, 1, True),
    _c(64, 128, 3, 3, 1), _c(128, 128, 3, 1, 1, True), _c(128, 128, 3, 1, 1, True),
    _c(128, 256, 3, (3, 2), 1), _c(256, 256, 3, 1, 1, True), _c(256, 256, 3, 1, 1, True),
    _c(256, 512, 3, 1, 0), _c(512, 512, 1, 1, 0),
]
# wav2lip.py:131-150
DISC_FACE_ENCODER: List[List[Row]] = [
    [_n(3, 32, 7, 1, 3)],
    [_n(32, 64, 5, (1, 2), 2), _n(64, 64, 5, 1, 2)],
    [_n(64, 128, 5, 2, 2), _n(128, 128, 5, 1, 2)],
    [_n(128, 256, 5, 2, 2), _n(256, 256, 5, 1, 2)],
    [_n(256, 512, 3, 2, 1), _n(512, 512, 3, 1, 1)],
    [_n(512, 512, 3, 2, 1), _n(512, 512, 3, 1, 1)],
    [_n(512, 512, 3, 1, 0), _n(512, 512, 1, 1, 0)],
]

BN_EPS = 1e-5  # nn.BatchNorm2d default, conv.py:10,38


def _pair(v) -> Tuple[int, int]:
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


# --------------------------------------------------------------------------------------
# Enumeration of (prefix, row) for every block of a network, with the reference's names.
# --------------------------------------------------------------------------------------
def generator_layers() -> List[Tuple[str, Row]]:
    out: List[Tuple[str, Row]] = []
    for i, blk in enumerate(GEN_FACE_ENCODER):
        for j, row in enumerate(blk):
            out.append((f"face_encoder_blocks.{i}.{j}", row))
    for i, row in enumerate(GEN_AUDIO_ENCODER):
        out.append((f"audio_encoder.{i}", row))
    for i, blk in enumerate(GEN_FACE_DECODER):
        for j, row in enumerate(blk):
            out.append((f"face_decoder_blocks.{i}.{j}", row))
    out.append(("output_block.0", GEN_OUTPUT_BLOCK0))
    return out


def syncnet_layers() -> List[Tuple[str, Row]]:
    out = [(f"face_encoder.{i}", r) for i, r in enumerate(SYNC_FACE_ENCODER)]
    out += [(f"audio_encoder.{i}", r) for i, r in enumerate(SYNC_AUDIO_ENCODER)]
    return out


def disc_layers() -> List[Tuple[str, Row]]:
    out: List[Tuple[str, Row]] = []
    for i, blk in enumerate(DISC_FACE_ENCODER):
        for j, row in enumerate(blk):
            out.append((f"face_encoder_blocks.{i}.{j}", row))
    return out


# --------------------------------------------------------------------------------------
# Seeded weights with the reference's key names
# --------------------------------------------------------------------------------------
def _block_tensors(prefix: str, row: Row, g: torch.Generator, gain: float, torch_default: bool = False) -> Dict[str, torch.Tensor]:
    kind, cin, cout, k, _s, _p, _op, _res = row
    kh, kw = _pair(k)
    sd: Dict[str, torch.Tensor] = {}
    fan_in = cin * kh * kw
    if kind == "t":  # nn.ConvTranspose2d weight is (Cin, Cout, kh, kw), conv.py:37
        wshape = (cin, cout, kh, kw)
        if torch_default:
            fan_in = cout * kh * kw  # torch computes fan_in from dim 1 of the stored weight
        else:
            fan_in = cin * kh * kw / 4.0 if kh == 3 else fan_in  # stride-2 convT touches ~K/4 taps per output
    else:
        wshape = (cout, cin, kh, kw)
    if torch_default:  # nn.Conv2d.reset_parameters: kaiming_uniform(a=sqrt(5)) -> U(+-1/sqrt(fan_in)) for weight and bias
        bound = 1.0 / math.sqrt(fan_in)
        bias_bound = bound
    else:
        bound = gain * math.sqrt(3.0 / fan_in)
        bias_bound = 0.1
    sd[f"{prefix}.conv_block.0.weight"] = (torch.rand(wshape, generator=g) * 2 - 1) * bound
    sd[f"{prefix}.conv_block.0.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bias_bound
    if kind != "n":
        sd[f"{prefix}.conv_block.1.weight"] = torch.rand(cout, generator=g) + 0.5          # gamma ~ U(0.5,1.5)
        sd[f"{prefix}.conv_block.1.bias"] = torch.randn(cout, generator=g) * 0.1           # beta  ~ N(0,0.1)
        sd[f"{prefix}.conv_block.1.running_mean"] = torch.randn(cout, generator=g) * 0.1
        sd[f"{prefix}.conv_block.1.running_var"] = torch.rand(cout, generator=g) + 0.5     # U(0.5,1.5)
        sd[f"{prefix}.conv_block.1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd


def make_state_dict(net: str, seed: int = 0, gain: float = 1.0, init: str = "stress") -> Dict[str, torch.Tensor]:
    """Deterministic fp32 weights (CPU generator) for net in {"generator","syncnet","disc"}.

    init="stress" (default): variance-preserving conv weights (`gain` around He-uniform) so that
        activations stay O(1..30) through ~50 layers and the pre-sigmoid logits have std ~2.5 — every
        layer's rounding error reaches the output.  Much harsher than anything the reference's own
        initialisation produces.
    init="default": the statistics of the reference's own constructor (torch's Conv2d default:
        U(+-1/sqrt(fan_in)) weights and biases) — the "random weights" of BASELINE.json's configs.
    In both cases BatchNorm affine and running statistics are randomised so that BN folding is
    actually exercised (a fresh BN is the identity).
    """
    td = init == "default"
    if init not in ("stress", "default"):
        raise ValueError(init)
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    if net == "generator":
        for prefix, row in generator_layers():
            sd.update(_block_tensors(prefix, row, g, gain, td))
        sd["output_block.1.weight"] = (torch.rand((3, 32, 1, 1), generator=g) * 2 - 1) * math.sqrt(3.0 / 32)
        sd["output_block.1.bias"] = (torch.rand(3, generator=g) * 2 - 1) * 0.1
    elif net == "syncnet":
        for prefix, row in syncnet_layers():
            sd.update(_block_tensors(prefix, row, g, gain, td))
    elif net == "disc":
        for prefix, row in disc_layers():
            sd.update(_block_tensors(prefix, row, g, gain, td))
        sd["binary_pred.0.weight"] = (torch.rand((1, 512, 1, 1), generator=g) * 2 - 1) * math.sqrt(3.0 / 512)
        sd["binary_pred.0.bias"] = (torch.rand(1, generator=g) * 2 - 1) * 0.1
    else:
        raise ValueError(net)
    return sd


# --------------------------------------------------------------------------------------
# Seeded inputs, as the reference's callers build them
# --------------------------------------------------------------------------------------
def make_generator_inputs(n: int, seed: int = 0, t: Optional[int] = None):
    """4-D (t=None): mel (n,1,80,16), face (n,6,96,96) as inference.py:136-140,259-260.
    5-D: mel (n,t,1,80,16), face (n,6,t,96,96) as wav2lip_train.py:153-163."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    if t is None:
        mel = torch.rand((n, 1, 80, 16), generator=g) * 8 - 4
        face = torch.rand((n, 6, 96, 96), generator=g)
        face[:, 0:3, 48:, :] = 0.0  # masked lower half of the target frame
    else:
        mel = torch.rand((n, t, 1, 80, 16), generator=g) * 8 - 4
        face = torch.rand((n, 6, t, 96, 96), generator=g)
        face[:, 0:3, :, 48:, :] = 0.0
    return mel, face


def make_syncnet_inputs(b: int, seed: int = 0):
    """mel (b,1,80,16), face (b,15,48,96) — color_syncnet_train.py:118-129."""
    g = torch.Generator(device="cpu")
    g.manual_seed(2000 + seed)
    mel = torch.rand((b, 1, 80, 16), generator=g) * 8 - 4
    face = torch.rand((b, 15, 48, 96), generator=g)
    return mel, face


def make_disc_inputs(b: int, t: int = 5, seed: int = 0):
    """frames (b,3,t,96,96) — hq_wav2lip_train.py:248,252."""
    g = torch.Generator(device="cpu")
    g.manual_seed(3000 + seed)
    return torch.rand((b, 3, t, 96, 96), generator=g)


# --------------------------------------------------------------------------------------
# The interpreter
# --------------------------------------------------------------------------------------
def block_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, row: Row, training: bool = False) -> torch.Tensor:
    """One conv.py block.  training=True: BatchNorm on batch statistics, running stats in `sd` updated in place
    (momentum 0.1, unbiased variance) exactly as nn.BatchNorm2d does in train mode."""
    kind, _cin, _cout, _k, s, p, op, res = row
    w = sd[f"{prefix}.conv_block.0.weight"]
    b = sd[f"{prefix}.conv_block.0.bias"]
    if kind == "t":  # conv.py:36-44
        y = F.conv_transpose2d(x, w, b, stride=_pair(s), padding=_pair(p), output_padding=_pair(op))
    else:
        y = F.conv2d(x, w, b, stride=_pair(s), padding=_pair(p))
    if kind == "n":  # conv.py:24-31
        return F.leaky_relu(y, 0.01)
    y = F.batch_norm(
        y,
        sd[f"{prefix}.conv_block.1.running_mean"],
        sd[f"{prefix}.conv_block.1.running_var"],
        sd[f"{prefix}.conv_block.1.weight"],
        sd[f"{prefix}.conv_block.1.bias"],
        training=training, momentum=0.1, eps=BN_EPS,
    )
    if res:  # conv.py:16-18: added after BN, before ReLU
        y = y + x
    return F.relu(y)


def generator_forward(sd, audio, face, taps: Optional[Dict[str, torch.Tensor]] = None,
                      return_logits: bool = False, training: bool = False):
    """Wav2Lip.forward — wav2lip.py:87-125.  `taps`, if given, receives every block output."""
    five_d = face.dim() > 4
    B = audio.size(0)
    if five_d:  # t-major flatten, wav2lip.py:93-94
        audio = torch.cat([audio[:, i] for i in range(audio.size(1))], dim=0)
        face = torch.cat([face[:, :, i] for i in range(face.size(2))], dim=0)

    a = audio
    for i, row in enumerate(GEN_AUDIO_ENCODER):
        a = block_forward(a, sd, f"audio_encoder.{i}", row, training)
        if taps is not None:
            taps[f"audio_encoder.{i}"] = a

    feats = []
    x = face
    for i, blk in enumerate(GEN_FACE_ENCODER):
        for j, row in enumerate(blk):
            x = block_forward(x, sd, f"face_encoder_blocks.{i}.{j}", row, training)
            if taps is not None:
                taps[f"face_encoder_blocks.{i}.{j}"] = x
        feats.append(x)

    x = a
    for i, blk in enumerate(GEN_FACE_DECODER):
        for j, row in enumerate(blk):
            x = block_forward(x, sd, f"face_decoder_blocks.{i}.{j}", row, training)
            if taps is not None:
                taps[f"face_decoder_blocks.{i}.{j}"] = x
        x = torch.cat((x, feats.pop()), dim=1)  # decoder channels first, wav2lip.py:108

    x = block_forward(x, sd, "output_block.0", GEN_OUTPUT_BLOCK0, training)
    if taps is not None:
        taps["output_block.0"] = x
    logits = F.conv2d(x, sd["output_block.1.weight"], sd["output_block.1.bias"])
    out = logits if return_logits else torch.sigmoid(logits)
    if five_d:  # wav2lip.py:118-120
        out = torch.stack(torch.split(out, B, dim=0), dim=2)
    return out


def syncnet_forward(sd, audio, face, taps: Optional[Dict[str, torch.Tensor]] = None, training: bool = False):
    """SyncNet_color.forward — syncnet.py:55-66.  Returns (audio_emb, face_emb)."""
    v = face
    for i, row in enumerate(SYNC_FACE_ENCODER):
        v = block_forward(v, sd, f"face_encoder.{i}", row, training)
        if taps is not None:
            taps[f"face_encoder.{i}"] = v
    a = audio
    for i, row in enumerate(SYNC_AUDIO_ENCODER):
        a = block_forward(a, sd, f"audio_encoder.{i}", row, training)
        if taps is not None:
            taps[f"audio_encoder.{i}"] = a
    a = F.normalize(a.reshape(a.size(0), -1), p=2, dim=1)
    v = F.normalize(v.reshape(v.size(0), -1), p=2, dim=1)
    return a, v


def disc_forward(sd, frames, taps: Optional[Dict[str, torch.Tensor]] = None, return_logits: bool = False):
    """Wav2Lip_disc_qual.forward — wav2lip.py:176-184 (to_2d :158-161, lower half :155-156)."""
    x = torch.cat([frames[:, :, i] for i in range(frames.size(2))], dim=0)
    x = x[:, :, x.size(2) // 2:]
    for i, blk in enumerate(DISC_FACE_ENCODER):
        for j, row in enumerate(blk):
            x = block_forward(x, sd, f"face_encoder_blocks.{i}.{j}", row)
            if taps is not None:
                taps[f"face_encoder_blocks.{i}.{j}"] = x
    logits = F.conv2d(x, sd["binary_pred.0.weight"], sd["binary_pred.0.bias"])
    out = logits if return_logits else torch.sigmoid(logits)
    return out.reshape(out.size(0), -1)


# Algorithmic work per unit (multiply-accumulates), used by bench.py's roofline.
def macs_per_unit(net: str) -> int:
    def conv_macs(row: Row, hin: int, win: int) -> Tuple[int, int, int]:
        kind, cin, cout, k, s, p, op, _ = row
        kh, kw = _pair(k); sh, sw = _pair(s); ph, pw = _pair(p)
        if kind == "t":
            ho = (hin - 1) * sh - 2 * ph + kh + op
            wo = (win - 1) * sw - 2 * pw + kw + op
            return hin * win * cin * cout * kh * kw, ho, wo
        ho = (hin + 2 * ph - kh) // sh + 1
        wo = (win + 2 * pw - kw) // sw + 1
        return ho * wo * cin * cout * kh * kw, ho, wo

    total = 0
    if net == "generator":
        h, w = 80, 16
        for row in GEN_AUDIO_ENCODER:
            m, h, w = conv_macs(row, h, w); total += m
        h, w = 96, 96
        for blk in GEN_FACE_ENCODER:
            for row in blk:
                m, h, w = conv_macs(row, h, w); total += m
        h, w = 1, 1
        for blk in GEN_FACE_DECODER:
            for row in blk:
                m, h, w = conv_macs(row, h, w); total += m
        m, h, w = conv_macs(GEN_OUTPUT_BLOCK0, h, w); total += m
        total += 96 * 96 * 32 * 3
    elif net == "syncnet":
        h, w = 48, 96
        for row in SYNC_FACE_ENCODER:
            m, h, w = conv_macs(row, h, w); total += m
        h, w = 80, 16
        for row in SYNC_AUDIO_ENCODER:
            m, h, w = conv_macs(row, h, w); total += m
    elif net == "disc":
        h, w = 48, 96
        for blk in DISC_FACE_ENCODER:
            for row in blk:
                m, h, w = conv_macs(row, h, w); total += m
        total += 512
    else:
        raise ValueError(net)
    return total
