"""Network-level parity of the CUDA path (through the C-ABI, via the models mirrors) against
(a) the committed outputs of the REAL reference modules (tests/golden/*.npz) and (b) the oracle.

Precision policy and tolerances (see DESIGN.md "Precision"): tensor-core operands and stored
activations are fp16 (10-bit mantissa = TF32's), accumulation, BatchNorm scale/shift, residual add and
the sigmoid head are fp32.  Per block that is ~2^-11 relative; through the ~20 blocks of the longest
path it grows to ~1.5e-3 relative on the last feature map.
  * BASELINE configs[0] ("random weights" = the reference constructor's init statistics, randomised BN):
    north_star's bar, max|out - ref| <= 1e-3, is asserted as is (measured 6.5e-5).
  * the "stress" weights (variance-preserving init, logits std 2.5 — far harsher than the bar's setting)
    are asserted at 8e-3 on the sigmoid output and 3e-3*max|ref| on every intermediate feature map;
    they exist so that a wrong tap, stride, phase or channel slice cannot hide behind sigmoid(~0)."""
import os

import numpy as np
import pytest
import torch

from oracle import w2l_oracle as O

pytestmark = pytest.mark.gpu

BAR = 1e-3          # north_star: outputs within 1e-3 max-abs of the reference
STRESS_OUT = 8e-3   # post-sigmoid, stress weights
FEAT_REL = 3e-3     # intermediate feature maps, relative to max|ref|


@pytest.fixture(scope="module")
def gen_stress():
    from wav2lip_b200.models import Wav2Lip
    g = Wav2Lip()
    g.load_state_dict(O.make_state_dict("generator", 0), strict=True)
    return g.cuda().eval()


def test_generator_default_init_meets_the_bar(golden_dir):
    from wav2lip_b200.models import Wav2Lip
    gold = np.load(os.path.join(golden_dir, "generator.npz"))
    g = Wav2Lip()
    g.load_state_dict(O.make_state_dict("generator", 0, init="default"), strict=True)
    g = g.cuda().eval()
    mel, face = O.make_generator_inputs(2, 0)
    with torch.no_grad():
        y = g(mel.cuda(), face.cuda()).cpu().numpy()
    assert y.shape == (2, 3, 96, 96)
    err = np.abs(y - gold["gen4_default_out"]).max()   # vs the REAL reference's output
    assert err <= BAR, err
    assert err <= 2e-4, f"default-init error regressed: {err}"  # measured 6.5e-5


def test_generator_4d_stress_vs_reference_golden(gen_stress, golden_dir):
    gold = np.load(os.path.join(golden_dir, "generator.npz"))
    mel, face = O.make_generator_inputs(2, 0)
    with torch.no_grad():
        y = gen_stress(mel.cuda(), face.cuda()).cpu().numpy()
    err = np.abs(y - gold["gen4_out"]).max()
    assert err <= STRESS_OUT, err
    assert np.abs(y - gold["gen4_out"]).mean() <= 6e-4


def test_generator_5d_and_odd_batch_vs_reference_golden(gen_stress, golden_dir):
    gold = np.load(os.path.join(golden_dir, "generator.npz"))
    mel5, face5 = O.make_generator_inputs(2, seed=1, t=5)
    with torch.no_grad():
        y5 = gen_stress(mel5.cuda(), face5.cuda()).cpu().numpy()
    assert y5.shape == (2, 3, 5, 96, 96)
    assert np.abs(y5 - gold["gen5_out"]).max() <= STRESS_OUT
    mel3, face3 = O.make_generator_inputs(3, seed=2)
    with torch.no_grad():
        y3 = gen_stress(mel3.cuda(), face3.cuda()).cpu().numpy()
    assert np.abs(y3 - gold["gen4n3_out"]).max() <= STRESS_OUT


def test_generator_every_block_vs_oracle():
    from wav2lip_b200.models import Wav2Lip
    sd = O.make_state_dict("generator", 0)
    mel, face = O.make_generator_inputs(2, 0)
    taps = {}
    with torch.no_grad():
        O.generator_forward(sd, mel, face, taps)
    g = Wav2Lip()
    g.load_state_dict(sd, strict=True)
    g = g.cuda().eval()
    with torch.no_grad():
        g._ensure(face.cuda()).set_debug(True)   # keep every block output addressable
        g(mel.cuda(), face.cuda())
    names = [n for n, _ in O.generator_layers()]
    for i, n in enumerate(names[:-1]):           # output_block.0 is fused with the head, never materialised
        got = g.debug_layer_output(i).cpu()
        ref = taps[n]
        assert tuple(got.shape) == tuple(ref.shape), n
        err = (got - ref).abs().max().item()
        assert err <= FEAT_REL * ref.abs().max().item(), f"{n}: {err}"


def test_generator_5d_is_tmajor_flatten_of_4d(gen_stress):
    """wav2lip.py:93-94,119-120: 5-D call == 4-D call on the t-major flattened batch, bit for bit."""
    mel5, face5 = O.make_generator_inputs(3, seed=4, t=5)
    mel4 = torch.cat([mel5[:, i] for i in range(5)], 0)
    face4 = torch.cat([face5[:, :, i] for i in range(5)], 0)
    with torch.no_grad():
        y5 = gen_stress(mel5.cuda(), face5.cuda())
        y4 = gen_stress(mel4.cuda(), face4.cuda())
    y5f = torch.cat([y5[:, :, i] for i in range(5)], 0)
    assert torch.equal(y5f, y4)


def test_generator_full_size_batch_independence(gen_stress):
    """BASELINE configs[1] size (N=128): every crop's result is independent of its batch neighbours and
    of its position in the batch (eval-mode forward has no cross-sample op) — checked bit-exactly against
    the same crops run as a small batch in a different order — and agrees with the oracle on a sample."""
    mel, face = O.make_generator_inputs(128, seed=9)
    with torch.no_grad():
        y = gen_stress(mel.cuda(), face.cuda())
        idx = torch.tensor([127, 0, 64, 5, 77])
        ys = gen_stress(mel[idx].cuda(), face[idx].cuda())
    assert torch.equal(y[idx.cuda()], ys)
    assert torch.isfinite(y).all() and (y >= 0).all() and (y <= 1).all()
    sd = O.make_state_dict("generator", 0)
    with torch.no_grad():
        ref = O.generator_forward(sd, mel[idx[:2]], face[idx[:2]])
    assert (ys[:2].cpu() - ref).abs().max().item() <= STRESS_OUT


def test_generator_reload_weights_and_module_prefix(gen_stress):
    """load_state_dict after first use re-packs; 'module.'-prefixed names are accepted by the C side."""
    from wav2lip_b200.models import Wav2Lip
    mel, face = O.make_generator_inputs(1, 3)
    g = Wav2Lip().cuda().eval()
    with torch.no_grad():
        y0 = g(mel.cuda(), face.cuda())
        g.load_state_dict(O.make_state_dict("generator", 0), strict=True)
        y1 = g(mel.cuda(), face.cuda())
        ref = gen_stress(mel.cuda(), face.cuda())
    assert not torch.equal(y0, y1)
    assert torch.equal(y1, ref)
    ctx = g._w2l_ctx
    tensors = {"module." + k: (v.data_ptr(), v.numel()) for k, v in g.state_dict(keep_vars=True).items()
               if v.dtype.is_floating_point}
    from wav2lip_b200 import _lib
    ctx.load_weights(_lib.NET_GENERATOR, tensors)
    with torch.no_grad():
        g._w2l_key = g._weights_key()
        assert torch.equal(g(mel.cuda(), face.cuda()), ref)
    bad = dict(tensors)
    bad.pop("module.output_block.1.bias")
    with pytest.raises(_lib.W2LError):
        ctx.load_weights(_lib.NET_GENERATOR, bad)


def test_generator_shape_errors(gen_stress):
    with pytest.raises(ValueError):
        gen_stress(torch.zeros(2, 1, 80, 16).cuda(), torch.zeros(2, 6, 64, 64).cuda())
    with pytest.raises(ValueError):
        gen_stress(torch.zeros(3, 1, 80, 16).cuda(), torch.zeros(2, 6, 96, 96).cuda())


def test_empty_and_single_item_batches(gen_stress):
    """Edge cases: an empty batch returns an empty tensor (as torch does), a 5-D call with T=1 and B=1 works."""
    from wav2lip_b200.models import SyncNet_color, Wav2Lip_disc_qual
    with torch.no_grad():
        y0 = gen_stress(torch.zeros(0, 1, 80, 16).cuda(), torch.zeros(0, 6, 96, 96).cuda())
        assert tuple(y0.shape) == (0, 3, 96, 96)
        y5 = gen_stress(torch.zeros(0, 5, 1, 80, 16).cuda(), torch.zeros(0, 6, 5, 96, 96).cuda())
        assert tuple(y5.shape) == (0, 3, 5, 96, 96)
        mel, face = O.make_generator_inputs(1, seed=8, t=1)
        y11 = gen_stress(mel.cuda(), face.cuda())
        assert tuple(y11.shape) == (1, 3, 1, 96, 96)
        y4 = gen_stress(mel[:, 0].cuda(), face[:, :, 0].cuda())
        assert torch.equal(y11[:, :, 0], y4)
        s = SyncNet_color().cuda().eval()
        a, v = s(torch.zeros(0, 1, 80, 16).cuda(), torch.zeros(0, 15, 48, 96).cuda())
        assert tuple(a.shape) == (0, 512) and tuple(v.shape) == (0, 512)
        d = Wav2Lip_disc_qual().cuda().eval()
        assert tuple(d(torch.zeros(0, 3, 5, 96, 96).cuda()).shape) == (0, 1)
        assert tuple(gen_stress.infer_u8(torch.zeros(0, 1, 80, 16).cuda(), torch.zeros(0, 96, 96, 3, dtype=torch.uint8).cuda()).shape) == (0, 96, 96, 3)


def test_syncnet_vs_reference_golden(golden_dir):
    from wav2lip_b200.models import SyncNet_color
    gold = np.load(os.path.join(golden_dir, "syncnet.npz"))
    s = SyncNet_color()
    s.load_state_dict(O.make_state_dict("syncnet", 0), strict=True)
    s = s.cuda().eval()
    mel, face = O.make_syncnet_inputs(3, 0)
    with torch.no_grad():
        a, v = s(mel.cuda(), face.cuda())
    a, v = a.cpu().numpy(), v.cpu().numpy()
    assert a.shape == (3, 512) and v.shape == (3, 512)
    assert np.abs(a - gold["sync_a"]).max() <= BAR       # unit-norm embeddings: entries <= 0.25
    assert np.abs(v - gold["sync_v"]).max() <= BAR
    np.testing.assert_allclose(np.linalg.norm(a, axis=1), 1.0, atol=1e-5)
    np.testing.assert_allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-5)
    assert (a >= 0).all() and (v >= 0).all()             # post-ReLU embeddings
    cos_gpu = (a * v).sum(1)
    cos_ref = (gold["sync_a"] * gold["sync_v"]).sum(1)
    assert np.abs(cos_gpu - cos_ref).max() <= 2e-3


def test_syncnet_full_size_batch_independence():
    from wav2lip_b200.models import SyncNet_color
    s = SyncNet_color()
    s.load_state_dict(O.make_state_dict("syncnet", 0), strict=True)
    s = s.cuda().eval()
    mel, face = O.make_syncnet_inputs(256, 1)
    with torch.no_grad():
        a, v = s(mel.cuda(), face.cuda())
        idx = torch.tensor([255, 3, 100])
        a2, v2 = s(mel[idx].cuda(), face[idx].cuda())
    assert torch.equal(a[idx.cuda()], a2) and torch.equal(v[idx.cuda()], v2)


def test_disc_vs_reference_golden(golden_dir):
    from wav2lip_b200.models import Wav2Lip_disc_qual
    gold = np.load(os.path.join(golden_dir, "disc.npz"))
    d = Wav2Lip_disc_qual()
    d.load_state_dict(O.make_state_dict("disc", 0), strict=True)
    d = d.cuda().eval()
    frames = O.make_disc_inputs(2, 5, 0)
    with torch.no_grad():
        p = d(frames.cuda())
        loss = d.perceptual_forward(frames.cuda())
    assert tuple(p.shape) == (10, 1)
    assert np.abs(p.cpu().numpy() - gold["disc_out"]).max() <= BAR
    ref_loss = float(-np.log(gold["disc_out"]).mean())   # BCE(pred, 1), wav2lip.py:171-172
    assert abs(loss.item() - ref_loss) <= 2e-3
    # rows are t-major: row t*B + b
    with torch.no_grad():
        p_b1 = d(frames[1:2].cuda())
    assert torch.allclose(p_b1.flatten(), p.flatten()[1::2], atol=0, rtol=0)


def test_fp32_faithful_mode_meets_1e3_on_stress_weights(golden_dir):
    """W2L_PREC_F32X: split fp16 operands (hi + lo, 3 MMAs per product): the stress weights that the fast mode
    can only hold to 8e-3 (TF32-class arithmetic, see tests/test_precision_model.py) meet north_star's 1e-3 bar."""
    from wav2lip_b200 import _lib
    from wav2lip_b200.models import SyncNet_color, Wav2Lip, Wav2Lip_disc_qual
    gold = np.load(os.path.join(golden_dir, "generator.npz"))
    g = Wav2Lip()
    g.precision = _lib.PREC_F32X
    g.load_state_dict(O.make_state_dict("generator", 0), strict=True)
    g = g.cuda().eval()
    mel, face = O.make_generator_inputs(2, 0)
    with torch.no_grad():
        y = g(mel.cuda(), face.cuda()).cpu().numpy()
    err = np.abs(y - gold["gen4_out"]).max()
    assert err <= BAR, err
    assert err <= 3e-4, err          # measured ~1e-4 (30x below the fast mode on these weights)
    mel5, face5 = O.make_generator_inputs(2, seed=1, t=5)
    with torch.no_grad():
        y5 = g(mel5.cuda(), face5.cuda()).cpu().numpy()
    assert np.abs(y5 - gold["gen5_out"]).max() <= 3e-4
    gs = np.load(os.path.join(golden_dir, "syncnet.npz"))
    s = SyncNet_color()
    s.precision = _lib.PREC_F32X
    s.load_state_dict(O.make_state_dict("syncnet", 0), strict=True)
    s = s.cuda().eval()
    mel, face = O.make_syncnet_inputs(3, 0)
    with torch.no_grad():
        a, v = s(mel.cuda(), face.cuda())
    assert np.abs(a.cpu().numpy() - gs["sync_a"]).max() <= 2e-5
    assert np.abs(v.cpu().numpy() - gs["sync_v"]).max() <= 2e-5
    gd = np.load(os.path.join(golden_dir, "disc.npz"))
    d = Wav2Lip_disc_qual()
    d.precision = _lib.PREC_F32X
    d.load_state_dict(O.make_state_dict("disc", 0), strict=True)
    d = d.cuda().eval()
    with torch.no_grad():
        p = d(O.make_disc_inputs(2, 5, 0).cuda())
    assert np.abs(p.cpu().numpy() - gd["disc_out"]).max() <= 2e-5


def test_bf16_precision_mode_runs(golden_dir):
    """The bf16-operand build of the same kernels (for checkpoints outside the fp16 range)."""
    from wav2lip_b200 import _lib
    from wav2lip_b200.models import Wav2Lip
    gold = np.load(os.path.join(golden_dir, "generator.npz"))
    g = Wav2Lip()
    g.precision = _lib.PREC_BF16
    g.load_state_dict(O.make_state_dict("generator", 0, init="default"), strict=True)
    g = g.cuda().eval()
    mel, face = O.make_generator_inputs(2, 0)
    with torch.no_grad():
        y = g(mel.cuda(), face.cuda()).cpu().numpy()
    assert np.abs(y - gold["gen4_default_out"]).max() <= BAR


def test_bf16_mode_syncnet_and_disc_vs_reference_golden(golden_dir):
    """BASELINE configs[3] names bf16 for SyncNet_color + Wav2Lip_disc_qual: the bf16-operand build of both nets against
    the REAL reference's fp32 outputs.  bf16 carries 8 mantissa bits (fp16: 11), so the bar is the dtype's own:
    unit-norm embeddings (entries <= 0.25) within 4e-3, disc probability within 4e-3 (measured ~1e-3 / ~1e-4)."""
    from wav2lip_b200 import _lib
    from wav2lip_b200.models import SyncNet_color, Wav2Lip_disc_qual
    gs = np.load(os.path.join(golden_dir, "syncnet.npz"))
    s = SyncNet_color()
    s.precision = _lib.PREC_BF16
    s.load_state_dict(O.make_state_dict("syncnet", 0), strict=True)
    s = s.cuda().eval()
    mel, face = O.make_syncnet_inputs(3, 0)
    with torch.no_grad():
        a, v = s(mel.cuda(), face.cuda())
    a, v = a.cpu().numpy(), v.cpu().numpy()
    assert np.abs(a - gs["sync_a"]).max() <= 4e-3, np.abs(a - gs["sync_a"]).max()
    assert np.abs(v - gs["sync_v"]).max() <= 4e-3, np.abs(v - gs["sync_v"]).max()
    np.testing.assert_allclose(np.linalg.norm(a, axis=1), 1.0, atol=1e-5)
    cos_gpu, cos_ref = (a * v).sum(1), (gs["sync_a"] * gs["sync_v"]).sum(1)
    assert np.abs(cos_gpu - cos_ref).max() <= 1e-2
    gd = np.load(os.path.join(golden_dir, "disc.npz"))
    d = Wav2Lip_disc_qual()
    d.precision = _lib.PREC_BF16
    d.load_state_dict(O.make_state_dict("disc", 0), strict=True)
    d = d.cuda().eval()
    with torch.no_grad():
        p = d(O.make_disc_inputs(2, 5, 0).cuda())
    assert np.abs(p.cpu().numpy() - gd["disc_out"]).max() <= 4e-3


def test_generator_headline_shape_b128_t5_vs_oracle():
    """The metric's own call — B=128, T=5, one 5-D batch of 640 crops — against the oracle on 8 sampled crops, logits
    included (BASELINE weights: reference default-init statistics + randomised BatchNorm; bar 1e-3 on the output)."""
    from wav2lip_b200.models import Wav2Lip
    sd = O.make_state_dict("generator", 0, init="default")
    g = Wav2Lip()
    g.load_state_dict(sd, strict=True)
    g = g.cuda().eval()
    B, T = 128, 5
    mel, face = O.make_generator_inputs(B, seed=21, t=T)
    with torch.no_grad():
        y = g(mel.cuda(), face.cuda()).cpu()
    assert tuple(y.shape) == (B, 3, T, 96, 96)
    picks = [(0, 0), (127, 4), (64, 2), (5, 1), (99, 3), (31, 0), (77, 4), (126, 2)]
    m4 = torch.stack([mel[b, t] for b, t in picks])
    f4 = torch.stack([face[b, :, t] for b, t in picks])
    with torch.no_grad():
        ref = O.generator_forward(sd, m4, f4)
    got = torch.stack([y[b, :, t] for b, t in picks])
    err = (got - ref).abs().max().item()
    assert err <= BAR, err
    # logits: sigmoid is monotone, so compare in logit space where the output is not saturated
    lg, lr = torch.logit(got.clamp(1e-6, 1 - 1e-6)), torch.logit(ref.clamp(1e-6, 1 - 1e-6))
    assert (lg - lr).abs().max().item() <= 5e-3


def test_fp16_range_guard_raises_instead_of_returning_inf():
    """A checkpoint whose activations leave the fp16 range (|v| > 65504) must not return silently poisoned results:
    the epilogues set a sticky device flag (w2l_f16_overflow), the mirror checks it after the first forward following
    a weight load and raises, naming the bf16 mode — which runs the same weights finitely."""
    from wav2lip_b200 import _lib
    from wav2lip_b200.models import Wav2Lip
    sd = O.make_state_dict("generator", 0)
    k = "face_encoder_blocks.0.0.conv_block.1.weight"   # BatchNorm gamma of the first block: scale its output by 1e6
    sd = {n: (t.clone() * 1.0e6 if n == k else t.clone()) for n, t in sd.items()}
    mel, face = O.make_generator_inputs(2, 0)
    g = Wav2Lip()
    g.load_state_dict(sd, strict=True)
    g = g.cuda().eval()
    with torch.no_grad():
        with pytest.raises(_lib.W2LError, match="fp16 range"):
            g(mel.cuda(), face.cuda())
    gb = Wav2Lip()
    gb.precision = _lib.PREC_BF16
    gb.load_state_dict(sd, strict=True)
    gb = gb.cuda().eval()
    with torch.no_grad():
        y = gb(mel.cuda(), face.cuda())
    assert torch.isfinite(y).all()
    # sane weights do not trip the guard (and clear nothing they should not)
    g2 = Wav2Lip()
    g2.load_state_dict(O.make_state_dict("generator", 0), strict=True)
    g2 = g2.cuda().eval()
    with torch.no_grad():
        g2(mel.cuda(), face.cuda())
    assert not g2._w2l_ctx.f16_overflow(clear=False)


def test_inputs_on_the_wrong_device_raise_cleanly(gen_stress):
    """a CPU mel next to a CUDA face used to be dereferenced on the device (sticky illegal-address error)."""
    from wav2lip_b200 import _lib, losses
    from wav2lip_b200.models import SyncNet_color
    mel, face = O.make_generator_inputs(2, 0)
    with torch.no_grad():
        with pytest.raises(_lib.W2LError):
            gen_stress(mel, face.cuda())
        with pytest.raises(_lib.W2LError):
            gen_stress.infer_u8(mel, torch.zeros(2, 96, 96, 3, dtype=torch.uint8).cuda())
        s = SyncNet_color().cuda().eval()
        with pytest.raises(_lib.W2LError):
            s(torch.zeros(2, 1, 80, 16), torch.zeros(2, 15, 48, 96).cuda())
        with pytest.raises(_lib.W2LError):
            losses.recon_loss(torch.zeros(8, 8).cuda(), torch.zeros(8, 8))
        y = gen_stress(mel.cuda(), face.cuda())   # the process is still healthy
    assert torch.isfinite(y).all()
