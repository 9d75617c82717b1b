"""Evaluation-loop pieces (wav2lip_train.py:178-198, 262-292) through the C-ABI against torch's own CPU arithmetic
(the reference calls nn.BCELoss / F.cosine_similarity / nn.L1Loss directly, so torch CPU fp32 IS the reference here).
Tolerances: losses are means of fp32 terms -> 2e-6 relative (summation order); the SyncNet embeddings carry the
tensor-core operand rounding documented in DESIGN.md section 4."""
import numpy as np
import pytest
import torch

from oracle import loss_oracle as LO
from oracle import w2l_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,D", [(1, 512), (7, 512), (256, 512), (1000, 64), (5000, 33)])
def test_cosine_bce_loss(B, D):
    from wav2lip_b200 import losses
    g = torch.Generator().manual_seed(B + D)
    a = torch.rand((B, D), generator=g)
    v = torch.rand((B, D), generator=g)
    y = (torch.rand((B, 1), generator=g) > 0.5).float()
    ref = LO.cosine_loss(a, v, y).item()
    got = losses.cosine_loss(a.cuda(), v.cuda(), y.cuda()).item()
    assert abs(got - ref) <= 2e-6 * max(1.0, abs(ref))
    ref1 = LO.cosine_loss(a, v, torch.ones(B, 1)).item()
    got1 = losses.cosine_loss(a.cuda(), v.cuda()).item()
    assert abs(got1 - ref1) <= 2e-6 * max(1.0, abs(ref1))


def test_cosine_bce_extremes():
    """identical one-hot rows (d == 1 exactly -> log(1-d) clamps at -100 for y = 0), orthogonal rows (d == 0 -> log(d)
    clamps), zero rows.  (Nearly parallel rows are ill-conditioned: log(1-d) amplifies the last ulp of d.)"""
    from wav2lip_b200 import losses
    a = torch.zeros((4, 8)); v = torch.zeros((4, 8))
    a[0, 0] = 1; v[0, 0] = 1            # d = 1
    a[1, 0] = 1; v[1, 1] = 1            # d = 0
    a[2, 3] = 2.0; v[2, 3] = 0.5        # d = 1 exactly (powers of two)
    # row 3: all zeros -> d = 0 / eps^2-guard = 0
    for y in (torch.ones(4, 1), torch.zeros(4, 1), torch.tensor([[1.], [0.], [0.], [1.]])):
        ref = LO.cosine_loss(a, v, y).item()
        got = losses.cosine_loss(a.cuda(), v.cuda(), y.cuda()).item()
        assert abs(got - ref) <= 1e-4 * max(1.0, abs(ref)), (y.flatten().tolist(), got, ref)


@pytest.mark.parametrize("shape", [(1,), (3,), (5, 7), (2, 3, 5, 96, 96), (128, 3, 5, 96, 96)])
def test_l1_loss(shape):
    from wav2lip_b200 import losses
    g = torch.Generator().manual_seed(len(shape))
    x = torch.rand(shape, generator=g)
    y = torch.rand(shape, generator=g)
    ref = LO.recon_loss(x.double(), y.double()).item()     # exact mean; fp32 summation error is the tolerance
    got = losses.recon_loss(x.cuda(), y.cuda()).item()
    assert abs(got - ref) <= 2e-6 * abs(ref) + 1e-9
    assert losses.recon_loss(x.cuda(), x.cuda()).item() == 0.0


def test_syncnet_on_frames_equals_syncnet_on_stacked_halves():
    """w2l_syncnet_forward_frames == slice + cat + w2l_syncnet_forward, bit for bit (same kernels, the stack is addressing)."""
    from wav2lip_b200.models import SyncNet_color
    sd = O.make_state_dict("syncnet", 3)
    s = SyncNet_color()
    s.load_state_dict(sd, strict=True)
    s = s.cuda().eval()
    gen = torch.Generator().manual_seed(11)
    for B in (1, 6, 33):
        frames = torch.rand((B, 3, 5, 96, 96), generator=gen)
        mel = torch.rand((B, 1, 80, 16), generator=gen) * 8 - 4
        with torch.no_grad():
            a0, v0 = s(mel.cuda(), LO.stack_lower_halves(frames).contiguous().cuda())
            a1, v1 = s.forward_frames(mel.cuda(), frames.cuda())
        assert torch.equal(a0, a1) and torch.equal(v0, v1)
    with pytest.raises(ValueError):
        s.forward_frames(mel.cuda(), torch.rand((B, 3, 4, 96, 96)).cuda())


def test_get_sync_loss_matches_oracle():
    from wav2lip_b200 import losses
    from wav2lip_b200.models import SyncNet_color
    sd = O.make_state_dict("syncnet", 4, init="default")
    s = SyncNet_color()
    s.load_state_dict(sd, strict=True)
    s = s.cuda().eval()
    gen = torch.Generator().manual_seed(12)
    B = 6
    frames = torch.rand((B, 3, 5, 96, 96), generator=gen)
    mel = torch.rand((B, 1, 80, 16), generator=gen) * 8 - 4
    with torch.no_grad():
        ref = LO.get_sync_loss(sd, mel, frames).item()
        got = losses.get_sync_loss(s, mel.cuda(), frames.cuda()).item()
    # d = cos(a, v) of unit-norm embeddings that each carry ~3e-4 of operand rounding; loss = -mean log d
    assert abs(got - ref) <= 2e-3 * max(1.0, abs(ref)), (got, ref)
