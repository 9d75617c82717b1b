"""Scope row f4 on the GPU: the S3FD network through the C-ABI (w2l_s3fd_forward, via the mirror
wav2lip_b200.face_detection) against the REAL reference's outputs (tests/golden/s3fd.npz) and the oracle; the detector around
it against the oracle's detections."""
import os

import numpy as np
import pytest
import torch

from oracle import s3fd_oracle as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net():
    from wav2lip_b200.face_detection.detection.sfd.net_s3fd import s3fd
    m = s3fd()
    m.load_state_dict(S.make_state_dict(0), strict=True)
    return m.cuda().eval()


def test_s3fd_maps_vs_reference_golden(net, golden_dir):
    gold = np.load(os.path.join(golden_dir, "s3fd.npz"))
    x = S.preprocess(S.make_images(2, 96, 128, seed=1))
    with torch.no_grad():
        outs = net(x.cuda())
    assert len(outs) == 12
    for i, o in enumerate(outs):
        ref = gold[f"o{i}"]
        assert tuple(o.shape) == ref.shape
        err = np.abs(o.cpu().numpy() - ref).max()
        # fp16 operands through up to 19 conv layers + L2Norm + head: a few 1e-3 of the map's range
        assert err <= 8e-3 * max(np.abs(ref).max(), 1e-3), (i, err, np.abs(ref).max())


def test_s3fd_every_backbone_layer_and_odd_sizes(net):
    sd = S.make_state_dict(0)
    imgs = S.make_images(1, 150, 210, seed=2)
    x = S.preprocess(imgs)
    taps = {}
    with torch.no_grad():
        ref = S.forward(sd, x, taps)
        net._ensure(x.cuda()).set_debug(True)
        outs = net(x.cuda())
    from wav2lip_b200 import _lib
    names = [l["name"] for l in _lib.net_layers(_lib.NET_S3FD)][:19]
    for li, n in enumerate(names):
        got = net.debug_layer_output(li).cpu()
        r = taps[n]
        assert tuple(got.shape) == tuple(r.shape), n
        assert (got - r).abs().max().item() <= 6e-3 * r.abs().max().item(), n
    for i, (o, r) in enumerate(zip(outs, ref)):
        assert tuple(o.shape) == tuple(r.shape)
        assert (o.cpu() - r).abs().max().item() <= 8e-3 * max(r.abs().max().item(), 1e-3), i
    net._ensure(x.cuda()).set_debug(False)


def test_detector_end_to_end_vs_oracle():
    """FaceAlignment.get_detections_for_batch (inference.py:85) on the GPU core vs the oracle's detector on the same weights."""
    from wav2lip_b200.face_detection import FaceAlignment, LandmarksType
    fa = FaceAlignment(LandmarksType._2D, flip_input=False, device="cuda")
    sd = S.make_state_dict(0)
    fa.face_detector.face_detector.load_state_dict(sd, strict=True)
    imgs_bgr = S.make_images(2, 96, 128, seed=1)
    got = fa.face_detector.detect_from_batch(imgs_bgr)
    ref = S.detect_from_batch(sd, imgs_bgr)
    assert len(got) == len(ref) == 2
    for g, r in zip(got, ref):
        # random weights put hundreds of boxes near the thresholds: the sets agree up to a few borderline boxes
        assert abs(len(g) - len(r)) <= max(3, 0.05 * len(r)), (len(g), len(r))
        if len(r):
            # (random weights decode to boxes of any size — exp(0.2 loc) — so "near" is relative to the box's own scale)
            gb, rb = np.array(g)[:, :4].astype(np.float64), np.array(r)[:, :4].astype(np.float64)
            d = (np.abs(gb[:, None, :] - rb[None, :, :]).max(axis=2) / (1.0 + np.abs(gb).max(axis=1))[:, None]).min(axis=1)
            assert np.mean(d <= 2e-2) >= 0.9, np.mean(d <= 2e-2)          # every box of ours has a near twin in the oracle's list
    res = fa.get_detections_for_batch(imgs_bgr[..., ::-1])
    assert len(res) == 2 and all(r is None or len(r) == 4 for r in res)
