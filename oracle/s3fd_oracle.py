"""CPU oracle for scope row f4 (the S3FD face detector) — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Functional fp32 restatement of
  network      /root/reference/face_detection/detection/sfd/net_s3fd.py:71-128 (forward; L2Norm :16-19)
  candidates   /root/reference/face_detection/detection/sfd/detect.py:58-94     (batch_detect: softmax, 0.05 threshold, decode)
  decode       /root/reference/face_detection/detection/sfd/bbox.py:110-129     (batch_decode, variances 0.1 / 0.2)
  nms          /root/reference/face_detection/detection/sfd/bbox.py:44-64
  detector     /root/reference/face_detection/detection/sfd/sfd_detector.py:40-46 (NMS 0.3, score > 0.5)
evaluated from a state_dict with the reference's own keys.  PINNED: tests/golden/s3fd.npz is produced by the REAL reference
modules (tests/golden/make_golden_s3fd.py imports net_s3fd.py, detect.py and bbox.py from /root/reference);
tests/test_s3fd_oracle.py checks this restatement against it (and against the live reference when it is present).
"""
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

BACKBONE = ["conv1_1", "conv1_2", "P", "conv2_1", "conv2_2", "P", "conv3_1", "conv3_2", "conv3_3", "T", "P",
            "conv4_1", "conv4_2", "conv4_3", "T", "P", "conv5_1", "conv5_2", "conv5_3", "T", "P",
            "fc6", "fc7", "T", "conv6_1", "conv6_2", "T", "conv7_1", "conv7_2", "T"]
GEOM = {"fc6": (1, 3), "fc7": (1, 0), "conv6_1": (1, 0), "conv6_2": (2, 1), "conv7_1": (1, 0), "conv7_2": (2, 1)}   # (stride, pad); default (1, 1)
TAPS = ["conv3_3_norm", "conv4_3_norm", "conv5_3_norm", "fc7", "conv6_2", "conv7_2"]
SHAPES = [("conv1_1", 3, 64, 3), ("conv1_2", 64, 64, 3), ("conv2_1", 64, 128, 3), ("conv2_2", 128, 128, 3), ("conv3_1", 128, 256, 3),
          ("conv3_2", 256, 256, 3), ("conv3_3", 256, 256, 3), ("conv4_1", 256, 512, 3), ("conv4_2", 512, 512, 3), ("conv4_3", 512, 512, 3),
          ("conv5_1", 512, 512, 3), ("conv5_2", 512, 512, 3), ("conv5_3", 512, 512, 3), ("fc6", 512, 1024, 3), ("fc7", 1024, 1024, 1),
          ("conv6_1", 1024, 256, 1), ("conv6_2", 256, 512, 3), ("conv7_1", 512, 128, 1), ("conv7_2", 128, 256, 3)]
TAP_C = [256, 512, 512, 1024, 512, 256]


def make_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded weights with the reference's keys, in its state_dict order: He-normal convs (activations keep their scale
    through the 19 ReLU layers), small biases, L2Norm weights at their constructor values 10 / 8 / 5."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, cin, cout, k in SHAPES:
        sd[name + ".weight"] = torch.randn((cout, cin, k, k), generator=g) * (2.0 / (cin * k * k)) ** 0.5
        sd[name + ".bias"] = 0.05 * torch.randn(cout, generator=g)
    for name, c, scale in (("conv3_3_norm", 256, 10.0), ("conv4_3_norm", 512, 8.0), ("conv5_3_norm", 512, 5.0)):
        sd[name + ".weight"] = torch.full((c,), scale) * (0.8 + 0.4 * torch.rand(c, generator=g))
    for i, t in enumerate(TAPS):
        for kind, cout in (("conf", 4 if i == 0 else 2), ("loc", 4)):
            sd[f"{t}_mbox_{kind}.weight"] = torch.randn((cout, TAP_C[i], 3, 3), generator=g) * (1.0 / (TAP_C[i] * 9)) ** 0.5
            sd[f"{t}_mbox_{kind}.bias"] = 0.1 * torch.randn(cout, generator=g)
    return sd


def make_images(B: int, H: int, W: int, seed: int = 0) -> np.ndarray:
    """(B,H,W,3) uint8 BGR frames."""
    return np.random.default_rng(seed).integers(0, 256, (B, H, W, 3), dtype=np.uint8)


def preprocess(images_bgr: np.ndarray) -> torch.Tensor:
    """detect.py:60-65: subtract the BGR mean, NHWC -> NCHW, float."""
    x = images_bgr - np.array([104, 117, 123])
    return torch.from_numpy(x.transpose(0, 3, 1, 2)).float()


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, taps_out: dict = None) -> List[torch.Tensor]:
    """net_s3fd.py:71-128."""
    h, feats = x, []
    last = None
    for item in BACKBONE:
        if item == "P":
            h = F.max_pool2d(h, 2, 2)
        elif item == "T":
            feats.append(h)
        else:
            s, p = GEOM.get(item, (1, 1))
            h = F.relu(F.conv2d(h, sd[item + ".weight"], sd[item + ".bias"], stride=s, padding=p))
            if taps_out is not None:
                taps_out[item] = h
            last = item
    out = []
    for i, t in enumerate(TAPS):
        f = feats[i]
        if i < 3:   # L2Norm, net_s3fd.py:16-19
            norm = f.pow(2).sum(dim=1, keepdim=True).sqrt() + 1e-10
            f = f / norm * sd[t + ".weight"].view(1, -1, 1, 1)
        cls = F.conv2d(f, sd[t + "_mbox_conf.weight"], sd[t + "_mbox_conf.bias"], padding=1)
        reg = F.conv2d(f, sd[t + "_mbox_loc.weight"], sd[t + "_mbox_loc.bias"], padding=1)
        if i == 0:  # max-out background label, net_s3fd.py:123-126
            c = torch.chunk(cls, 4, 1)
            cls = torch.cat([torch.max(torch.max(c[0], c[1]), c[2]), c[3]], dim=1)
        out += [cls, reg]
    return out


def batch_candidates(olist: List[torch.Tensor]) -> np.ndarray:
    """detect.py:66-93, loop for loop (small inputs only)."""
    olist = [o.clone() for o in olist]
    BB = olist[0].shape[0]
    for i in range(6):
        olist[2 * i] = F.softmax(olist[2 * i], dim=1)
    rows = []
    for i in range(6):
        ocls, oreg = olist[2 * i], olist[2 * i + 1]
        stride = 2 ** (i + 2)
        for _, hindex, windex in zip(*np.where(ocls[:, 1].numpy() > 0.05)):
            axc, ayc = stride / 2 + windex * stride, stride / 2 + hindex * stride
            score = ocls[:, 1, hindex, windex]
            loc = oreg[:, :, hindex, windex].contiguous().view(BB, 1, 4)
            pri = torch.tensor([[axc, ayc, stride * 4.0, stride * 4.0]]).view(1, 1, 4)
            box = torch.cat((pri[:, :, :2] + loc[:, :, :2] * 0.1 * pri[:, :, 2:], pri[:, :, 2:] * torch.exp(loc[:, :, 2:] * 0.2)), 2)
            box[:, :, :2] -= box[:, :, 2:] / 2
            box[:, :, 2:] += box[:, :, :2]
            rows.append(torch.cat([box[:, 0], score.unsqueeze(1)], 1).numpy())
    return np.array(rows) if rows else np.zeros((1, BB, 5))


def nms(dets: np.ndarray, thresh: float):
    """bbox.py:44-64."""
    if 0 == len(dets):
        return []
    x1, y1, x2, y2, scores = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3], dets[:, 4]
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = scores.argsort()[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(i)
        xx1, yy1 = np.maximum(x1[i], x1[order[1:]]), np.maximum(y1[i], y1[order[1:]])
        xx2, yy2 = np.minimum(x2[i], x2[order[1:]]), np.minimum(y2[i], y2[order[1:]])
        w, h = np.maximum(0.0, xx2 - xx1 + 1), np.maximum(0.0, yy2 - yy1 + 1)
        ovr = w * h / (areas[i] + areas[order[1:]] - w * h)
        order = order[np.where(ovr <= thresh)[0] + 1]
    return keep


def detect_from_batch(sd, images_bgr: np.ndarray):
    """sfd_detector.py:40-46."""
    with torch.no_grad():
        cand = batch_candidates(forward(sd, preprocess(images_bgr)))
    out = []
    for i in range(cand.shape[1]):
        d = cand[:, i, :]
        d = d[nms(d, 0.3), :]
        out.append([x for x in d if x[-1] > 0.5])
    return out
