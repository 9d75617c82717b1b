"""The detector around the network (face_detection/detection/sfd/sfd_detector.py:17-49, detect.py:58-94, bbox.py:44-64 of the
reference), host side: softmax of the class maps, the 0.05 candidate threshold, prior decoding with variances (0.1, 0.2),
greedy NMS at IoU 0.3 and the final 0.5 score cut — vectorised NumPy with the reference's semantics, including its batch
quirk: a location that passes the threshold in ANY image of the batch becomes a candidate in EVERY image (detect.py:77-88)."""
import os

import numpy as np
import torch

from .net_s3fd import s3fd

MEAN_BGR = np.array([104.0, 117.0, 123.0])


def decode_candidates(olist, thresh=0.05):
    """olist: the 12 maps as float32 numpy arrays.  Returns (n_candidates, B, 5) = x1, y1, x2, y2, score — the array
    batch_detect builds (detect.py:66-93)."""
    B = olist[0].shape[0]
    rows = []
    for i in range(6):
        cls, reg = olist[2 * i], olist[2 * i + 1]
        e = np.exp(cls - cls.max(axis=1, keepdims=True))
        prob = e[:, 1] / e.sum(axis=1)                                       # F.softmax(cls, dim=1)[:, 1]
        stride = 2 ** (i + 2)
        _, hh, ww = np.where(prob > thresh)                                  # one candidate per (image, y, x) hit, as the reference's zip
        if hh.size == 0:
            continue
        axc, ayc = stride / 2 + ww * stride, stride / 2 + hh * stride
        loc = reg[:, :, hh, ww]                                               # (B, 4, n)
        cx = axc[None] + loc[:, 0] * 0.1 * (stride * 4)
        cy = ayc[None] + loc[:, 1] * 0.1 * (stride * 4)
        w = stride * 4 * np.exp(loc[:, 2] * 0.2)
        h = stride * 4 * np.exp(loc[:, 3] * 0.2)
        x1, y1 = cx - w / 2, cy - h / 2
        box = np.stack([x1, y1, x1 + w, y1 + h, prob[:, hh, ww]], axis=2)    # (B, n, 5)
        rows.append(np.transpose(box, (1, 0, 2)))
    if not rows:
        return np.zeros((1, B, 5))
    return np.concatenate(rows, axis=0).astype(np.float32)


def nms(dets, thresh):
    """Greedy NMS of bbox.py:44-64 (boxes with +1 pixel extents, descending score, keep while overlap <= thresh)."""
    if len(dets) == 0:
        return []
    x1, y1, x2, y2, sc = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3], dets[:, 4]
    area = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = sc.argsort()[::-1]
    keep = []
    while order.size > 0:
        i, rest = order[0], order[1:]
        keep.append(i)
        iw = np.maximum(0.0, np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]) + 1)
        ih = np.maximum(0.0, np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]) + 1)
        ovr = iw * ih / (area[i] + area[rest] - iw * ih)
        order = rest[ovr <= thresh]
    return keep


class SFDDetector:
    def __init__(self, device="cuda", path_to_detector=None, verbose=False):
        self.device, self.verbose = device, verbose
        self.face_detector = s3fd()
        path = path_to_detector or os.path.join(os.path.dirname(os.path.abspath(__file__)), "s3fd.pth")
        if os.path.isfile(path):
            self.face_detector.load_state_dict(torch.load(path, map_location="cpu"))
        elif path_to_detector is not None:
            raise FileNotFoundError(path_to_detector)
        # (no network access here: without s3fd.pth the detector keeps its random initialisation; the reference downloads it)
        self.face_detector.to(device)
        self.face_detector.eval()

    def detect_from_batch(self, images):
        """images (B,H,W,3) uint8 BGR -> per image a list of [x1, y1, x2, y2, score] rows (sfd_detector.py:40-46)."""
        imgs = np.asarray(images).astype(np.float32) - MEAN_BGR.astype(np.float32)
        x = torch.from_numpy(np.ascontiguousarray(imgs.transpose(0, 3, 1, 2))).to(self.device)
        with torch.no_grad():
            olist = [o.cpu().numpy() for o in self.face_detector(x)]
        cand = decode_candidates(olist)
        out = []
        for i in range(cand.shape[1]):
            d = cand[:, i, :]
            d = d[nms(d, 0.3), :]
            out.append([r for r in d if r[-1] > 0.5])
        return out
