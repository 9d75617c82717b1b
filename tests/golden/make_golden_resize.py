"""Golden vectors for the resize rows from OpenCV ITSELF (cv2 is importable in the build
container; opencv-python is the reference's own dependency, requirements.txt): `cv2.resize(img, dsize)` with the default
INTER_LINEAR on seeded uint8 images, for the two call shapes of inference.py (:126 crop -> 96x96, :269 96x96 -> box) plus
odd cases (1-pixel sources, exact 2x, extreme aspect).  Run:  python tests/golden/make_golden_resize.py"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [  # (src_h, src_w, dst_h, dst_w)
    (137, 121, 96, 96), (96, 96, 137, 121), (192, 192, 96, 96), (48, 48, 96, 96), (96, 96, 96, 96), (1, 1, 96, 96),
    (96, 96, 1, 1), (5, 300, 96, 96), (96, 96, 301, 7), (250, 211, 96, 96), (96, 96, 211, 250), (33, 97, 96, 96),
]


def image(h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def main():
    out = {"cases": np.array(CASES, dtype=np.int32), "cv2_version": np.array(cv2.__version__)}
    for i, (sh, sw, dh, dw) in enumerate(CASES):
        out[f"out{i}"] = cv2.resize(image(sh, sw, 1000 + i), (dw, dh))
    # the paste: one frame, two boxes
    frames = np.random.default_rng(7).integers(0, 256, (2, 120, 160, 3), dtype=np.uint8)
    boxes = np.array([[0, 10, 100, 20, 95], [1, 0, 120, 33, 160], [1, 57, 58, 3, 4]], dtype=np.int32)
    pred = np.random.default_rng(8).integers(0, 256, (3, 96, 96, 3), dtype=np.uint8)
    crops, pasted = [], []
    for p, (f, y1, y2, x1, x2) in zip(pred, boxes):
        crops.append(cv2.resize(frames[f][y1:y2, x1:x2], (96, 96)))
        fr = frames[f].copy()
        fr[y1:y2, x1:x2] = cv2.resize(p, (x2 - x1, y2 - y1))
        pasted.append(fr)
    out["boxes"] = boxes
    out["crops"] = np.stack(crops)
    out["pasted"] = np.stack(pasted)
    np.savez_compressed(os.path.join(HERE, "resize.npz"), **out)
    print("wrote resize.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
