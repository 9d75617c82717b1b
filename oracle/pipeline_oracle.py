"""CPU oracle for the "next" scope rows around the generator call — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Restates, in NumPy exactly as the reference writes it:
  mel_chunks        /root/reference/inference.py:231-240  (chunk start int(i*80./fps), last chunk right-aligned)
  assemble_batch    /root/reference/inference.py:134-140  (mask rows 48.., concat [masked|full], /255.) and
                    :259-260 (transpose to NCHW, torch.FloatTensor)
  postprocess       /root/reference/inference.py:265, :269 (transpose to NHWC, * 255., astype(np.uint8))
The reference has no tests or fixtures for these lines; they are plain NumPy, restated verbatim.
"""
from __future__ import annotations

import numpy as np

MEL_STEP = 16  # mel_step_size, inference.py:54


def mel_chunks(mel: np.ndarray, fps: float):
    chunks = []
    mult = 80. / fps
    i = 0
    while 1:
        start = int(i * mult)
        if start + MEL_STEP > len(mel[0]):
            chunks.append(mel[:, len(mel[0]) - MEL_STEP:])
            break
        chunks.append(mel[:, start: start + MEL_STEP])
        i += 1
    return chunks


def assemble_batch(faces_u8: np.ndarray, mels) -> tuple:
    """faces_u8 (N,96,96,3) uint8 crops, mels: list of (80,16) -> (mel_batch (N,1,80,16) f32, img_batch (N,6,96,96) f32)."""
    img_batch, mel_batch = np.asarray(faces_u8), np.asarray(mels)
    img_masked = img_batch.copy()
    img_masked[:, 96 // 2:] = 0
    img_batch = np.concatenate((img_masked, img_batch), axis=3) / 255.
    mel_batch = np.reshape(mel_batch, [len(mel_batch), mel_batch.shape[1], mel_batch.shape[2], 1])
    img = np.transpose(img_batch, (0, 3, 1, 2)).astype(np.float32)   # torch.FloatTensor(...)
    mel = np.transpose(mel_batch, (0, 3, 1, 2)).astype(np.float32)
    return mel, img


def postprocess(pred_nchw: np.ndarray) -> np.ndarray:
    """pred (N,3,96,96) float32 in (0,1) -> (N,96,96,3) uint8, inference.py:265,269."""
    p = pred_nchw.transpose(0, 2, 3, 1) * 255.
    return p.astype(np.uint8)


# ----------------------------------------------------------------------------------------------------------------------
# cv2.resize(..., interpolation=INTER_LINEAR) on uint8 images — the two resizes around the generator call:
#   /root/reference/inference.py:126   face = cv2.resize(face, (img_size, img_size))          crop -> 96 x 96
#   /root/reference/inference.py:269   p = cv2.resize(p.astype(np.uint8), (x2 - x1, y2 - y1)) prediction -> box size
#   /root/reference/inference.py:271   f[y1:y2, x1:x2] = p                                    paste
# OpenCV is a third-party dependency of the reference (requirements.txt: opencv-python==4.1.0.25, not vendored).  Its
# 8-bit bilinear resize is fixed-point: coefficients are rounded to 11 bits (INTER_RESIZE_COEF_BITS), the horizontal pass
# keeps 32-bit sums, the vertical pass is ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.  Restated from
# the published algorithm (modules/imgproc/src/resize.cpp: resizeGeneric_ / HResizeLinear / VResizeLinear<uchar,...>).
# PINNED: cv2 4.13 is importable in the build container; tests/test_pipeline_rows.py compares this restatement with
# cv2.resize bit for bit on random shapes, and tests/golden/resize.npz (made by tests/golden/make_golden_resize.py from
# cv2 itself) travels to the GPU box.
# ----------------------------------------------------------------------------------------------------------------------
def _linear_coeffs(dst: int, src: int):
    """Per destination index: source index s (second tap s+1 where it exists), 11-bit weights (a0, a1)."""
    inv_scale = np.float64(dst) / np.float64(src)
    scale = np.float64(1.0) / inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)        # (float)((dx + 0.5) * scale_x - 0.5)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def _round_half_even_short(v: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(v.astype(np.float32)), -32768, 32767).astype(np.int64)   # saturate_cast<short>(float): cvRound


def resize_linear_u8(img: np.ndarray, dsize) -> np.ndarray:
    """cv2.resize(img, dsize) for an (H, W, C) uint8 image, dsize = (width, height), INTER_LINEAR (the default)."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, _ = img.shape
    dw, dh = int(dsize[0]), int(dsize[1])
    sx, fx = _linear_coeffs(dw, W)
    # x: out-of-range source columns are handled by forcing the fraction to 0 (resize.cpp: "if( sx < ksize2-1 ) ... fx = 0, sx = 0")
    lo, hi = sx < 0, sx >= W - 1
    fx = np.where(lo | hi, np.float32(0), fx).astype(np.float32)
    sx = np.where(lo, 0, np.where(hi, W - 1, sx))
    a0 = _round_half_even_short((np.float32(1.0) - fx) * np.float32(2048))
    a1 = _round_half_even_short(fx * np.float32(2048))
    sx1 = np.minimum(sx + 1, W - 1)
    # y: indices are clamped, weights are kept (resizeGeneric_Invoker: clip(sy0 - ksize2 + 1 + k, 0, ssize.height))
    sy, fy = _linear_coeffs(dh, H)
    b0 = _round_half_even_short((np.float32(1.0) - fy) * np.float32(2048))
    b1 = _round_half_even_short(fy * np.float32(2048))
    sy0 = np.clip(sy, 0, H - 1)
    sy1 = np.clip(sy + 1, 0, H - 1)
    src = img.astype(np.int64)
    rows = src[:, sx, :] * a0[None, :, None] + src[:, sx1, :] * a1[None, :, None]          # (H, dw, C) 32-bit sums
    r0, r1 = rows[sy0], rows[sy1]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def crop_resize_batch(frames: np.ndarray, boxes, size: int = 96) -> np.ndarray:
    """inference.py:102 + :126 for a batch: boxes rows are (frame index, y1, y2, x1, x2); -> (N, size, size, 3) uint8."""
    return np.stack([resize_linear_u8(frames[f][y1:y2, x1:x2], (size, size)) for f, y1, y2, x1, x2 in boxes])


def paste_batch(pred_u8: np.ndarray, frames: np.ndarray, boxes) -> np.ndarray:
    """inference.py:123 (frame copy), :267-271: one output frame per item = its frame with the resized prediction pasted."""
    out = []
    for p, (f, y1, y2, x1, x2) in zip(pred_u8, boxes):
        fr = frames[f].copy()
        fr[y1:y2, x1:x2] = resize_linear_u8(p, (x2 - x1, y2 - y1))
        out.append(fr)
    return np.stack(out)
