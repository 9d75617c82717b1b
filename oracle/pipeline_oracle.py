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
