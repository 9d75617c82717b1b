"""`FaceAlignment` as inference.py uses it (face_detection/api.py:46-79 of the reference): a face detector behind
`get_detections_for_batch`.  Only the 'sfd' detector exists."""
from enum import Enum

import numpy as np


class LandmarksType(Enum):
    _2D = 1
    _2halfD = 2
    _3D = 3


class NetworkSize(Enum):
    LARGE = 4

    def __int__(self):
        return self.value


class FaceAlignment:
    def __init__(self, landmarks_type, network_size=NetworkSize.LARGE, device="cuda", flip_input=False, face_detector="sfd",
                 verbose=False, path_to_detector=None):
        if face_detector != "sfd":
            raise ValueError("only the S3FD ('sfd') detector is provided")
        from .detection.sfd import FaceDetector
        self.device, self.flip_input, self.landmarks_type, self.verbose = device, flip_input, landmarks_type, verbose
        self.face_detector = FaceDetector(device=device, verbose=verbose, path_to_detector=path_to_detector)

    def get_detections_for_batch(self, images):
        """images: (B,H,W,3) uint8 RGB (inference.py:85 passes frames converted to RGB; api.py:64 flips them back to BGR).
        Returns, per image, the first surviving detection as (x1, y1, x2, y2) ints clipped at 0, or None."""
        bgr = np.ascontiguousarray(np.asarray(images)[..., ::-1])
        out = []
        for dets in self.face_detector.detect_from_batch(bgr):
            if len(dets) == 0:
                out.append(None)
                continue
            d = np.clip(dets[0], 0, None)
            out.append(tuple(int(v) for v in d[:4]))
        return out
