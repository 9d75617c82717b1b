"""Mirrors of models.Wav2Lip and models.Wav2Lip_disc_qual (/root/reference/models/wav2lip.py).

Same class names, zero-arg constructors, forward signatures and state_dict keys; the forward
pass is ONE call into libw2l.so (w2l_generator_forward / w2l_disc_forward)."""
import ctypes as C

import torch
from torch import nn
from torch.nn import functional as F

from ._bridge import lib as _lib
from ._net import NativeNet, build_tree


class Wav2Lip(NativeNet):
    """wav2lip.py:8-125.  forward(audio_sequences, face_sequences): audio first.
       4-D: (N,1,80,16), (N,6,96,96) -> (N,3,96,96); 5-D: (B,T,1,80,16), (B,6,T,96,96) -> (B,3,T,96,96)."""
    NET = _lib.NET_GENERATOR

    def __init__(self):
        super().__init__()
        build_tree(self, self.NET)  # face_encoder_blocks, audio_encoder, face_decoder_blocks, output_block.0
        # wav2lip.py:83-85: output_block = Sequential(Conv2d(80,32,3,1,1), nn.Conv2d(32,3,1,1,0), Sigmoid)
        self.output_block.add_module("1", nn.Conv2d(32, 3, kernel_size=1, stride=1, padding=0))
        self.output_block.add_module("2", nn.Sigmoid())

    def forward(self, audio_sequences, face_sequences):
        if self.training:
            # wav2lip_train.py:211,220: BatchNorm on batch statistics + an autograd node whose backward is the native
            # dgrad / wgrad pass (wav2lip_b200/training.py, include/w2l.h w2l_train_forward / w2l_train_backward)
            from .. import training
            self._check_shapes(audio_sequences, face_sequences)
            return training.train_forward(self, "gen", audio_sequences, face_sequences)
        ctx = self._ensure(face_sequences)
        self._same_device(ctx, audio_sequences, face_sequences)
        mel, face = self._in(audio_sequences), self._in(face_sequences)
        if face.dim() > 4:  # wav2lip.py:91-94
            B, _, T, H, W = face.shape
            if tuple(mel.shape) != (B, T, 1, 80, 16) or face.shape[1] != 6 or (H, W) != (96, 96):
                raise ValueError(f"expected (B,T,1,80,16) and (B,6,T,96,96), got {tuple(mel.shape)} and {tuple(face.shape)}")
            out = torch.empty((B, 3, T, 96, 96), device=face.device, dtype=torch.float32)
        else:
            B, T = face.shape[0], 0
            if tuple(mel.shape) != (B, 1, 80, 16) or tuple(face.shape[1:]) != (6, 96, 96):
                raise ValueError(f"expected (N,1,80,16) and (N,6,96,96), got {tuple(mel.shape)} and {tuple(face.shape)}")
            out = torch.empty((B, 3, 96, 96), device=face.device, dtype=torch.float32)
        if out.numel() == 0:  # empty batch: torch returns an empty tensor, and so do we (no launch)
            return out
        stream = torch.cuda.current_stream(face.device).cuda_stream
        _lib.check(ctx.lib.w2l_generator_forward(ctx.h, self._p(mel), self._p(face), self._p(out), B, T, C.c_void_p(stream)))
        self._range_guard(ctx, stream)
        return out


    @staticmethod
    def _check_shapes(mel, face):
        if face.dim() > 4:
            B, _, T, H, W = face.shape
            ok = tuple(mel.shape) == (B, T, 1, 80, 16) and face.shape[1] == 6 and (H, W) == (96, 96)
        else:
            ok = face.dim() == 4 and tuple(mel.shape) == (face.shape[0], 1, 80, 16) and tuple(face.shape[1:]) == (6, 96, 96)
        if not ok or face.shape[0] == 0:
            raise ValueError(f"expected (N,1,80,16)+(N,6,96,96) or (B,T,1,80,16)+(B,6,T,96,96) with N > 0, got "
                             f"{tuple(mel.shape)} and {tuple(face.shape)}")

    def infer_u8(self, mel_batch, face_crops_u8):
        """The inner loop of inference.py with the batch assembly fused in (scope row f):
        mel_batch (N,1,80,16) float, face_crops_u8 (N,96,96,3) uint8 BGR crops already resized to 96x96
        (inference.py:126) -> (N,96,96,3) uint8 BGR predictions, i.e. `p.astype(np.uint8)` of inference.py:269.
        Equivalent to inference.py:134-140 + :259-265 on the GPU."""
        ctx = self._ensure(face_crops_u8)
        if face_crops_u8.dtype != torch.uint8 or face_crops_u8.dim() != 4 or tuple(face_crops_u8.shape[1:]) != (96, 96, 3):
            raise ValueError(f"expected uint8 (N,96,96,3) crops, got {face_crops_u8.dtype} {tuple(face_crops_u8.shape)}")
        N = face_crops_u8.shape[0]
        self._same_device(ctx, mel_batch, face_crops_u8)
        mel = self._in(mel_batch)
        if tuple(mel.shape) != (N, 1, 80, 16):
            raise ValueError(f"expected mel (N,1,80,16), got {tuple(mel.shape)}")
        faces = face_crops_u8.contiguous()
        out = torch.empty((N, 96, 96, 3), device=faces.device, dtype=torch.uint8)
        if N == 0:
            return out
        stream = torch.cuda.current_stream(faces.device).cuda_stream
        _lib.check(ctx.lib.w2l_generator_forward_u8(ctx.h, self._p(mel), self._p(faces), self._p(out), N, C.c_void_p(stream)))
        self._range_guard(ctx, stream)
        return out


    @staticmethod
    def _boxes(boxes, n_expected=None):
        import numpy as np
        b = np.ascontiguousarray(np.asarray(boxes.cpu() if isinstance(boxes, torch.Tensor) else boxes, dtype=np.int32))
        if b.ndim != 2 or b.shape[1] != 5 or (n_expected is not None and b.shape[0] != n_expected):
            raise ValueError(f"expected boxes of shape (N,5) = (frame index, y1, y2, x1, x2), got {b.shape}")
        return b

    def crop_resize(self, frames_u8, boxes):
        """inference.py:102 + :126 on the GPU: frames (F,H,W,3) uint8 BGR (CUDA), boxes (N,5) rows (frame index, y1, y2, x1, x2)
        -> (N,96,96,3) uint8, bit-identical to `cv2.resize(frames[f][y1:y2, x1:x2], (96, 96))`."""
        ctx = self._ensure(frames_u8)
        self._same_device(ctx, frames_u8)
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[3] != 3:
            raise ValueError(f"expected uint8 (F,H,W,3) frames, got {frames_u8.dtype} {tuple(frames_u8.shape)}")
        b = self._boxes(boxes)
        fr = frames_u8.contiguous()
        out = torch.empty((b.shape[0], 96, 96, 3), device=fr.device, dtype=torch.uint8)
        if b.shape[0] == 0:
            return out
        stream = torch.cuda.current_stream(fr.device).cuda_stream
        _lib.check(ctx.lib.w2l_crop_resize_u8(ctx.h, self._p(fr), fr.shape[0], fr.shape[1], fr.shape[2],
                                              b.ctypes.data_as(C.POINTER(C.c_int32)), b.shape[0], self._p(out), C.c_void_p(stream)))
        return out

    def paste(self, pred_u8, frames_u8, boxes):
        """inference.py:267-271 on the GPU: pred (N,96,96,3) uint8 -> (N,H,W,3) uint8 = copies of frames[box.frame] with the
        prediction resized to the box (cv2.resize arithmetic, bit-identical) and pasted."""
        ctx = self._ensure(frames_u8)
        self._same_device(ctx, frames_u8, pred_u8)
        b = self._boxes(boxes, pred_u8.shape[0])
        if pred_u8.dtype != torch.uint8 or tuple(pred_u8.shape[1:]) != (96, 96, 3) or frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4:
            raise ValueError("expected uint8 pred (N,96,96,3) and uint8 frames (F,H,W,3)")
        fr, pr = frames_u8.contiguous(), pred_u8.contiguous()
        out = torch.empty((b.shape[0],) + tuple(fr.shape[1:]), device=fr.device, dtype=torch.uint8)
        if b.shape[0] == 0:
            return out
        stream = torch.cuda.current_stream(fr.device).cuda_stream
        _lib.check(ctx.lib.w2l_paste_u8(ctx.h, self._p(pr), self._p(fr), fr.shape[0], fr.shape[1], fr.shape[2],
                                        b.ctypes.data_as(C.POINTER(C.c_int32)), b.shape[0], self._p(out), C.c_void_p(stream)))
        return out

    def infer_frames(self, mel_batch, frames_u8, boxes):
        """The whole inner loop of inference.py (:120-140, :259-271) in one native call: for every (mel chunk, box) pair crop the
        face box out of its frame, resize to 96x96, mask / concat / normalise, run the generator, scale to uint8, resize the
        prediction back to the box and paste it into a copy of the frame.
        mel_batch (N,1,80,16) float, frames (F,H,W,3) uint8 BGR, boxes (N,5) rows (frame index, y1, y2, x1, x2) -> (N,H,W,3) uint8."""
        ctx = self._ensure(frames_u8)
        self._same_device(ctx, frames_u8, mel_batch)
        b = self._boxes(boxes, mel_batch.shape[0])
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[3] != 3:
            raise ValueError(f"expected uint8 (F,H,W,3) frames, got {frames_u8.dtype} {tuple(frames_u8.shape)}")
        mel = self._in(mel_batch)
        if tuple(mel.shape) != (b.shape[0], 1, 80, 16):
            raise ValueError(f"expected mel (N,1,80,16), got {tuple(mel.shape)}")
        fr = frames_u8.contiguous()
        out = torch.empty((b.shape[0],) + tuple(fr.shape[1:]), device=fr.device, dtype=torch.uint8)
        if b.shape[0] == 0:
            return out
        stream = torch.cuda.current_stream(fr.device).cuda_stream
        _lib.check(ctx.lib.w2l_lipsync_frames_u8(ctx.h, self._p(mel), self._p(fr), fr.shape[0], fr.shape[1], fr.shape[2],
                                                 b.ctypes.data_as(C.POINTER(C.c_int32)), b.shape[0], self._p(out), C.c_void_p(stream)))
        self._range_guard(ctx, stream)
        return out

    def infer_stream(self, batches, device=None):
        """Serving loop over HOST batches with the copies overlapped with the kernels (w2l_generator_submit_*_host /
        w2l_host_wait, include/w2l.h): `batches` yields (mel, faces) CPU tensors — fp32 (N,1,80,16) with either fp32
        (N,6,96,96) assembled inputs (inference.py:259-260) or uint8 (N,96,96,3) crops (inference.py:126) — and the
        generator yields, in order, fp32 (N,3,96,96) or uint8 (N,96,96,3) CPU tensors.  Inputs are staged through two
        sets of pinned buffers that are reused from batch to batch; batch k+1 is uploaded while batch k computes and
        batch k-1 downloads."""
        dev = torch.device(device if device is not None else next(self.parameters()).device)
        if dev.type != "cuda":
            raise RuntimeError("wav2lip_b200 has no CPU path: move the module to a CUDA device first")
        ctx = self._ensure(torch.empty(0, device=dev))
        lib = ctx.lib
        pool = {}

        def pinned(slot, kind, shape, dtype):
            n = 1
            for d in shape:
                n *= int(d)
            t = pool.get((slot, kind, dtype))
            if t is None or t.numel() < n:
                t = pool[(slot, kind, dtype)] = torch.empty(max(n, 1), dtype=dtype).pin_memory()
            return t[:n].view(shape)

        pending = []   # out buffers of the submissions in flight, oldest first
        seq = 0
        try:
            for mel, faces in batches:
                u8 = faces.dtype == torch.uint8
                N = int(faces.shape[0])
                if tuple(mel.shape) != (N, 1, 80, 16) or tuple(faces.shape[1:]) != ((96, 96, 3) if u8 else (6, 96, 96)):
                    raise ValueError(f"bad batch shapes mel {tuple(mel.shape)} faces {tuple(faces.shape)}")
                if N == 0:
                    raise ValueError("empty batch")
                slot = seq & 1     # the batch that used this slot two submissions ago has been retired (and its output copied)
                mel_p = pinned(slot, "mel", (N, 1, 80, 16), torch.float32)
                mel_p.copy_(mel)
                faces_p = pinned(slot, "faces", tuple(faces.shape), faces.dtype if u8 else torch.float32)
                faces_p.copy_(faces)
                out = pinned(slot, "out", (N, 96, 96, 3) if u8 else (N, 3, 96, 96), torch.uint8 if u8 else torch.float32)
                if u8:
                    _lib.check(lib.w2l_generator_submit_u8_host(ctx.h, self._p(mel_p), self._p(faces_p), self._p(out), N))
                else:
                    _lib.check(lib.w2l_generator_submit_host(ctx.h, self._p(mel_p), self._p(faces_p), self._p(out), N, 0))
                pending.append(out)
                seq += 1
                if len(pending) == 2:
                    _lib.check(lib.w2l_host_wait(ctx.h, 1))
                    yield pending.pop(0).clone()
            while pending:
                _lib.check(lib.w2l_host_wait(ctx.h, len(pending) - 1))
                yield pending.pop(0).clone()
        finally:
            _lib.check(lib.w2l_host_wait(ctx.h, 0))


class Wav2Lip_disc_qual(NativeNet):
    """wav2lip.py:127-184.  forward((B,3,T,96,96)) -> (B*T, 1), rows t-major."""
    NET = _lib.NET_DISC

    def __init__(self):
        super().__init__()
        build_tree(self, self.NET)
        self.binary_pred = nn.Sequential(nn.Conv2d(512, 1, kernel_size=1, stride=1, padding=0), nn.Sigmoid())
        self.label_noise = .0

    def get_lower_half(self, face_sequences):
        return face_sequences[:, :, face_sequences.size(2) // 2:]

    def to_2d(self, face_sequences):
        return torch.cat([face_sequences[:, :, i] for i in range(face_sequences.size(2))], dim=0)

    def forward(self, face_sequences):
        if self.training and torch.is_grad_enabled():
            # hq_wav2lip_train.py:225-253: the perceptual loss back-propagates THROUGH the discriminator into the
            # generator, and the discriminator's own step differentiates w.r.t. its parameters
            from .. import training
            if face_sequences.dim() != 5 or face_sequences.shape[1] != 3 or tuple(face_sequences.shape[3:]) != (96, 96) or face_sequences.shape[0] == 0:
                raise ValueError(f"expected (B,3,T,96,96), got {tuple(face_sequences.shape)}")
            return training.train_forward(self, "disc", face_sequences, None)
        ctx = self._ensure(face_sequences)
        self._same_device(ctx, face_sequences)
        x = self._in(face_sequences)
        if x.dim() != 5 or x.shape[1] != 3 or tuple(x.shape[3:]) != (96, 96):
            raise ValueError(f"expected (B,3,T,96,96), got {tuple(x.shape)}")
        B, T = x.shape[0], x.shape[2]
        out = torch.empty((B * T, 1), device=x.device, dtype=torch.float32)
        if out.numel() == 0:
            return out
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(ctx.lib.w2l_disc_forward(ctx.h, self._p(x), self._p(out), B, T, C.c_void_p(stream)))
        self._range_guard(ctx, stream)
        return out

    def perceptual_forward(self, false_face_sequences):
        # wav2lip.py:163-174: BCE(pred, 1) on the device the input lives on
        pred = self.forward(false_face_sequences)
        return F.binary_cross_entropy(pred, torch.ones_like(pred))
