"""Mirror of models.SyncNet_color (/root/reference/models/syncnet.py:7-66)."""
import ctypes as C

import torch

from ._bridge import lib as _lib
from ._net import NativeNet, build_tree


class SyncNet_color(NativeNet):
    """forward(audio (B,1,80,16), face (B,15,48,96)) -> (audio_embedding, face_embedding), each (B,512),
    L2-normalised (syncnet.py:59-66)."""
    NET = _lib.NET_SYNCNET

    def __init__(self):
        super().__init__()
        build_tree(self, self.NET)  # face_encoder, audio_encoder

    def forward(self, audio_sequences, face_sequences):
        if self.training:
            # color_syncnet_train.py:146-163 (the expert's own training) and get_sync_loss inside the generator's step
            # (wav2lip_train.py:192-198: the scripts never call .eval() on the frozen expert, :187-189)
            from .. import training
            B = face_sequences.shape[0]
            if B == 0 or tuple(audio_sequences.shape) != (B, 1, 80, 16) or tuple(face_sequences.shape[1:]) != (15, 48, 96):
                raise ValueError(f"expected (B,1,80,16) and (B,15,48,96), got {tuple(audio_sequences.shape)} and {tuple(face_sequences.shape)}")
            return training.train_forward(self, "stacked", audio_sequences, face_sequences)
        ctx = self._ensure(face_sequences)
        self._same_device(ctx, audio_sequences, face_sequences)
        mel, face = self._in(audio_sequences), self._in(face_sequences)
        B = face.shape[0]
        if tuple(mel.shape) != (B, 1, 80, 16) or tuple(face.shape[1:]) != (15, 48, 96):
            raise ValueError(f"expected (B,1,80,16) and (B,15,48,96), got {tuple(mel.shape)} and {tuple(face.shape)}")
        a = torch.empty((B, 512), device=face.device, dtype=torch.float32)
        v = torch.empty((B, 512), device=face.device, dtype=torch.float32)
        if B == 0:
            return a, v
        stream = torch.cuda.current_stream(face.device).cuda_stream
        _lib.check(ctx.lib.w2l_syncnet_forward(ctx.h, self._p(mel), self._p(face), self._p(a), self._p(v), B, C.c_void_p(stream)))
        self._range_guard(ctx, stream)
        return a, v

    def forward_frames(self, audio_sequences, frames):
        """The expert-discriminator call of `get_sync_loss` (wav2lip_train.py:192-196) on generator output / ground truth:
        frames (B,3,T=5,96,96) -> lower half, T frames stacked on channels -> (audio_embedding, face_embedding).
        Same result as forward(mel, cat([frames[:, :, i, 48:] for i in range(5)], 1)) without materialising the stack."""
        if self.training:
            from .. import training
            B = frames.shape[0]
            if B == 0 or frames.dim() != 5 or tuple(frames.shape[1:]) != (3, 5, 96, 96) or tuple(audio_sequences.shape) != (B, 1, 80, 16):
                raise ValueError(f"expected (B,1,80,16) and (B,3,5,96,96), got {tuple(audio_sequences.shape)} and {tuple(frames.shape)}")
            return training.train_forward(self, "frames", audio_sequences, frames)
        ctx = self._ensure(frames)
        self._same_device(ctx, audio_sequences, frames)
        mel, fr = self._in(audio_sequences), self._in(frames)
        B = fr.shape[0]
        if fr.dim() != 5 or tuple(fr.shape[1:]) != (3, 5, 96, 96) or tuple(mel.shape) != (B, 1, 80, 16):
            raise ValueError(f"expected (B,1,80,16) and (B,3,5,96,96), got {tuple(mel.shape)} and {tuple(fr.shape)}")
        a = torch.empty((B, 512), device=fr.device, dtype=torch.float32)
        v = torch.empty((B, 512), device=fr.device, dtype=torch.float32)
        if B == 0:
            return a, v
        stream = torch.cuda.current_stream(fr.device).cuda_stream
        _lib.check(ctx.lib.w2l_syncnet_forward_frames(ctx.h, self._p(mel), self._p(fr), self._p(a), self._p(v), B, 5, C.c_void_p(stream)))
        self._range_guard(ctx, stream)
        return a, v
