"""GPU mirrors of the script-level loss functions the evaluation loops call (forward values only):

  cosine_loss(a, v, y)        wav2lip_train.py:178-183, hq_wav2lip_train.py, color_syncnet_train.py:133-138
  recon_loss(g, gt)           nn.L1Loss(), wav2lip_train.py:191, :281
  get_sync_loss(syncnet, mel, g)   wav2lip_train.py:192-198 (the reference reads `syncnet` from a module global)

They return 0-dim CUDA tensors like the reference's, bound to libw2l.so (include/w2l.h); there is no autograd graph
behind them — the training step (backward, Adam) is the next scope row."""
import ctypes as C

import torch

from . import _lib


def _ctx(t: torch.Tensor, *others) -> "_lib.Context":
    """Context of t's device; every other tensor whose pointer crosses the C-ABI must live on that same device."""
    if not t.is_cuda:
        raise _lib.W2LError("wav2lip_b200 has no CPU path: tensors must be on a CUDA device")
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    for o in others:
        if o is not None and (not o.is_cuda or (o.device.index if o.device.index is not None else torch.cuda.current_device()) != dev):
            raise _lib.W2LError(f"expected every tensor on cuda:{dev}, got one on {o.device}")
    c = _CTX.get(dev)
    if c is None:
        c = _CTX[dev] = _lib.Context(dev)
    return c


_CTX = {}


def _f32(t):
    return t.detach().contiguous().float()


def cosine_loss(a, v, y=None):
    a, v = _f32(a), _f32(v)
    if a.dim() != 2 or a.shape != v.shape:
        raise ValueError(f"expected two (B,D) embeddings, got {tuple(a.shape)} and {tuple(v.shape)}")
    B, D = a.shape
    if y is not None:
        y = _f32(y).reshape(-1)
        if y.numel() != B:
            raise ValueError(f"expected {B} targets, got {y.numel()}")
    out = torch.empty((), device=a.device, dtype=torch.float32)
    if B == 0:
        return out.fill_(float("nan"))   # nn.BCELoss over an empty batch: mean of nothing
    ctx = _ctx(a, v, y)
    stream = torch.cuda.current_stream(a.device).cuda_stream
    _lib.check(ctx.lib.w2l_cosine_bce_loss(ctx.h, C.c_void_p(a.data_ptr()), C.c_void_p(v.data_ptr()),
                                           C.c_void_p(y.data_ptr()) if y is not None else None, B, D,
                                           C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
    return out


def recon_loss(g, gt):
    g, gt = _f32(g), _f32(gt)
    if g.shape != gt.shape:
        raise ValueError(f"shape mismatch {tuple(g.shape)} vs {tuple(gt.shape)}")
    out = torch.empty((), device=g.device, dtype=torch.float32)
    if g.numel() == 0:
        return out.fill_(float("nan"))
    ctx = _ctx(g, gt)
    stream = torch.cuda.current_stream(g.device).cuda_stream
    _lib.check(ctx.lib.w2l_l1_loss(ctx.h, C.c_void_p(g.data_ptr()), C.c_void_p(gt.data_ptr()), g.numel(),
                                   C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
    return out


def get_sync_loss(syncnet, mel, g, expert_training: bool = False):
    """syncnet: a wav2lip_b200.models.SyncNet_color in eval mode; mel (B,1,80,16); g (B,3,5,96,96).

    DIVERGENCE from the reference, made explicit in the signature: the reference scripts never call .eval() on their
    expert (wav2lip_train.py:187-189), so THEIR get_sync_loss runs the expert's BatchNorm on batch statistics (and moves
    its running averages).  This forward-value function runs the expert on its running statistics (`expert_training=
    False`), so the number it returns is NOT the one the reference logs as `sync_loss`.  The reference's behaviour
    (batch statistics, differentiable w.r.t. g) is what wav2lip_b200.training.TrainStep computes inside the training
    step; asking for it here raises."""
    if expert_training:
        raise NotImplementedError(
            "get_sync_loss(expert_training=True) — the reference's batch-statistics expert — is computed inside "
            "wav2lip_b200.training.TrainStep (which also back-propagates it); this evaluation helper runs the expert "
            "in eval mode only")
    a, v = syncnet.forward_frames(mel, g)
    return cosine_loss(a, v, None)
