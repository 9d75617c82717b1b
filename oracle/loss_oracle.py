"""ORACLE (test infrastructure only — used by tests/ and nothing else): the script-level loss functions of the
evaluation loops, restated on torch CPU.  The reference defines them at module level of scripts that parse argv on
import (wav2lip_train.py, color_syncnet_train.py), so they cannot be imported; their arithmetic IS torch's
(nn.BCELoss, F.cosine_similarity, nn.L1Loss), which is what is called here — fp32, CPU.

  cosine_loss      /root/reference/wav2lip_train.py:178-183  (same text in color_syncnet_train.py:133-138, hq_wav2lip_train.py)
  recon_loss       /root/reference/wav2lip_train.py:191
  get_sync_loss    /root/reference/wav2lip_train.py:192-198  (syncnet_T = 5, hparams.py)

Mode of the expert: wav2lip_train.py:187-189 / hq_wav2lip_train.py:189-191 construct the expert SyncNet, freeze its
parameters and never call .eval() on it, so in the scripts its BatchNorm layers run on BATCH statistics (and keep
updating their running averages) even inside eval_model.  This restatement — like the product — evaluates the expert in
eval mode (running statistics), i.e. what the same function computes once the caller has put the expert in eval mode;
batch-statistics BatchNorm belongs to the training row and is not built.
"""
import torch
from torch import nn

from . import w2l_oracle as O

syncnet_T = 5
logloss = nn.BCELoss()
recon_loss = nn.L1Loss()


def cosine_loss(a, v, y):
    d = nn.functional.cosine_similarity(a, v)
    return logloss(d.unsqueeze(1), y)


def stack_lower_halves(g):
    """wav2lip_train.py:193-194: (B,3,T,H,W) -> (B, 3*T, H//2, W)"""
    g = g[:, :, :, g.size(3) // 2:]
    return torch.cat([g[:, :, i] for i in range(syncnet_T)], dim=1)


def get_sync_loss(syncnet_sd, mel, g):
    a, v = O.syncnet_forward(syncnet_sd, mel, stack_lower_halves(g))
    y = torch.ones(g.size(0), 1).float()
    return cosine_loss(a, v, y)
