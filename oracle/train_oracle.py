"""ORACLE for the NEXT scope row (the training step) — test infrastructure only; no product
code exists for this row yet.  torch-CPU restatement of one optimisation step of

  wav2lip_train.py:210-231        generator: g = model(indiv_mels, x); sync_loss = get_sync_loss(mel, g) (:192-198);
                                  l1 = L1(g, gt); loss = wt*sync + (1-wt)*l1; backward; Adam step (:357-360, lr 1e-4)
  color_syncnet_train.py:146-163  expert: a, v = model(mel, x); loss = cosine_loss(a, v, y); backward; Adam step
  hq_wav2lip_train.py:213-255     generator with the quality discriminator: loss = wt*sync + disc_wt*perceptual +
                                  (1-wt-disc_wt)*l1 (:229-240), Adam betas (0.5, 0.999) (:418-421); then the
                                  discriminator on gt (target 1) and on g.detach() (target 0), two backward calls, one step

on the functional nets of oracle/w2l_oracle.py (train-mode BatchNorm: batch statistics, running stats updated with
momentum 0.1), with autograd for the gradients and a spelled-out Adam (torch.optim.Adam defaults: betas (0.9, 0.999),
eps 1e-8, no weight decay, no amsgrad).  Pinned by tests/golden/train.npz, produced by tests/golden/make_golden_train.py
from the REAL reference modules + torch.optim.Adam (tests/test_train_oracle.py).

Quirk kept: the scripts never call .eval() on the frozen expert (wav2lip_train.py:187-189), so inside get_sync_loss its
BatchNorm layers use batch statistics and update their running averages; `expert_training=True` reproduces that.
"""
from typing import Dict, Optional

import torch

from . import loss_oracle as LO
from . import w2l_oracle as O

BETAS = (0.9, 0.999)
ADAM_EPS = 1e-8


def is_param(name: str) -> bool:
    return name.endswith(".weight") or name.endswith(".bias")


def leaves(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Copy of sd whose parameters are autograd leaves; buffers (running stats) are plain clones."""
    out = {}
    for k, v in sd.items():
        t = v.clone()
        if is_param(k) and t.dtype.is_floating_point:
            t.requires_grad_(True)
        out[k] = t
    return out


def adam_step(sd, grads, state: Optional[dict], lr: float, betas=BETAS):
    """torch.optim.Adam.step for the parameters that have a gradient; returns (new sd tensors in place, state)."""
    if state is None:
        state = {"step": 0, "m": {}, "v": {}}
    state["step"] += 1
    t = state["step"]
    bc1 = 1 - betas[0] ** t
    bc2 = 1 - betas[1] ** t
    with torch.no_grad():
        for k, g in grads.items():
            m = state["m"].setdefault(k, torch.zeros_like(g))
            v = state["v"].setdefault(k, torch.zeros_like(g))
            m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
            v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
            denom = (v.sqrt() / (bc2 ** 0.5)).add_(ADAM_EPS)
            sd[k].addcdiv_(m, denom, value=-(lr / bc1))
    return state


def wav2lip_train_step(gen_sd, sync_sd, x, indiv_mels, mel, gt, syncnet_wt: float, lr: float = 1e-4,
                       state: Optional[dict] = None, expert_training: bool = True):
    """One iteration of wav2lip_train.py:210-231.  gen_sd / sync_sd are updated IN PLACE (parameters by Adam, BatchNorm
    buffers by the train-mode forward).  Returns losses, gradients and the optimizer state."""
    g_sd = leaves(gen_sd)
    g = O.generator_forward(g_sd, indiv_mels, x, training=True)
    if syncnet_wt > 0.0:
        e_sd = {k: v for k, v in sync_sd.items()}           # frozen: no leaves, buffers updated in place if training
        a, v = O.syncnet_forward(e_sd, mel, LO.stack_lower_halves(g), training=expert_training)
        sync_loss = LO.cosine_loss(a, v, torch.ones(g.size(0), 1))
    else:
        sync_loss = torch.zeros(())
    l1 = LO.recon_loss(g, gt)
    loss = syncnet_wt * sync_loss + (1 - syncnet_wt) * l1
    names = [k for k, t in g_sd.items() if t.requires_grad]
    gs = torch.autograd.grad(loss, [g_sd[k] for k in names])
    grads = dict(zip(names, gs))
    with torch.no_grad():
        for k, t in g_sd.items():                           # running stats moved by the forward; bump the counters
            if not t.requires_grad:
                gen_sd[k].copy_(t)
            if k.endswith("num_batches_tracked"):
                gen_sd[k] += 1
        if syncnet_wt > 0.0 and expert_training:
            for k in sync_sd:
                if k.endswith("num_batches_tracked"):
                    sync_sd[k] += 1
    state = adam_step(gen_sd, grads, state, lr)
    return {"loss": loss.detach(), "sync_loss": sync_loss.detach(), "l1": l1.detach(), "g": g.detach(), "grads": grads,
            "state": state}


def syncnet_train_step(sd, x, mel, y, lr: float = 1e-4, state: Optional[dict] = None):
    """One iteration of color_syncnet_train.py:146-163 (x (B,15,48,96), mel (B,1,80,16), y (B,1) in {0,1})."""
    s_sd = leaves(sd)
    a, v = O.syncnet_forward(s_sd, mel, x, training=True)
    loss = LO.cosine_loss(a, v, y)
    names = [k for k, t in s_sd.items() if t.requires_grad]
    gs = torch.autograd.grad(loss, [s_sd[k] for k in names])
    grads = dict(zip(names, gs))
    with torch.no_grad():
        for k, t in s_sd.items():
            if not t.requires_grad:
                sd[k].copy_(t)
            if k.endswith("num_batches_tracked"):
                sd[k] += 1
    state = adam_step(sd, grads, state, lr)
    return {"loss": loss.detach(), "grads": grads, "state": state}


HQ_BETAS = (0.5, 0.999)   # hq_wav2lip_train.py:418-421


def _bce(pred, target_value: float):
    return torch.nn.functional.binary_cross_entropy(pred, torch.full((len(pred), 1), target_value))


def hq_train_step(gen_sd, disc_sd, sync_sd, x, indiv_mels, mel, gt, syncnet_wt: float, disc_wt: float, lr: float = 1e-4,
                  disc_lr: float = 1e-4, states: Optional[dict] = None, expert_training: bool = True):
    """One iteration of hq_wav2lip_train.py:213-255.  perceptual_forward (wav2lip.py:163-174) == BCE(disc(g), 1).
    gen_sd / disc_sd (/ sync_sd buffers) are updated in place."""
    if states is None:
        states = {"gen": None, "disc": None}
    g_sd = leaves(gen_sd)
    d_sd = leaves(disc_sd)                                   # the perceptual loss back-propagates THROUGH the disc ...
    g = O.generator_forward(g_sd, indiv_mels, x, training=True)
    if syncnet_wt > 0.0:
        a, v = O.syncnet_forward(dict(sync_sd), mel, LO.stack_lower_halves(g), training=expert_training)
        sync_loss = LO.cosine_loss(a, v, torch.ones(g.size(0), 1))
    else:
        sync_loss = torch.zeros(())
    perceptual = _bce(O.disc_forward(d_sd, g), 1.0) if disc_wt > 0.0 else torch.zeros(())
    l1 = LO.recon_loss(g, gt)
    loss = syncnet_wt * sync_loss + disc_wt * perceptual + (1.0 - syncnet_wt - disc_wt) * l1
    names = [k for k, t in g_sd.items() if t.requires_grad]
    gs = torch.autograd.grad(loss, [g_sd[k] for k in names])   # ... but its gradients w.r.t. the disc are zeroed (:243)
    g_grads = dict(zip(names, gs))
    with torch.no_grad():
        for k, t in g_sd.items():
            if not t.requires_grad:
                gen_sd[k].copy_(t)
            if k.endswith("num_batches_tracked"):
                gen_sd[k] += 1
        if syncnet_wt > 0.0 and expert_training:
            for k in sync_sd:
                if k.endswith("num_batches_tracked"):
                    sync_sd[k] += 1
    states["gen"] = adam_step(gen_sd, g_grads, states["gen"], lr, HQ_BETAS)
    # discriminator: real then fake, gradients accumulate, one step (:245-253)
    d_sd = leaves(disc_sd)
    real = _bce(O.disc_forward(d_sd, gt), 1.0)
    fake = _bce(O.disc_forward(d_sd, g.detach()), 0.0)
    dnames = [k for k, t in d_sd.items() if t.requires_grad]
    dgs = torch.autograd.grad(real + fake, [d_sd[k] for k in dnames])   # sum of the two backward() calls
    d_grads = dict(zip(dnames, dgs))
    states["disc"] = adam_step(disc_sd, d_grads, states["disc"], disc_lr, HQ_BETAS)
    return {"loss": loss.detach(), "sync_loss": sync_loss.detach(), "perceptual": perceptual.detach(), "l1": l1.detach(),
            "disc_real": real.detach(), "disc_fake": fake.detach(), "g": g.detach(), "gen_grads": g_grads,
            "disc_grads": d_grads, "states": states}
