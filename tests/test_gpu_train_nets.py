"""Training row, network level, through the public surface: the mirrors in train mode + autograd
(the reference scripts' `loss.backward()`), and the fused native step `Wav2LipTrainStep` (w2l_wav2lip_train_step).

What can be asserted, and why (tests/test_precision_model.py): on the reference's initialisation the
generator's end-to-end gradients amplify relative perturbations by ~1e5 (fp32 vs fp64: 0.6 %; TF32-class operands, i.e.
the reference's own GPU arithmetic: 24 %; two bf16 runs whose inputs differ by 1e-7: 50 %, cosine 0.87) — ReLU-mask flips
of near-zero pre-activations under batch-statistics BatchNorm.  So:
  * kernels are held to the per-block bar in tests/test_gpu_train_blocks.py (2e-2, every geometry);
  * the WIRING of the networks (skip-concat gradient split, residual and skip accumulation, transposed-conv phases, the
    expert's and the discriminator's input gradients, t-major flatten) is checked here on weights whose BatchNorm shift
    keeps (almost) every ReLU active (beta = 3) and at a batch size that gives the 1x1-resolution BatchNorms at least 8 samples
    per channel: the backward is then smooth and bf16 rounding stays at the per-cent level for most tensors (the CPU model
    with the same rounding points, oracle/backward_recipe.py + bf16 hooks, measures median 2.8 %, 90th percentile 13 %, worst
    22 % against float64 at N = 8; at N = 4 the audio encoder's tiny gradients are already 85 % off in that CPU model, and
    the kernels reproduce exactly that pattern), against float64 autograd through the oracle, every parameter tensor;
  * the reference's own step (tests/golden/train.npz, produced by the real modules + torch.optim.Adam) pins what IS well
    conditioned: losses, the generator output, the head's gradient, the Adam update size, the BatchNorm buffers."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import loss_oracle as LO
from oracle import w2l_oracle as O

pytestmark = pytest.mark.gpu


def rel_l2(got, ref):
    return ((got.double().cpu() - ref.double()).norm() / (ref.double().norm() + 1e-30)).item()


def fp3(t):
    f = t.detach().double().flatten().cpu()
    return np.array([f.sum().item(), f.abs().sum().item(), f.abs().max().item()])


def relu_active(sd, beta=3.0):
    return {k: (torch.full_like(v, beta) if k.endswith("conv_block.1.bias") else v.clone()) for k, v in sd.items()}


def autograd_reference(forward, sd, scalar_of_outputs):
    """float64 autograd through the oracle: returns {param name: grad} for every weight / bias."""
    leaves = {k: v.double().clone().requires_grad_(k.endswith(".weight") or k.endswith(".bias"))
              for k, v in sd.items() if v.dtype.is_floating_point}
    out = forward(leaves)
    loss = scalar_of_outputs(out)
    names = [k for k, v in leaves.items() if v.requires_grad]
    return out, dict(zip(names, torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)))


def param_grad_errors(module, ref):
    """{parameter name: relative L2 error of .grad} for every parameter with a structurally non-zero reference gradient."""
    scale = max(g.double().norm().item() for g in ref.values() if g is not None)
    errs = {}
    for name, p in module.named_parameters():
        r = ref[name]
        if name.endswith("conv_block.0.bias") and module.NET != 2:
            continue                      # conv bias under a BatchNorm: the true gradient is 0 (we return exactly 0)
        if r is None or r.double().norm().item() < 1e-6 * scale:
            continue                      # structurally zero gradients (a constant shift removed by the next BatchNorm)
        assert p.grad is not None, name
        errs[name] = rel_l2(p.grad, r)
    return errs


def summarize(errs):
    v = sorted(errs.values())
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:8]
    return {"n": len(v), "median": round(v[len(v) // 2], 4), "p90": round(v[int(len(v) * 0.9)], 4), "max": round(v[-1], 4),
            "worst": [(k, round(e, 4)) for k, e in worst]}


def report(name, rep):
    """Measured numbers of a test, one JSON line on stdout (pytest -s) — the tolerances below were set from these."""
    import json
    print("REPORT " + json.dumps({"test": name, **rep}, default=str), flush=True)


def _train_inputs(B, seed):
    g = torch.Generator().manual_seed(seed)
    indiv_mels, x = O.make_generator_inputs(B, seed=seed, t=5)
    mel = torch.rand((B, 1, 80, 16), generator=g) * 8 - 4
    gt = torch.rand((B, 3, 5, 96, 96), generator=g)
    return x, indiv_mels, mel, gt


def test_generator_train_mode_forward_matches_oracle():
    """model.train(); model(indiv_mels, x): BatchNorm on batch statistics over the T*B flatten, running averages and
    num_batches_tracked updated as nn.BatchNorm2d does."""
    from wav2lip_b200.models import Wav2Lip
    sd = O.make_state_dict("generator", 0, init="default")
    mel, face = O.make_generator_inputs(2, seed=7, t=5)
    ref_sd = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        ref = O.generator_forward(ref_sd, mel, face, training=True)
    g = Wav2Lip()
    g.load_state_dict(sd, strict=True)
    g = g.cuda().train()
    with torch.no_grad():
        y = g(mel.cuda(), face.cuda())
    assert tuple(y.shape) == (2, 3, 5, 96, 96)
    # batch statistics over only N = 10 crops (10 values per channel at the 1x1 layers) amplify the bf16 rounding of ~50
    # blocks: the CPU model with the same rounding points (oracle/backward_recipe.py, bf16 hooks) differs from float64 by
    # 0.092 max / 0.0096 mean on this very input; the kernels measure 0.13 / 0.01
    err = (y.cpu() - ref).abs().max().item()
    assert err <= 0.25, err
    assert (y.cpu() - ref).abs().mean().item() <= 0.03, (y.cpu() - ref).abs().mean().item()
    got = g.state_dict()
    for k, v in ref_sd.items():
        if k.endswith("running_mean"):
            assert (got[k].cpu() - v).abs().max().item() <= 2e-2 * max(1.0, v.abs().max().item()), k
        elif k.endswith("running_var"):
            assert rel_l2(got[k], v) <= 3e-2, k
        elif k.endswith("num_batches_tracked"):
            assert int(got[k]) == 1, k


def test_generator_backward_wiring_against_float64_autograd():
    """Every parameter gradient of the generator (4-D call) through the autograd bridge, all ReLUs active."""
    from wav2lip_b200.models import Wav2Lip
    sd = relu_active(O.make_state_dict("generator", 0, init="default"))
    N = 8
    mel, face = O.make_generator_inputs(N, seed=1)
    dout = torch.randn((N, 3, 96, 96), generator=torch.Generator().manual_seed(4)) / (N * 3 * 9216)
    out_ref, ref = autograd_reference(lambda s: O.generator_forward(s, mel.double(), face.double(), training=True), sd,
                                      lambda o: (o * dout.double()).sum())
    g = Wav2Lip()
    g.load_state_dict(sd, strict=True)
    g = g.cuda().train()
    y = g(mel.cuda(), face.cuda())
    assert y.requires_grad
    fwd = (y.detach().cpu() - out_ref.detach()).abs().max().item()      # CPU bf16 model vs float64 on this input: 0.018
    (y * dout.cuda()).sum().backward()
    errs = param_grad_errors(g, ref)
    rep = summarize(errs)
    rep["fwd_max_abs"] = round(fwd, 4)
    rep["head"] = round(errs["output_block.1.weight"], 4)
    rep["all"] = {k: round(v, 3) for k, v in errs.items() if k.endswith("conv_block.0.weight")}
    report("generator_wiring", rep)
    assert fwd <= 0.05 and rep["max"] <= 0.5 and rep["p90"] <= 0.25 and rep["median"] <= 0.06 and rep["head"] <= 3e-2 and rep["n"] >= 120, rep


def test_generator_backward_5d_is_the_tmajor_flatten():
    """wav2lip.py:93-94,119-120 in the backward: a 5-D training call equals the 4-D call on the t-major flattened batch."""
    from wav2lip_b200.models import Wav2Lip
    sd = relu_active(O.make_state_dict("generator", 0, init="default"))
    mel5, face5 = O.make_generator_inputs(2, seed=3, t=3)
    mel4 = torch.cat([mel5[:, i] for i in range(3)], 0)
    face4 = torch.cat([face5[:, :, i] for i in range(3)], 0)
    d5 = torch.randn((2, 3, 3, 96, 96), generator=torch.Generator().manual_seed(5))
    d4 = torch.cat([d5[:, :, i] for i in range(3)], 0)
    grads = []
    for mel, face, d in ((mel5, face5, d5), (mel4, face4, d4)):
        g = Wav2Lip()
        g.load_state_dict(sd, strict=True)
        g = g.cuda().train()
        y = g(mel.cuda(), face.cuda())
        (y * d.cuda()).sum().backward()
        grads.append({n: p.grad.clone() for n, p in g.named_parameters()})
    for n in grads[0]:
        assert torch.equal(grads[0][n], grads[1][n]), n


def test_syncnet_training_through_autograd_bridge():
    """color_syncnet_train.py:146-163: a, v = model(mel, x); loss = cosine_loss(a, v, y); loss.backward() — both encoders,
    every parameter, against float64 autograd (all ReLUs active)."""
    from wav2lip_b200.models import SyncNet_color
    sd = relu_active(O.make_state_dict("syncnet", 2, init="default"))
    mel, face = O.make_syncnet_inputs(16, seed=5)
    y = torch.tensor([[1.0], [0.0]] * 8)

    def loss_of(av):
        d = F.cosine_similarity(av[0], av[1])
        return F.binary_cross_entropy(d.unsqueeze(1), y.to(d.dtype))

    (a_ref, v_ref), ref = autograd_reference(lambda s: O.syncnet_forward(s, mel.double(), face.double(), training=True), sd, loss_of)
    s = SyncNet_color()
    s.load_state_dict(sd, strict=True)
    s = s.cuda().train()
    a, v = s(mel.cuda(), face.cuda())
    ea = (a.detach().cpu() - a_ref.detach()).abs().max().item()
    ev = (v.detach().cpu() - v_ref.detach()).abs().max().item()
    d = F.cosine_similarity(a, v)
    loss = F.binary_cross_entropy(d.unsqueeze(1), y.cuda())
    loss.backward()
    rep = summarize(param_grad_errors(s, ref))
    rep["emb_max_abs"] = (round(ea, 4), round(ev, 4))     # unit-norm embeddings (entries ~0.05), batch statistics over 4 samples
    report("syncnet_bridge", rep)
    # measured: median 17 %, p90 23 %; the worst tensors are the BatchNorm shifts of the last 1x1 blocks, whose true gradient
    # nearly cancels (78 % of a tiny number); every gradient of this net passes through those 16-sample BatchNorms
    assert ea <= 0.05 and ev <= 0.05 and rep["max"] <= 1.0 and rep["p90"] <= 0.35 and rep["median"] <= 0.25, rep


@pytest.mark.parametrize("B", [16, 48])
def test_sync_loss_gradient_reaches_the_generator_output(B):
    """get_sync_loss (wav2lip_train.py:192-198) with the frozen expert in train mode: dL/dg through the slice + channel
    stack (torch ops) and the expert's input gradient (native), against float64 autograd."""
    from wav2lip_b200.models import SyncNet_color
    sd = relu_active(O.make_state_dict("syncnet", 1, init="default"))
    gen = torch.Generator().manual_seed(11)
    mel = torch.rand((B, 1, 80, 16), generator=gen) * 8 - 4
    g0 = torch.rand((B, 3, 5, 96, 96), generator=gen)
    g64 = g0.double().requires_grad_(True)
    a, v = O.syncnet_forward({k: t.double() for k, t in sd.items() if t.dtype.is_floating_point}, mel.double(),
                             LO.stack_lower_halves(g64), training=True)
    LO.cosine_loss(a, v, torch.ones(B, 1, dtype=torch.float64)).backward()
    s = SyncNet_color()
    s.load_state_dict(sd, strict=True)
    s = s.cuda().train()
    for p in s.parameters():
        p.requires_grad_(False)                                   # wav2lip_train.py:188-189
    gg = g0.cuda().requires_grad_(True)
    half = gg[:, :, :, gg.size(3) // 2:]
    stacked = torch.cat([half[:, :, i] for i in range(5)], dim=1)  # :193-194, torch ops, as the script does
    a2, v2 = s(mel.cuda(), stacked)
    d = F.cosine_similarity(a2, v2)
    F.binary_cross_entropy(d.unsqueeze(1), torch.ones(B, 1).cuda()).backward()
    assert gg.grad is not None and gg.grad[:, :, :, :48].abs().max().item() == 0.0
    assert all(p.grad is None for p in s.parameters())
    e = rel_l2(gg.grad, g64.grad)
    cos = F.cosine_similarity(gg.grad.flatten().double().cpu(), g64.grad.flatten(), dim=0).item()
    report("sync_loss_input_grad", {"B": B, "rel_l2": e, "cos": cos})
    assert e <= (0.35 if B == 16 else 0.2) and cos >= 0.93, {"B": B, "rel_l2": e, "cos": cos, "emb_err": (a2.detach().cpu() - a.detach()).abs().max().item(),
                       "|got|": gg.grad.norm().item(), "|ref|": g64.grad.norm().item()}


def test_disc_training_through_autograd_bridge():
    """hq_wav2lip_train.py:245-253: BCE(disc(gt), 1) + BCE(disc(g.detach()), 0), two backward() calls accumulating."""
    from wav2lip_b200.models import Wav2Lip_disc_qual
    sd = O.make_state_dict("disc", 3, init="default")
    real = O.make_disc_inputs(2, 5, 0)
    fake = O.make_disc_inputs(2, 5, 1)

    def loss_of(_):
        return None

    leaves = {k: v.double().clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point}
    pr = O.disc_forward(leaves, real.double())
    pf = O.disc_forward(leaves, fake.double())
    lref = F.binary_cross_entropy(pr, torch.ones_like(pr)) + F.binary_cross_entropy(pf, torch.zeros_like(pf))
    names = list(leaves.keys())
    ref = dict(zip(names, torch.autograd.grad(lref, [leaves[k] for k in names])))
    d = Wav2Lip_disc_qual()
    d.load_state_dict(sd, strict=True)
    d = d.cuda().train()
    p1 = d(real.cuda())
    F.binary_cross_entropy(p1, torch.ones_like(p1)).backward()
    p2 = d(fake.cuda())
    F.binary_cross_entropy(p2, torch.zeros_like(p2)).backward()
    rep = summarize({n: rel_l2(p.grad, ref[n]) for n, p in d.named_parameters()})
    rep["prob_max_abs"] = round((p1.detach().cpu() - pr.detach()).abs().max().item(), 5)
    report("disc_bridge", rep)
    # LeakyReLU slope flips of the pre-activations within bf16 rounding of zero (no BatchNorm to amplify them)
    assert rep["prob_max_abs"] <= 5e-3 and rep["max"] <= 0.35 and rep["median"] <= 0.25, rep


def test_fused_train_step_against_the_reference_golden(golden_dir):
    """Two iterations of wav2lip_train.py:210-231 as native calls (B=2, T=5, syncnet_wt 0.03, lr 1e-4) against the REAL
    reference + torch.optim.Adam (tests/golden/train.npz)."""
    from wav2lip_b200.models import SyncNet_color, Wav2Lip
    from wav2lip_b200.training import Wav2LipTrainStep
    gold = np.load(os.path.join(golden_dir, "train.npz"))
    gen = Wav2Lip()
    gen.load_state_dict(O.make_state_dict("generator", 0, init="default"), strict=True)
    gen = gen.cuda().train()
    expert = SyncNet_color()
    expert.load_state_dict(O.make_state_dict("syncnet", 1, init="default"), strict=True)
    expert = expert.cuda().train()
    before = {k: v.clone() for k, v in gen.state_dict().items()}
    x, indiv_mels, mel, gt = _train_inputs(2, seed=7)
    step = Wav2LipTrainStep(gen, expert, lr=1e-4, syncnet_wt=0.03)
    for it in range(2):
        losses = step(x.cuda(), indiv_mels.cuda(), mel.cuda(), gt.cuda()).cpu().numpy()   # [sync, l1, 0, total]
        ref = gold[f"gen{it}_losses"]                                                   # [loss, sync, l1]
        assert abs(losses[1] - ref[2]) <= 2e-3 * ref[2], (it, losses, ref)              # L1: mean over 276 k values
        # sync: -log cos of two embeddings whose BatchNorm statistics come from B = 2 samples (zhat = +-1 at the 1x1
        # layers): bf16 moves it by several per cent (0.762 vs 0.818 measured at step 0)
        assert abs(losses[0] - ref[1]) <= 0.12 * ref[1], (it, losses, ref)
        assert abs(losses[3] - (0.03 * losses[0] + 0.97 * losses[1])) <= 1e-6, losses      # the weighted sum itself
        assert abs(losses[3] - ref[0]) <= 1e-2 * ref[0], (it, losses, ref)
        gfp = fp3(step.last_output(2, 5))
        assert abs(gfp[1] - gold[f"gen{it}_g_fp"][1]) <= 2e-3 * gold[f"gen{it}_g_fp"][1]
        if it == 0:
            # (the head's gradient is NOT compared with the golden here: with syncnet_wt = 0.03 it is dominated 100:1 by the
            #  sync term, which runs through the expert's BatchNorm over B = 2 samples — zhat = +-1 — and is not reproducible
            #  in 16-bit arithmetic; test_fused_train_step_l1_only_against_the_oracle pins it on the well-posed L1 path)
            ref_hw = gold["gen0_grad_head_w"]
            # first Adam step: every weight with a non-negligible gradient moves by lr
            after = gen.state_dict()
            k = "output_block.1.weight"
            dlt = (after[k] - before[k]).abs().flatten().cpu()
            big = torch.from_numpy(np.abs(ref_hw) > 1e-6)
            assert torch.allclose(dlt[big], torch.full_like(dlt[big], 1e-4), rtol=2e-2)
            moved = sum(int(((after[n] - before[n]).abs() > 0.5e-4).sum()) for n in after if n.endswith("conv_block.0.weight"))
            total = sum(after[n].numel() for n in after if n.endswith("conv_block.0.weight"))
            report("fused_golden_step0", {"losses": losses.tolist(), "ref_losses": ref.tolist(), "moved_frac": moved / total})
            assert moved >= 0.9 * total, moved / total
        # parameters and BatchNorm buffers after the step: abs-sum fingerprints of all 352 tensors
        names = list(gold[f"gen{it}_sd_names"])
        sd_now = gen.state_dict()
        assert names == list(sd_now.keys())
        got = np.stack([fp3(sd_now[n]) for n in names])
        ref_fp = gold[f"gen{it}_sd_fp"]
        is_stat = np.array(["running" in n for n in names])
        is_cnt = np.array(["num_batches" in n for n in names])
        rel = np.abs(got[:, 1] - ref_fp[:, 1]) / np.maximum(ref_fp[:, 1], 1e-12)
        assert np.all(rel[is_cnt] == 0)
        assert np.all(rel[is_stat & ~is_cnt] <= 6e-2), rel[is_stat & ~is_cnt].max()     # measured up to 3.2e-2 (running_var of the 10-sample 1x1 layers)
        # Adam moves every element by +-lr; where the gradient's sign is noise (deep layers, bf16) the sign differs from the
        # reference's: abs-sum fingerprints of small tensors (16-element BatchNorm shifts) then differ by a few 1e-3
        assert np.all(rel[~is_stat & ~is_cnt] <= 6e-3), rel[~is_stat & ~is_cnt].max()
        # the frozen expert ran in train mode (the scripts' quirk): its running averages moved, its weights did not
        ex = np.stack([fp3(v) for k, v in expert.state_dict().items() if "running" in k or "num_batches" in k])
        rel_e = np.abs(ex[:, 1] - gold[f"gen{it}_expert_buf_fp"][:, 1]) / np.maximum(gold[f"gen{it}_expert_buf_fp"][:, 1], 1e-12)
        assert rel_e.max() <= 6e-2, rel_e.max()        # same bar as the generator's own 10-sample statistics above (measured 1.8e-2 .. 3.2e-2 across kernel builds)


def test_fused_train_step_l1_only_against_the_oracle():
    """syncnet_wt = 0 (wav2lip_train.py:222-225 takes that branch until the sync loss is switched on, :286-288): loss = L1.
    Against oracle/train_oracle.py (fp32 CPU, pinned to the real reference by tests/test_train_oracle.py): loss, output, the
    gradients of the LAST layers (head, output block, decoder stage 6 — a handful of blocks from the loss, before the
    amplification sets in), Adam's first update."""
    from oracle import train_oracle as T
    from wav2lip_b200.models import Wav2Lip
    from wav2lip_b200.training import Wav2LipTrainStep
    gen_sd = O.make_state_dict("generator", 0, init="default")
    x, indiv_mels, mel, gt = _train_inputs(4, seed=9)
    gen = Wav2Lip()
    gen.load_state_dict(gen_sd, strict=True)
    gen = gen.cuda().train()
    ref_sd = {k: v.clone() for k, v in gen_sd.items()}
    r = T.wav2lip_train_step(ref_sd, {}, x, indiv_mels, mel, gt, syncnet_wt=0.0, lr=1e-4)
    step = Wav2LipTrainStep(gen, None, lr=1e-4, syncnet_wt=0.0)
    losses = step(x.cuda(), indiv_mels.cuda(), mel.cuda(), gt.cuda()).cpu().numpy()
    rep = {"l1": float(losses[1]), "l1_ref": float(r["l1"]), "g_max_abs": (step.last_output(4, 5).cpu() - r["g"]).abs().max().item()}
    for k in ("output_block.1.weight", "output_block.1.bias", "output_block.0.conv_block.0.weight", "output_block.0.conv_block.1.weight",
              "face_decoder_blocks.6.2.conv_block.0.weight", "face_decoder_blocks.6.1.conv_block.0.weight",
              "face_decoder_blocks.6.0.conv_block.0.weight", "face_decoder_blocks.5.2.conv_block.0.weight"):
        rep[k] = round(rel_l2(step.b.grads[k], r["grads"][k]), 4)
    report("fused_l1_only", rep)
    assert abs(rep["l1"] - rep["l1_ref"]) <= 2e-3 * rep["l1_ref"], rep
    assert losses[0] == 0.0 and abs(losses[3] - losses[1]) <= 1e-7
    assert rep["output_block.1.weight"] <= 0.1 and rep["output_block.1.bias"] <= 0.1, rep
    assert rep["output_block.0.conv_block.0.weight"] <= 0.25 and rep["face_decoder_blocks.6.2.conv_block.0.weight"] <= 0.4, rep
    # parameters moved by Adam exactly as the oracle's (|delta| = lr wherever the gradient is not negligible)
    after = gen.state_dict()
    k = "output_block.1.weight"
    dlt = (after[k].cpu() - ref_sd[k]).abs().flatten()                   # 0 where the update has the oracle's sign, 2 lr where not
    assert (dlt <= 2.5e-5).float().mean().item() >= 0.9, (dlt <= 2.5e-5).float().mean().item()


def test_hq_step_through_the_autograd_bridge_against_the_reference_golden(golden_dir):
    """Two iterations of hq_wav2lip_train.py:213-255 written exactly as the script writes them — mirrors in train mode,
    `loss.backward()`, two `torch.optim.Adam(betas=(0.5, 0.999))` — against the REAL reference's step (tests/golden/train.npz,
    keys hq0_* / hq1_*).  This is the composite the separate bridge tests do not cover: the perceptual loss back-propagates
    THROUGH the discriminator into the generator (input gradient of the disc plan), the discriminator's own gradients from
    that pass are discarded by `disc_optimizer.zero_grad()`, then real + fake accumulate into one disc step."""
    from torch import nn, optim
    from wav2lip_b200.models import SyncNet_color, Wav2Lip, Wav2Lip_disc_qual
    logloss, recon_loss = nn.BCELoss(), nn.L1Loss()                 # hq_wav2lip_train.py:170, :191

    def cosine_loss(a, v, y):                                       # :171-175
        return logloss(F.cosine_similarity(a, v).unsqueeze(1), y)

    def get_sync_loss(mel, g):                                      # :181-189 (the expert stays in train mode: never .eval()'d)
        g = g[:, :, :, g.size(3) // 2:]
        g = torch.cat([g[:, :, i] for i in range(T)], dim=1)
        a, v = syncnet(mel, g)
        return cosine_loss(a, v, torch.ones(g.size(0), 1, device=g.device))
    gold = np.load(os.path.join(golden_dir, "train.npz"))
    SYNCNET_WT, DISC_WT, T = 0.03, 0.07, 5
    model = Wav2Lip()
    model.load_state_dict(O.make_state_dict("generator", 0, init="default"), strict=True)
    disc = Wav2Lip_disc_qual()
    disc.load_state_dict(O.make_state_dict("disc", 3, init="default"), strict=True)
    syncnet = SyncNet_color()
    syncnet.load_state_dict(O.make_state_dict("syncnet", 1, init="default"), strict=True)
    model, disc, syncnet = model.cuda(), disc.cuda(), syncnet.cuda()
    for p in syncnet.parameters():
        p.requires_grad = False
    optimizer = optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4, betas=(0.5, 0.999))
    disc_optimizer = optim.Adam([p for p in disc.parameters() if p.requires_grad], lr=1e-4, betas=(0.5, 0.999))
    x, indiv_mels, mel, gt = (t.cuda() for t in _train_inputs(2, seed=8))
    for it in range(2):
        disc.train(); model.train()
        optimizer.zero_grad(); disc_optimizer.zero_grad()
        g = model(indiv_mels, x)
        sync_loss = get_sync_loss(mel, g)
        pred = disc(g)
        perceptual_loss = F.binary_cross_entropy(pred, torch.ones((g.size(0) * T, 1), device=g.device))
        l1loss = recon_loss(g, gt)
        loss = SYNCNET_WT * sync_loss + DISC_WT * perceptual_loss + (1. - SYNCNET_WT - DISC_WT) * l1loss
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
        gnorm = {n: p.grad.norm().item() for n, p in model.named_parameters()}
        optimizer.step()
        disc_optimizer.zero_grad()
        pr = disc(gt)
        disc_real_loss = F.binary_cross_entropy(pr, torch.ones((len(pr), 1), device=g.device))
        disc_real_loss.backward()
        pf = disc(g.detach())
        disc_fake_loss = F.binary_cross_entropy(pf, torch.zeros((len(pf), 1), device=g.device))
        disc_fake_loss.backward()
        dgot = {n: p.grad.detach().clone() for n, p in disc.named_parameters()}
        disc_optimizer.step()
        got = np.array([loss.item(), sync_loss.item(), perceptual_loss.item(), l1loss.item(), disc_real_loss.item(), disc_fake_loss.item()])
        ref = gold[f"hq{it}_losses"]
        rel = np.abs(got - ref) / np.abs(ref)
        # discriminator gradients (no BatchNorm: LeakyReLU slope flips only): abs-sum fingerprints per tensor
        dnames = list(gold[f"hq{it}_disc_grad_names"])
        assert dnames == list(dgot.keys())
        dfp = np.stack([fp3(dgot[n]) for n in dnames])
        drel = np.abs(dfp[:, 1] - gold[f"hq{it}_disc_grad_fp"][:, 1]) / np.maximum(gold[f"hq{it}_disc_grad_fp"][:, 1], 1e-12)
        # the generator's gradient norms against the reference's abs-sum fingerprints would need the raw tensors; what the
        # golden allows is the head: its gradient is a 16-bit-stable quantity only on the L1 path (see the test above), so here
        # only the post-step fingerprints are asserted
        gsd = np.stack([fp3(v) for v in model.state_dict().values()])
        dsd = np.stack([fp3(v) for v in disc.state_dict().values()])
        gnames = list(model.state_dict().keys())
        is_stat = np.array([("running" in n) or ("num_batches" in n) for n in gnames])
        grel = np.abs(gsd[:, 1] - gold[f"hq{it}_gen_sd_fp"][:, 1]) / np.maximum(gold[f"hq{it}_gen_sd_fp"][:, 1], 1e-12)
        dsrel = np.abs(dsd[:, 1] - gold[f"hq{it}_disc_sd_fp"][:, 1]) / np.maximum(gold[f"hq{it}_disc_sd_fp"][:, 1], 1e-12)
        report(f"hq_bridge_step{it}", {"losses": got.tolist(), "ref": ref.tolist(), "loss_rel": rel.tolist(),
                                       "disc_grad_fp_rel_max": float(drel.max()), "disc_grad_fp_rel_median": float(np.median(drel)),
                                       "gen_sd_rel_max": float(grel[~is_stat].max()), "gen_stat_rel_max": float(grel[is_stat].max()),
                                       "disc_sd_rel_max": float(dsrel.max()), "min_gen_grad_norm": min(gnorm.values())})
        assert rel[3] <= 2e-3, (it, got, ref)                      # L1
        assert rel[1] <= 0.12, (it, got, ref)                      # sync (B = 2 BatchNorm in the expert, see above)
        assert rel[2] <= 2e-2 and rel[4] <= 2e-2 and rel[5] <= 2e-2, (it, got, ref)   # the three BCE terms of the disc
        assert rel[0] <= 1e-2, (it, got, ref)
        assert np.median(drel) <= 0.1 and drel.max() <= 0.5, (it, drel)
        assert grel[~is_stat].max() <= 6e-3 and dsrel.max() <= 6e-3, (it, grel[~is_stat].max(), dsrel.max())
        assert grel[is_stat].max() <= 6e-2, (it, grel[is_stat].max())


def test_perceptual_loss_gradient_reaches_the_generator_output():
    """wav2lip.py:163-174 (perceptual_forward) inside hq_wav2lip_train.py:233-242: d BCE(disc(g), 1) / dg through the
    bridge's input gradient of the discriminator plan — lower half only (wav2lip.py:152-153), t-major flatten (:155-161) —
    against float64 autograd through the oracle.  No BatchNorm in this network: only LeakyReLU slope flips."""
    from wav2lip_b200.models import Wav2Lip_disc_qual
    sd = O.make_state_dict("disc", 3, init="default")
    frames = O.make_disc_inputs(3, 5, 2)
    g64 = frames.double().clone().requires_grad_(True)
    pr = O.disc_forward({k: v.double() for k, v in sd.items() if v.dtype.is_floating_point}, g64)
    F.binary_cross_entropy(pr, torch.ones_like(pr)).backward()
    d = Wav2Lip_disc_qual()
    d.load_state_dict(sd, strict=True)
    d = d.cuda().train()
    for p in d.parameters():
        p.requires_grad = False                      # only the input gradient is wanted (the plan without wgrad)
    gg = frames.cuda().requires_grad_(True)
    p1 = d(gg)
    F.binary_cross_entropy(p1, torch.ones_like(p1)).backward()
    assert gg.grad is not None and gg.grad.shape == frames.shape
    assert float(gg.grad[:, :, :, :48].abs().max()) == 0.0          # the upper half never reaches the discriminator
    e = rel_l2(gg.grad, g64.grad)
    cos = F.cosine_similarity(gg.grad.double().cpu().flatten(), g64.grad.flatten(), dim=0).item()
    report("perceptual_input_grad", {"rel_l2": e, "cos": cos})
    assert e <= 0.2 and cos >= 0.98, (e, cos)
