"""GPU bring-up diagnostics: runs the kernels against the oracle and PRINTS error structure instead of
asserting, so that one gpurun round-trip tells as much as possible.  Test infrastructure, not product.

    python tools/gpu_diag.py [--quick] > gpurun_out/diag.txt
"""
import argparse
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mel_oracle as M  # noqa: E402
from oracle import w2l_oracle as O  # noqa: E402
from wav2lip_b200 import _lib, audio  # noqa: E402
from wav2lip_b200.models import SyncNet_color, Wav2Lip, Wav2Lip_disc_qual  # noqa: E402
from wav2lip_b200.models.conv import Conv2d, Conv2dTranspose, nonorm_Conv2d  # noqa: E402


def err_report(name, got, ref, tol_rel=3e-3):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    d = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    bad = d > tol_rel * scale
    line = (f"{name}: shape={tuple(ref.shape)} max|ref|={scale:.4g} max|err|={d.max().item():.4g} "
            f"rel={d.max().item() / scale:.3g} mean|err|={d.mean().item():.3g} bad={bad.float().mean().item():.4f}")
    nan = torch.isnan(got).sum().item()
    if nan:
        line += f" NaN={nan}"
    print(line, flush=True)
    ok = bad.sum().item() == 0 and nan == 0
    if not ok and ref.dim() == 4:
        # structure of the error: by channel, by row, by column, by sample
        for dim, label in ((0, "n"), (1, "c"), (2, "y"), (3, "x")):
            other = tuple(i for i in range(4) if i != dim)
            frac = bad.float().mean(dim=other)
            idx = torch.nonzero(frac > 0).flatten().tolist()
            print(f"    bad along {label}: {len(idx)}/{ref.shape[dim]} first={idx[:24]}")
        i = torch.nonzero(bad)[:6]
        for r in i:
            r = tuple(r.tolist())
            print(f"    at {r}: got {got[r].item():.5g} ref {ref[r].item():.5g}")
    return ok


def block_case(name, kind, cin, cout, k, s, p, op, res, N, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    row = (kind, cin, cout, k, s, p, op, res)
    sd = O._block_tensors("b", row, g, 1.0)
    x = torch.rand((N, cin, H, W), generator=g) * 2 - 0.5
    with torch.no_grad():
        ref = O.block_forward(x, sd, "b", row)
    cls = {"c": Conv2d, "t": Conv2dTranspose, "n": nonorm_Conv2d}[kind]
    if kind == "t":
        m = cls(cin, cout, k, s, p, op)
    else:
        m = cls(cin, cout, k, s, p, residual=res)
    m.load_state_dict({kk[2:]: v for kk, v in sd.items()}, strict=True)
    m = m.cuda().eval()
    try:
        with torch.no_grad():
            y = m(x.cuda())
        torch.cuda.synchronize()
        return err_report(f"block[{name}] {kind} {cin}->{cout} k{k} s{s} p{p} N{N} {H}x{W} res={res}", y, ref)
    except Exception as e:  # noqa: BLE001
        print(f"block[{name}] FAILED: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--skip-nets", action="store_true")
    args = ap.parse_args()
    print("device:", torch.cuda.get_device_name(0), "| lib:", _lib.lib_path(), flush=True)
    results = {}

    # ---------------- mel ----------------
    try:
        for kind in ("noise", "sweep", "mix"):
            wav = M.make_wav(48000 + 137, seed=7, kind=kind)
            ref = M.melspectrogram(wav)
            got = audio.melspectrogram(wav)
            d = np.abs(got - ref)
            print(f"mel[{kind}]: shape={got.shape} max|err|={d.max():.3g} mean={d.mean():.3g} "
                  f"n>1e-4={(d > 1e-4).sum()} floor_frac={np.mean(ref == -4.0):.3f}", flush=True)
            results["mel_" + kind] = bool(d.max() <= 1e-4)
            if d.max() > 1e-4:
                bad = np.argwhere(d > 1e-4)[:8]
                for b in bad:
                    print(f"    at {tuple(b)}: got {got[tuple(b)]:.6f} ref {ref[tuple(b)]:.6f}")
    except Exception as e:  # noqa: BLE001
        print("mel FAILED:", type(e).__name__, e, flush=True)
        traceback.print_exc()

    # ---------------- conv blocks, simplest first ----------------
    cases = [
        ("3x3-64", "c", 64, 64, 3, 1, 1, 0, False, 2, 24, 24),
        ("3x3-64-res", "c", 64, 64, 3, 1, 1, 0, True, 2, 24, 24),
        ("1x1-512", "c", 512, 512, 1, 1, 0, 0, False, 5, 1, 1),
        ("3x3-32(BK32)", "c", 32, 32, 3, 1, 1, 0, True, 2, 48, 48),
        ("3x3-16in(BK16)", "c", 16, 32, 3, 2, 1, 0, False, 2, 96, 96),
        ("7x7-6in", "c", 6, 16, 7, 1, 3, 0, False, 2, 96, 96),
        ("3x3-128", "c", 128, 128, 3, 1, 1, 0, True, 3, 12, 12),
        ("3x3-256", "c", 256, 256, 3, 1, 1, 0, True, 3, 6, 6),
        ("3x3-384", "c", 384, 384, 3, 1, 1, 0, True, 2, 12, 12),
        ("3x3-512", "c", 512, 512, 3, 1, 1, 0, True, 3, 3, 3),
        ("s2", "c", 64, 128, 3, 2, 1, 0, False, 2, 24, 24),
        ("s(3,1)", "c", 32, 64, 3, (3, 1), 1, 0, False, 2, 80, 16),
        ("s3", "c", 64, 128, 3, 3, 1, 0, False, 2, 27, 16),
        ("s(3,2)", "c", 128, 256, 3, (3, 2), 1, 0, False, 2, 9, 6),
        ("p0", "c", 512, 512, 3, 1, 0, 0, False, 3, 3, 3),
        ("audio0", "c", 1, 32, 3, 1, 1, 0, False, 2, 80, 16),
        ("80->32", "c", 80, 32, 3, 1, 1, 0, False, 1, 96, 96),
        ("convT-1x1", "t", 1024, 512, 3, 1, 0, 0, False, 3, 1, 1),
        ("convT-s2", "t", 1024, 512, 3, 2, 1, 1, False, 2, 3, 3),
        ("convT-s2-160", "t", 160, 64, 3, 2, 1, 1, False, 1, 48, 48),
        ("convT-s2-320", "t", 320, 128, 3, 2, 1, 1, False, 1, 24, 24),
        ("sync-k5", "c", 32, 64, 5, (1, 2), 1, 0, False, 2, 48, 96),
        ("sync-7x7", "c", 15, 32, 7, 1, 3, 0, False, 2, 48, 96),
        ("sync-46x47", "c", 64, 64, 3, 1, 1, 0, True, 2, 46, 47),
        ("sync-23x24-s2", "c", 64, 128, 3, 2, 1, 0, False, 2, 46, 47),
        ("disc-7x7", "n", 3, 32, 7, 1, 3, 0, False, 2, 48, 96),
        ("disc-k5s(1,2)", "n", 32, 64, 5, (1, 2), 2, 0, False, 2, 48, 96),
        ("disc-k5", "n", 64, 64, 5, 1, 2, 0, False, 2, 48, 48),
        ("disc-k5s2", "n", 128, 256, 5, 2, 2, 0, False, 2, 24, 24),
    ]
    if args.quick:
        cases = cases[:6]
    for c in cases:
        results["block_" + c[0]] = block_case(*c)

    if not args.skip_nets:
        # ---------------- generator, layer by layer ----------------
        try:
            sd = O.make_state_dict("generator", 0)
            mel, face = O.make_generator_inputs(2, 0)
            taps = {}
            with torch.no_grad():
                ref_logits = O.generator_forward(sd, mel, face, taps, return_logits=True)
            ref = torch.sigmoid(ref_logits)
            g = Wav2Lip()
            g.load_state_dict(sd, strict=True)
            g = g.cuda().eval()
            with torch.no_grad():
                g._ensure(face.cuda())
                g._w2l_ctx.set_debug(True)
                y = g(mel.cuda(), face.cuda())
            torch.cuda.synchronize()
            results["gen4"] = err_report("generator N=2 output (post-sigmoid)", y, ref, tol_rel=1e-3)
            print("   max abs err on output:", (y.cpu() - ref).abs().max().item())
            names = [n for n, _ in O.generator_layers()]
            for i, n in enumerate(names[:-1]):
                try:
                    got = g.debug_layer_output(i)
                    err_report(f"  layer {i:2d} {n}", got, taps[n])
                except Exception as e:  # noqa: BLE001
                    print(f"  layer {i} {n}: {e}")
            mel5, face5 = O.make_generator_inputs(2, seed=1, t=5)
            with torch.no_grad():
                ref5 = O.generator_forward(sd, mel5, face5)
                y5 = g(mel5.cuda(), face5.cuda())
            results["gen5"] = err_report("generator 5-D B=2 T=5", y5.reshape(2, 3, -1, 96), ref5.reshape(2, 3, -1, 96), tol_rel=1e-3)
        except Exception as e:  # noqa: BLE001
            print("generator FAILED:", type(e).__name__, e, flush=True)
            traceback.print_exc()
        # ---------------- syncnet / disc ----------------
        try:
            sd = O.make_state_dict("syncnet", 0)
            mel, face = O.make_syncnet_inputs(3, 0)
            with torch.no_grad():
                a0, v0 = O.syncnet_forward(sd, mel, face)
            s = SyncNet_color()
            s.load_state_dict(sd, strict=True)
            s = s.cuda().eval()
            with torch.no_grad():
                a1, v1 = s(mel.cuda(), face.cuda())
            results["sync_a"] = err_report("syncnet audio emb", a1, a0, tol_rel=1e-2)
            results["sync_v"] = err_report("syncnet face emb", v1, v0, tol_rel=1e-2)
        except Exception as e:  # noqa: BLE001
            print("syncnet FAILED:", type(e).__name__, e, flush=True)
            traceback.print_exc()
        try:
            sd = O.make_state_dict("disc", 0)
            fr = O.make_disc_inputs(2, 5, 0)
            with torch.no_grad():
                p0 = O.disc_forward(sd, fr)
            d = Wav2Lip_disc_qual()
            d.load_state_dict(sd, strict=True)
            d = d.cuda().eval()
            with torch.no_grad():
                p1 = d(fr.cuda())
            results["disc"] = err_report("disc prob", p1, p0, tol_rel=1e-2)
        except Exception as e:  # noqa: BLE001
            print("disc FAILED:", type(e).__name__, e, flush=True)
            traceback.print_exc()

        # ---------------- quick throughput look ----------------
        try:
            sd = O.make_state_dict("generator", 0)
            g = Wav2Lip()
            g.load_state_dict(sd, strict=True)
            g = g.cuda().eval()
            for n in (128,):
                mel, face = O.make_generator_inputs(n, 0)
                mel, face = mel.cuda(), face.cuda()
                with torch.no_grad():
                    for _ in range(3):
                        g(mel, face)
                    torch.cuda.synchronize()
                    t0 = time.time()
                    for _ in range(5):
                        g(mel, face)
                    torch.cuda.synchronize()
                dt = (time.time() - t0) / 5
                print(f"generator N={n}: {dt * 1e3:.2f} ms/forward, {n / dt:.0f} crops/s, "
                      f"{n / dt * 7.934e9 / 1e12:.1f} TFLOP/s", flush=True)
                prof = g._w2l_ctx.profile_plan(_lib.NET_GENERATOR, iters=5)
                tot = sum(ms for _, ms, _ in prof)
                print(f"per-launch profile (sum {tot:.3f} ms):")
                for nm, ms, fl in prof:
                    print(f"   {nm:34s} {ms * 1e3:9.1f} us  {fl / ms / 1e9 if ms > 0 else 0:8.1f} TFLOP/s  {100 * ms / tot:5.1f}%")
        except Exception as e:  # noqa: BLE001
            print("throughput FAILED:", type(e).__name__, e, flush=True)
            traceback.print_exc()

    print("SUMMARY:", {k: v for k, v in results.items()})
    print("ALL_OK" if all(results.values()) else "SOME_FAILED")


if __name__ == "__main__":
    main()
