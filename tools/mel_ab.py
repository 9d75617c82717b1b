"""A/B timing of the two mel kernels (mel_kernel_v2 = default, mel_kernel with W2L_DISABLE_MELV2=1) on 1 M frames, CUDA events.
Measurement infrastructure."""
import ctypes as C
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wav2lip_b200 import _lib

wav = (0.1 * torch.randn(200 * 1000000 - 200)).cuda()
out = torch.empty((80, 1000000), device="cuda")
for name, flag in (("mel_kernel_v2 (registers)", None), ("mel_kernel (shared-memory Stockham)", "1")):
    if flag is None:
        os.environ.pop("W2L_DISABLE_MELV2", None)
    else:
        os.environ["W2L_DISABLE_MELV2"] = flag
    ctx = _lib.Context(0)
    run = lambda: _lib.check(ctx.lib.w2l_melspectrogram(ctx.h, C.c_void_p(wav.data_ptr()), wav.numel(), C.c_void_p(out.data_ptr()), None))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.3f} ms per 1M frames = {1e3 / ms:.1f} M frames/s = {1120e6 / ms / 1e6:.1f} GB/s algorithmic")
