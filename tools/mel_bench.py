"""Runs the mel kernel on 10 k and 1 M frames (for ncu / timing). Test infrastructure."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wav2lip_b200 import audio
for nfr in (10000, 1000000):
    wav = (0.1 * torch.randn((nfr - 1) * 200, device="cuda")).float()
    for _ in range(3):
        m = audio.melspectrogram(wav)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m = audio.melspectrogram(wav)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{nfr} frames: {ms:.3f} ms, {nfr/ms/1e3:.1f} M frames/s, algorithmic {nfr*1120/ms/1e6:.1f} GB/s")
