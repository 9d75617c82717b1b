"""Per-launch CUDA-event profile of the SyncNet / disc plans. Test infrastructure."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wav2lip_b200 import _lib
from wav2lip_b200.models import SyncNet_color, Wav2Lip_disc_qual
torch.manual_seed(0)
with torch.no_grad():
    s = SyncNet_color().cuda().eval()
    mel, face = (torch.rand(256, 1, 80, 16) * 8 - 4).cuda(), torch.rand(256, 15, 48, 96).cuda()
    for _ in range(3): s(mel, face)
    prof = s._w2l_ctx.profile_plan(_lib.NET_SYNCNET, iters=5)
    tot = sum(m for _, m, _ in prof)
    print(f"syncnet B=256: sum {tot:.3f} ms")
    for n, m, f in prof: print(f"   {n:38s} {m*1e3:8.1f} us {f/m/1e9 if m>0 else 0:8.1f} TFLOP/s {100*m/tot:5.1f}%")
    d = Wav2Lip_disc_qual().cuda().eval()
    fr = torch.rand(256, 3, 5, 96, 96).cuda()
    for _ in range(3): d(fr)
    prof = d._w2l_ctx.profile_plan(_lib.NET_DISC, iters=5)
    tot = sum(m for _, m, _ in prof)
    print(f"disc B=256 T=5: sum {tot:.3f} ms")
    for n, m, f in prof: print(f"   {n:38s} {m*1e3:8.1f} us {f/m/1e9 if m>0 else 0:8.1f} TFLOP/s {100*m/tot:5.1f}%")
