"""One forward of SyncNet_color (B=256), Wav2Lip_disc_qual (B=256, T=5) and the evaluation-loop losses inside a
cudaProfilerStart/Stop window, for `ncu --profile-from-start off`.  Measurement infrastructure."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wav2lip_b200 import losses
from wav2lip_b200.models import SyncNet_color, Wav2Lip_disc_qual

torch.manual_seed(0)
with torch.no_grad():
    s = SyncNet_color().cuda().eval()
    mel, face = (torch.rand(256, 1, 80, 16) * 8 - 4).cuda(), torch.rand(256, 15, 48, 96).cuda()
    d = Wav2Lip_disc_qual().cuda().eval()
    fr = torch.rand(256, 3, 5, 96, 96).cuda()
    g, gt = torch.rand(128, 3, 5, 96, 96).cuda(), torch.rand(128, 3, 5, 96, 96).cuda()
    for _ in range(2):
        a, v = s(mel, face); d(fr); losses.cosine_loss(a, v); losses.recon_loss(g, gt)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    a, v = s(mel, face)
    d(fr)
    losses.cosine_loss(a, v)
    losses.recon_loss(g, gt)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok")
