"""Soak: alternating batch sizes / entry points for a while; checks determinism and memory stability. Test infrastructure."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wav2lip_b200.models import SyncNet_color, Wav2Lip, Wav2Lip_disc_qual
torch.manual_seed(0)
with torch.no_grad():
    g = Wav2Lip().cuda().eval(); s = SyncNet_color().cuda().eval(); d = Wav2Lip_disc_qual().cuda().eval()
    sizes = [128, 7, 64, 1, 33, 128, 200, 16]
    ins = {n: ((torch.rand(n, 1, 80, 16) * 8 - 4).cuda(), torch.rand(n, 6, 96, 96).cuda()) for n in set(sizes)}
    ref = {}
    t0 = time.time(); it = 0
    while time.time() - t0 < 40:
        for n in sizes:
            y = g(*ins[n])
            if n not in ref: ref[n] = y.clone()
            assert torch.equal(y, ref[n]), (it, n)
        if it % 5 == 0:   # the asynchronous host pipeline with changing batch sizes (staging buffers grow, plans get evicted)
            hb = [(ins[n][0].cpu(), ins[n][1].cpu()) for n in (7, 200, 33, 128)]
            for (mel_c, face_c), yo in zip(hb, g.infer_stream(iter(hb))):
                assert torch.equal(yo, ref[mel_c.shape[0]].cpu()), (it, mel_c.shape[0])
        a, v = s(ins[16][0], torch.rand(16, 15, 48, 96).cuda())
        p = d(torch.rand(4, 3, 5, 96, 96).cuda())
        assert torch.isfinite(a).all() and torch.isfinite(p).all()
        it += 1
        if it % 20 == 0:
            torch.cuda.synchronize()
            print(it, "iters;", torch.cuda.memory_allocated() // 2**20, "MiB torch;", g._w2l_ctx.device_bytes() // 2**20, "MiB w2l generator ctx", flush=True)
    torch.cuda.synchronize()
print("soak ok:", it, "iterations,", it * len(sizes), "generator forwards")
