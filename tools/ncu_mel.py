"""audio.melspectrogram on 1 M frames inside a cudaProfilerStart/Stop window, for `ncu --profile-from-start off`.
Measurement infrastructure."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wav2lip_b200 import audio

torch.manual_seed(0)
wav = (0.1 * torch.randn(200 * 1000000 - 200)).cuda()
for _ in range(2):
    m = audio.melspectrogram(wav)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m = audio.melspectrogram(wav)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", tuple(m.shape))
