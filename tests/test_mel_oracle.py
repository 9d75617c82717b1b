"""oracle/mel_oracle.py against its committed vectors and against two independent
implementations installed in the image (torch.stft, torchaudio melscale_fbanks).
The librosa boundary itself is unpinned (see oracle/mel_oracle.py header)."""
import os

import numpy as np
import pytest
import torch

from oracle import mel_oracle as M


def test_mel_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "mel.npz"))
    np.testing.assert_array_equal(M.mel_basis(), g["mel_basis"])
    for kind in ("noise", "sweep", "mix"):
        wav = M.make_wav(48000 + 137, seed=7, kind=kind)
        assert abs(np.abs(wav.astype(np.float64)).sum() - float(g["wav_checksum_" + kind])) < 1e-6
        mel = M.melspectrogram(wav)
        assert mel.dtype == np.float32 and mel.shape == (80, 1 + (48000 + 137) // 200)
        np.testing.assert_allclose(mel, g["mel_" + kind], atol=2e-5, rtol=0)


def test_mel_basis_properties():
    b = M.mel_basis()
    assert b.shape == (80, 401) and b.dtype == np.float32
    assert int((b != 0).sum()) == 739
    assert int((b != 0).sum(1).max()) == 27
    cols = np.nonzero(b.any(0))[0]
    assert cols[0] == 3 and cols[-1] == 379


def test_mel_basis_vs_torchaudio():
    ta = pytest.importorskip("torchaudio")
    fb = ta.functional.melscale_fbanks(401, 55.0, 7600.0, 80, 16000, norm="slaney", mel_scale="slaney").T.numpy()
    np.testing.assert_allclose(M.mel_basis(), fb, atol=2e-7, rtol=0)


@pytest.mark.parametrize("kind", ["noise", "sweep", "mix"])
def test_mel_vs_torch_stft(kind):
    wav = M.make_wav(16000 * 2 + 55, seed=3, kind=kind)
    mel = M.melspectrogram(wav)
    y = torch.from_numpy(M.preemphasis(wav))
    D = torch.stft(y, 800, 200, 800, window=torch.hann_window(800, periodic=True, dtype=torch.float64),
                   center=True, pad_mode="reflect", return_complex=True)
    S = torch.from_numpy(M.mel_basis()).double() @ D.abs()
    S = 20 * torch.log10(torch.clamp(S, min=1e-5)) - 20
    S = torch.clamp(8 * ((S + 100) / 100) - 4, -4, 4)
    np.testing.assert_allclose(mel, S.numpy(), atol=5e-5, rtol=0)


def test_mel_edges():
    # silence hits the 1e-5 floor -> exactly -4.0 everywhere
    mel = M.melspectrogram(np.zeros(4000, dtype=np.float32))
    assert mel.shape == (80, 21) and np.all(mel == -4.0)
    assert M.melspectrogram(np.ones(401, dtype=np.float32)).shape == (80, 3)
    # librosa 0.7.0 does not check the length: shorter clips are reflected more than once by np.pad
    assert M.melspectrogram(np.ones(400, dtype=np.float32)).shape == (80, 3)
    assert M.melspectrogram(np.ones(57, dtype=np.float32)).shape == (80, 1)
    with pytest.raises(ValueError):
        M.melspectrogram(np.ones(1, dtype=np.float32))
    # preemphasis: zero initial state
    y = M.preemphasis(np.array([1.0, 1.0, 1.0], dtype=np.float32))
    np.testing.assert_allclose(y, [1.0, 0.03, 0.03], atol=1e-12)
