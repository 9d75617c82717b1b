"""CPU oracle for `audio.melspectrogram` — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.

PARITY UNPINNED at the librosa boundary: the reference's /root/reference/audio.py:1-2
imports librosa (pinned librosa==0.7.0, /root/reference/requirements.txt:1), which is a
third-party dependency that is neither vendored in the reference nor installed in this
image, and the reference holds no test, fixture or stored spectrogram for this path.
This file therefore restates (a) the reference's own pipeline and constants and (b) the
published librosa-0.7.0 algorithm for the two calls the reference makes into it, and it is
cross-checked in tests/test_mel_oracle.py against two independent implementations that
ARE installed (torch.stft and torchaudio.functional.melscale_fbanks).

Reference call sites restated here:
  melspectrogram      /root/reference/audio.py:45-51
  preemphasis         /root/reference/audio.py:20-23   (scipy.signal.lfilter([1,-k],[1],wav))
  _stft               /root/reference/audio.py:57-61   (librosa.stft(y, n_fft=800, hop_length=200, win_length=800))
  _linear_to_mel      /root/reference/audio.py:92-96
  _build_mel_basis    /root/reference/audio.py:98-101  (librosa.filters.mel(16000, 800, n_mels=80, fmin=55, fmax=7600))
  _amp_to_db          /root/reference/audio.py:103-105
  _normalize          /root/reference/audio.py:110-114 (symmetric, clipping branch)
  constants           /root/reference/hparams.py:33-73

librosa 0.7.0 semantics used (from its documented behaviour):
  stft: window = scipy.signal.get_window('hann', 800, fftbins=True) (periodic Hann, float64);
        center=True -> np.pad(y, 400, mode='reflect'); frame t = y_pad[200 t : 200 t + 800];
        n_frames = 1 + len(y)//200; FFT of the windowed float64 frame, first 401 bins,
        stored as complex64.
  filters.mel: Slaney mel scale (htk=False): linear below 1 kHz with f_sp = 200/3 Hz/mel,
        log above with step ln(6.4)/27; n_mels+2 band edges linearly spaced in mel between
        fmin and fmax; triangular weights from ramps against np.linspace(0, sr/2, 401);
        Slaney area normalisation 2/(f[i+2]-f[i]); float32 result.
Arithmetic dtypes follow NumPy 1.17 value-based casting (the pinned numpy==1.17.1): the
float64 python scalars in _amp_to_db/_normalize do NOT promote the float32 arrays.
"""
from __future__ import annotations

import numpy as np

# hparams.py:33-73
NUM_MELS = 80
N_FFT = 800
HOP = 200
WIN = 800
SR = 16000
PREEMPH = 0.97
MIN_LEVEL_DB = -100.0
REF_LEVEL_DB = 20.0
FMIN = 55.0
FMAX = 7600.0
MAX_ABS = 4.0
N_BINS = 1 + N_FFT // 2


def preemphasis(wav: np.ndarray, k: float = PREEMPH) -> np.ndarray:
    """y[n] = x[n] - k x[n-1], y[0] = x[0]; float64 result, zero initial state (audio.py:20-23)."""
    x = np.asarray(wav, dtype=np.float64)
    y = x.copy()
    y[1:] -= k * x[:-1]
    return y


def hann_periodic(n: int = WIN) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def stft(y: np.ndarray) -> np.ndarray:
    """librosa.stft(y, 800, 200, 800) -> complex64 (401, 1 + len(y)//200)."""
    y = np.asarray(y, dtype=np.float64)
    if y.shape[0] < 2:
        raise ValueError("reflect padding needs at least 2 samples")
    # librosa 0.7.0 pads with np.pad(mode="reflect") and does not check the length: clips shorter than n_fft//2 are
    # reflected more than once (np.pad's own rule), they do not raise
    yp = np.pad(y, N_FFT // 2, mode="reflect")
    n_frames = 1 + (yp.shape[0] - N_FFT) // HOP
    idx = np.arange(N_FFT)[:, None] + HOP * np.arange(n_frames)[None, :]
    frames = yp[idx] * hann_periodic()[:, None]
    return np.fft.rfft(frames, axis=0).astype(np.complex64)


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_basis() -> np.ndarray:
    """librosa.filters.mel(16000, 800, n_mels=80, fmin=55, fmax=7600) -> float32 (80, 401)."""
    weights = np.zeros((NUM_MELS, N_BINS), dtype=np.float32)
    fftfreqs = np.linspace(0, float(SR) / 2, N_BINS, endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(FMIN), _hz_to_mel(FMAX), NUM_MELS + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(NUM_MELS):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:NUM_MELS + 2] - mel_f[:NUM_MELS])
    # `weights *= enorm[:, None]` on a float32 array: product taken in float64, rounded to float32
    return (weights.astype(np.float64) * enorm[:, np.newaxis]).astype(np.float32)


_BASIS = None


def linear_to_mel(mag: np.ndarray) -> np.ndarray:
    global _BASIS
    if _BASIS is None:
        _BASIS = mel_basis()
    return np.dot(_BASIS, mag.astype(np.float32, copy=False))


def amp_to_db(x: np.ndarray) -> np.ndarray:
    min_level = np.float32(np.exp(MIN_LEVEL_DB / 20 * np.log(10)))  # 1e-5
    return np.float32(20) * np.log10(np.maximum(min_level, x.astype(np.float32, copy=False)))


def normalize(S: np.ndarray) -> np.ndarray:
    S = S.astype(np.float32, copy=False)
    v = np.float32(2 * MAX_ABS) * ((S - np.float32(MIN_LEVEL_DB)) / np.float32(-MIN_LEVEL_DB)) - np.float32(MAX_ABS)
    return np.clip(v, np.float32(-MAX_ABS), np.float32(MAX_ABS))


def melspectrogram(wav: np.ndarray) -> np.ndarray:
    """audio.py:45-51 -> float32 (80, 1 + len(wav)//200) in [-4, 4]."""
    D = stft(preemphasis(wav))
    S = amp_to_db(linear_to_mel(np.abs(D))) - np.float32(REF_LEVEL_DB)
    return normalize(S)


def make_wav(n_samples: int, seed: int = 0, kind: str = "noise") -> np.ndarray:
    """Deterministic float32 test signals: noise, sweep (sine sweep + tone), silence-gapped mix."""
    rng = np.random.RandomState(seed)
    t = np.arange(n_samples, dtype=np.float64) / SR
    if kind == "noise":
        x = 0.1 * rng.randn(n_samples)
    elif kind == "sweep":
        f0, f1 = 80.0, 7000.0
        dur = max(t[-1], 1e-3)
        phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / dur)
        x = 0.5 * np.sin(phase) + 0.01 * rng.randn(n_samples)
    elif kind == "mix":
        x = 0.1 * rng.randn(n_samples) + 0.3 * np.sin(2 * np.pi * 440.0 * t)
        seg = n_samples // 4
        x[seg:2 * seg] = 0.0  # silence -> clipped floor -4.0 exactly
    else:
        raise ValueError(kind)
    return x.astype(np.float32)
