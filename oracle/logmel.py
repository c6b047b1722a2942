"""oracle/logmel.py — numpy restatement of the Whisper feature extractor (stage 1).  TEST INFRASTRUCTURE ONLY.

Follows HF/models/whisper/feature_extraction_whisper.py (transformers 5.5.0):
  mel_filters            :95-103  (audio_utils.mel_filter_bank, HF/audio_utils.py:453-545, slaney scale + norm)
  _torch_extract_fbank_features :135-164
  padding / attention mask      :296-337
The arithmetic is done in float64 and rounded once to float32: it is the infinitely-precise version of what the
reference computes in float32 (torch.stft -> pocketfft/MKL); tests compare with a stated tolerance.
"""
from __future__ import annotations

import numpy as np

SR, N_FFT, HOP, N_SAMPLES, N_FRAMES = 16000, 400, 160, 480000, 3000


def hertz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    min_log_hertz, min_log_mel, logstep = 1000.0, 15.0, 27.0 / np.log(6.4)
    mels = 3.0 * f / 200.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_region = f >= min_log_hertz
        mels = np.where(log_region, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hertz) * logstep, mels)
    return mels


def mel_to_hertz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    min_log_hertz, min_log_mel, logstep = 1000.0, 15.0, np.log(6.4) / 27.0
    f = 200.0 * m / 3.0
    log_region = m >= min_log_mel
    return np.where(log_region, min_log_hertz * np.exp(logstep * (m - min_log_mel)), f)


def mel_filter_bank(n_mels: int, n_freq: int = 201, fmin: float = 0.0, fmax: float = 8000.0, sr: int = SR) -> np.ndarray:
    """[n_freq, n_mels] float64 — HF audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney")."""
    mel_min, mel_max = hertz_to_mel_slaney(fmin), hertz_to_mel_slaney(fmax)
    mel_freqs = np.linspace(mel_min, mel_max, n_mels + 2)
    filter_freqs = mel_to_hertz_slaney(mel_freqs)
    fft_freqs = np.linspace(0, sr // 2, n_freq)
    filter_diff = np.diff(filter_freqs)
    slopes = np.expand_dims(filter_freqs, 0) - np.expand_dims(fft_freqs, 1)
    down = -slopes[:, :-2] / filter_diff[:-1]
    up = slopes[:, 2:] / filter_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (filter_freqs[2:n_mels + 2] - filter_freqs[:n_mels])
    return fb * np.expand_dims(enorm, 0)


def pad_or_trim(wave: np.ndarray) -> np.ndarray:
    """feature_extraction_whisper.py:296-303 — zero-pad / truncate to 480000 samples."""
    w = np.zeros(N_SAMPLES, np.float32)
    n = min(len(wave), N_SAMPLES)
    w[:n] = wave[:n]
    return w


def log_mel(wave: np.ndarray, mel_filters_T: np.ndarray) -> np.ndarray:
    """wave float32 [480000]; mel_filters_T float32 [n_mels, 201] -> float32 [n_mels, 3000]."""
    x = wave.astype(np.float64)
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    n = np.arange(N_FFT)
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)           # torch.hann_window(400): periodic
    idx = np.arange(N_FRAMES)[:, None] * HOP + n[None, :]        # frame 3000 is dropped (:150)
    spec = np.fft.rfft(xp[idx] * win[None, :], axis=1)            # [3000, 201]
    power = (spec.real ** 2 + spec.imag ** 2).T                   # [201, 3000]
    mel = mel_filters_T.astype(np.float64) @ power
    log_spec = np.log10(np.maximum(mel, 1e-10))
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    return ((log_spec + 4.0) / 4.0).astype(np.float32)


def num_frames(n_valid: int) -> int:
    """attention_mask[:, ::160].sum(-1) (:328-337)."""
    n_valid = max(0, min(int(n_valid), N_SAMPLES))
    return (n_valid + HOP - 1) // HOP
