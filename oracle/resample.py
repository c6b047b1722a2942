"""oracle/resample.py — numpy restatement of the resampler the reference's pipeline applies to inputs that are not at
16 kHz.  TEST INFRASTRUCTURE ONLY (never imported by the product).

The reference path: AutomaticSpeechRecognitionPipeline.preprocess calls torchaudio.functional.resample(x, sr_in, 16000)
(HF/pipelines/automatic_speech_recognition.py:394-408) = band-limited sinc interpolation with a Hann window,
lowpass_filter_width 6, rolloff 0.99 (torchaudio 2.11 functional/functional.py: _get_sinc_resample_kernel,
_apply_sinc_resample_kernel).  Restated here:

  o, n   = sr_in / gcd, sr_out / gcd;  base = min(o, n) * rolloff;  width = ceil(lpw * o / base);  taps = 2 width + o
  kern[p][k] = sinc(pi t) * cos^2(pi t / (2 lpw)) * base / o,   t = clamp((-p / n + (k - width) / o) * base, -lpw, lpw)
  y[q n + p] = sum_k xpad[q o + k] kern[p][k],  xpad = x zero-padded by `width` on the left, width + o on the right
  len(y)     = ceil(n len(x) / o)

The table is evaluated in float64 and rounded once to float32 (torchaudio evaluates it in the waveform dtype, float32);
the sums are float64.  Pinned against torchaudio's output in tests/golden/resample_ta.npz (tests/test_oracle_pins.py)."""
from __future__ import annotations

import math

import numpy as np


def sinc_kernel(sr_in: int, sr_out: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """(kern float32 [n, taps], width, o, n)."""
    g = math.gcd(int(sr_in), int(sr_out))
    o, n = int(sr_in) // g, int(sr_out) // g
    base = min(o, n) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    t = (np.arange(0, -n, -1, dtype=np.float64)[:, None] / n + idx) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * window * (base / o)
    return k.astype(np.float32), width, o, n


def resample(x: np.ndarray, sr_in: int, sr_out: int = 16000) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    if sr_in == sr_out:
        return x
    kern, width, o, n = sinc_kernel(sr_in, sr_out)
    taps = kern.shape[1]
    length = x.shape[-1]
    xpad = np.zeros(length + 2 * width + o, dtype=np.float64)
    xpad[width:width + length] = x
    frames = length // o + 1
    win = np.lib.stride_tricks.sliding_window_view(xpad, taps)[::o][:frames]  # [frames, taps]
    y = win @ kern.astype(np.float64).T                                        # [frames, n]
    target = int(math.ceil(n * length / o))
    return y.reshape(-1)[:target].astype(np.float32)
