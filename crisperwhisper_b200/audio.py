"""Audio front-end of the host: input normalisation and the 30 s / 5 s-stride chunker.

Mirrors AutomaticSpeechRecognitionPipeline.preprocess + chunk_iter
(HF/pipelines/automatic_speech_recognition.py:341-477, :61-84): accepted inputs, mono mix-down, chunk/stride
arithmetic.  File decoding uses scipy's WAV reader (ffmpeg, which HF shells out to at HF/pipelines/audio_utils.py:9-45,
is not part of the hot path and absent from this image); resampling to 16 kHz runs on the GPU (cw_resample)."""
from __future__ import annotations

import io
from typing import Dict, Iterator, List, Tuple

import numpy as np

SAMPLING_RATE = 16000
N_SAMPLES = 480000


def read_wav(src) -> Tuple[np.ndarray, int]:
    """path / bytes -> (float32 mono waveform in [-1, 1], sampling rate). PCM16/32, float WAV."""
    from scipy.io import wavfile
    if isinstance(src, (bytes, bytearray)):
        src = io.BytesIO(bytes(src))
    sr, data = wavfile.read(src)
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1)
    return x, int(sr)


def process_audio_bytes(audio_bytes, resampler=None) -> np.ndarray:
    """The demo app's front-end (REF/app.py:85-96): WAV bytes -> raw sample values as float32 (no PCM scaling), standardised
    to zero mean / unit std over the whole file, divided by 8, resampled to 16 kHz -> float32 waveform `[1, n]` (the caller
    transcribes row 0, REF/app.py:99-103).  `resampler(x, sr_in) -> x16k` is Engine.resample (cw_resample = the kernel
    torchaudio.transforms.Resample applies); the reference's side effect of writing sample.wav is not reproduced."""
    from scipy.io import wavfile
    sr, y = wavfile.read(io.BytesIO(bytes(audio_bytes)))
    y = y.astype(np.float32)
    y = (y - np.mean(y)) / np.std(y)
    y = np.ascontiguousarray(y / 8, dtype=np.float32)
    if int(sr) != SAMPLING_RATE:
        if resampler is None:
            raise ValueError(f"input is at {sr} Hz: pass a resampler (the pipeline uses Engine.resample) or 16 kHz audio")
        y = np.ascontiguousarray(resampler(y, int(sr)), dtype=np.float32)
    return y[None, :]


def normalize_input(inputs, resampler=None) -> np.ndarray:
    """str (wav path) | bytes (wav file) | np.ndarray | {"array"|"raw", "sampling_rate"} -> float32 mono @16 kHz
    (automatic_speech_recognition.py:342-417).  Inputs at another rate go through `resampler(x, sr_in) -> x16k`, which the
    pipeline binds to Engine.resample (cw_resample, the reference's torchaudio.functional.resample :394-408 on the GPU);
    there is no host resampler: without one such an input is an error."""
    sr = SAMPLING_RATE
    if isinstance(inputs, str):
        inputs, sr = read_wav(inputs)
    elif isinstance(inputs, (bytes, bytearray)):
        inputs, sr = read_wav(inputs)
    elif isinstance(inputs, dict):
        d = dict(inputs)
        if not ("sampling_rate" in d and ("raw" in d or "array" in d)):
            raise ValueError('When passing a dictionary, it needs a "raw" or "array" key with the numpy audio and a '
                             '"sampling_rate" key')
        arr = d.pop("raw", None)
        if arr is None:
            d.pop("path", None)
            arr = d.pop("array", None)
        sr = int(d.pop("sampling_rate"))
        inputs = arr
    try:
        import torch
        if isinstance(inputs, torch.Tensor):
            inputs = inputs.detach().cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    if not isinstance(inputs, np.ndarray):
        raise TypeError(f"We expect a numpy ndarray or torch tensor as input, got `{type(inputs)}`")
    x = inputs
    if x.ndim != 1:
        x = x.mean(axis=0)
    x = np.ascontiguousarray(x, dtype=np.float32)
    if sr != SAMPLING_RATE:
        if resampler is None:
            raise ValueError(f"input is at {sr} Hz: pass a resampler (the pipeline uses Engine.resample) or 16 kHz audio")
        x = np.ascontiguousarray(resampler(x, sr), dtype=np.float32)
    return x


def chunk_plan(n_samples: int, chunk_length_s: float = 30.0, stride_length_s=None, sampling_rate: int = SAMPLING_RATE):
    """[(start, length, stride_left, stride_right, is_last)] in samples — chunk_iter (:61-84) with the default stride
    chunk_length/6 on each side (:428-438)."""
    if stride_length_s is None:
        stride_length_s = chunk_length_s / 6
    if isinstance(stride_length_s, (int, float)):
        stride_length_s = [stride_length_s, stride_length_s]
    chunk_len = int(round(chunk_length_s * sampling_rate))
    sl = int(round(stride_length_s[0] * sampling_rate))
    sr_ = int(round(stride_length_s[1] * sampling_rate))
    if chunk_len < sl + sr_:
        raise ValueError("Chunk length must be superior to stride length")
    step = chunk_len - sl - sr_
    plan = []
    for start in range(0, n_samples, step):
        end = start + chunk_len
        length = min(end, n_samples) - start
        left = 0 if start == 0 else sl
        is_last = end >= n_samples
        right = 0 if is_last else sr_
        if length > left:
            plan.append((start, length, left, right, is_last))
        if is_last:
            break
    return plan
