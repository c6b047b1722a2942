"""oracle/align.py — CPU restatement of the token-timestamp stage (stage 3).  TEST INFRASTRUCTURE ONLY.

Follows HF/models/whisper/generation_whisper.py (transformers 5.5.0):
  _median_filter            :43-61
  _dynamic_time_warping     :64-115      (inner loops in oracle/dtw.c, pure-Python fallback below)
  _extract_token_timestamps :241-381     (per-utterance branch :352-379)

The restatement is *bit-exact* w.r.t. the CPU reference, including the summation orders ATen uses:
  - torch.mean(dim=-2) and .mean(dim=0) are `sum / n` with ATen's cascade summation
    (aten/src/ATen/native/cpu/SumKernel.cpp: multi_row_sum with level_step 16, and the 4-way interleaved
    `row_sum` for the column remainder that does not fill a SIMD group of 4 vectors) — `cascade_sum` below;
  - torch.std(unbiased=False) accumulates in float64 (Welford) and rounds once to float32.
The SIMD group width is a property of the host CPU that ran the reference (64 columns with AVX-512, 32 with
AVX2); `simd_cols` makes it explicit.  tests/test_oracle_pins.py pins all of this against the installed
transformers on this machine.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB = None


def build_dtw_lib(force: bool = False) -> str:
    """gcc-compile oracle/dtw.c into oracle/_build/libdtw_oracle.so (idempotent)."""
    os.makedirs(_BUILD, exist_ok=True)
    so = os.path.join(_BUILD, "libdtw_oracle.so")
    src = os.path.join(_HERE, "dtw.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_dtw_lib())
        lib.dtw_oracle.restype = ctypes.c_int
        lib.dtw_oracle.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p]
        _LIB = lib
    return _LIB


# ----------------------------------------------------------------------------------------------------------
# ATen CPU summation order
# ----------------------------------------------------------------------------------------------------------
def _multi_row_sum(x: np.ndarray) -> np.ndarray:
    """ATen multi_row_sum along axis 0 in float32: 4 accumulator levels, level_step = 16 (valid for n < 2**19)."""
    n = x.shape[0]
    zero = np.zeros(x.shape[1:], np.float32)
    acc = [zero.copy() for _ in range(4)]
    i = 0
    while i + 16 <= n:
        for _ in range(16):
            acc[0] = acc[0] + x[i]
            i += 1
        for j in range(1, 4):
            acc[j] = acc[j] + acc[j - 1]
            acc[j - 1] = zero.copy()
            if (i & (15 << (4 * j))) != 0:
                break
    while i < n:
        acc[0] = acc[0] + x[i]
        i += 1
    for j in range(1, 4):
        acc[0] = acc[0] + acc[j]
    return acc[0]


def _row_sum_ilp4(x: np.ndarray) -> np.ndarray:
    """ATen row_sum along axis 0: 4 interleaved partial sums (rows k mod 4), each a multi_row_sum."""
    n = x.shape[0]
    n4 = n // 4
    parts = [_multi_row_sum(x[k:4 * n4:4]) if n4 > 0 else np.zeros(x.shape[1:], np.float32) for k in range(4)]
    for i in range(4 * n4, n):
        parts[0] = parts[0] + x[i]
    for k in range(1, 4):
        parts[0] = parts[0] + parts[k]
    return parts[0]


def cascade_sum(x: np.ndarray, simd_cols: int = 64) -> np.ndarray:
    """float32 sum over axis 0 of x[n, inner...] with inner flattened, in ATen's vectorized_outer_sum order:
    full groups of `simd_cols` inner elements use multi_row_sum, the remainder uses the ilp-4 row_sum."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = x.shape[0]
    flat = x.reshape(n, -1)
    m = flat.shape[1]
    main = (m // simd_cols) * simd_cols if simd_cols > 0 else m  # simd_cols == 0: pure multi_row_sum order
    out = np.empty(m, np.float32)
    if main:
        out[:main] = _multi_row_sum(flat[:, :main])
    if main < m:
        out[main:] = _row_sum_ilp4(flat[:, main:])
    return out.reshape(x.shape[1:])


# ----------------------------------------------------------------------------------------------------------
# stage 3 proper
# ----------------------------------------------------------------------------------------------------------
def median_filter(x: np.ndarray, width: int) -> np.ndarray:
    """generation_whisper.py:43-61 — reflect pad (edge excluded), window median along the last axis."""
    if width <= 0 or width % 2 != 1:
        raise ValueError("`filter_width` should be an odd number")
    pad = width // 2
    if x.shape[-1] <= pad:
        return x
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad, pad)], mode="reflect")
    win = np.lib.stride_tricks.sliding_window_view(xp, width, axis=-1)
    # torch.sort puts NaN last; np.sort does too.
    return np.sort(win, axis=-1)[..., pad]


def normalize(w: np.ndarray, simd_cols: int = 64) -> np.ndarray:
    """generation_whisper.py:357-359 — (w - mean_T) / std_T, population std, w: [H, T, F] float32."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    H, T, F = w.shape
    mean = np.empty((H, 1, F), np.float32)
    for h in range(H):
        mean[h, 0] = cascade_sum(w[h], simd_cols) / np.float32(T)
    w64 = w.astype(np.float64)
    m64 = w64.mean(axis=1, keepdims=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        std = np.sqrt(((w64 - m64) ** 2).mean(axis=1, keepdims=True)).astype(np.float32)
        return (w - mean) / std


def cost_matrix(w: np.ndarray, median_width: int = 7, simd_cols: int = 64) -> np.ndarray:
    """normalise -> median filter -> head mean; returns the [T, F] float32 matrix HF hands (negated) to DTW."""
    x = normalize(w, simd_cols)
    x = median_filter(x, median_width)
    H = x.shape[0]
    return cascade_sum(x, simd_cols) / np.float32(H)


def dtw_python(matrix: np.ndarray):
    """Literal pure-Python restatement of generation_whisper.py:64-115 (small cases only)."""
    T, F = matrix.shape
    cost = np.ones((T + 1, F + 1), dtype=np.float32) * np.inf
    trace = -np.ones((T + 1, F + 1), dtype=np.float32)
    cost[0, 0] = 0
    for j in range(1, F + 1):
        for i in range(1, T + 1):
            c0, c1, c2 = cost[i - 1, j - 1], cost[i - 1, j], cost[i, j - 1]
            if c0 < c1 and c0 < c2:
                c, t = c0, 0
            elif c1 < c0 and c1 < c2:
                c, t = c1, 1
            else:
                c, t = c2, 2
            cost[i, j] = matrix[i - 1, j - 1] + c
            trace[i, j] = t
    i, j = T, F
    trace[0, :] = 2
    trace[:, 0] = 1
    ti, tj = [], []
    while i > 0 or j > 0:
        ti.append(i - 1)
        tj.append(j - 1)
        if trace[i, j] == 0:
            i -= 1
            j -= 1
        elif trace[i, j] == 1:
            i -= 1
        else:
            j -= 1
    return np.array(ti)[::-1], np.array(tj)[::-1]


def dtw(matrix: np.ndarray):
    """DTW on a float64 [T, F] matrix (callers pass -cost.astype(float64), :367). Returns
    (text_indices, time_indices, jump_index[T])."""
    m = np.ascontiguousarray(matrix, dtype=np.float64)
    T, F = m.shape
    ti = np.empty(T + F + 1, np.int32)
    tj = np.empty(T + F + 1, np.int32)
    jump = np.zeros(T, np.int32)
    n = _lib().dtw_oracle(m.ctypes.data, T, F, ti.ctypes.data, tj.ctypes.data, jump.ctypes.data)
    if n < 0:
        raise MemoryError
    return ti[:n].copy(), tj[:n].copy(), jump


def jump_indices(w: np.ndarray, median_width: int = 7, simd_cols: int = 64) -> np.ndarray:
    """[H, T, F] float32 alignment-head attention rows -> int32 [T] frame index per token
    (jump_times = idx * 0.02, generation_whisper.py:368-369).  T == 0 -> empty."""
    if w.shape[1] == 0:
        return np.zeros(0, np.int32)
    c = cost_matrix(w, median_width, simd_cols)
    _, _, jump = dtw(-c.astype(np.float64))
    return jump


def extract_token_timestamps(align: np.ndarray, T_len, F_len, n_prompt: int, median_width: int = 7,
                             time_precision: float = 0.02, simd_cols: int = 64) -> np.ndarray:
    """align: [N, H, T_max, F_max] float32 rows of the generated part (prompt rows already dropped).
    Returns float32 [N, n_prompt + T_max + 1] laid out like HF: zeros for the prompt, jump_times, last value
    duplicated (:377-379); rows shorter than T_max are zero-padded on the right."""
    N, H, T_max, _ = align.shape
    out = np.zeros((N, n_prompt + T_max + 1), np.float32)
    for n in range(N):
        T, F = int(T_len[n]), int(F_len[n])
        if T == 0:
            continue
        j = jump_indices(align[n, :, :T, :F], median_width, simd_cols)
        jt = (j * time_precision).astype(np.float32)  # python float * int64 in HF -> float64 -> torch.tensor f32
        out[n, n_prompt:n_prompt + T] = jt
        out[n, n_prompt + T] = jt[-1]
    return out
