"""oracle/logits.py — numpy restatement of the Whisper logits processors.  TEST INFRASTRUCTURE ONLY.

Follows HF/generation/logits_process.py (transformers 5.5.0), applied in the order generation_whisper.py:1774-1812
builds them:  SuppressTokensAtBegin (:1847-1862) -> SuppressTokens (:1894-1902) -> WhisperTimeStamp (:1963-2043).
"""
from __future__ import annotations

import numpy as np


def process(scores: np.ndarray, sampled: list, *, begin: bool, eos: int, no_ts: int, suppress=(), begin_suppress=(),
            max_initial_timestamp_index=None, timestamp_rules: bool = True) -> np.ndarray:
    """scores float32 [V] for one sequence; `sampled` = tokens generated so far (input_ids[begin_index:]);
    `begin` = (input_ids.shape[1] == begin_index)."""
    s = scores.astype(np.float32).copy()
    NEG = np.float32(-np.inf)
    if begin and len(begin_suppress):
        s[np.asarray(begin_suppress, dtype=np.int64)] = NEG
    if len(suppress):
        s[np.asarray(suppress, dtype=np.int64)] = NEG
    if not timestamp_rules:
        return s
    ts_begin = no_ts + 1
    s[no_ts] = NEG
    seq = list(sampled)
    last_was_ts = len(seq) >= 1 and seq[-1] >= ts_begin
    penult_was_ts = len(seq) < 2 or seq[-2] >= ts_begin
    if last_was_ts:
        if penult_was_ts:
            s[ts_begin:] = NEG
        else:
            s[:eos] = NEG
    ts = [t for t in seq if t >= ts_begin]
    if ts:
        if last_was_ts and not penult_was_ts:
            ts_last = ts[-1]
        else:
            ts_last = ts[-1] + 1
        s[ts_begin:ts_last] = NEG
    if begin:
        s[:ts_begin] = NEG
        if max_initial_timestamp_index is not None:
            s[ts_begin + max_initial_timestamp_index + 1:] = NEG
    # log_softmax in float32, then logsumexp over the timestamp slice (as torch does: max-shifted)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        m = s.max()
        lse = m + np.log(np.exp(s - m, dtype=np.float32).sum(dtype=np.float32), dtype=np.float32)
        lp = (s - lse).astype(np.float32)
        tl = lp[ts_begin:]
        tm = tl.max() if tl.size else NEG
        ts_lp = (np.log(np.exp(tl - tm, dtype=np.float32).sum(dtype=np.float32), dtype=np.float32) + tm) if tm > NEG else NEG
        mx_text = lp[:ts_begin].max()
        if ts_lp > mx_text:
            s[:ts_begin] = NEG
    return s
