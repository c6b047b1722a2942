"""Token ids + token timestamps -> text and word chunks: the host stage after cw_align (first "next" row).

The reference gets this from the tokenizer the caller hands to `pipeline(...)`: `tokenizer._decode_asr`
(HF/models/whisper/tokenization_whisper.py:901-1150) with its helpers `_find_longest_common_sequence` (:1153-1270),
`_collate_word_timestamps` / `_combine_tokens_into_words` (:1273-1318), `_split_tokens_on_unicode` (:1321-1350),
`_split_tokens_on_spaces` (:1353-1376) and `_merge_punctuations` (:1379-1405).  That code calls
`tokenizer.decode(...)` once per text token (a growing list every time) and builds numpy arrays per overlap shift;
at 8 x 445 tokens it costs ~195 ms of host time per batch, which is more than the whole B200 decode of those tokens.

This module restates the same algorithm with the same results (tests/test_decode_asr.py checks it against the HF
functions on randomised token streams, strides, prompts and language switches):

  * text of a run of byte-level BPE tokens = UTF-8 decode (errors="replace") of the concatenated token bytes, taken
    from a per-token byte table built once per tokenizer; tokens that are not byte-level (added tokens) and tokenizers
    that post-process decoded text fall back to `tokenizer.decode` — same answer, slower;
  * overlap merge: all shifts scored at once from one [left, right] equality matrix (diagonal sums), same score
    `matches / i + i / 10000`, same first-maximum tie rule, same `matches > 1` gate;
  * word grouping and punctuation merging as in the reference, one pass each.

The tokenizer stays the caller's: ids of special tokens, language names and the eos boundary are read from it.

The hot configuration (`return_timestamps="word"`, byte-level vocabulary) runs natively: `cw_words_from_tokens`
(csrc/postproc.cu, declared in include/crisper.h) is the same algorithm in C++ behind the C-ABI, fed with the byte table and
special-token tables built here once per tokenizer.  It answers CW_POST_PUNT for anything it does not model (and for inputs on
which the reference raises); this module then runs the Python path below, which is also what `CW_POSTPROC=python` forces.
tests/test_postproc_native_cpu.py holds the two to identical results on the randomised streams of tests/test_decode_asr.py.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_REPL = "�"
_ASCII_PUNCT = "!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~"
_PREPENDED = "\"'“¡¿([{-"
_APPENDED = "\"'.。,，!！?？:：”)]}、"
_UNSPACED_LANGUAGES = frozenset({"chinese", "japanese", "thai", "lao", "myanmar", "cantonese"})


def _byte_alphabet() -> Dict[str, int]:
    """Inverse of the GPT-2 byte<->printable-character table every byte-level BPE vocabulary is written in."""
    keep = list(range(0x21, 0x7F)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    table, spill = {}, 0
    for b in range(256):
        if b in keep:
            table[chr(b)] = b
        else:
            table[chr(256 + spill)] = b
            spill += 1
    return table


_CHAR_TO_BYTE = _byte_alphabet()


def merge_overlaps(sequences: Sequence[List[int]], stamps: Optional[Sequence[List[Tuple[float, float]]]] = None):
    """Stitch consecutive token runs whose ends overlap in audio (tokenization_whisper.py:1153-1270).

    For every shift i of the right run over the left one, `matches` counts positions where the tokens are equal (and,
    with `stamps`, the left (start, end) pair is <= the right one); the shift with the largest matches / i + i / 1e4
    among those with more than one match wins, the earliest on ties; each side keeps its half of the overlap.
    Returns the merged ids (and merged stamps when `stamps` is not None)."""
    use_stamps = bool(stamps)
    left = list(sequences[0])
    left_st = list(stamps[0]) if use_stamps else None
    total: List[int] = []
    total_st: List[Tuple[float, float]] = []
    for si in range(1, len(sequences)):
        right = sequences[si]
        right_st = stamps[si] if use_stamps else None
        L, R = len(left), len(right)
        pick = (L, L, 0, 0)
        if L and R:
            eq = np.asarray(left, dtype=np.int64)[:, None] == np.asarray(right, dtype=np.int64)[None, :]
            if use_stamps:
                ls = np.asarray(left_st, dtype=np.float64).reshape(L, 2)
                rs = np.asarray(right_st, dtype=np.float64).reshape(R, 2)
                a0, a1, b0, b1 = ls[:, 0:1], ls[:, 1:2], rs[:, 0][None, :], rs[:, 1][None, :]
                eq &= (a0 < b0) | ((a0 == b0) & (a1 <= b1))
            li, ri = np.nonzero(eq)
            # shift i pairs left[l] with right[r] where r - l == i - L: one diagonal of `eq` per shift
            counts = np.bincount(ri - li + L, minlength=L + R)[: L + R]
            shifts = np.arange(L + R, dtype=np.int64)
            ok = counts > 1
            ok[0] = False
            if ok.any():
                with np.errstate(divide="ignore", invalid="ignore"):
                    score = counts.astype(np.float64) / shifts + shifts / 10000.0
                score[~ok] = -1.0
                i = int(np.argmax(score))
                pick = (max(0, L - i), min(L, L + R - i), max(0, i - L), min(R, i))
        l_mid = (pick[1] + pick[0]) // 2
        r_mid = (pick[3] + pick[2]) // 2
        total.extend(left[:l_mid])
        left = list(right[r_mid:])
        if use_stamps:
            total_st.extend(left_st[:l_mid])
            left_st = list(right_st[r_mid:])
    total.extend(left)
    if stamps is None:
        return total
    if use_stamps:
        total_st.extend(left_st)
        return total, total_st
    return total, []


def attach_punctuation(words: List[str], tokens: List[List[int]], indices: List[List[int]],
                       prepended: str = _PREPENDED, appended: str = _APPENDED) -> None:
    """In place: glue opening punctuation to the following word and closing punctuation to the preceding one
    (tokenization_whisper.py:1379-1405; note both membership tests are substring tests, as there)."""
    n = len(words)
    j = n - 1
    for i in range(n - 2, -1, -1):
        w = words[i]
        if w.startswith(" ") and w.strip() in prepended:
            words[j] = w + words[j]
            tokens[j] = tokens[i] + tokens[j]
            indices[j] = indices[i] + indices[j]
            words[i], tokens[i], indices[i] = "", [], []
        else:
            j = i
    i = 0
    for j in range(1, n):
        if not words[i].endswith(" ") and words[j] in appended:
            words[i] += words[j]
            tokens[i] += tokens[j]
            indices[i] += indices[j]
            words[j], tokens[j], indices[j] = "", [], []
        else:
            i = j
    words[:] = [w for w in words if w]
    tokens[:] = [t for t in tokens if t]
    indices[:] = [x for x in indices if x]


class WordDecoder:
    """Per-tokenizer state for `decode_asr`: special-token ids, language names, the token byte table."""

    def __init__(self, tokenizer):
        self.tok = tokenizer
        self.timestamp_begin = tokenizer.convert_tokens_to_ids("<|notimestamps|>") + 1
        self.special_ids = frozenset(tokenizer.all_special_ids)
        self.prompt_id = tokenizer.convert_tokens_to_ids("<|startofprev|>")
        self.sot_id = tokenizer.convert_tokens_to_ids("<|startoftranscript|>")
        self.eos_id = tokenizer.eos_token_id
        self._lang: Dict[int, Optional[str]] = {}
        self._bytes: Dict[int, Optional[bytes]] = {}
        self._bytes_ok = self._probe_byte_level()
        self._native = None          # vocabulary tables for cw_words_from_tokens, built on first use
        self.native_calls = 0        # outputs answered by the native path / handed back to the Python path
        self.native_punts = 0

    # -- token text -----------------------------------------------------------------------------------------------
    def _token_bytes(self, tid: int) -> Optional[bytes]:
        try:
            return self._bytes[tid]
        except KeyError:
            pass
        out: Optional[bytes] = None
        if 0 <= tid < self.eos_id and tid not in self.special_ids:
            piece = self.tok.convert_ids_to_tokens(int(tid))
            if isinstance(piece, str):
                try:
                    out = bytes(_CHAR_TO_BYTE[c] for c in piece)
                except KeyError:
                    out = None
        self._bytes[tid] = out
        return out

    def _probe_byte_level(self) -> bool:
        """The byte table is only used if it reproduces tokenizer.decode on a spread of vocabulary ids (it does for
        byte-level BPE without text clean-up, i.e. Whisper's); otherwise every decode goes through the tokenizer."""
        if getattr(self.tok, "clean_up_tokenization_spaces", False):
            return False
        n = int(self.eos_id or 0)
        if n <= 0:
            return False
        rng = np.random.default_rng(0)
        for width in (1, 2, 5, 16):
            for _ in range(8):
                ids = [int(t) for t in rng.integers(0, n, width)]
                tb = [self._token_bytes(t) for t in ids]
                if any(b is None for b in tb):
                    continue
                want = self.tok.decode(ids, decode_with_timestamps=True)
                if b"".join(tb).decode("utf-8", errors="replace") != want or want != self.tok.decode(ids):
                    return False
        return True

    def _byte_table(self, ids: Sequence[int]) -> Optional[List[bytes]]:
        if not self._bytes_ok:
            return None
        tb = [self._token_bytes(t) for t in ids]
        return None if any(b is None for b in tb) else tb

    def text(self, ids: Sequence[int]) -> str:
        tb = self._byte_table(ids)
        if tb is None:
            return self.tok.decode(list(ids))
        return b"".join(tb).decode("utf-8", errors="replace")

    def language_of(self, tid: int) -> Optional[str]:
        """Language name of a special token like <|en|>, else None (tokenization_whisper.py:998-1003)."""
        try:
            return self._lang[tid]
        except KeyError:
            from transformers.models.whisper.tokenization_whisper import LANGUAGES
            name = LANGUAGES.get(self.tok.decode([tid])[2:-2])
            self._lang[tid] = name
            return name

    # -- words ----------------------------------------------------------------------------------------------------
    def split_units(self, ids: List[int]):
        """Smallest runs of tokens that decode to complete unicode (tokenization_whisper.py:1321-1350): a run is
        closed when its text holds no U+FFFD, or when the U+FFFD is also there in the text of the whole sequence."""
        tb = self._byte_table(ids)
        if tb is None:
            def dec(lo, hi):
                return self.tok.decode(ids[lo:hi], decode_with_timestamps=True)
        else:
            def dec(lo, hi):
                return b"".join(tb[lo:hi]).decode("utf-8", errors="replace")
        whole = dec(0, len(ids))
        units, unit_tokens, unit_indices = [], [], []
        lo, offset = 0, 0
        for k in range(len(ids)):
            s = dec(lo, k + 1)
            p = s.find(_REPL)
            if p < 0 or whole[offset + p] == _REPL:
                units.append(s)
                unit_tokens.append(list(ids[lo:k + 1]))
                unit_indices.append(list(range(lo, k + 1)))
                lo = k + 1
                offset += len(s)
        return units, unit_tokens, unit_indices

    def split_words(self, ids: List[int], language: Optional[str]):
        """(words, word tokens, word token indices) — tokenization_whisper.py:1286-1318,:1353-1376."""
        if language is None:
            language = getattr(self.tok, "language", None)
        if language is None:
            language = "english"
        units, unit_tokens, unit_indices = self.split_units(ids)
        if language in _UNSPACED_LANGUAGES:
            words, word_tokens, word_indices = units, unit_tokens, unit_indices
        else:
            words, word_tokens, word_indices = [], [], []
            for u, ut, ui in zip(units, unit_tokens, unit_indices):
                opens = (ut[0] >= self.eos_id) or u.startswith(" ") or (u.strip() in _ASCII_PUNCT) or not words
                if opens:
                    words.append(u)
                    word_tokens.append(ut)
                    word_indices.append(ui)
                else:
                    words[-1] = words[-1] + u
                    word_tokens[-1].extend(ut)
                    word_indices[-1].extend(ui)
        attach_punctuation(words, word_tokens, word_indices)
        return words, word_tokens, word_indices

    def word_chunks(self, ids, stamps, language, return_language):
        words, _, idx = self.split_words(ids, language)
        extra = {"language": language} if return_language else {}
        return [{"text": w, "timestamp": (stamps[ix[0]][0], stamps[ix[-1]][1]), **extra} for w, ix in zip(words, idx)]

    # -- native path (csrc/postproc.cu) -----------------------------------------------------------------------------
    def _native_tables(self):
        """Byte spelling of every text token + special / language tables, as flat arrays for cw_words_from_tokens."""
        if self._native is not None:
            return self._native
        from transformers.models.whisper.tokenization_whisper import LANGUAGES
        from . import _lib as L
        lib = L.load()
        eos = int(self.eos_id)
        n_ids = max(len(self.tok), max(self.special_ids) + 1 if self.special_ids else 0)
        pieces = self.tok.convert_ids_to_tokens(list(range(eos)))
        special = np.zeros(n_ids, dtype=np.uint8)
        special[[t for t in self.special_ids if 0 <= t < n_ids]] = 1
        off = np.zeros(eos + 1, dtype=np.int64)
        has = np.zeros(max(eos, 1), dtype=np.uint8)
        blob = bytearray()
        for t, piece in enumerate(pieces):
            if not special[t] and isinstance(piece, str):
                try:
                    blob += bytes(_CHAR_TO_BYTE[c] for c in piece)
                    has[t] = 1
                except KeyError:
                    pass
            off[t + 1] = len(blob)
        names = sorted(set(LANGUAGES.values()))
        index = {nm: k for k, nm in enumerate(names)}
        lang_of = np.full(n_ids, -1, dtype=np.int32)
        for t in self.special_ids:
            nm = self.language_of(t)
            if nm is not None and 0 <= t < n_ids:
                lang_of[t] = index[nm]
        unspaced = np.asarray([nm in _UNSPACED_LANGUAGES for nm in names], dtype=np.uint8)
        default = getattr(self.tok, "language", None) or "english"
        self._native = dict(lib=lib, bytes=np.frombuffer(bytes(blob) or b"\0", dtype=np.uint8).copy(), off=off, has=has, eos=eos,
                            special=special, lang_of=lang_of, n_ids=n_ids, unspaced=unspaced, names=names,
                            default_unspaced=int(default in _UNSPACED_LANGUAGES),
                            max_len=int(np.diff(off).max()) if eos else 1)
        return self._native

    def _decode_words_native(self, model_outputs, return_language, time_precision, segment_size):
        """(text, {"chunks": words}) from cw_words_from_tokens, or None when the native path hands the input back."""
        nt = self._native_tables()
        runs, times, strides, has_stride = [], [], [], []
        for output in model_outputs:
            ids = np.asarray(output["tokens"])
            tt = np.asarray(output["token_timestamps"])
            if ids.ndim != 2 or tt.ndim != 2 or ids.shape[1] != tt.shape[1]:
                return None
            runs.append(np.ascontiguousarray(ids[0], dtype=np.int64))
            times.append(np.ascontiguousarray(tt[0], dtype=np.float64))
            st = output.get("stride")
            has_stride.append(st is not None)
            strides.append([float(x) for x in st] if st is not None else [0.0, 0.0, 0.0])
        n_out = len(runs)
        lens = np.asarray([len(r) for r in runs], dtype=np.int64)
        out_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        flat = np.concatenate(runs) if n_out else np.zeros(0, np.int64)
        if flat.size and (flat.min() < 0 or flat.max() >= nt["n_ids"]):
            return None
        tokens = np.ascontiguousarray(flat, dtype=np.int32)
        ttimes = np.ascontiguousarray(np.concatenate(times) if n_out else np.zeros(0), dtype=np.float64)
        strides_a = np.ascontiguousarray(np.asarray(strides, dtype=np.float64).reshape(-1, 3) if n_out else np.zeros((0, 3)))
        has_a = np.asarray(has_stride, dtype=np.uint8)
        n_tok = int(tokens.size)
        text_cap = n_tok * nt["max_len"] + 16
        word_cap = 3 * text_cap + 16
        text_buf = np.empty(text_cap, dtype=np.uint8)
        word_buf = np.empty(word_cap, dtype=np.uint8)
        chunk_off = np.empty(n_tok + n_out + 3, dtype=np.int64)
        word_off = np.empty(n_tok + 3, dtype=np.int64)
        w_start = np.empty(n_tok + 2, dtype=np.float64)
        w_end = np.empty(n_tok + 2, dtype=np.float64)
        w_lang = np.empty(n_tok + 2, dtype=np.int32)
        n_chunks, n_words, flags = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int32(0)

        def p(a):
            return a.ctypes.data_as(ctypes.c_void_p)
        rc = nt["lib"].cw_words_from_tokens(
            p(nt["bytes"]), p(nt["off"]), p(nt["has"]), nt["eos"], p(nt["special"]), p(nt["lang_of"]), nt["n_ids"], p(nt["unspaced"]),
            len(nt["names"]), nt["default_unspaced"], int(self.timestamp_begin), int(self.prompt_id), int(self.sot_id), n_out,
            p(tokens), p(ttimes), p(out_off), p(strides_a), p(has_a), float(time_precision), int(segment_size),
            p(text_buf), text_cap, p(chunk_off), int(chunk_off.size - 1), ctypes.byref(n_chunks),
            p(word_buf), word_cap, p(word_off), p(w_start), p(w_end), p(w_lang), int(w_lang.size), ctypes.byref(n_words),
            ctypes.byref(flags))
        if rc != 0:
            return None
        if flags.value & 1:
            import logging
            logging.getLogger(__name__).warning(
                "no closing timestamp token: the audio may be cut mid-word, or the timestamp rules were off")
        tb = text_buf.tobytes()
        co = chunk_off[: n_chunks.value + 1].tolist()
        text = "".join(tb[co[k]:co[k + 1]].decode("utf-8", errors="replace") for k in range(n_chunks.value))
        wb = word_buf.tobytes()
        wo = word_off[: n_words.value + 1].tolist()
        ws, we, wl = w_start[: n_words.value].tolist(), w_end[: n_words.value].tolist(), w_lang[: n_words.value].tolist()
        names = nt["names"]
        words = []
        for k in range(n_words.value):
            item = {"text": wb[wo[k]:wo[k + 1]].decode("utf-8"), "timestamp": (ws[k], we[k])}
            if return_language:
                item["language"] = names[wl[k]] if wl[k] >= 0 else None
            words.append(item)
        return text, {"chunks": words}

    # -- the pass over the model outputs --------------------------------------------------------------------------
    def decode_asr(self, model_outputs, *, return_timestamps, return_language, time_precision, segment_size=1500):
        """Same contract as tokenizer._decode_asr (tokenization_whisper.py:901-1150): returns (text, optional) where
        optional is {} or {"chunks": [...]}.  `model_outputs` items hold "tokens" [1, L], optionally
        "token_timestamps" [1, L] (seconds, cumulative) and "stride" (chunk_len, left, right) in seconds."""
        word_mode = return_timestamps == "word"
        if word_mode and self._bytes_ok and os.environ.get("CW_POSTPROC", "native") != "python":
            try:
                got = self._decode_words_native(model_outputs, return_language, time_precision, segment_size)
            except (KeyError, TypeError, ValueError, IndexError):
                got = None           # malformed outputs: the Python path raises what the reference raises
            if got is not None:
                self.native_calls += 1
                return got
            self.native_punts += 1
        ts0 = self.timestamp_begin
        language: Optional[str] = None

        def fresh():
            return {"language": language, "timestamp": [None, None], "text": ""}

        chunks: List[Dict] = []
        chunk = fresh()
        time_offset = 0.0
        held: List[List[int]] = []          # token runs waiting to be merged into the open chunk
        held_st: List[List[Tuple[float, float]]] = []
        skip = False

        def close(current, current_st):
            nonlocal chunk, held, held_st
            held.append(current)
            if word_mode:
                held_st.append(current_st)
            ids, st = merge_overlaps(held, held_st)
            chunk["text"] = self.text(ids)
            if word_mode:
                chunk["words"] = self.word_chunks(ids, st, language, return_language)
            chunks.append(chunk)
            held, held_st = [], []
            chunk = fresh()

        for output in model_outputs:
            ids = np.asarray(output["tokens"])[0].tolist()
            if ids and ids[0] == self.prompt_id:  # drop a <|startofprev|> prompt (tokenization_whisper.py:_strip_prompt)
                ids = ids[ids.index(self.sot_id):] if self.sot_id in ids else []
            if word_mode:
                token_times = np.asarray(output["token_timestamps"])[0].tolist()
            stride_end_token = None         # timestamp tokens at/after this one lie in the right stride
            first_timestamp = ts0
            seg_max, seg_prev_max, segs_before = 0.0, 0.0, 0.0
            stride = output.get("stride")
            if stride is not None:
                chunk_len, stride_left, stride_right = stride
                time_offset -= stride_left
                right_start = chunk_len - stride_right
                if stride_left:
                    first_timestamp = stride_left / time_precision + ts0
                if stride_right:
                    for t in reversed(ids):
                        if t >= ts0:
                            if stride_end_token is not None and (t - ts0) * time_precision < right_start:
                                break
                            stride_end_token = t
            current: List[int] = []
            current_st: List[Tuple[float, float]] = []
            for i, t in enumerate(ids):
                if t in self.special_ids:
                    name = self.language_of(t)
                    if name is not None:
                        if language and name != language and not return_timestamps:
                            held.append(current)
                            chunk["text"] = self.text(merge_overlaps(held))
                            chunks.append(chunk)
                            held, current = [], []
                            chunk = fresh()
                        chunk["language"] = name
                        language = name
                elif t >= ts0:
                    stamp = float((t - ts0) * time_precision)
                    if stamp < seg_max:     # timestamps restarted: generate() concatenated another 30 s segment
                        single_ending = i >= 2 and not (ids[i - 1] >= ts0 and ids[i - 2] >= ts0)
                        if single_ending:
                            segs_before += time_precision * segment_size
                        else:
                            seg_max = seg_prev_max
                            segs_before += seg_prev_max
                    seg_prev_max = seg_max
                    seg_max = stamp
                    when = round((t - ts0) * time_precision + time_offset + segs_before, 2)
                    if stride_end_token and t >= stride_end_token:
                        skip = True         # inside the right stride: resolved by the next output's left stride
                    elif skip or (held and t < first_timestamp):
                        skip = False
                    elif chunk["timestamp"][0] is None:
                        chunk["timestamp"][0] = when
                    elif when != chunk["timestamp"][0]:
                        chunk["timestamp"][1] = when
                        close(current, current_st)
                        current, current_st = [], []
                else:
                    current.append(t)
                    if word_mode:
                        start = round(0.0 + time_offset, 2) if i == 0 else round(token_times[i - 1] + time_offset, 2)
                        current_st.append((start, round(token_times[i] + time_offset, 2)))
            if stride is not None:
                time_offset += chunk_len - stride_right
            if current:
                held.append(current)
                if word_mode:
                    held_st.append(current_st)
            elif not any(held):
                chunk = fresh()
                held, held_st = [], []

        if held:
            if return_timestamps:
                import logging
                logging.getLogger(__name__).warning(
                    "no closing timestamp token: the audio may be cut mid-word, or the timestamp rules were off")
            ids, st = merge_overlaps(held, held_st)
            chunk["text"] = self.text(ids)
            if word_mode:
                chunk["words"] = self.word_chunks(ids, st, language, return_language)
            chunks.append(chunk)

        text = "".join(c["text"] for c in chunks)
        if not (return_timestamps or return_language):
            return text, {}
        for c in chunks:
            if return_timestamps:
                c["timestamp"] = tuple(c["timestamp"])
            else:
                c.pop("timestamp")
            if not return_language:
                c.pop("language")
        if word_mode:
            return text, {"chunks": [w for c in chunks for w in c["words"]]}
        return text, {"chunks": chunks}


def decode_asr(tokenizer, model_outputs, *, return_timestamps, return_language, time_precision, segment_size=1500):
    """Functional form; builds (and caches on the tokenizer object) a WordDecoder."""
    wd = getattr(tokenizer, "_cw_word_decoder", None)
    if wd is None or wd.tok is not tokenizer:
        wd = WordDecoder(tokenizer)
        try:
            tokenizer._cw_word_decoder = wd
        except Exception:
            pass
    return wd.decode_asr(model_outputs, return_timestamps=return_timestamps, return_language=return_language,
                         time_precision=time_precision, segment_size=segment_size)
