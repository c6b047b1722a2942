"""Host orchestration of Whisper's `generate(..., return_timestamps=True, return_token_timestamps=True,
return_segments=True)` for batches of <= 30 s chunks, on top of the libcrisper kernels.

Restates the control flow (no arithmetic) of HF/models/whisper/generation_whisper.py (transformers 5.5.0):
  generate                :649-968   seek loop :785-903, time offsets in float64 :800-802
  _maybe_reduce_batch     :1821-1850 (active-row bookkeeping)
  _get_input_segment      :1853-1880 (slice the *features* at `seek`, right-pad with zeros to 3000 frames)
  generate_with_fallback  :970-1116  (pad stripping :1064-1084; the fallback itself is inactive: thresholds unset)
  _postprocess_outputs    :1129-1192 (num_frames - seek for the alignment crop :1147-1150)
  _retrieve_segment       :1976-2073 (split on consecutive timestamp tokens, seek advance)
  _pad_to_max_length      :126-237   (concatenate segment tokens / token timestamps)
and, for what the ASR pipeline then consumes, HF/pipelines/automatic_speech_recognition.py:519-535.

Batch-composition note: HF aligns every sample over all `T_batch - 1` decoder rows of the batch it
was decoded in, including the rows a finished sample produces while it is fed pad tokens.  `hf_batch_compat=True`
(default, = what the reference CLI with batch_size=16 computes) reproduces that; `False` aligns each sample over its
own rows only (= HF at batch size 1).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L

TIME_PRECISION = 0.02
TIME_PRECISION_FEATURES = 0.01
INPUT_STRIDE = 2
NUM_SEGMENT_FRAMES = 3000
# Rows per cw_decode_greedy call (the step kernel's shared-memory plan covers B <= 16). The reference CLI's batch_size is
# 16 (REF/transcribe.py:27), so its batches map 1:1 onto decode calls. A pipeline batch_size above 16 is decoded in groups
# of 16 and the hf_batch_compat T_batch semantics then apply per group of 16, not per pipeline batch — a documented
# difference from HF at batch_size > 16 (asr_pipeline warns once).
MAX_DECODE_BATCH = 16


@dataclass
class GenOptions:
    max_new_tokens: Optional[int] = None   # generation_config.max_new_tokens
    max_length: Optional[int] = 448        # generation_config.max_length
    hf_batch_compat: bool = True
    force_unique_generate_call: bool = False
    suppress_eos: bool = False             # benchmark mode (fixed decode length)
    init_tokens: Optional[List[int]] = None
    return_timestamps: bool = True         # timestamp tokens + WhisperTimeStampLogitsProcessor (pipeline: True or "word")
    return_token_timestamps: bool = True   # alignment-head DTW word timing (pipeline: "word")
    language: Optional[object] = None      # generate_kwargs["language"]: str or per-row list (HF :1530-1552)
    task: Optional[str] = None             # generate_kwargs["task"]


def _max_new(cfg: Dict, opts: GenOptions, n_prompt: int) -> int:
    """_set_max_new_tokens_and_length (generation_whisper.py:1713-1744)."""
    n_ctx = cfg["n_text_ctx"]
    if opts.max_new_tokens is not None:
        return max(1, min(int(opts.max_new_tokens), n_ctx - n_prompt))
    max_length = opts.max_length if opts.max_length is not None else n_ctx
    num_initial = min(n_ctx // 2 - 1, n_prompt - 1)
    max_length = min(max_length + num_initial, n_ctx)
    return max(1, max_length - n_prompt)


def retrieve_segment(seek_sequence: np.ndarray, token_timestamps: np.ndarray, time_offset: float, timestamp_begin: int,
                     seek_num_frames: int, idx_offset: int):
    """_retrieve_segment (generation_whisper.py:1976-2073) for one sample.  seek_sequence: generated ids (eos stripped);
    token_timestamps: float32 row [n_prompt + G] of this generate call.  Returns (segments, segment_offset)."""
    seq = np.asarray(seek_sequence, dtype=np.int64)
    is_ts = seq >= timestamp_begin
    single_timestamp_ending = is_ts[-2:].tolist() == [False, True]
    consec = np.nonzero(is_ts[:-1] & is_ts[1:])[0] + 1
    segments = []
    if len(consec) > 0:
        slices = consec.tolist()
        if single_timestamp_ending:
            slices.append(len(seq))
        else:
            slices[-1] += 1
        last_slice = 0
        for i, cur in enumerate(slices):
            is_last = i == len(slices) - 1
            sl = seq[last_slice:cur]
            start_pos = int(sl[0]) - timestamp_begin
            end_pos = int(sl[-1 if (not is_last or single_timestamp_ending) else -2]) - timestamp_begin
            segments.append({
                "start": time_offset + np.float64(start_pos) * TIME_PRECISION,
                "end": time_offset + np.float64(end_pos) * TIME_PRECISION,
                "tokens": sl,
                "idxs": (idx_offset + last_slice, idx_offset + cur),
                "token_timestamps": _add_offset(token_timestamps[idx_offset + last_slice: idx_offset + cur], time_offset),
            })
            last_slice = cur
        if single_timestamp_ending:
            segment_offset = int(seek_num_frames)
        else:
            last_timestamp_pos = int(seq[last_slice - 2]) - timestamp_begin
            segment_offset = last_timestamp_pos * INPUT_STRIDE
    else:
        ts = seq[is_ts]
        last_timestamp_pos = int(seek_num_frames * TIME_PRECISION_FEATURES / TIME_PRECISION)
        if ts.size > 0 and ts[-1] != timestamp_begin:
            last_timestamp_pos = np.float64(ts[-1] - timestamp_begin)
        segments.append({
            "start": time_offset,
            "end": time_offset + last_timestamp_pos * TIME_PRECISION,
            "tokens": seq,
            "idxs": (idx_offset, idx_offset + len(seq)),
            "token_timestamps": _add_offset(token_timestamps[idx_offset: idx_offset + len(seq)], time_offset),
        })
        segment_offset = int(seek_num_frames)
    return segments, segment_offset


def _add_offset(ts_f32: np.ndarray, time_offset) -> np.ndarray:
    # HF: float32 tensor + float64 0-dim tensor -> stays float32 (0-dim tensors do not promote)
    return (ts_f32.astype(np.float32) + np.float32(time_offset)).astype(np.float32)


def token_timestamps_row(jump: np.ndarray, T: int, n_prompt: int, total_len: int) -> np.ndarray:
    """[0]*n_prompt ++ jump*0.02 ++ [last] (generation_whisper.py:368-379), float32, padded with zeros to total_len."""
    row = np.zeros(total_len, np.float32)
    if T > 0:
        jt = (jump[:T].astype(np.float64) * TIME_PRECISION).astype(np.float32)
        row[n_prompt:n_prompt + T] = jt
        if n_prompt + T < total_len:
            row[n_prompt + T] = jt[-1]
    return row


def generate(engine, feats_tm: torch.Tensor, num_frames: np.ndarray, opts: GenOptions, stats: Optional[Dict] = None):
    """feats_tm bf16 [B, 3002, 128] (cw_logmel layout) on the engine's device; num_frames int [B].
    Returns per chunk {"tokens": int64 [L], "token_timestamps": float32 [L], "segments": [...]} — the generated ids
    of all seek passes concatenated (no prompt), as HF's `sequences` / `segments` before batch padding."""
    cfg = engine.desc
    B = feats_tm.shape[0]
    init_rows, need_detect = init_token_template(cfg, opts, B)   # per-row prompts; language slot filled on the first pass
    n_prompt = len(init_rows[0])
    ts_begin = cfg["no_timestamps_id"] + 1
    eos = cfg["eos_id"]
    max_new = _max_new(cfg, opts, n_prompt)
    seek = np.zeros(B, np.int64)
    # short-form input (<= 3000 frames): HF seeks over the whole padded window whatever the attention mask says
    # (_retrieve_max_frames_and_seek, generation_whisper.py:1760-1772); `num_frames` only crops the alignment
    max_frames = np.full(B, NUM_SEGMENT_FRAMES, np.int64)
    num_frames = np.asarray(num_frames, np.int64)
    current_segments: List[List[Dict]] = [[] for _ in range(B)]
    flags = L.CW_DEC_SUPPRESS_EOS if opts.suppress_eos else 0
    if not opts.return_timestamps:
        flags |= L.CW_DEC_NO_TIMESTAMP_RULES
    n_pass = 0
    while (seek < max_frames).any():
        active = [i for i in range(B) if seek[i] < max_frames[i]]
        time_offset = seek.astype(np.float64) * TIME_PRECISION / INPUT_STRIDE
        seek_num_frames = np.minimum(max_frames - seek, NUM_SEGMENT_FRAMES)
        if len(active) == B and (seek == 0).all() and (seek_num_frames == NUM_SEGMENT_FRAMES).all():
            seg_in = feats_tm
        else:
            seg_in = torch.zeros(len(active), feats_tm.shape[1], feats_tm.shape[2], dtype=feats_tm.dtype, device=feats_tm.device)
            for k, i in enumerate(active):
                n = int(seek_num_frames[i])
                seg_in[k, 1:1 + n] = feats_tm[i, 1 + int(seek[i]): 1 + int(seek[i]) + n]
        # decode in groups of <= 16 rows (the reference's batch_size, REF/transcribe.py:27); HF batch semantics
        # (T_batch) apply per group
        toks_l, G_l, T_l, jump_l = [], [], [], []
        for g0 in range(0, len(active), MAX_DECODE_BATCH):
            sl = slice(g0, min(g0 + MAX_DECODE_BATCH, len(active)))
            nb = sl.stop - sl.start
            xkv, _ = engine.encode(seg_in[sl])
            rows_g = active[sl]
            if need_detect and n_pass == 0:   # HF detects once per generate call, on the first window of every row
                for i, lid in zip(rows_g, detect_language(engine, xkv, cfg, nb)):
                    init_rows[i][1] = lid
                if stats is not None:
                    stats["language_detect_calls"] = stats.get("language_detect_calls", 0) + 1
            prompt = torch.tensor([init_rows[i] for i in rows_g], dtype=torch.int32, device=engine.device)
            out = engine.decode(xkv, prompt, max_new, flags=flags, want_align=opts.return_token_timestamps)
            engine.sync()
            if stats is not None:
                stats["decode_steps"] = stats.get("decode_steps", 0) + out["steps"]
            toks_g = out["tokens"].cpu().numpy()
            gen_counts = out["lengths"].cpu().numpy() - n_prompt   # generated tokens incl. eos
            if stats is not None:
                stats["d2h_bytes"] = stats.get("d2h_bytes", 0) + toks_g.nbytes + gen_counts.nbytes
            G_g = int(gen_counts.max())                             # HF: the batch decodes until its longest row ends
            # alignment rows = G - 1: the last generated token is never fed back (generation_whisper.py:371-376)
            T_rows = np.full(nb, G_g - 1) if opts.hf_batch_compat else gen_counts - 1
            idx = np.asarray(active[sl])
            # weights[..., : (num_frames - seek) // 2] with Python slice semantics (generation_whisper.py:1147-1150):
            # a negative bound counts from the end and 0 leaves no frame (-> every jump index is -1).
            # When the bound is the same for the whole batch HF slices twice (:315-329 and again :354).
            k = (num_frames[idx] - seek[idx]) // 2
            F_full = cfg["n_audio_ctx"]
            crop = lambda width, kk: np.where(kk >= 0, np.minimum(kk, width), np.maximum(width + kk, 0))
            F_len = crop(np.full(nb, F_full), k)
            if len(np.unique(k)) == 1:
                F_len = crop(F_len, k)
            if opts.return_token_timestamps and out["align"] is not None and T_rows.max() > 0:
                j = engine.align(out["align"], torch.from_numpy(T_rows.astype(np.int32)),
                                 torch.from_numpy(np.maximum(F_len, 1).astype(np.int32)), cfg["median_filter_width"])
                engine.sync()
                jump_g = j.cpu().numpy().copy()
                jump_g[F_len == 0] = -1
                if stats is not None:
                    stats["d2h_bytes"] = stats.get("d2h_bytes", 0) + jump_g.nbytes
            else:
                jump_g = np.zeros((nb, max_new), np.int32)
            toks_l.append(toks_g); G_l += [G_g] * nb; T_l += T_rows.tolist(); jump_l.append(jump_g)
        toks = np.concatenate(toks_l)
        jump = np.concatenate(jump_l)
        for k, i in enumerate(active):
            G_batch, T_k = G_l[k], int(T_l[k])
            L_row = n_prompt + G_batch
            tt = token_timestamps_row(jump[k], T_k, n_prompt, L_row)
            if not opts.hf_batch_compat and n_prompt + T_k + 1 < L_row:
                tt[n_prompt + T_k + 1:] = tt[n_prompt + T_k]
            seq = toks[k, n_prompt:n_prompt + G_batch].astype(np.int64)
            # generate_with_fallback: strip paddings but one eos, then the eos itself (:1064-1084)
            if len(seq) and seq[-1] == eos:
                n_pad = int((seq == eos).sum()) - 1
                if n_pad:
                    seq = seq[:-n_pad]
            if len(seq) and seq[-1] == eos:
                seq = seq[:-1]
            if len(seq) == 0:  # HF would fail on an empty sequence; an immediately-finished row advances the window
                seek[i] += int(seek_num_frames[i])
                continue
            segs, seg_off = retrieve_segment(seq, tt, time_offset[i], ts_begin, int(seek_num_frames[i]), n_prompt)
            seek[i] += seg_off
            current_segments[i] += segs
        n_pass += 1
        if opts.force_unique_generate_call:
            break
    if stats is not None:
        stats["generate_passes"] = stats.get("generate_passes", 0) + n_pass
    results = []
    for i in range(B):
        segs = current_segments[i]
        if segs:
            tokens = np.concatenate([s["tokens"] for s in segs])
            tts = np.concatenate([s["token_timestamps"] for s in segs])
        else:
            tokens, tts = np.zeros(0, np.int64), np.zeros(0, np.float32)
        results.append({"tokens": tokens, "token_timestamps": tts, "segments": segs, "init_tokens": list(init_rows[i])})
    return results


def language_to_id(cfg: Dict, language: str) -> int:
    """language_to_id of HF's _retrieve_init_tokens (generation_whisper.py:1465-1487): accepts "<|en|>", "en" or "english"."""
    table = cfg.get("lang_to_id") or {}
    lang = str(language).lower()
    if lang in table:
        token = lang
    else:
        try:
            from transformers.models.whisper.tokenization_whisper import TO_LANGUAGE_CODE
        except Exception:  # transformers is the caller's dependency (tokenizer); without it only codes are understood
            TO_LANGUAGE_CODE = {}
        if lang in TO_LANGUAGE_CODE:
            token = f"<|{TO_LANGUAGE_CODE[lang]}|>"
        elif lang in TO_LANGUAGE_CODE.values() or (not TO_LANGUAGE_CODE and f"<|{lang}|>" in table):
            token = f"<|{lang}|>"
        else:
            raise ValueError(f"Unsupported language: {language}.")
    if token not in table:
        raise ValueError(f"{token} is not supported by this specific model as it is not in the `generation_config.lang_to_id`.")
    return int(table[token])


def init_token_template(cfg: Dict, opts: GenOptions, batch_size: int):
    """Decoder prompt per row, following HF's _retrieve_init_tokens (generation_whisper.py:1455-1608):
    [<|startoftranscript|>, language, task, (<|notimestamps|>)].  Returns (rows, detect): `rows` is a list of
    `batch_size` id lists; when `detect` is True the language slot (index 1) holds None and must be filled with the
    ids cw_decode_greedy's language-detection step returns (HF detect_language :1612-1672).
    Synthetic configs without the generation_config tables may pin `lang_id` / `task_id` directly."""
    if opts.init_tokens is not None:
        return [list(opts.init_tokens) for _ in range(batch_size)], False
    sot = cfg.get("decoder_start_token_id")
    if sot is None:
        raise ValueError("config lacks decoder_start_token_id")
    lang_to_id, task_to_id = cfg.get("lang_to_id") or {}, cfg.get("task_to_id") or {}
    language = opts.language if opts.language is not None else cfg.get("language")
    task = opts.task if opts.task is not None else cfg.get("task")
    init: List[Optional[int]] = [int(sot)]
    if task is None and language is None and cfg.get("lang_id") is None:
        fdi = cfg.get("forced_decoder_ids")   # deprecated HF flag, kept for old checkpoints (:1494-1527)
        if fdi and fdi[0][0] == 1:
            fdi = [list(x) for x in fdi]
            i = 1
            while fdi and fdi[0][0] == i:
                init.append(fdi[0][1])
                fdi = fdi[1:]
                i += 1
            if fdi:
                raise ValueError("forced_decoder_ids do not follow the prompt pattern of Whisper")
    undefined = len(init) <= 1 or init[1] is None
    if isinstance(language, (list, tuple)):
        if any(l is None for l in language) or len(language) != batch_size:
            raise ValueError("a list of languages must hold one language per batch row")
        languages = list(language)
    elif language is None:
        languages = [None] * batch_size
    else:
        languages = [language] * batch_size
    rows = [list(init) for _ in range(batch_size)]
    detect = False
    lang_ids = None
    if language is not None:
        lang_ids = [language_to_id(cfg, l) for l in languages]
    elif cfg.get("lang_id") is not None and undefined:
        lang_ids = [int(cfg["lang_id"])] * batch_size
    elif lang_to_id and undefined:
        lang_ids, detect = [None] * batch_size, True
    elif cfg.get("is_multilingual") and undefined:
        raise ValueError("multilingual model without a language: set generation_config.language / lang_to_id "
                         "(or pass generate_kwargs={'language': ...}); no language token can be resolved")
    if lang_ids is not None:
        for r, lid in zip(rows, lang_ids):
            if len(r) > 1:
                r[1] = lid
            else:
                r.append(lid)
    for r in rows:
        if task is not None:
            if task in task_to_id:
                tid = int(task_to_id[task])
            elif cfg.get("task_id") is not None:
                tid = int(cfg["task_id"])
            else:
                raise ValueError(f"The `{task}` task is not supported.")
            if any(t in task_to_id.values() for t in r if t is not None):
                r[:] = [tid if (t is not None and t in task_to_id.values()) else t for t in r]
            else:
                r.append(tid)
        elif language is not None and task_to_id:
            if not any(t in task_to_id.values() for t in r if t is not None):
                r.append(int(task_to_id["transcribe"]))
        elif cfg.get("task_id") is not None and not task_to_id:
            r.append(int(cfg["task_id"]))
        no_ts = cfg["no_timestamps_id"]
        if not opts.return_timestamps and r[-1] != no_ts:
            r.append(int(no_ts))
        elif opts.return_timestamps and r[-1] == no_ts:
            r.pop()
    keep = [[k for k, t in enumerate(r) if t is not None or (detect and k == 1)] for r in rows]
    rows = [[r[k] for k in ks] for r, ks in zip(rows, keep)]
    return rows, detect


def detect_language(engine, xkv, cfg: Dict, n_rows: int) -> List[int]:
    """HF detect_language (generation_whisper.py:1612-1672): one decoder step on [<|startoftranscript|>], logits restricted
    to the language ids, argmax — per row.  Runs on the step kernel with the raw (unsuppressed, rule-free) logits."""
    sot = int(cfg["decoder_start_token_id"])
    prompt = torch.tensor([[sot]] * n_rows, dtype=torch.int32, device=engine.device)
    out = engine.decode(xkv, prompt, 1, flags=L.CW_DEC_NO_TIMESTAMP_RULES | L.CW_DEC_NO_SUPPRESS | L.CW_DEC_SUPPRESS_EOS * 0,
                        want_logits=True, want_align=False)
    engine.sync()
    ids = sorted(int(v) for v in (cfg.get("lang_to_id") or {}).values())
    logits = out["logits"][:, 0, :].float().cpu().numpy()
    sel = logits[:, ids]
    return [ids[int(k)] for k in sel.argmax(-1)]


def default_init_tokens(cfg: Dict) -> List[int]:
    """Prompt of a batch-size-1 call with default options (kept for callers that only need the prompt length)."""
    rows, detect = init_token_template(cfg, GenOptions(), 1)
    return [t if t is not None else -1 for t in rows[0]]
