"""Drop-in host for the reference's `pipeline("automatic-speech-recognition", ..., return_timestamps="word")` call
(REF/transcribe.py:21-33, REF/README.md:159-174): same keyword arguments, same call contract, same output dict —
with the three compute stages running on libcrisper.so (B200, sm_100a) instead of HF transformers.

    pipe = pipeline("automatic-speech-recognition", model=model, tokenizer=processor.tokenizer,
                    feature_extractor=processor.feature_extractor, chunk_length_s=30, batch_size=16,
                    return_timestamps="word", torch_dtype=torch.float16, device="cuda:0")
    out = pipe(audio)      # {"text": str, "chunks": [{"text": str, "timestamp": (start, end)}, ...]}

`model` may be a HF WhisperForConditionalGeneration (its state dict is repacked once into the bf16 layout of
include/crisper.h), a `PackedWeights`, or an `Engine` that already holds weights.  The tokenizer is the caller's
(as in the reference); token ids + token timestamps become words in decode_asr.py, a restatement of the tokenizer's
`_decode_asr` (HF/models/whisper/tokenization_whisper.py:901-1150) that reads the vocabulary from that tokenizer and
is ~20x cheaper on the host.  Host flow per call (HF/pipelines/automatic_speech_recognition.py):
preprocess/chunk_iter (:341-477,:61-84) -> batches of `batch_size` chunks -> cw_logmel -> generate.generate
(cw_encode / cw_decode_greedy / cw_align) -> postprocess (:562-656).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import audio as A
from . import decode_asr as D
from . import generate as G
from . import weights as Wt
from .engine import Engine


def mel_filters_slaney(n_mels: int) -> np.ndarray:
    """[n_mels, 201] float32 slaney mel filter bank == WhisperFeatureExtractor.mel_filters.T
    (HF/models/whisper/feature_extraction_whisper.py:95-103, HF/audio_utils.py:453-545)."""
    def hz2mel(f):
        f = np.asarray(f, np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * (27.0 / np.log(6.4)), 3.0 * f / 200.0)

    def mel2hz(m):
        m = np.asarray(m, np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), 200.0 * m / 3.0)

    ff = mel2hz(np.linspace(hz2mel(0.0), hz2mel(8000.0), n_mels + 2))
    fft = np.linspace(0, 8000, 201)
    diff = np.diff(ff)
    slopes = ff[None, :] - fft[:, None]
    fb = np.maximum(0.0, np.minimum(-slopes[:, :-2] / diff[:-1], slopes[:, 2:] / diff[1:]))
    fb *= (2.0 / (ff[2:n_mels + 2] - ff[:n_mels]))[None, :]
    return np.ascontiguousarray(fb.T.astype(np.float32))


class AutomaticSpeechRecognitionPipeline:
    def __init__(self, model, tokenizer=None, feature_extractor=None, chunk_length_s: float = 0, stride_length_s=None,
                 batch_size: int = 1, return_timestamps=None, torch_dtype=None, dtype=None, device=None,
                 generate_kwargs: Optional[Dict] = None, hf_batch_compat: bool = True, **_ignored):
        if isinstance(model, Engine) or (hasattr(model, "desc") and hasattr(model, "decode") and hasattr(model, "align")):
            self.engine = model
        else:
            dev = device if device is not None else 0
            self.engine = Engine(dev)
            if isinstance(model, Wt.PackedWeights):
                pw = model if model.arena_bf16.device == self.engine.device else model.to(self.engine.device)
            else:  # HF WhisperForConditionalGeneration (any dtype; REF/transcribe.py loads fp16 on GPU)
                pw = Wt.pack_hf_model(model, device=self.engine.device)
                gc = getattr(model, "generation_config", None)
                # language / task / forced ids travel in pw.config (weights.config_from_hf); the prompt is resolved per
                # call by generate.init_token_template, including per-chunk language detection when none is set
            self.engine.load_weights(pw)
        if self.engine.desc is None:
            raise RuntimeError("pipeline: the engine has no weights loaded")
        self.tokenizer = tokenizer
        self._words = D.WordDecoder(tokenizer) if tokenizer is not None else None
        self.feature_extractor = feature_extractor
        self.chunk_length_s = chunk_length_s
        self.stride_length_s = stride_length_s
        self.batch_size = max(1, int(batch_size))
        self.return_timestamps = return_timestamps
        self.hf_batch_compat = hf_batch_compat
        self.generate_kwargs = dict(generate_kwargs or {})
        cfg = self.engine.desc
        if feature_extractor is not None and hasattr(feature_extractor, "mel_filters"):
            filt = np.ascontiguousarray(np.asarray(feature_extractor.mel_filters, dtype=np.float32).T)
        else:
            filt = mel_filters_slaney(cfg["n_mels"])
        if filt.shape != (cfg["n_mels"], 201):
            raise ValueError(f"feature extractor has {filt.shape[0]} mel bins, the model expects {cfg['n_mels']}")
        self.mel_filters = torch.from_numpy(filt).to(self.engine.device)
        self.last_stats: Dict = {}

    # ------------------------------------------------------------------------------------------------------
    def __call__(self, inputs, return_timestamps=None, generate_kwargs: Optional[Dict] = None, batch_size=None,
                 chunk_length_s=None, return_language=None, **kw):
        """Single input -> dict; list of inputs -> list of dicts.  Like HF's DataLoader-backed pipeline, chunks of
        different inputs share batches of `batch_size` (HF/pipelines/base.py:1298-1318)."""
        is_list = isinstance(inputs, (list, tuple))
        items = list(inputs) if is_list else [inputs]
        rt = return_timestamps if return_timestamps is not None else self.return_timestamps
        gk = dict(self.generate_kwargs)
        gk.update(generate_kwargs or {})
        # the reference pipeline's other call arguments (_sanitize_parameters, automatic_speech_recognition.py:262-300)
        if kw.get("max_new_tokens") is not None:
            gk["max_new_tokens"] = kw["max_new_tokens"]
        stride = kw.get("stride_length_s", self.stride_length_s)
        unknown = sorted(set(kw) - {"max_new_tokens", "stride_length_s", "ignore_warning", "decoder_kwargs", "num_workers"})
        if unknown:
            raise TypeError(f"pipeline call: unexpected keyword arguments {unknown}")
        bs = self.batch_size if batch_size is None else max(1, int(batch_size))
        cl = self.chunk_length_s if chunk_length_s is None else chunk_length_s
        waves = [A.normalize_input(x, self._resample) for x in items]
        per_input = self._run(waves, cl, bs, gk, rt, stride_length_s=stride)
        results = [self._postprocess(mo, rt, return_language) for mo in per_input]
        return results if is_list else results[0]

    def forward(self, inputs, return_timestamps=None, generate_kwargs: Optional[Dict] = None, batch_size=None, chunk_length_s=None):
        """Stages 1-3 only (HF preprocess + _forward): per input, the list of per-chunk model outputs
        {"tokens", "token_timestamps", "stride", "is_last"} that `postprocess` turns into {"text","chunks"}.  A multi-GPU job
        gathers these small records (distributed.gather_results) and runs `postprocess` on one rank (BASELINE cfg 4)."""
        items = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        rt = return_timestamps if return_timestamps is not None else self.return_timestamps
        gk = dict(self.generate_kwargs)
        gk.update(generate_kwargs or {})
        bs = self.batch_size if batch_size is None else max(1, int(batch_size))
        cl = self.chunk_length_s if chunk_length_s is None else chunk_length_s
        waves = [A.normalize_input(x, self._resample) for x in items]
        return self._run(waves, cl, bs, gk, rt)

    def postprocess(self, model_outputs: List[Dict], return_timestamps=None, return_language=None):
        rt = return_timestamps if return_timestamps is not None else self.return_timestamps
        return self._postprocess(model_outputs, rt, return_language)

    def transcribe_bytes(self, audio_bytes: bytes) -> Dict:
        """REF/app.py:99-103 `transcribe`: the demo's byte front-end (standardise, / 8, resample) + word timestamps."""
        waveform = A.process_audio_bytes(audio_bytes, self._resample)
        return self(waveform[0, :], return_timestamps="word")

    def _resample(self, x: np.ndarray, sr_in: int) -> np.ndarray:
        """Host waveform at sr_in -> 16 kHz through cw_resample (HF preprocess :394-408 calls torchaudio here)."""
        dev = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(self.engine.device)
        out = self.engine.resample(dev, sr_in, A.SAMPLING_RATE)
        self.engine.sync()
        return out.cpu().numpy()

    # ------------------------------------------------------------------------------------------------------
    _GENERATE_KEYS = {"max_new_tokens", "max_length", "hf_batch_compat", "force_unique_generate_call", "suppress_eos",
                      "init_tokens", "language", "task", "return_timestamps", "return_token_timestamps", "num_beams", "do_sample",
                      "temperature", "compression_ratio_threshold", "logprob_threshold", "no_speech_threshold", "prompt_ids",
                      "condition_on_prev_tokens", "num_frames"}

    @classmethod
    def _check_generate_kwargs(cls, gk: Dict):
        """The device path is the reference's default decode: greedy, one pass per window (REF/transcribe.py passes no
        generate_kwargs; HF generation_whisper.py:641-700 then uses temperature 0 and no fallback thresholds).  Anything that
        asks for another strategy fails here instead of being silently ignored."""
        unknown = sorted(set(gk) - cls._GENERATE_KEYS)
        if unknown:
            raise ValueError(f"generate_kwargs not understood by crisperwhisper_b200: {unknown}")
        t = gk.get("temperature")
        temps = list(t) if isinstance(t, (list, tuple)) else [t]
        if any(x not in (None, 0, 0.0) for x in temps) or gk.get("do_sample"):
            raise NotImplementedError("sampling / temperature fallback is not implemented: cw_decode_greedy is the greedy "
                                      "decode the reference runs by default (temperature 0)")
        if gk.get("num_beams") not in (None, 1):
            raise NotImplementedError("beam search is not implemented (the reference decodes with num_beams=1)")
        for k in ("compression_ratio_threshold", "logprob_threshold", "no_speech_threshold"):
            if gk.get(k) is not None:
                raise NotImplementedError(f"{k}: the fallback / no-speech heuristics of HF generate_with_fallback are not implemented")
        if gk.get("prompt_ids") is not None or gk.get("condition_on_prev_tokens"):
            raise NotImplementedError("prompt_ids / condition_on_prev_tokens are not implemented")

    def _run(self, waves: List[np.ndarray], chunk_length_s, batch_size: int, gk: Dict, return_timestamps="word",
             stride_length_s="default") -> List[List[Dict]]:
        self._check_generate_kwargs(gk)
        eng = self.engine
        plan = []  # (input index, start, length, left, right, is_last, with_stride)
        for wi, wave in enumerate(waves):
            if chunk_length_s:
                stride = self.stride_length_s if isinstance(stride_length_s, str) else stride_length_s
                plan += [(wi,) + p + (True,) for p in A.chunk_plan(len(wave), chunk_length_s, stride)]
            else:
                if len(wave) > A.N_SAMPLES:
                    raise NotImplementedError("inputs longer than 30 s need chunk_length_s (the reference always sets 30)")
                plan.append((wi, 0, len(wave), 0, 0, True, False))
        opts = G.GenOptions(max_new_tokens=gk.get("max_new_tokens"), max_length=gk.get("max_length", 448),
                            hf_batch_compat=gk.get("hf_batch_compat", self.hf_batch_compat),
                            force_unique_generate_call=bool(gk.get("force_unique_generate_call", False)),
                            suppress_eos=bool(gk.get("suppress_eos", False)), init_tokens=gk.get("init_tokens"),
                            # HF _forward (:503-508): "word" -> token timestamps + timestamp tokens; True -> timestamp tokens
                            # only; None/False -> <|notimestamps|> prompt and no timestamp rules
                            return_timestamps=bool(return_timestamps), return_token_timestamps=(return_timestamps == "word"),
                            language=gk.get("language"), task=gk.get("task"))
        if batch_size > G.MAX_DECODE_BATCH and opts.hf_batch_compat and opts.return_token_timestamps:
            import warnings
            warnings.warn(f"batch_size={batch_size} > {G.MAX_DECODE_BATCH}: chunks are decoded in groups of {G.MAX_DECODE_BATCH} and "
                          "hf_batch_compat's batch-wide alignment length applies per group (HF applies it per pipeline batch)",
                          stacklevel=3)
        stats = {"chunks": len(plan), "decode_steps": 0, "generate_passes": 0, "h2d_bytes": 0, "d2h_bytes": 0}
        outputs: List[List[Dict]] = [[] for _ in waves]
        for b0 in range(0, len(plan), batch_size):
            items = plan[b0:b0 + batch_size]
            host = torch.zeros(len(items), A.N_SAMPLES, dtype=torch.float32, pin_memory=(eng.device.type == "cuda"))
            n_valid = []
            for k, (wi, start, length, _, _, _, _) in enumerate(items):
                n = min(length, A.N_SAMPLES)
                host[k, :n] = torch.from_numpy(waves[wi][start:start + n])
                n_valid.append(n)
            dev_wave = host.to(eng.device, non_blocking=True)
            stats["h2d_bytes"] += host.numel() * 4
            nv = torch.tensor(n_valid, dtype=torch.int32, device=eng.device)
            _, tm, frames = eng.logmel(dev_wave, self.mel_filters, nv, want_f32=False, want_tm=True)
            eng.sync()
            res = G.generate(eng, tm, frames.cpu().numpy(), opts, stats)
            for (wi, start, length, left, right, is_last, with_stride), r in zip(items, res):
                out = {"tokens": r["tokens"][None, :], "is_last": is_last}
                if opts.return_token_timestamps:
                    out["token_timestamps"] = r["token_timestamps"][None, :]
                if with_stride:
                    out["stride"] = (length, left, right)
                outputs[wi].append(out)
        self.last_stats = stats
        return outputs

    # ------------------------------------------------------------------------------------------------------
    def _postprocess(self, model_outputs: List[Dict], return_timestamps, return_language=None):
        """postprocess (automatic_speech_recognition.py:562-656) for the seq2seq_whisper type; `return_language` adds the
        language of every chunk / word, as the reference pipeline's call argument of the same name (:296-299)."""
        if self.tokenizer is None:
            return {"tokens": [o["tokens"][0] for o in model_outputs],
                    "token_timestamps": [o["token_timestamps"][0] for o in model_outputs if "token_timestamps" in o]}
        time_precision = 30.0 / self.engine.desc["n_audio_ctx"]
        for o in model_outputs:
            if "stride" in o:
                cl, sl, sr = o["stride"]
                o["stride"] = (cl / A.SAMPLING_RATE, sl / A.SAMPLING_RATE, sr / A.SAMPLING_RATE)
        if self._words is None or self._words.tok is not self.tokenizer:
            self._words = D.WordDecoder(self.tokenizer)
        text, optional = self._words.decode_asr(model_outputs, return_timestamps=return_timestamps,
                                                return_language=return_language, time_precision=time_precision)
        return {"text": text, **optional}


def pipeline(task: str = "automatic-speech-recognition", model=None, **kwargs) -> AutomaticSpeechRecognitionPipeline:
    """Same signature the reference uses for transformers.pipeline (REF/transcribe.py:21-31)."""
    if task != "automatic-speech-recognition":
        raise ValueError("crisperwhisper_b200.pipeline only implements 'automatic-speech-recognition'")
    if model is None:
        raise ValueError("pipeline: `model` is required")
    return AutomaticSpeechRecognitionPipeline(model, **kwargs)
