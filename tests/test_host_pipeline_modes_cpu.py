"""CPU tests of the pipeline modes the reference surface allows besides return_timestamps="word": language
auto-detection per chunk (HF _retrieve_init_tokens + detect_language, the reference CLI's default flow: no language is
passed, REF/transcribe.py:21-33), generate_kwargs language/task, segment-level timestamps (return_timestamps=True, as
REF/app.py's setup) and no timestamps.  Checker: the HF pipeline itself, live, on the same random-init model; the compute
stages of the product host are supplied by the CPU oracle engine."""
import warnings

import numpy as np
import pytest


def _setup(seed=3, language="en", n_lang=3):
    from oracle_engine import OracleEngine
    from oracle import hf_harness as H
    from crisperwhisper_b200 import weights as Wt
    m = H.build_model(H.tiny_hf_config(n_mels=80), seed=seed, logit_scale=8.0, pos_scale=20.0, max_new_tokens=12)
    gc = m.generation_config
    gc.lang_to_id = {f"<|{c}|>": H.TOK_IDS["en"] + k for k, c in enumerate(["en", "zh", "de"][:n_lang])}
    gc.language = language
    gc.task = None
    cfg = Wt.config_from_hf(m)
    eng = OracleEngine({k: v.float() for k, v in m.state_dict().items()}, cfg)
    return m, eng, H


def _both(m, eng, H, wave, return_timestamps, generate_kwargs=None, batch_size=2):
    from crisperwhisper_b200 import pipeline
    tok = H.synthetic_tokenizer()
    ref = H.build_pipeline(m, tok, batch_size=batch_size)
    ours = pipeline("automatic-speech-recognition", model=eng, tokenizer=tok, chunk_length_s=30, batch_size=batch_size)
    gk = dict(generate_kwargs or {})
    gk.setdefault("max_new_tokens", 12)   # both sides decode the same number of steps (HF aligns over the batch's rows)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = ref(wave.copy(), return_timestamps=return_timestamps, generate_kwargs=dict(gk))
        got = ours(wave.copy(), return_timestamps=return_timestamps, generate_kwargs=dict(gk))
    return want, got, ours


def _same(want, got):
    assert got["text"] == want["text"]
    if "chunks" in want:
        assert [(c["text"], tuple(c["timestamp"])) for c in got["chunks"]] == [(c["text"], tuple(c["timestamp"])) for c in want["chunks"]]
    else:
        assert "chunks" not in got


def test_language_is_detected_per_chunk_when_none_is_configured():
    """generation_config.language = None on a multilingual model: one decoder step on <|startoftranscript|> per chunk,
    argmax over the language ids, and that id becomes prompt token 1 — as the reference CLI (no language passed) does."""
    m, eng, H = _setup(language=None)
    wave = np.concatenate([H.speechlike(11, 200000), H.noise(12, 480000), H.speechlike(13, 300000)])
    want, got, pipe = _both(m, eng, H, wave, "word")
    _same(want, got)
    assert pipe.last_stats["language_detect_calls"] >= 1
    # the prompt really carries a language id (never the malformed [sot, task] prompt)
    import crisperwhisper_b200.generate as G
    rows, detect = G.init_token_template(eng.desc, G.GenOptions(), 2)
    assert detect and all(len(r) == 2 and r[1] is None for r in rows)


def test_generate_kwargs_language_and_task_names():
    """generate_kwargs={"language": "german", "task": "translate"} resolve through TO_LANGUAGE_CODE / task_to_id like HF."""
    m, eng, H = _setup(language="en")
    wave = H.speechlike(21, 250000)
    for gk in ({"language": "german"}, {"language": "zh", "task": "translate"}, {"language": "<|en|>"}):
        want, got, _ = _both(m, eng, H, wave, "word", generate_kwargs=gk, batch_size=1)
        _same(want, got)
    import crisperwhisper_b200.generate as G
    rows, _ = G.init_token_template(eng.desc, G.GenOptions(language="german", task="translate"), 1)
    assert rows[0] == [H.TOK_IDS["sot"], H.TOK_IDS["en"] + 2, H.TOK_IDS["translate"]]
    with pytest.raises(ValueError):
        G.init_token_template(eng.desc, G.GenOptions(language="klingon"), 1)


@pytest.mark.parametrize("rt", [True, False])
def test_segment_level_and_no_timestamp_modes(rt):
    """return_timestamps=True (REF/app.py): timestamp tokens, segment chunks, no DTW; False: <|notimestamps|> prompt, no
    timestamp rules, text only."""
    m, eng, H = _setup(language="en")
    wave = np.concatenate([H.speechlike(31, 300000), H.noise(32, 400000)])
    want, got, pipe = _both(m, eng, H, wave, rt)
    _same(want, got)
    import crisperwhisper_b200.generate as G
    rows, _ = G.init_token_template(eng.desc, G.GenOptions(return_timestamps=bool(rt)), 1)
    assert (rows[0][-1] == H.TOK_IDS["no_timestamps"]) == (not rt)


def test_multilingual_model_without_any_language_source_raises():
    """No language, no lang_to_id table, no pinned id: refuse instead of decoding with a prompt that lacks the language token."""
    import crisperwhisper_b200.generate as G
    from crisperwhisper_b200 import weights as Wt
    cfg = Wt.make_config(d_model=128, n_heads=2, enc_layers=1, dec_layers=1, ffn_dim=256, vocab=1865, n_mels=80, eos_id=256,
                         no_timestamps_id=363, alignment_heads=[[0, 0]], decoder_start_token_id=257, is_multilingual=True)
    with pytest.raises(ValueError):
        G.init_token_template(cfg, G.GenOptions(), 1)


def test_unsupported_generate_kwargs_fail_loudly():
    """Decoding strategies the device path does not implement are refused, not ignored (HF would change the transcript)."""
    import pytest
    from crisperwhisper_b200.asr_pipeline import AutomaticSpeechRecognitionPipeline as P
    P._check_generate_kwargs({"max_new_tokens": 8, "language": "en", "task": "transcribe", "temperature": 0.0, "num_beams": 1})
    for bad in ({"temperature": (0.0, 0.2, 0.4)}, {"temperature": 0.7}, {"do_sample": True}, {"num_beams": 5},
                {"logprob_threshold": -1.0}, {"compression_ratio_threshold": 1.35}, {"no_speech_threshold": 0.6},
                {"prompt_ids": [1, 2]}, {"condition_on_prev_tokens": True}):
        with pytest.raises(NotImplementedError):
            P._check_generate_kwargs(bad)
    with pytest.raises(ValueError):
        P._check_generate_kwargs({"top_k": 5})


def test_call_arguments_of_the_reference_pipeline(monkeypatch):
    """return_language / max_new_tokens / stride_length_s are call arguments of the reference pipeline
    (automatic_speech_recognition.py:262-300); anything else is a TypeError there and here."""
    from crisperwhisper_b200.asr_pipeline import AutomaticSpeechRecognitionPipeline as P
    seen = {}

    class Fake(P):
        def __init__(self):
            self.generate_kwargs, self.batch_size, self.chunk_length_s, self.return_timestamps = {}, 2, 30, "word"
            self.stride_length_s = None

        def _resample(self, x, sr):
            return x

        def _run(self, waves, cl, bs, gk, rt, stride_length_s="default"):
            seen.update(gk=gk, stride=stride_length_s, rt=rt)
            return [[] for _ in waves]

        def _postprocess(self, mo, rt, return_language=None):
            seen["lang"] = return_language
            return {"text": ""}

    pipe = Fake()
    pipe(np.zeros(16000, np.float32), max_new_tokens=7, stride_length_s=(4, 2), return_language=True, ignore_warning=True)
    assert seen["gk"]["max_new_tokens"] == 7 and seen["stride"] == (4, 2) and seen["lang"] is True and seen["rt"] == "word"
    import pytest
    with pytest.raises(TypeError):
        pipe(np.zeros(16000, np.float32), top_k=3)
