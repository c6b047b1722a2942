"""CPU end-to-end test of the product's HOST logic (chunker, seek loop with HF batch semantics, segment retrieval,
stride merge through the tokenizer, pause adjustment) against the reference pipeline's golden outputs.  The compute
stages are supplied by the CPU oracle engine (tests/oracle_engine.py) — the product package itself never imports it."""
import copy
import json
import os
import warnings

import numpy as np
import pytest

from conftest import GOLDEN


def _waves():
    from oracle import hf_harness as H
    long70 = np.concatenate([H.speechlike(3), H.noise(4), H.noise(5, 160000)])
    return {"clip5s": H.noise(0, 80000), "clip70s": long70, "clip70s_bs1": long70, "clip12s_80": H.speechlike(6, 12 * 16000)}


@pytest.mark.parametrize("name", ["clip5s", "clip70s", "clip70s_bs1", "clip12s_80"])
def test_host_logic_on_oracle_engine_matches_reference_pipeline(name):
    from oracle_engine import OracleEngine
    from oracle import hf_harness as H
    from crisperwhisper_b200 import adjust_pauses_for_hf_pipeline_output, pipeline
    from crisperwhisper_b200 import weights as Wt
    from transformers import WhisperFeatureExtractor
    with open(os.path.join(GOLDEN, "pipeline_hf.json")) as f:
        g = json.load(f)[name]
    m = H.build_model(H.tiny_hf_config(n_mels=g["n_mels"]), seed=g["seed"], logit_scale=g["logit_scale"], pos_scale=g["pos_scale"])
    cfg = Wt.config_from_hf(m)
    cfg["lang_id"], cfg["task_id"] = H.TOK_IDS["en"], H.TOK_IDS["transcribe"]
    eng = OracleEngine({k: v.float() for k, v in m.state_dict().items()}, cfg)
    pipe = pipeline("automatic-speech-recognition", model=eng, tokenizer=H.synthetic_tokenizer(),
                    feature_extractor=WhisperFeatureExtractor(feature_size=g["n_mels"]), chunk_length_s=30,
                    batch_size=g["batch_size"], return_timestamps="word")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = pipe(_waves()[name].copy(), generate_kwargs={"max_new_tokens": g["max_new_tokens"]})
    want = g["pipeline"]
    assert out["text"] == want["text"]
    assert [(c["text"], tuple(c["timestamp"])) for c in out["chunks"]] == [(c["text"], tuple(c["timestamp"])) for c in want["chunks"]]
    adj = adjust_pauses_for_hf_pipeline_output(copy.deepcopy(out))
    assert [tuple(c["timestamp"]) for c in adj["chunks"]] == [tuple(c["timestamp"]) for c in g["adjusted"]["chunks"]]
    assert pipe.last_stats["generate_passes"] >= 1


def test_inputs_at_other_rates_go_through_the_engine_resampler():
    """{"array", "sampling_rate": 8000} and a WAV file at 8 kHz are resampled by engine.resample (cw_resample on the GPU;
    here the oracle) before chunking: same result as feeding the resampled 16 kHz array (HF preprocess :394-408)."""
    import io
    from scipy.io import wavfile
    from oracle_engine import OracleEngine
    from oracle import hf_harness as H
    from oracle import resample as RS
    from crisperwhisper_b200 import pipeline
    from crisperwhisper_b200 import weights as Wt
    m = H.build_model(H.tiny_hf_config(n_mels=80), seed=1, logit_scale=8.0, pos_scale=20.0)
    cfg = Wt.config_from_hf(m)
    cfg["lang_id"], cfg["task_id"] = H.TOK_IDS["en"], H.TOK_IDS["transcribe"]
    eng = OracleEngine({k: v.float() for k, v in m.state_dict().items()}, cfg)
    pipe = pipeline("automatic-speech-recognition", model=eng, tokenizer=H.synthetic_tokenizer(), chunk_length_s=30,
                    batch_size=1, return_timestamps="word")
    x8 = H.speechlike(9, 3 * 8000)
    gk = {"max_new_tokens": 6}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = pipe({"array": x8.copy(), "sampling_rate": 8000}, generate_kwargs=gk)
        b = pipe(RS.resample(x8, 8000, 16000), generate_kwargs=gk)
        buf = io.BytesIO()
        wavfile.write(buf, 8000, x8)
        c = pipe(buf.getvalue(), generate_kwargs=gk)
    assert a == b == c
