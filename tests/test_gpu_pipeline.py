"""GPU end-to-end tests (-m gpu): the drop-in pipeline against the reference pipeline's golden outputs
(tests/golden/pipeline_hf.json: HF transformers 5.5.0 CPU fp32 + REF/utils.py on a seeded random-init Whisper)."""
import copy
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _waves():
    from oracle import hf_harness as H
    long70 = np.concatenate([H.speechlike(3), H.noise(4), H.noise(5, 160000)])
    return {"clip5s": H.noise(0, 80000), "clip70s": long70, "clip70s_bs1": long70, "clip12s_80": H.speechlike(6, 12 * 16000)}


@pytest.mark.parametrize("name", ["clip5s", "clip70s", "clip70s_bs1", "clip12s_80"])
def test_pipeline_matches_reference_golden(engine, name):
    """text and word chunks identical; word timestamps identical (they are multiples of 20 ms, so the ±10 ms bar of
    BASELINE.json means exact equality); adjust_pauses output identical."""
    from crisperwhisper_b200 import adjust_pauses_for_hf_pipeline_output, pipeline
    from crisperwhisper_b200 import weights as Wt
    from oracle import hf_harness as H
    from transformers import WhisperFeatureExtractor
    with open(os.path.join(GOLDEN, "pipeline_hf.json")) as f:
        g = json.load(f)[name]
    m = H.build_model(H.tiny_hf_config(n_mels=g["n_mels"]), seed=g["seed"], logit_scale=g["logit_scale"], pos_scale=g["pos_scale"])
    pw = Wt.pack_hf_model(m, device=engine.device)
    pw.config["lang_id"], pw.config["task_id"] = H.TOK_IDS["en"], H.TOK_IDS["transcribe"]
    engine.load_weights(pw)
    pipe = pipeline("automatic-speech-recognition", model=engine, tokenizer=H.synthetic_tokenizer(),
                    feature_extractor=WhisperFeatureExtractor(feature_size=g["n_mels"]), chunk_length_s=30,
                    batch_size=g["batch_size"], return_timestamps="word")
    out = pipe(_waves()[name].copy(), generate_kwargs={"max_new_tokens": g["max_new_tokens"]})
    want = g["pipeline"]
    assert out["text"] == want["text"]
    got_chunks = [(c["text"], tuple(c["timestamp"])) for c in out["chunks"]]
    want_chunks = [(c["text"], tuple(c["timestamp"])) for c in want["chunks"]]
    assert got_chunks == want_chunks
    adj = adjust_pauses_for_hf_pipeline_output(copy.deepcopy(out))
    assert [tuple(c["timestamp"]) for c in adj["chunks"]] == [tuple(c["timestamp"]) for c in g["adjusted"]["chunks"]]
