"""CPU tests (-m "not gpu") of the host logic against the reference's own functions (installed transformers):
chunk plan vs chunk_iter, retrieve_segment vs WhisperGenerationMixin._retrieve_segment, mel filters, utils."""
import copy
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


def test_chunk_plan_matches_hf_chunk_iter():
    from transformers.pipelines.automatic_speech_recognition import chunk_iter
    from crisperwhisper_b200 import audio as A

    class FE:  # records what chunk_iter would feed the feature extractor
        sampling_rate = 16000

        def __call__(self, chunk, **kw):
            return {"n": len(chunk)}

    for n in (80000, 480000, 480001, 800000, 1120000, 9600000, 320000 * 3 + 80000, 320000 * 3 + 80001):
        x = np.zeros(n, np.float32)
        ref = [(it["stride"], it["is_last"]) for it in chunk_iter(x, FE(), 480000, 80000, 80000)]
        mine = [((ln, l, r), last) for (_, ln, l, r, last) in A.chunk_plan(n, 30.0)]
        assert mine == ref, n
    plan = A.chunk_plan(9600000, 30.0)  # 10 min -> 30 chunks
    assert len(plan) == 30 and plan[1][:4] == (320000, 480000, 80000, 80000) and plan[-1][:4] == (9280000, 320000, 80000, 0)


@pytest.mark.parametrize("seq", [
    [364, 70, 71, 400, 400, 72, 73, 500],          # consecutive timestamps, single timestamp ending
    [364, 70, 71, 400, 400, 72, 73],               # consecutive timestamps, unfinished tail
    [364, 70, 71, 400, 400, 72, 500, 500],         # ends with a double timestamp
    [364, 70, 71, 72],                             # no closing timestamp
    [364, 70, 71, 420],                            # single timestamp ending, no consecutive pair
    [364, 364],                                    # only timestamps
    [70, 71],                                      # no timestamp at all
    [364, 70, 380, 380, 71, 390, 390, 72, 400, 400, 73],
])
def test_retrieve_segment_matches_hf(seq):
    from transformers.models.whisper.generation_whisper import WhisperGenerationMixin
    from crisperwhisper_b200 import generate as G
    ts_begin, n_prompt = 364, 3
    rng = np.random.default_rng(len(seq))
    tt = np.round(rng.uniform(0, 30, n_prompt + len(seq)) / 0.02) * 0.02
    tt = tt.astype(np.float32)
    for seek_frames, off in ((3000, 0.0), (1800, 12.0)):
        time_offset = torch.tensor([off], dtype=torch.float64)
        ref_segs, ref_off = WhisperGenerationMixin._retrieve_segment(
            seek_sequence=torch.tensor(seq), seek_outputs=[{"token_timestamps": torch.from_numpy(tt)}],
            time_offset=time_offset, timestamp_begin=ts_begin, seek_num_frames=torch.tensor([seek_frames]),
            time_precision=0.02, time_precision_features=0.01, input_stride=2, prev_idx=0, idx=0,
            return_token_timestamps=True, decoder_input_ids=torch.zeros(1, n_prompt, dtype=torch.long))
        segs, seg_off = G.retrieve_segment(np.array(seq), tt, off, ts_begin, seek_frames, n_prompt)
        assert int(ref_off) == int(seg_off)
        assert len(segs) == len(ref_segs)
        for a, b in zip(segs, ref_segs):
            assert a["tokens"].tolist() == b["tokens"].tolist()
            assert tuple(a["idxs"]) == tuple(b["idxs"])
            assert float(a["start"]) == float(b["start"]) and float(a["end"]) == float(b["end"])
            assert np.array_equal(a["token_timestamps"], b["token_timestamps"].numpy())


def test_mel_filters_match_hf():
    from transformers import WhisperFeatureExtractor
    from crisperwhisper_b200.asr_pipeline import mel_filters_slaney
    for nm in (80, 128):
        ref = WhisperFeatureExtractor(feature_size=nm).mel_filters.T.astype(np.float32)
        assert np.array_equal(mel_filters_slaney(nm), ref)


def test_adjust_pauses_matches_reference_golden():
    from crisperwhisper_b200 import adjust_pauses_for_hf_pipeline_output
    with open(os.path.join(GOLDEN, "pauses_ref.json")) as f:
        g = json.load(f)
    for key, case in g.items():
        thr = float(key.split("@")[1])
        inp = copy.deepcopy(case["input"])
        for c in inp["chunks"]:
            c["timestamp"] = tuple(c["timestamp"])
        first = inp["chunks"][0] if inp["chunks"] else None
        out = adjust_pauses_for_hf_pipeline_output(inp, split_threshold=thr)
        assert [list(c["timestamp"]) for c in out["chunks"]] == [list(c["timestamp"]) for c in case["output"]["chunks"]], key
        assert out is inp and (first is None or out["chunks"][0] is first)  # mutates its argument like the reference


def test_weight_packing_roundtrip_shapes():
    from crisperwhisper_b200 import weights as Wt
    from oracle import hf_harness as H
    m = H.build_model(H.tiny_hf_config(n_mels=80), seed=0)
    cfg = Wt.config_from_hf(m)
    pw = Wt.pack_state_dict(m.state_dict(), cfg)
    shapes = Wt.slot_shapes(cfg)
    assert len(pw.tensors) == len(shapes) == 13 + 2 * 12 + 2 * 18
    slots = {n: t for (n, _, _), t in zip(shapes, pw.tensors)}
    sd = m.state_dict()
    d = cfg["d_model"]
    # q is pre-scaled by 1/8 exactly; k has no bias
    assert torch.equal(slots["enc0.WQKV"][:d].float(), (sd["model.encoder.layers.0.self_attn.q_proj.weight"] * 0.125).to(torch.bfloat16).float())
    assert torch.count_nonzero(slots["dec1.BQKV"][d:2 * d]) == 0
    # conv1 is tap-major with channels padded to 128
    w = sd["model.encoder.conv1.weight"]
    assert torch.equal(slots["CONV1_W"].view(d, 3, 128)[:, 1, :80].float(), w[:, :, 1].to(torch.bfloat16).float())
    assert torch.count_nonzero(slots["CONV1_W"].view(d, 3, 128)[:, :, 80:]) == 0
    assert slots["TOK_EMB"].shape[0] % 128 == 0 and torch.count_nonzero(slots["TOK_EMB"][cfg["vocab"]:]) == 0


def test_timestamps_to_vtt_matches_reference_golden():
    """crisperwhisper_b200.utils.timestamps_to_vtt == REF/app.py:74-82 on tests/golden/vtt_ref.json (generated from the
    reference function): unpadded hours, per-field rounding of the seconds, one cue per word."""
    import json
    import os
    from conftest import GOLDEN
    from crisperwhisper_b200.utils import timestamps_to_vtt
    with open(os.path.join(GOLDEN, "vtt_ref.json")) as f:
        g = json.load(f)
    assert len(g) >= 5
    for name, case in g.items():
        assert timestamps_to_vtt(case["input"]) == case["output"], name


def test_chunk_plan_matches_hf_chunk_iter_randomised():
    """Random lengths, chunk lengths and stride pairs. HF's preprocess rounds chunk and strides to multiples of
    `_align_to` = getattr(model.config, "inputs_to_logits_ratio", 1), which is 1 for a WhisperConfig
    (automatic_speech_recognition.py:328-338,435-438): plain rounding to samples, as chunk_plan does."""
    from transformers.pipelines.automatic_speech_recognition import chunk_iter
    from crisperwhisper_b200 import audio as A

    class FE:
        sampling_rate = 16000

        def __call__(self, chunk, **kw):
            return {"n": len(chunk)}

    rng = np.random.default_rng(7)
    for _ in range(200):
        n = int(rng.integers(1, 3_000_000))
        chunk_s = float(rng.choice([5.0, 12.5, 30.0]))
        sl = float(rng.choice([0.5, 1.0, chunk_s / 6]))
        sr_ = float(rng.choice([0.5, 1.0, chunk_s / 6]))
        align_to = 1
        chunk_len = int(round(chunk_s * 16000 / align_to) * align_to)
        left = int(round(sl * 16000 / align_to) * align_to)
        right = int(round(sr_ * 16000 / align_to) * align_to)
        x = np.zeros(n, np.float32)
        ref = [(it["stride"], it["is_last"]) for it in chunk_iter(x, FE(), chunk_len, left, right)]
        mine = [((ln, l, r), last) for (_, ln, l, r, last) in A.chunk_plan(n, chunk_s, (sl, sr_))]
        assert mine == ref, (n, chunk_s, sl, sr_)


def test_retrieve_segment_matches_hf_randomised():
    """Random token streams over {text, timestamp} with random timestamp placement, seek window and time offset."""
    from transformers.models.whisper.generation_whisper import WhisperGenerationMixin
    from crisperwhisper_b200 import generate as G
    ts_begin, n_prompt = 364, 3
    rng = np.random.default_rng(11)
    for case in range(300):
        L = int(rng.integers(1, 24))
        t = int(rng.integers(0, 200))
        seq = []
        for _ in range(L):
            if rng.random() < 0.35:
                t = min(1500, t + int(rng.integers(0, 120)))
                seq.append(ts_begin + t)
                if rng.random() < 0.5:
                    seq.append(ts_begin + t)
            else:
                seq.append(int(rng.integers(0, 256)))
        tt = (np.round(rng.uniform(0, 30, n_prompt + len(seq)) / 0.02) * 0.02).astype(np.float32)
        seek_frames = int(rng.choice([3000, 2400, 1801, 600]))
        off = float(rng.choice([0.0, 12.0, 29.98]))
        ref_segs, ref_off = WhisperGenerationMixin._retrieve_segment(
            seek_sequence=torch.tensor(seq), seek_outputs=[{"token_timestamps": torch.from_numpy(tt)}],
            time_offset=torch.tensor([off], dtype=torch.float64), timestamp_begin=ts_begin,
            seek_num_frames=torch.tensor([seek_frames]), time_precision=0.02, time_precision_features=0.01, input_stride=2,
            prev_idx=0, idx=0, return_token_timestamps=True, decoder_input_ids=torch.zeros(1, n_prompt, dtype=torch.long))
        segs, seg_off = G.retrieve_segment(np.array(seq), tt, off, ts_begin, seek_frames, n_prompt)
        assert int(ref_off) == int(seg_off), (case, seq)
        assert len(segs) == len(ref_segs), (case, seq)
        for a, b in zip(segs, ref_segs):
            assert a["tokens"].tolist() == b["tokens"].tolist(), (case, seq)
            assert tuple(a["idxs"]) == tuple(b["idxs"])
            assert float(a["start"]) == float(b["start"]) and float(a["end"]) == float(b["end"]), (case, seq)
            assert np.array_equal(a["token_timestamps"], b["token_timestamps"].numpy()), (case, seq)


def test_process_audio_bytes_matches_reference_app():
    """REF/app.py:85-96 restated: raw sample values, (y - mean) / std, / 8, resample to 16 kHz, shape [1, n]."""
    import io
    from scipy.io import wavfile
    from crisperwhisper_b200 import audio as A
    rs = np.random.RandomState(4)
    pcm = (rs.randn(4000) * 3000).astype(np.int16)
    buf = io.BytesIO()
    wavfile.write(buf, 16000, pcm)
    got = A.process_audio_bytes(buf.getvalue())
    y = pcm.astype(np.float32)
    want = ((y - np.mean(y)) / np.std(y)) / 8
    assert got.shape == (1, 4000) and got.dtype == np.float32 and np.array_equal(got[0], want.astype(np.float32))
    buf = io.BytesIO()
    wavfile.write(buf, 8000, pcm)
    calls = []
    got = A.process_audio_bytes(buf.getvalue(), lambda x, sr: (calls.append((len(x), sr)), np.repeat(x, 2))[1])
    assert calls == [(4000, 8000)] and got.shape == (1, 8000)
    with pytest.raises(ValueError):
        A.process_audio_bytes(buf.getvalue())
