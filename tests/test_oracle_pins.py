"""CPU tests (-m "not gpu"): every oracle restatement is pinned against (a) the committed golden vectors produced
by running the reference (installed transformers 5.5.0 + REF/utils.py) and (b) the live transformers functions."""
import copy
import json
import os
import warnings

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import align as A
from oracle import logits as LG
from oracle import logmel as LM
from oracle import postprocess as PP

ALIGN_CASES = ["basic", "w3", "crop", "t1", "t2", "tinyF", "F4", "nancols", "t33", "w1"]


@pytest.mark.parametrize("name", ALIGN_CASES)
def test_align_oracle_vs_hf_golden(name):
    g = np.load(os.path.join(GOLDEN, "align_hf.npz"))
    w, nf = g[f"{name}.w"], g[f"{name}.num_frames"]
    mw, n_prompt, ts = int(g[f"{name}.median"]), int(g[f"{name}.n_prompt"]), g[f"{name}.ts"]
    N, H, T, F = w.shape
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mine = A.extract_token_timestamps(w, [T] * N, [int(x) // 2 for x in nf], n_prompt, median_width=mw)
    assert mine.shape == ts.shape
    assert np.array_equal(mine, ts, equal_nan=True), (mine, ts)


def test_dtw_c_matches_literal_python():
    rng = np.random.default_rng(0)
    for T, F in [(1, 1), (1, 9), (7, 1), (12, 40), (5, 5)]:
        m = rng.standard_normal((T, F))
        ti, tj = A.dtw_python(m)
        ci, cj, jump = A.dtw(m)
        assert np.array_equal(ti, ci) and np.array_equal(tj, cj)
        jumps = np.pad(np.diff(ti), (1, 0), constant_values=1).astype(bool)
        assert np.array_equal(tj[jumps], jump)
    # all-equal costs: ties go left, column 0 goes up; all-NaN: time index -1
    _, _, j = A.dtw(np.zeros((4, 6)))
    assert j.tolist() == [0, 0, 0, 0]
    _, _, j = A.dtw(np.full((3, 5), np.nan))
    assert j.tolist() == [-1, -1, -1]


def test_dtw_matches_hf_function():
    from transformers.models.whisper.generation_whisper import _dynamic_time_warping
    rng = np.random.default_rng(1)
    m = rng.standard_normal((20, 60)).astype(np.float32).astype(np.float64)
    ti, tj = _dynamic_time_warping(m)
    ci, cj, _ = A.dtw(m)
    assert np.array_equal(ti, ci) and np.array_equal(tj, cj)


def test_median_filter_matches_hf():
    from transformers.models.whisper.generation_whisper import _median_filter
    rng = np.random.default_rng(2)
    for shape, w in [((3, 5, 40), 7), ((2, 4, 9), 3), ((2, 3, 3), 7), ((1, 2, 4), 7)]:
        x = rng.standard_normal(shape).astype(np.float32)
        ref = _median_filter(torch.from_numpy(x), w).numpy()
        assert np.array_equal(A.median_filter(x, w), ref)


def test_cost_matrix_bit_exact_vs_torch():
    """ATen summation order emulation (cascade_sum) + float64 std: bit-exact cost matrix on a contiguous case."""
    from transformers.models.whisper.generation_whisper import _median_filter
    rng = np.random.default_rng(3)
    w = torch.softmax(torch.from_numpy(rng.standard_normal((20, 64, 320)).astype(np.float32)) * 3, -1)
    std = torch.std(w, dim=-2, keepdim=True, unbiased=False)
    mean = torch.mean(w, dim=-2, keepdim=True)
    ref = _median_filter((w - mean) / std, 7).mean(dim=0).numpy()
    mine = A.cost_matrix(w.numpy(), 7)
    assert (mine == ref).mean() > 0.999  # thread-partition dependent remainders in ATen: see oracle/align.py header


@pytest.mark.parametrize("nm", [80, 128])
def test_logmel_oracle_vs_hf_golden(nm):
    from oracle import hf_harness as H
    g = np.load(os.path.join(GOLDEN, "logmel_hf.npz"))
    filt = g[f"filters{nm}"]
    assert np.abs(LM.mel_filter_bank(nm) - filt).max() < 1e-7
    waves = {"noise7s": H.noise(11, 7 * 16000), "speech30s": H.speechlike(12), "short1s": H.noise(13, 16000),
             "long31s": H.noise(14, 31 * 16000)}
    for name, wv in waves.items():
        out = LM.log_mel(LM.pad_or_trim(wv), filt.T.astype(np.float32))
        assert np.abs(out[:, ::7] - g[f"{name}.{nm}.feats_sub"]).max() < 5e-5, name
        assert np.abs(out[:, :4] - g[f"{name}.{nm}.first"]).max() < 5e-5
        assert np.abs(out[:, -4:] - g[f"{name}.{nm}.last"]).max() < 5e-5
        assert LM.num_frames(len(wv)) == int(g[f"{name}.{nm}.frames"])


def test_logits_oracle_vs_hf_golden():
    from oracle import hf_harness as H
    g = np.load(os.path.join(GOLDEN, "logits_hf.npz"))
    sup, bsup = g["suppress"].tolist(), g["begin_suppress"].tolist()
    for i in range(10):
        hist = g[f"c{i}.hist"].tolist()
        mi = int(g[f"c{i}.mi"])
        for v in range(2):
            out = LG.process(g[f"c{i}.{v}.in"], hist, begin=(len(hist) == 0), eos=H.TOK_IDS["eos"],
                             no_ts=H.TOK_IDS["no_timestamps"], suppress=sup, begin_suppress=bsup,
                             max_initial_timestamp_index=None if mi < 0 else mi)
            assert np.array_equal(out, g[f"c{i}.{v}.out"]), (i, v)


def test_pauses_oracle_vs_reference_golden():
    with open(os.path.join(GOLDEN, "pauses_ref.json")) as f:
        g = json.load(f)
    for key, case in g.items():
        thr = float(key.split("@")[1])
        inp = copy.deepcopy(case["input"])
        for c in inp["chunks"]:
            c["timestamp"] = tuple(c["timestamp"])
        out = PP.adjust_pauses(inp, split_threshold=thr)
        got = [list(c["timestamp"]) for c in out["chunks"]]
        want = [list(c["timestamp"]) for c in case["output"]["chunks"]]
        assert got == want, key  # exact float equality (REF/utils.py does not round)


def test_whisper_ref_vs_hf_module():
    """The torch fp32 restatement reproduces the HF module: encoder states, teacher-forced logits, cross-attention
    weights and the greedy token ids of generate()."""
    from transformers import WhisperFeatureExtractor
    from oracle import hf_harness as H
    from oracle import whisper_ref as R
    m = H.build_model(H.tiny_hf_config(), seed=0)
    sd = {k: v.float() for k, v in m.state_dict().items()}
    cfg = dict(n_heads=2, enc_layers=2, dec_layers=2, eos_id=H.TOK_IDS["eos"], no_timestamps_id=H.TOK_IDS["no_timestamps"],
               alignment_heads=m.generation_config.alignment_heads)
    fe = WhisperFeatureExtractor(feature_size=128)
    feats = torch.from_numpy(fe(H.noise(0, 80000), sampling_rate=16000, return_tensors="np")["input_features"])
    with torch.no_grad():
        enc_hf = m.model.encoder(feats).last_hidden_state
    enc = R.encoder_forward(sd, cfg, feats)
    assert (enc - enc_hf).abs().max() < 1e-5
    toks = torch.tensor([[257, 258, 359, 400, 70, 71, 500]])
    m.model.config._attn_implementation = "eager"  # what generate() forces for token timestamps (generation_whisper.py:706-707)
    with torch.no_grad():
        o = m(input_features=feats, decoder_input_ids=toks, output_attentions=True)
    logits, cross = R.decoder_forward(sd, cfg, enc, toks)
    assert (logits - o.logits).abs().max() < 1e-4
    assert (cross[1] - o.cross_attentions[1]).abs().max() < 1e-6
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = m.generate(feats, attention_mask=torch.ones(1, 3000, dtype=torch.long), return_timestamps=True,
                         return_token_timestamps=True, return_dict_in_generate=True, max_new_tokens=12,
                         force_unique_generate_call=True)
    ref = R.greedy_decode(sd, cfg, enc, np.array([[257, 258, 359]]), 12)
    assert ref["tokens"][0].tolist() == out["sequences"][0].tolist()
    # token timestamps from the oracle's alignment rows == HF's
    T = ref["tokens"].shape[1] - 3 - 1
    ts = A.extract_token_timestamps(ref["align"][:, :, :T], [T], [1500], 3)
    assert np.array_equal(ts[0], out["token_timestamps"][0].numpy())


def test_resample_oracle_vs_torchaudio_golden():
    """oracle/resample.py (float64 table and sums) vs torchaudio.functional.resample (float32 table, conv1d) on the
    committed vectors: max abs err < 5e-5 on signals of amplitude ~0.5; output lengths equal."""
    from oracle import resample as RS
    g = np.load(os.path.join(GOLDEN, "resample_ta.npz"))
    n_cases = len([k for k in g.files if k.endswith("_x")])
    assert n_cases >= 8
    for i in range(n_cases):
        x, y, sr = g[f"c{i}_x"], g[f"c{i}_y"], g[f"c{i}_sr"]
        mine = RS.resample(x, int(sr[0]), int(sr[1]))
        assert mine.shape == y.shape, (i, mine.shape, y.shape)
        if y.size:
            assert np.abs(mine - y).max() < 5e-5, (i, np.abs(mine - y).max())
