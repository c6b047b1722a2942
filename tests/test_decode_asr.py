"""CPU parity of crisperwhisper_b200/decode_asr.py (token ids + token timestamps -> text / word chunks) against the
functions the reference's pipeline call reaches in the caller's tokenizer: HF tokenization_whisper._decode_asr and its
helpers.  Randomised token streams: text bytes (ASCII words, punctuation, multi-byte UTF-8, stray continuation bytes),
timestamp pairs, language switches, <|startofprev|> prompts, strided chunks whose text overlaps the previous chunk.
The bar is equality of the returned (text, optional) — floats included — or the same exception type."""
import numpy as np
import pytest

from oracle import hf_harness as H

TW = pytest.importorskip("transformers.models.whisper.tokenization_whisper")

IDS = H.TOK_IDS
TS0 = IDS["no_timestamps"] + 1


@pytest.fixture(scope="module")
def tok():
    return H.synthetic_tokenizer()


@pytest.fixture(scope="module")
def byte_id(tok):
    from crisperwhisper_b200.decode_asr import _CHAR_TO_BYTE
    inv = {b: c for c, b in _CHAR_TO_BYTE.items()}
    return [tok.convert_tokens_to_ids(inv[b]) for b in range(256)]


_SNIPPETS = [" the", " quick", " brown", " fox", ",", ".", " (", ")", " \"", "\"", "'s", " -", " ¿", "?", "!", " 你好", "世界",
             " émoji", " 🙂", "。", "，", " [UH]", " [UM]", "  ", " ", "\n", " a-b", " x", "yz", " über", " 1,5", ":", " {", "}"]


def _text_ids(rng, byte_id, n_items):
    out = []
    for _ in range(n_items):
        r = rng.random()
        if r < 0.88:
            out += [byte_id[b] for b in _SNIPPETS[int(rng.integers(len(_SNIPPETS)))].encode("utf-8")]
        elif r < 0.94:  # a multi-byte character cut short, or a stray continuation byte
            full = "世界🙂é"[int(rng.integers(4))].encode("utf-8")
            out += [byte_id[b] for b in full[: int(rng.integers(1, len(full) + 1))]]
        else:
            out.append(byte_id[int(rng.integers(0x80, 0x100))])
    return out


def _make_outputs(rng, byte_id, word_mode):
    n_out = int(rng.integers(1, 5))
    strided = n_out > 1 and rng.random() < 0.8
    outputs, tail = [], []
    for k in range(n_out):
        ids = []
        if rng.random() < 0.15:
            ids += [IDS["startofprev"]] + _text_ids(rng, byte_id, 2)
            if rng.random() < 0.8:
                ids += [IDS["sot"]]
        elif rng.random() < 0.9:
            ids += [IDS["sot"]]
        if rng.random() < 0.8:
            ids += [IDS["en"] + int(rng.integers(0, 3)) * int(rng.random() < 0.3), IDS["transcribe"]]
        t = int(rng.integers(0, 50))
        n_seg = int(rng.integers(1, 5))
        with_ts = rng.random() < 0.8
        for s in range(n_seg):
            if with_ts and rng.random() < 0.95:
                ids.append(TS0 + t)
            body = _text_ids(rng, byte_id, int(rng.integers(0, 9)))
            if s == 0 and tail and rng.random() < 0.8:  # repeat the end of the previous chunk (audio overlap)
                body = tail[-int(rng.integers(2, len(tail) + 1)):] + body if len(tail) >= 2 else body
                if rng.random() < 0.3 and body:
                    body[int(rng.integers(len(body)))] = byte_id[int(rng.integers(0x41, 0x5B))]
            ids += body
            if body:
                tail = body
            t += int(rng.integers(0, 400))
            if rng.random() < 0.1:
                t = int(rng.integers(0, 100))  # timestamps restart: concatenated 30 s segments
            t = min(t, 1500)
            if with_ts and rng.random() < 0.9:
                ids.append(TS0 + t)
                if rng.random() < 0.15:
                    ids.append(TS0 + t)
        if rng.random() < 0.7:
            ids.append(IDS["eos"])
        out = {"tokens": np.asarray([ids], dtype=np.int64)}
        if word_mode:
            steps = rng.integers(0, 40, len(ids)) * 0.02
            out["token_timestamps"] = np.cumsum(steps).astype(np.float32 if rng.random() < 0.5 else np.float64)[None, :]
        if strided:
            left = 0.0 if k == 0 else float(rng.integers(0, 6))
            right = 0.0 if k == n_out - 1 else float(rng.integers(0, 6))
            out["stride"] = (30.0 if rng.random() < 0.8 else float(rng.integers(8, 30)), left, right)
        outputs.append(out)
    return outputs


def _copy(outputs):
    return [{k: (v.copy() if hasattr(v, "copy") else v) for k, v in o.items()} for o in outputs]


def _run(fn, *a, **kw):
    try:
        return ("ok", fn(*a, **kw))
    except Exception as e:  # noqa: BLE001 — the exception type is part of the behaviour compared
        return ("err", type(e).__name__)


@pytest.mark.parametrize("return_timestamps,return_language", [("word", None), ("word", True), (True, None), (None, None),
                                                               (None, True), (True, True)])
def test_decode_asr_matches_hf_on_random_streams(tok, byte_id, return_timestamps, return_language):
    from crisperwhisper_b200.decode_asr import WordDecoder
    wd = WordDecoder(tok)
    assert wd._bytes_ok, "byte-level fast path must be active for a byte-level BPE vocabulary"
    rng = np.random.default_rng(1234 + 17 * len(str(return_timestamps)) + (3 if return_language else 0))
    n_ok = 0
    for case in range(250):
        outputs = _make_outputs(rng, byte_id, return_timestamps == "word")
        want = _run(TW._decode_asr, tok, _copy(outputs), return_timestamps=return_timestamps,
                    return_language=return_language, time_precision=0.02)
        got = _run(wd.decode_asr, _copy(outputs), return_timestamps=return_timestamps, return_language=return_language,
                   time_precision=0.02)
        assert got == want, (case, [o["tokens"].tolist() for o in outputs], got, want)
        n_ok += want[0] == "ok"
    assert n_ok > 200


def test_merge_overlaps_matches_hf(tok):
    from crisperwhisper_b200.decode_asr import merge_overlaps
    rng = np.random.default_rng(5)
    for case in range(400):
        n = int(rng.integers(1, 5))
        seqs, stamps = [], []
        base = rng.integers(0, 6, 64).tolist()
        pos = 0
        for _ in range(n):
            ln = int(rng.integers(0, 14))
            s = base[pos:pos + ln]
            if s and rng.random() < 0.4:
                s[int(rng.integers(len(s)))] = int(rng.integers(0, 6))
            seqs.append(list(s))
            t0 = np.round(np.cumsum(rng.integers(0, 3, len(s))) * 0.02 + pos * 0.02 * rng.integers(0, 2), 2)
            stamps.append([(float(a), float(a + 0.02 * rng.integers(0, 3))) for a in t0])
            pos += max(0, ln - int(rng.integers(0, 6)))
        assert merge_overlaps(seqs) == TW._find_longest_common_sequence([list(s) for s in seqs])
        got = merge_overlaps(seqs, stamps)
        want = TW._find_longest_common_sequence([list(s) for s in seqs], [list(s) for s in stamps])
        assert got == want, (case, seqs, stamps)
        assert merge_overlaps(seqs, []) == TW._find_longest_common_sequence([list(s) for s in seqs], [])


def test_word_split_matches_hf(tok, byte_id):
    from crisperwhisper_b200.decode_asr import WordDecoder
    wd = WordDecoder(tok)
    rng = np.random.default_rng(9)
    for case in range(400):
        ids = _text_ids(rng, byte_id, int(rng.integers(0, 12)))
        for lang in (None, "english", "chinese"):
            want = _run(TW._combine_tokens_into_words, tok, list(ids), lang)
            got = _run(wd.split_words, list(ids), lang)
            assert got == want, (case, lang, ids)


def test_tokenizer_fallback_path(tok, byte_id):
    """With the byte table disabled every decode goes through tokenizer.decode; results are unchanged."""
    from crisperwhisper_b200.decode_asr import WordDecoder
    fast, slow = WordDecoder(tok), WordDecoder(tok)
    slow._bytes_ok = False
    rng = np.random.default_rng(21)
    for _ in range(40):
        outputs = _make_outputs(rng, byte_id, True)
        kw = dict(return_timestamps="word", return_language=None, time_precision=0.02)
        assert _run(fast.decode_asr, _copy(outputs), **kw) == _run(slow.decode_asr, _copy(outputs), **kw)
