"""cw_words_from_tokens (csrc/postproc.cu, the C++ word decoder behind the C-ABI) against the Python module it mirrors
(crisperwhisper_b200/decode_asr.py, itself pinned to HF's tokenizer._decode_asr by tests/test_decode_asr.py) and against
HF directly, on the same randomised token streams: identical (text, chunks) — floats included — whenever the reference
answers; on inputs where the reference raises, the native path must hand the input back so that the same exception type
comes out.  Also: the UTF-8 replacement decoder and Python-style rounding inside the C++ file, probed through word texts."""
import numpy as np
import pytest

from oracle import hf_harness as H
from test_decode_asr import IDS, TS0, _copy, _make_outputs, _run

TW = pytest.importorskip("transformers.models.whisper.tokenization_whisper")


@pytest.fixture(scope="module")
def tok():
    return H.synthetic_tokenizer()


@pytest.fixture(scope="module")
def byte_id(tok):
    from crisperwhisper_b200.decode_asr import _CHAR_TO_BYTE
    inv = {b: c for c, b in _CHAR_TO_BYTE.items()}
    return [tok.convert_tokens_to_ids(inv[b]) for b in range(256)]


@pytest.mark.parametrize("return_language", [None, True])
def test_native_words_match_python_and_hf(tok, byte_id, return_language, monkeypatch):
    from crisperwhisper_b200.decode_asr import WordDecoder
    native, python = WordDecoder(tok), WordDecoder(tok)
    rng = np.random.default_rng(4321 + (7 if return_language else 0))
    n_ok = 0
    for case in range(400):
        outputs = _make_outputs(rng, byte_id, True)
        want = _run(TW._decode_asr, tok, _copy(outputs), return_timestamps="word", return_language=return_language,
                    time_precision=0.02)
        monkeypatch.setenv("CW_POSTPROC", "python")
        py = _run(python.decode_asr, _copy(outputs), return_timestamps="word", return_language=return_language, time_precision=0.02)
        monkeypatch.setenv("CW_POSTPROC", "native")
        got = _run(native.decode_asr, _copy(outputs), return_timestamps="word", return_language=return_language, time_precision=0.02)
        assert py == want and got == want, (case, [o["tokens"].tolist() for o in outputs], got, want)
        n_ok += want[0] == "ok"
    assert n_ok > 300
    assert python.native_calls == 0
    # the native path answered every stream the reference answers and handed back exactly those on which it raises
    assert native.native_calls == n_ok and native.native_punts == 400 - n_ok, (native.native_calls, native.native_punts, n_ok)


def test_native_utf8_replacement_and_units(tok, byte_id, monkeypatch):
    """Byte soup: truncated / overlong / surrogate / out-of-range sequences, stray continuation bytes, genuine U+FFFD."""
    from crisperwhisper_b200.decode_asr import WordDecoder
    native, python = WordDecoder(tok), WordDecoder(tok)
    rng = np.random.default_rng(99)
    pool = [b"\xe4\xb8", b"\xe4", b"\xb8", b"\xf0\x9f\x99", b"\xf0\x9f", b"\xc0\x80", b"\xe0\x80\x80", b"\xed\xa0\x80",
            b"\xf4\x90\x80\x80", b"\xf5", b"\xff", b"\xc2", b"\xef\xbf\xbd", b"\xe0\xa0", b"\xf0\x90\x80", b"\xc1\xbf", b" a", b"b",
            b" ", b".", b" (", b"\xe4\xb8\x96", b"\xf0\x9f\x99\x82", b"\xc2\xa0", b"\xe3\x80\x80", b"\xe2\x80\x89", b"\x1c", b"\t"]
    n_ok = 0
    for case in range(300):
        ids = [IDS["sot"], IDS["en"], IDS["transcribe"], TS0]
        for _ in range(int(rng.integers(1, 12))):
            ids += [byte_id[b] for b in pool[int(rng.integers(len(pool)))]]
        ids += [TS0 + 100, IDS["eos"]]
        out = [{"tokens": np.asarray([ids], dtype=np.int64),
                "token_timestamps": np.cumsum(rng.integers(0, 30, len(ids)) * 0.02).astype(np.float32)[None, :]}]
        want = _run(TW._decode_asr, tok, _copy(out), return_timestamps="word", return_language=None, time_precision=0.02)
        monkeypatch.setenv("CW_POSTPROC", "native")
        got = _run(native.decode_asr, _copy(out), return_timestamps="word", return_language=None, time_precision=0.02)
        assert got == want, (case, ids, got, want)
        n_ok += want[0] == "ok"
    # (the reference itself raises IndexError on some of this soup: those are the inputs the native path hands back)
    assert n_ok > 200 and native.native_calls == n_ok and native.native_punts == 300 - n_ok


def test_native_handles_pipeline_sized_batch_and_is_faster(tok, byte_id, monkeypatch):
    """8 chunks x 445 tokens (the benchmark's post-processing load): same answer, and the native path is the cheaper one."""
    import time
    from crisperwhisper_b200.decode_asr import WordDecoder
    wd = WordDecoder(tok)
    rng = np.random.default_rng(5)
    outputs = []
    for k in range(8):
        ids = [IDS["sot"], IDS["en"], IDS["transcribe"], TS0]
        t = 0
        while len(ids) < 440:
            ids += [byte_id[b] for b in (" w%d" % int(rng.integers(1000))).encode()]
            if rng.random() < 0.1:
                t = min(t + int(rng.integers(50, 200)), 1400)
                ids += [TS0 + t, TS0 + t]
        ids += [TS0 + 1500, IDS["eos"]]
        outputs.append({"tokens": np.asarray([ids], dtype=np.int64), "stride": (30.0, 0.0 if k == 0 else 5.0, 0.0 if k == 7 else 5.0),
                        "token_timestamps": np.minimum(np.cumsum(rng.integers(0, 8, len(ids)) * 0.02), 30.0).astype(np.float32)[None, :]})
    res = {}
    for mode in ("python", "native"):
        monkeypatch.setenv("CW_POSTPROC", mode)
        wd.decode_asr(_copy(outputs), return_timestamps="word", return_language=None, time_precision=0.02)   # warm (tables)
        t0 = time.perf_counter()
        for _ in range(5):
            res[mode] = wd.decode_asr(_copy(outputs), return_timestamps="word", return_language=None, time_precision=0.02)
        res[mode + "_s"] = (time.perf_counter() - t0) / 5
    assert res["native"] == res["python"]
    assert wd.native_punts == 0
    print(f"post-processing of 8 x 445 tokens: python {res['python_s'] * 1e3:.2f} ms, native {res['native_s'] * 1e3:.2f} ms")
    assert res["native_s"] < res["python_s"]
