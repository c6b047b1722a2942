"""CPU tests: the C-ABI library builds, loads and exports every symbol include/crisper.h declares (no compute)."""
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "crisper.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cw_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from crisperwhisper_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libcrisper.so does not export {n}"
        assert n in _lib.EXPORTS, f"{n} has no ctypes signature in crisperwhisper_b200/_lib.py"
    assert lib.cw_abi_version() == 2


def test_product_never_imports_oracle():
    """The product package must not route through the oracle or any CPU fallback."""
    pkg = os.path.join(ROOT, "crisperwhisper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "from oracle" not in txt and "import oracle" not in txt, f


def test_missing_library_fails_loudly(monkeypatch):
    from crisperwhisper_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libcrisper.so")
    try:
        _lib.load()
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("load() must raise when the CUDA library is missing")


def test_cross_attention_plan_covers_and_balances():
    """cw_decode_cross_plan (host-only): the step kernel's cross-attention stream plan. Every (sample, head) task's frames
    are covered exactly once by chunks of <= chunk_rows frames; a task is cut into splits[t] contiguous segments, every
    segment belongs to exactly one consumer group (first/last flags on its first/last chunk, consecutive segment indices),
    and while tasks <= 4 x n_cta every group owns exactly one segment. At the bench shape (160 tasks x 1500 frames, 80-row
    chunks, 148 CTAs) every CTA streams 20 or 21 chunks and a task has 3 or 4 segments."""
    import ctypes as C
    import numpy as np
    from crisperwhisper_b200 import _lib as L
    lib = L.load()
    for tasks, F, cr, n_cta in ((160, 1500, 80, 148), (6, 1500, 8, 148), (100, 1500, 80, 132), (320, 1500, 80, 148),
                                (1, 1500, 80, 1), (3, 37, 8, 5), (640, 1500, 80, 148)):
        cpt = -(-F // cr)
        items = np.full((tasks * cpt, 6), -7, dtype=np.int32)
        off = np.zeros(n_cta + 1, dtype=np.int32)
        splits = np.zeros(tasks, dtype=np.int32)
        rc = lib.cw_decode_cross_plan(tasks, F, cr, n_cta, items.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                      splits.ctypes.data_as(C.c_void_p))
        assert rc == 0, lib.cw_last_error()
        assert off[0] == 0 and off[-1] == tasks * cpt and (np.diff(off) >= 0).all()
        cover = np.zeros((tasks, F), dtype=np.int32)
        segs = {}
        group_segs = {}
        for c in range(n_cta):
            for t, f0, nf, g, sg, fl in items[off[c]:off[c + 1]]:
                assert 0 <= t < tasks and 0 < nf <= cr and f0 % cr == 0 and f0 + nf <= F and 0 <= g < 4
                cover[t, f0:f0 + nf] += 1
                segs.setdefault((t, sg), []).append((c, g, f0, nf, fl))
                group_segs.setdefault((c, g), set()).add((t, sg))
        assert (cover == 1).all()
        for t in range(tasks):
            assert sorted(sg for (tt, sg) in segs if tt == t) == list(range(splits[t]))
        for (t, sg), lst in segs.items():
            assert len({(c, g) for c, g, *_ in lst}) == 1                      # one group owns the whole segment
            f = [x[2] for x in lst]
            assert f == sorted(f) and all(f[i + 1] == f[i] + lst[i][3] for i in range(len(f) - 1))   # contiguous frames
            assert lst[0][4] & 1 and lst[-1][4] & 2 and all(not (x[4] & 1) for x in lst[1:]) and all(not (x[4] & 2) for x in lst[:-1])
        if tasks <= 4 * n_cta:
            assert max(len(v) for v in group_segs.values()) == 1               # one segment, one finalisation per group
        if (tasks, n_cta) == (160, 148):
            per_cta = np.diff(off)
            assert per_cta.min() == 20 and per_cta.max() == 21 and set(np.unique(splits)) == {3, 4}
    bad = np.zeros((10, 6), dtype=np.int32)
    assert lib.cw_decode_cross_plan(100, 1500, 0, 10, bad.ctypes.data_as(C.c_void_p), bad.ctypes.data_as(C.c_void_p),
                                    bad.ctypes.data_as(C.c_void_p)) != 0


def test_header_is_valid_c99_and_links_from_plain_c(tmp_path):
    """include/crisper.h compiles as C99 (-Wall -Wextra -pedantic, no warnings) and libcrisper.so answers a plain-C
    consumer (examples/c_abi_host_only.c) through its host-only entry points; a device entry point called without a
    context returns an error code and a message instead of crashing."""
    import os
    import shutil
    import subprocess
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    from crisperwhisper_b200 import _lib as L
    L.load()  # builds / checks the library is present
    exe = str(tmp_path / "c_abi_host_only")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                        os.path.join(root, "examples", "c_abi_host_only.c"), "-L" + os.path.join(root, "crisperwhisper_b200"),
                        "-lcrisper", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(root, "crisperwhisper_b200"))
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "resample_out_len 480000" in r.stdout and "worst_cta_chunks 21" in r.stdout and "ctx is NULL" in r.stdout
    assert "words 2: [ hi, 0.00-0.36] [ there 0.36-0.80]" in r.stdout, r.stdout
