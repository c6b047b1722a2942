"""GPU parity tests (-m gpu) for the individual kernels, through the C-ABI (crisperwhisper_b200.engine -> libcrisper.so).
Checker: the CPU oracle (oracle/) on the same seeded inputs + the committed golden vectors from the reference."""
import os
import warnings

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _align_inputs(seed, N, H, T, F, peak=6.0, noise=3.0):
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((N, H, T, F))
    t = np.arange(T)[None, None, :, None]
    f = np.arange(F)[None, None, None, :]
    logit = noise * z + peak * np.exp(-(((f - t * F / max(T, 1)) / 20.0) ** 2))
    return torch.softmax(torch.from_numpy(logit.astype(np.float32)), -1).numpy()


# ---------------------------------------------------------------------------------------------- stage 3
@pytest.mark.parametrize("name", ["basic", "w3", "crop", "t1", "t2", "tinyF", "F4", "nancols", "t33", "w1"])
def test_align_vs_hf_golden(engine, name):
    """bit-exact token timestamps vs the reference's _extract_token_timestamps (golden vectors)."""
    g = np.load(os.path.join(GOLDEN, "align_hf.npz"))
    w, nf = g[f"{name}.w"], g[f"{name}.num_frames"]
    mw, n_prompt, ts = int(g[f"{name}.median"]), int(g[f"{name}.n_prompt"]), g[f"{name}.ts"]
    N, H, T, F = w.shape
    jump = engine.align(torch.from_numpy(w).cuda(), torch.full((N,), T), torch.from_numpy(nf.astype(np.int64) // 2), mw)
    engine.sync()
    jump = jump.cpu().numpy()
    got = np.zeros_like(ts)
    jt = (jump.astype(np.float64) * 0.02).astype(np.float32)
    got[:, n_prompt:n_prompt + T] = jt[:, :T]
    got[:, n_prompt + T] = jt[:, T - 1]
    assert np.array_equal(got, ts), (jump, ts)


@pytest.mark.parametrize("shape", [(3, 20, 100, 1500), (2, 15, 448, 1500), (4, 6, 37, 700), (2, 20, 129, 1499), (5, 1, 8, 33)])
def test_align_vs_oracle(engine, shape):
    from oracle import align as OA
    N, H, T, F = shape
    w = _align_inputs(7 + T, N, H, T, F)
    T_len = np.array([T - (n % 3) * (T // 5) for n in range(N)])
    F_len = np.array([F - (n % 2) * (F // 3) for n in range(N)])
    jump = engine.align(torch.from_numpy(w).cuda(), torch.from_numpy(T_len), torch.from_numpy(F_len), 7)
    engine.sync()
    jump = jump.cpu().numpy()
    for n in range(N):
        want = OA.jump_indices(w[n, :, :T_len[n], :F_len[n]], 7)
        assert np.array_equal(jump[n, :T_len[n]], want), f"utt {n}: {np.nonzero(jump[n, :T_len[n]] != want)[0][:10]}"
        assert (jump[n, T_len[n]:] == 0).all()


def test_align_full_size_properties(engine):
    """BASELINE cfg-5 shape (H=20, T=448, F=1500): monotone, in-range jump indices; equals the oracle on 1 utterance."""
    from oracle import align as OA
    w = _align_inputs(99, 2, 20, 448, 1500)
    jump = engine.align(torch.from_numpy(w).cuda(), torch.tensor([448, 448]), torch.tensor([1500, 1500]), 7)
    engine.sync()
    jump = jump.cpu().numpy()
    assert (np.diff(jump, axis=1) >= 0).all() and jump.min() >= 0 and jump.max() <= 1499
    assert np.array_equal(jump[0], OA.jump_indices(w[0], 7))


# ---------------------------------------------------------------------------------------------- stage 1
@pytest.mark.parametrize("nm", [80, 128])
def test_logmel_vs_hf_golden_and_oracle(engine, nm):
    """tolerance: 2e-4 abs on the (log10 + 4)/4 output scale (fp32 FFT vs the reference's fp32 pocketfft; the reference
    documents 1e-5 between its own two implementations, feature_extraction_whisper.py:107-108)."""
    from oracle import hf_harness as H
    from oracle import logmel as LM
    g = np.load(os.path.join(GOLDEN, "logmel_hf.npz"))
    filt = g[f"filters{nm}"]
    waves = {"noise7s": H.noise(11, 7 * 16000), "speech30s": H.speechlike(12), "short1s": H.noise(13, 16000),
             "long31s": H.noise(14, 31 * 16000)}
    names = list(waves)
    batch = np.stack([LM.pad_or_trim(waves[k]) for k in names])
    n_valid = torch.tensor([min(len(waves[k]), 480000) for k in names], dtype=torch.int32).cuda()
    feats, tm, frames = engine.logmel(torch.from_numpy(batch).cuda(), torch.from_numpy(np.ascontiguousarray(filt.T)).cuda(),
                                      n_valid)
    engine.sync()
    feats, tm, frames = feats.cpu().numpy(), tm.float().cpu().numpy(), frames.cpu().numpy()
    for i, k in enumerate(names):
        assert np.abs(feats[i][:, ::7] - g[f"{k}.{nm}.feats_sub"]).max() < 2e-4, k
        assert np.abs(feats[i][:, :4] - g[f"{k}.{nm}.first"]).max() < 2e-4, k
        assert np.abs(feats[i][:, -4:] - g[f"{k}.{nm}.last"]).max() < 2e-4, k
        assert frames[i] == int(g[f"{k}.{nm}.frames"])
        ref = LM.log_mel(batch[i], filt.T.astype(np.float32))
        assert np.abs(feats[i] - ref).max() < 2e-4, k
        # bf16 time-major copy: rows 1..3000 = frames, row 0 / 3001 and channels >= n_mels are zero
        assert np.abs(tm[i, 1:3001, :nm] - ref.T).max() < 1.2e-2
        assert (tm[i, 0] == 0).all() and (tm[i, 3001] == 0).all() and (tm[i, :, nm:] == 0).all()


# ---------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (300, 384, 192), (1500, 1280, 1280), (3000, 5120, 1280),
                                   (1000, 1280, 5120), (77, 16, 64), (129, 272, 320)])
def test_gemm_tcgen05_vs_check_kernel_and_torch(engine, M, N, K):
    """fp32-accumulate bf16 GEMM: tcgen05 kernel vs the CUDA-core checker (same arithmetic, different order)
    and vs torch fp32 matmul. Tolerance 2e-2 relative to the output scale (bf16 output rounding dominates)."""
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = (torch.randn(M, K, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device="cuda")
    res = torch.randn(M, N, generator=g, device="cuda")
    ref = A.float() @ W.float().T
    scale = ref.abs().max().item()
    for gelu, out_f32, use_bias, use_res in [(0, 1, 0, 0), (0, 0, 1, 0), (1, 0, 1, 0), (0, 1, 1, 1)]:
        c = engine.gemm(A, W, bias if use_bias else None, res if use_res else None, gelu, out_f32)
        k = engine.gemm(A, W, bias if use_bias else None, res if use_res else None, gelu, out_f32, check_kernel=True)
        engine.sync()
        r = ref + (bias if use_bias else 0)
        if gelu:
            r = torch.nn.functional.gelu(r)
        if use_res:
            r = r + res
        tol = (2e-5 if out_f32 else 1e-2) * max(scale, 1.0)
        assert (k.float() - r).abs().max().item() < tol, "checker kernel disagrees with torch"
        err = (c.float() - r).abs().max().item()
        assert err < tol, f"tcgen05 GEMM max err {err} (scale {scale}) cfg {(gelu, out_f32, use_bias, use_res)}"


def test_layernorm(engine):
    g = torch.Generator(device="cuda").manual_seed(0)
    for d in (128, 1280):
        x = torch.randn(333, d, generator=g, device="cuda") * 3 + 1
        gm, bt = torch.randn(d, generator=g, device="cuda"), torch.randn(d, generator=g, device="cuda")
        out = engine.layernorm(x, gm, bt)
        engine.sync()
        ref = torch.nn.functional.layer_norm(x, (d,), gm, bt, 1e-5)
        # bf16 output: half an ulp of the result (2^-9 relative) plus fp32 noise
        assert ((out.float() - ref).abs() <= ref.abs() * 2 ** -8 + 2e-3).all()


@pytest.mark.parametrize("B,S,H", [(1, 1500, 2), (2, 300, 3), (1, 64, 1), (2, 129, 2)])
def test_attention_enc(engine, B, S, H):
    """flash attention vs torch fp32 softmax(QK^T)V on the same bf16 inputs; tolerance 2e-2 abs (bf16 P and output)."""
    g = torch.Generator(device="cuda").manual_seed(S)
    d = H * 64
    qkv = (torch.randn(B * S, 3 * d, generator=g, device="cuda")).to(torch.bfloat16)
    qkv[:, :d] *= 0.25
    out = engine.attention_enc(qkv, B, S, H)
    engine.sync()
    x = qkv.float().view(B, S, 3, H, 64)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    ref = (torch.softmax(q @ k.transpose(2, 3), -1) @ v).transpose(1, 2).reshape(B * S, d)
    assert (out.float() - ref).abs().max().item() < 2e-2


def test_resample_vs_torchaudio_golden_and_oracle(engine):
    """cw_resample vs (a) torchaudio.functional.resample vectors (tests/golden/resample_ta.npz; the reference's preprocess
    call) and (b) oracle/resample.py: max abs err < 5e-5 / 2e-6 on amplitude ~0.5 signals; lengths equal;
    sr_in == sr_out and empty inputs pass through."""
    from oracle import resample as RS
    g = np.load(os.path.join(GOLDEN, "resample_ta.npz"))
    n_cases = len([k for k in g.files if k.endswith("_x")])
    for i in range(n_cases):
        x, y, sr = g[f"c{i}_x"], g[f"c{i}_y"], g[f"c{i}_sr"]
        got = engine.resample(torch.from_numpy(x).cuda(), int(sr[0]), int(sr[1]))
        engine.sync()
        got = got.cpu().numpy()
        assert got.shape == y.shape, (i, got.shape, y.shape)
        if y.size:
            assert np.abs(got - y).max() < 5e-5, (i, "vs torchaudio", np.abs(got - y).max())
            assert np.abs(got - RS.resample(x, int(sr[0]), int(sr[1]))).max() < 2e-6, (i, "vs oracle")
    e = engine.resample(torch.zeros(0, dtype=torch.float32).cuda(), 44100, 16000)
    assert e.numel() == 0


def test_resample_full_size_sine_property(engine):
    """30 s at 44.1 kHz -> 16 kHz (1 323 000 -> 480 000 samples): an in-band sine comes out as the same sine sampled at
    16 kHz (band-limited interpolation), |err| < 2e-3 away from the zero-padded edges; length = ceil(n * 160 / 441)."""
    sr, n = 44100, 44100 * 30
    t = np.arange(n, dtype=np.float64) / sr
    x = (0.5 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)
    y = engine.resample(torch.from_numpy(x).cuda(), sr, 16000)
    engine.sync()
    y = y.cpu().numpy()
    assert y.shape == (480000,)
    want = 0.5 * np.sin(2 * np.pi * 1000.0 * np.arange(480000) / 16000.0)
    assert np.abs(y[200:-200] - want[200:-200]).max() < 2e-3


def test_pipeline_resamples_other_rates(engine):
    """normalize_input routes a non-16 kHz dict input through cw_resample (HF preprocess :394-408)."""
    from crisperwhisper_b200 import audio as A
    from oracle import resample as RS
    x = (np.random.default_rng(3).standard_normal(8000) * 0.1).astype(np.float32)

    def rs(w, sr_in):
        out = engine.resample(torch.from_numpy(w).cuda(), sr_in, 16000)
        engine.sync()
        return out.cpu().numpy()
    got = A.normalize_input({"array": x, "sampling_rate": 8000}, rs)
    assert got.shape == (16000,) and np.abs(got - RS.resample(x, 8000, 16000)).max() < 2e-6
    with pytest.raises(ValueError):
        A.normalize_input({"array": x, "sampling_rate": 8000})
