"""Engine: the Python owner of one cw_ctx (one per process / GPU).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every compute call goes through the C-ABI of
libcrisper.so with raw pointers.  All methods raise RuntimeError on failure; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib as L


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Engine:
    N_FRAMES = 3000
    CHUNK = 480000
    F_ENC = 1500

    def __init__(self, device: int | str | torch.device = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("crisperwhisper_b200 needs a CUDA device (B200, sm_100a); none is visible")
        self.lib = L.load()
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise RuntimeError("crisperwhisper_b200 runs on CUDA devices only")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        h = C.c_void_p()
        L.check(self.lib.cw_init(self.device.index, C.byref(h)), "cw_init")
        self._h = h
        self.stream = torch.cuda.Stream(device=self.device)  # non-default stream: CUDA-graph capture needs one
        self._ws = {}
        self.weights = None
        self.desc = None

    def close(self):
        if getattr(self, "_h", None):
            self.lib.cw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -----------------------------------------------------------------------------------------
    def _workspace(self, key: str, nbytes: int) -> torch.Tensor:
        t = self._ws.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws[key] = t
        return t

    def _sp(self):
        return C.c_void_p(self.stream.cuda_stream)

    def _on_stream(self):
        """Context: work is enqueued on (and tensors allocated for) the engine stream, ordered after everything
        already queued on the caller's current stream; on exit the caller's stream waits for the engine stream."""
        eng = self

        class _Ctx:
            def __enter__(self_inner):
                self_inner.outer = torch.cuda.current_stream(eng.device)
                eng.stream.wait_stream(self_inner.outer)
                self_inner.cm = torch.cuda.stream(eng.stream)
                self_inner.cm.__enter__()

            def __exit__(self_inner, *a):
                self_inner.cm.__exit__(*a)
                self_inner.outer.wait_stream(eng.stream)
                return False

        return _Ctx()

    def launch_count(self) -> int:
        return int(self.lib.cw_launch_count(self._h))

    # -- weights -----------------------------------------------------------------------------------------
    def load_weights(self, packed) -> None:
        """packed: crisperwhisper_b200.weights.PackedWeights already on this device."""
        ptrs = (C.c_void_p * len(packed.tensors))(*[t.data_ptr() for t in packed.tensors])
        c = packed.config
        ah = np.ascontiguousarray(np.array(c["alignment_heads"], dtype=np.int32).reshape(-1, 2))
        sup = np.ascontiguousarray(np.array(c.get("suppress_tokens") or [], dtype=np.int32))
        bsup = np.ascontiguousarray(np.array(c.get("begin_suppress_tokens") or [], dtype=np.int32))
        d = L.ModelDesc()
        d.d_model, d.n_heads = c["d_model"], c["n_heads"]
        d.enc_layers, d.dec_layers, d.ffn_dim = c["enc_layers"], c["dec_layers"], c["ffn_dim"]
        d.vocab, d.vocab_padded, d.n_mels = c["vocab"], c["vocab_padded"], c["n_mels"]
        d.n_audio_ctx, d.n_text_ctx = c["n_audio_ctx"], c["n_text_ctx"]
        d.eos_id, d.no_timestamps_id = c["eos_id"], c["no_timestamps_id"]
        mi = c.get("max_initial_timestamp_index")
        d.max_initial_timestamp_index = -1 if mi is None else int(mi)
        d.median_filter_width = c["median_filter_width"]
        d.n_align_heads = ah.shape[0]
        d.align_heads_host = ah.ctypes.data_as(C.POINTER(C.c_int32))
        d.n_suppress = sup.size
        d.suppress_host = sup.ctypes.data_as(C.POINTER(C.c_int32))
        d.n_begin_suppress = bsup.size
        d.begin_suppress_host = bsup.ctypes.data_as(C.POINTER(C.c_int32))
        with torch.cuda.device(self.device):
            L.check(self.lib.cw_load_weights(self._h, ptrs, len(packed.tensors), C.byref(d)), "cw_load_weights")
        self.weights = packed  # keep the tensors alive: the library borrows the pointers
        self.desc = c
        # fragment-major copies of the decoder matrices for the streaming step kernel (caller-owned, like the weights)
        nb = int(self.lib.cw_decode_pack_bytes(self._h))
        with self._on_stream():
            self._pack = torch.empty(nb, dtype=torch.uint8, device=self.device)
            L.check(self.lib.cw_decode_pack(self._h, _p(self._pack), nb, self._sp()), "cw_decode_pack")
        self.stream.synchronize()

    # -- stage 1 -----------------------------------------------------------------------------------------
    def logmel(self, wave: torch.Tensor, mel_filters: torch.Tensor, n_valid: Optional[torch.Tensor] = None,
               want_f32: bool = True, want_tm: bool = True):
        """wave f32 [B, 480000] (device) -> (input_features f32 [B, n_mels, 3000] | None,
        feats_tm bf16 [B, 3002, 128] | None, frames i32 [B])."""
        assert wave.is_cuda and wave.dtype == torch.float32 and wave.dim() == 2 and wave.shape[1] == self.CHUNK
        wave = wave.contiguous()
        B = wave.shape[0]
        n_mels = mel_filters.shape[0]
        assert mel_filters.shape[1] == 201 and mel_filters.dtype == torch.float32 and mel_filters.is_cuda
        mel_filters = mel_filters.contiguous()
        with self._on_stream():
            feats = torch.empty(B, n_mels, self.N_FRAMES, dtype=torch.float32, device=self.device) if want_f32 else None
            tm = torch.empty(B, self.N_FRAMES + 2, 128, dtype=torch.bfloat16, device=self.device) if want_tm else None
            frames = torch.empty(B, dtype=torch.int32, device=self.device)
            nb = self.lib.cw_logmel_workspace_bytes(B, n_mels)
            ws = self._workspace("logmel", nb)
            L.check(self.lib.cw_logmel(self._h, _p(wave), _p(n_valid), _p(mel_filters), B, n_mels, _p(feats), _p(tm),
                                       _p(frames), _p(ws), ws.numel(), self._sp()), "cw_logmel")
        return feats, tm, frames

    # -- audio front-end -------------------------------------------------------------------------------------
    def resample(self, wave: torch.Tensor, sr_in: int, sr_out: int = 16000) -> torch.Tensor:
        """f32 [n] at sr_in -> f32 [ceil(n * sr_out / sr_in)] at sr_out on the device (cw_resample: the sinc/Hann
        interpolation torchaudio.functional.resample applies in the reference's preprocess)."""
        assert wave.is_cuda and wave.dtype == torch.float32 and wave.dim() == 1
        wave = wave.contiguous()
        n_in = wave.numel()
        n_out = int(self.lib.cw_resample_out_len(n_in, int(sr_in), int(sr_out)))
        if n_out < 0:
            raise ValueError(f"resample: bad rates {sr_in} -> {sr_out}")
        with self._on_stream():
            out = torch.empty(n_out, dtype=torch.float32, device=self.device)
            nb = self.lib.cw_resample_workspace_bytes(int(sr_in), int(sr_out))
            ws = self._workspace("resample", max(nb, 256))
            L.check(self.lib.cw_resample(self._h, _p(wave), n_in, int(sr_in), int(sr_out), _p(out), n_out, _p(ws), ws.numel(),
                                         self._sp()), "cw_resample")
        return out

    # -- stage 2a ----------------------------------------------------------------------------------------
    def encode(self, feats_tm: torch.Tensor, want_enc_out: bool = False):
        """feats_tm bf16 [B, 3002, 128] -> xkv bf16 [L_dec, B, H, 2, 1500, 64] (head-major; and enc_out bf16 [B,1500,d])."""
        c = self.desc
        B = feats_tm.shape[0]
        assert feats_tm.dtype == torch.bfloat16 and feats_tm.shape[1:] == (self.N_FRAMES + 2, 128)
        feats_tm = feats_tm.contiguous()
        with self._on_stream():
            xkv = torch.empty(c["dec_layers"], B, c["n_heads"], 2, self.F_ENC, 64, dtype=torch.bfloat16, device=self.device)
            enc = torch.empty(B, self.F_ENC, c["d_model"], dtype=torch.bfloat16, device=self.device) if want_enc_out else None
            nb = self.lib.cw_encode_workspace_bytes(self._h, B)
            ws = self._workspace("encode", nb)
            L.check(self.lib.cw_encode(self._h, _p(feats_tm), B, _p(enc), _p(xkv), _p(ws), ws.numel(), self._sp()),
                    "cw_encode")
        return xkv, enc

    # -- stage 2b ----------------------------------------------------------------------------------------
    def decode(self, xkv: torch.Tensor, prompt: torch.Tensor, max_new: int, flags: int = 0,
               forced: Optional[torch.Tensor] = None, want_logits: bool = False, want_align: bool = True):
        """Greedy decode.  Returns dict(tokens i32 [B, n_prompt+max_new], lengths i32 [B], align f32
        [B, H_a, max_new, 1500] | None, logits f32 [B, max_new, V] | None, argmax i32 [B, max_new], steps int)."""
        c = self.desc
        B, n_prompt = prompt.shape
        assert prompt.dtype == torch.int32 and prompt.is_cuda
        prompt = prompt.contiguous()
        H_a = len(c["alignment_heads"])
        if forced is not None:
            assert forced.dtype == torch.int32 and forced.shape == (B, max_new) and forced.is_cuda
            forced = forced.contiguous()
        steps = C.c_int(0)
        with self._on_stream():
            tokens = torch.empty(B, n_prompt + max_new, dtype=torch.int32, device=self.device)
            lens = torch.empty(B, dtype=torch.int32, device=self.device)
            align = (torch.empty(B, H_a, max_new, self.F_ENC, dtype=torch.float32, device=self.device)
                     if (want_align and H_a > 0) else None)
            logits = torch.empty(B, max_new, c["vocab"], dtype=torch.float32, device=self.device) if want_logits else None
            argmax = torch.zeros(B, max_new, dtype=torch.int32, device=self.device)
            nb = self.lib.cw_decode_workspace_bytes(self._h, B, max_new)
            ws = self._workspace("decode", nb)
            L.check(self.lib.cw_decode_greedy(self._h, _p(xkv), B, _p(prompt), n_prompt, max_new, flags, _p(forced),
                                              _p(tokens), _p(lens), _p(align), _p(logits), _p(argmax), C.byref(steps),
                                              _p(ws), ws.numel(), self._sp()), "cw_decode_greedy")
        return dict(tokens=tokens, lengths=lens, align=align, logits=logits, argmax=argmax, steps=steps.value)

    def decode_profile(self):
        """(ms[4], launches[4]) of the last CW_DEC_PROFILE decode: gemv, self-attn, cross-attn, other."""
        ms = (C.c_double * 4)()
        n = (C.c_longlong * 4)()
        L.check(self.lib.cw_decode_profile(self._h, ms, n), "cw_decode_profile")
        return list(ms), list(n)

    # -- stage 3 -----------------------------------------------------------------------------------------
    def align(self, align: torch.Tensor, T_len: torch.Tensor, F_len: torch.Tensor, median_width: int = 7) -> torch.Tensor:
        """align f32 [N, H, T_max, F_max]; T_len/F_len i32 [N] -> jump index i32 [N, T_max]."""
        assert align.is_cuda and align.dtype == torch.float32 and align.dim() == 4
        align = align.contiguous()
        N, H, T_max, F_max = align.shape
        T_len = T_len.to(device=self.device, dtype=torch.int32).contiguous()
        F_len = F_len.to(device=self.device, dtype=torch.int32).contiguous()
        nb = self.lib.cw_align_workspace_bytes(N, T_max, F_max)
        if nb == 0:
            raise RuntimeError(f"cw_align: unsupported shape T_max={T_max} F_max={F_max}")
        with self._on_stream():
            out = torch.empty(N, T_max, dtype=torch.int32, device=self.device)
            ws = self._workspace("align", nb)
            L.check(self.lib.cw_align(self._h, _p(align), _p(T_len), _p(F_len), N, H, T_max, F_max, median_width, _p(out),
                                      _p(ws), ws.numel(), self._sp()), "cw_align")
        return out

    # -- building blocks (tests / roofline) ----------------------------------------------------------------
    def gemm(self, A, W, bias=None, residual=None, gelu=False, out_f32=False, check_kernel=False):
        M, K = A.shape
        N = W.shape[0]
        A, W = A.contiguous(), W.contiguous()
        fn = self.lib.cw_gemm_bf16_check if check_kernel else self.lib.cw_gemm_bf16
        with self._on_stream():
            Cc = torch.empty(M, N, dtype=torch.float32 if out_f32 else torch.bfloat16, device=self.device)
            L.check(fn(self._h, _p(A), _p(W), _p(bias), _p(residual), _p(Cc), M, N, K, int(gelu), int(out_f32), self._sp()),
                    "cw_gemm_bf16")
        return Cc

    def attention_enc(self, qkv, B, S, n_heads):
        with self._on_stream():
            out = torch.empty(B * S, n_heads * 64, dtype=torch.bfloat16, device=self.device)
            L.check(self.lib.cw_attention_enc(self._h, _p(qkv.contiguous()), _p(out), B, S, n_heads, self._sp()),
                    "cw_attention_enc")
        return out

    def layernorm(self, x, g, b):
        M, d = x.shape
        with self._on_stream():
            out = torch.empty(M, d, dtype=torch.bfloat16, device=self.device)
            L.check(self.lib.cw_layernorm(self._h, _p(x.contiguous()), _p(g), _p(b), _p(out), M, d, self._sp()),
                    "cw_layernorm")
        return out

    def sync(self):
        self.stream.synchronize()
