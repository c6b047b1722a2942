"""A CPU stand-in for crisperwhisper_b200.engine.Engine built entirely from the oracle (TEST INFRASTRUCTURE).

It lets the CPU test-suite drive the product's *host* logic (chunker, generate seek loop, segment retrieval,
post-processing) end to end against the reference pipeline's golden outputs, and it reports the smallest
top-1/top-2 score margin met during greedy decoding, so that golden cases free of near-ties can be selected for
the bf16 GPU end-to-end test.  It is never imported by the product."""
import numpy as np
import torch

from oracle import align as OA
from oracle import logmel as LM
from oracle import resample as RS
from oracle import whisper_ref as R


class OracleEngine:
    def __init__(self, sd, cfg):
        self.sd, self.desc = sd, cfg
        self.device = torch.device("cpu")
        self.min_margin = float("inf")

    def sync(self):
        pass

    def resample(self, wave, sr_in, sr_out=16000):
        return torch.from_numpy(RS.resample(wave.numpy(), int(sr_in), int(sr_out)))

    def logmel(self, wave, mel_filters, n_valid=None, want_f32=True, want_tm=True):
        w = wave.numpy()
        B = w.shape[0]
        nm = mel_filters.shape[0]
        feats = np.stack([LM.log_mel(w[b], mel_filters.numpy()) for b in range(B)])
        tm = torch.zeros(B, 3002, 128, dtype=torch.float32)
        tm[:, 1:3001, :nm] = torch.from_numpy(feats).permute(0, 2, 1)
        nv = n_valid.numpy() if n_valid is not None else np.full(B, 480000)
        frames = torch.tensor([LM.num_frames(int(x)) for x in nv], dtype=torch.int32)
        return (torch.from_numpy(feats) if want_f32 else None), tm, frames

    def encode(self, feats_tm, want_enc_out=False):
        nm = self.desc["n_mels"]
        feats = feats_tm[:, 1:3001, :nm].permute(0, 2, 1).contiguous().float()
        enc = R.encoder_forward(self.sd, self.desc, feats)
        return enc, enc

    def decode(self, xkv, prompt, max_new, flags=0, forced=None, want_logits=False, want_align=True):
        out = R.greedy_decode(self.sd, self.desc, xkv, prompt.numpy().astype(np.int64), max_new,
                              suppress_eos=bool(flags & 1), timestamp_rules=not (flags & 2), no_suppress=bool(flags & 64))
        B, n_prompt = prompt.shape
        toks = out["tokens"]
        n_gen = toks.shape[1] - n_prompt
        eos = self.desc["eos_id"]
        full = np.full((B, n_prompt + max_new), eos, np.int64)
        full[:, : toks.shape[1]] = toks
        lens = np.zeros(B, np.int32)
        for b in range(B):
            g = toks[b, n_prompt:]
            e = np.nonzero(g == eos)[0]
            lens[b] = n_prompt + (int(e[0]) + 1 if len(e) else n_gen)
        sc = np.sort(out["scores"], axis=-1)
        for b in range(B):  # margins only while the row is alive
            alive = int(lens[b] - n_prompt)
            m = (sc[b, :alive, -1] - sc[b, :alive, -2])
            if m.size:
                self.min_margin = min(self.min_margin, float(m.min()))
        H_a = len(self.desc["alignment_heads"])
        align = torch.zeros(B, H_a, max_new, 1500)
        align[:, :, :n_gen] = torch.from_numpy(out["align"])
        logits = None
        if want_logits:
            lg = np.full((B, max_new, out["scores"].shape[-1]), -np.inf, np.float32)
            lg[:, :n_gen] = out["scores"]
            logits = torch.from_numpy(lg)
        return dict(tokens=torch.from_numpy(full.astype(np.int32)), lengths=torch.from_numpy(lens),
                    align=align if want_align else None, logits=logits, argmax=None, steps=n_gen)

    def align(self, align, T_len, F_len, median_width=7):
        a = align.numpy()
        N, H, T_max, _ = a.shape
        out = np.zeros((N, T_max), np.int32)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for n in range(N):
                T, F = int(T_len[n]), int(F_len[n])
                if T > 0:
                    out[n, :T] = OA.jump_indices(a[n, :, :T, :F], median_width)
        return torch.from_numpy(out)
