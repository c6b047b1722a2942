"""oracle/whisper_ref.py — plain PyTorch fp32 restatement of the Whisper encoder / decoder forward and of the greedy
loop with the Whisper logits processors.  TEST INFRASTRUCTURE ONLY.

Follows HF/models/whisper/modeling_whisper.py (transformers 5.5.0):
  eager_attention_forward :215-238, WhisperAttention.forward :284-357 (q scaled by head_dim**-0.5 before QK^T :310,
  k_proj without bias :279), WhisperEncoderLayer :380-414, WhisperEncoder :593-647 (conv1/GELU/conv2(stride 2)/GELU,
  + embed_positions :625, final layer_norm :643), WhisperDecoderLayer :449-506, WhisperDecoder :691-796
  (embed_tokens + embed_positions :738-763, layer_norm :791), tied proj_out :966,1081;
and HF/generation/utils.py:2743-2800 (fp32 logits, processors, argmax, eos->pad bookkeeping).
It works directly on a WhisperForConditionalGeneration state dict (fp32 tensors) and is pinned against the HF
module itself in tests/test_oracle_pins.py.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import logits as logits_oracle


def _pre(sd):
    return "model." if any(k.startswith("model.") for k in sd) else ""


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _heads(x, H):
    B, T, D = x.shape
    return x.view(B, T, H, D // H).transpose(1, 2)  # [B, H, T, hd]


def _attn(sd, p, xq, xkv, H, mask=None, kv=None):
    """WhisperAttention.forward: returns (output, attn_weights[B,H,Tq,Tk]).  `kv` = (k, v) already projected and split
    into heads (the cross-attention cache of HF, modeling_whisper.py:326-336) — same arithmetic, computed once."""
    hd = xq.shape[-1] // H
    q = (F.linear(xq, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])) * (hd ** -0.5)
    q = _heads(q, H)
    if kv is None:
        k = _heads(F.linear(xkv, sd[p + "k_proj.weight"]), H)
        v = _heads(F.linear(xkv, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"]), H)
    else:
        k, v = kv
    w = torch.matmul(q, k.transpose(2, 3))
    if mask is not None:
        w = w + mask
    w = torch.softmax(w, dim=-1)
    o = torch.matmul(w, v).transpose(1, 2).reshape(xq.shape[0], xq.shape[1], -1)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"]), w


@torch.no_grad()
def encoder_forward(sd: Dict[str, torch.Tensor], cfg: Dict, feats: torch.Tensor) -> torch.Tensor:
    """feats f32 [B, n_mels, 3000] -> last_hidden_state f32 [B, 1500, d]."""
    E = _pre(sd) + "encoder."
    x = F.gelu(F.conv1d(feats, sd[E + "conv1.weight"], sd[E + "conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, sd[E + "conv2.weight"], sd[E + "conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1) + sd[E + "embed_positions.weight"]
    H = cfg["n_heads"]
    for l in range(cfg["enc_layers"]):
        P = f"{E}layers.{l}."
        h = _ln(x, sd, P + "self_attn_layer_norm")
        a, _ = _attn(sd, P + "self_attn.", h, h, H)
        x = x + a
        h = _ln(x, sd, P + "final_layer_norm")
        x = x + F.linear(F.gelu(F.linear(h, sd[P + "fc1.weight"], sd[P + "fc1.bias"])), sd[P + "fc2.weight"], sd[P + "fc2.bias"])
    return _ln(x, sd, E + "layer_norm")


@torch.no_grad()
def cross_kv(sd, cfg, enc_out: torch.Tensor):
    """The cross-attention K/V of every decoder layer (HF caches them after the first step, modeling_whisper.py:326-336):
    list over layers of (k, v), each f32 [B, H, 1500, 64]."""
    D = _pre(sd) + "decoder."
    H = cfg["n_heads"]
    out = []
    for l in range(cfg["dec_layers"]):
        P = f"{D}layers.{l}.encoder_attn."
        out.append((_heads(F.linear(enc_out, sd[P + "k_proj.weight"]), H),
                    _heads(F.linear(enc_out, sd[P + "v_proj.weight"], sd[P + "v_proj.bias"]), H)))
    return out


@torch.no_grad()
def decoder_forward(sd, cfg, enc_out: torch.Tensor, tokens: torch.Tensor, xkv_cache=None):
    """Full-sequence (teacher-forced) decoder pass.  tokens i64 [B, T].
    Returns logits f32 [B, T, V] and cross-attention weights: list over layers of [B, H, T, 1500]."""
    D = _pre(sd) + "decoder."
    B, T = tokens.shape
    x = F.embedding(tokens, sd[D + "embed_tokens.weight"]) + sd[D + "embed_positions.weight"][:T]
    mask = torch.full((T, T), float("-inf")).triu(1)[None, None]
    H = cfg["n_heads"]
    cross = []
    for l in range(cfg["dec_layers"]):
        P = f"{D}layers.{l}."
        h = _ln(x, sd, P + "self_attn_layer_norm")
        a, _ = _attn(sd, P + "self_attn.", h, h, H, mask)
        x = x + a
        h = _ln(x, sd, P + "encoder_attn_layer_norm")
        a, w = _attn(sd, P + "encoder_attn.", h, enc_out, H, kv=None if xkv_cache is None else xkv_cache[l])
        cross.append(w)
        x = x + a
        h = _ln(x, sd, P + "final_layer_norm")
        x = x + F.linear(F.gelu(F.linear(h, sd[P + "fc1.weight"], sd[P + "fc1.bias"])), sd[P + "fc2.weight"], sd[P + "fc2.bias"])
    x = _ln(x, sd, D + "layer_norm")
    return F.linear(x, sd[D + "embed_tokens.weight"]).float(), cross


@torch.no_grad()
def greedy_decode(sd, cfg, enc_out: torch.Tensor, prompt: np.ndarray, max_new: int, *, suppress_eos: bool = False,
                  timestamp_rules: bool = True, forced: Optional[np.ndarray] = None, xkv_cache=None, no_suppress: bool = False):
    """Greedy loop with the Whisper logits processors.  Returns dict with
       tokens  i64 [B, n_prompt + n_gen]   (finished rows padded with eos, like HF `sequences`)
       scores  f32 [B, n_gen, V]           processed scores of every step (HF `scores`)
       argmax  i64 [B, n_gen]
       align   f32 [B, H_a, n_gen, 1500]   alignment-head cross-attention rows; row s = query whose input is
                                           generated token s (rows of the prompt already dropped, row n_gen-1 of the
                                           last step is present here although HF never computes it)
    Every step recomputes the full prefix (O(T^2)) — fine for the small oracle configurations."""
    B, n_prompt = prompt.shape
    eos, no_ts = cfg["eos_id"], cfg["no_timestamps_id"]
    suppress = ([] if no_suppress else list(cfg.get("suppress_tokens") or [])) + ([eos] if suppress_eos else [])
    seq = torch.from_numpy(prompt.astype(np.int64))
    finished = np.zeros(B, bool)
    scores_all, argmax_all = [], []
    for step in range(max_new):
        logits, _ = decoder_forward(sd, cfg, enc_out, seq, xkv_cache)
        last = logits[:, -1].numpy()
        proc = np.stack([
            logits_oracle.process(last[b], seq[b, n_prompt:].tolist(), begin=(seq.shape[1] == n_prompt), eos=eos,
                                  no_ts=no_ts, suppress=suppress, begin_suppress=[] if no_suppress else (cfg.get("begin_suppress_tokens") or []),
                                  max_initial_timestamp_index=cfg.get("max_initial_timestamp_index"),
                                  timestamp_rules=timestamp_rules)
            for b in range(B)])
        am = proc.argmax(-1)
        nxt = am.copy() if forced is None else forced[:, step].astype(np.int64)
        nxt[finished] = eos
        scores_all.append(proc)
        argmax_all.append(am)
        seq = torch.cat([seq, torch.from_numpy(nxt.astype(np.int64))[:, None]], 1)
        finished |= (nxt == eos)
        if finished.all() and forced is None and not suppress_eos:
            break
    n_gen = seq.shape[1] - n_prompt
    # alignment rows: one teacher-forced pass over the final sequence (identical to the incremental rows)
    _, cross = decoder_forward(sd, cfg, enc_out, seq, xkv_cache)
    heads = cfg["alignment_heads"]
    if heads:
        align = torch.stack([cross[l][:, h] for l, h in heads], 1)[:, :, n_prompt:, :].numpy()
    else:
        align = np.zeros((B, 0, n_gen, enc_out.shape[1]), np.float32)
    return dict(tokens=seq.numpy(), scores=np.stack(scores_all, 1), argmax=np.stack(argmax_all, 1), align=align)
