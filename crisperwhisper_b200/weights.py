"""Weight packing: HF Whisper state dict -> the slot table libcrisper.so borrows (include/crisper.h, CW_W_* enums).

Replaces `model.to(device)` of the reference (REF/transcribe.py:14-17).  All matrices are stored bf16 in one arena and
all vectors f32 in another, so that multi-GPU start-up is two NCCL broadcasts (distributed.py).  Packing rules (each
one exact, no rounding beyond the bf16 cast the model dtype implies):
  * q_proj weight and bias are multiplied by head_dim**-0.5 = 0.125 — a power of two, so `(xW+b)*0.125`
    (HF/models/whisper/modeling_whisper.py:310) equals `x(0.125W)+0.125b` bit for bit in floating point;
  * q/k/v weights of self-attention are concatenated row-wise, k_proj has no bias (:279) -> zero rows;
  * conv weights [out, in, 3] become tap-major [out, 3 * in_padded] (in_padded = 128 for conv1);
  * the cross-attention k_proj / v_proj of all decoder layers are stacked: [dec_layers * 2 * d, d];
  * the tied token embedding is padded with zero rows to a multiple of 128.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from ._lib import W_DEC_LAYER, W_ENC_LAYER, W_GLOBAL

MELS_PADDED = 128


@dataclass
class PackedWeights:
    tensors: List[torch.Tensor]  # slot order of include/crisper.h
    arena_bf16: torch.Tensor
    arena_f32: torch.Tensor
    config: Dict

    def to(self, device) -> "PackedWeights":
        return _rebind(self, self.arena_bf16.to(device), self.arena_f32.to(device))


def slot_shapes(c: Dict):
    """[(name, dtype, shape)] in slot order — shared by the packer, the synthetic generator and the broadcaster."""
    d, ffn, L_e, L_d = c["d_model"], c["ffn_dim"], c["enc_layers"], c["dec_layers"]
    bf, f32 = torch.bfloat16, torch.float32
    g = {
        "CONV1_W": (bf, (d, 3 * MELS_PADDED)), "CONV1_B": (f32, (d,)),
        "CONV2_W": (bf, (d, 3 * d)), "CONV2_B": (f32, (d,)),
        "ENC_POS": (f32, (c["n_audio_ctx"], d)), "ENC_LNF_G": (f32, (d,)), "ENC_LNF_B": (f32, (d,)),
        "XKV_W": (bf, (L_d * 2 * d, d)), "XKV_B": (f32, (L_d * 2 * d,)),
        "TOK_EMB": (bf, (c["vocab_padded"], d)), "DEC_POS": (f32, (c["n_text_ctx"], d)),
        "DEC_LNF_G": (f32, (d,)), "DEC_LNF_B": (f32, (d,)),
    }
    el = {
        "LN1_G": (f32, (d,)), "LN1_B": (f32, (d,)), "WQKV": (bf, (3 * d, d)), "BQKV": (f32, (3 * d,)),
        "WO": (bf, (d, d)), "BO": (f32, (d,)), "LN2_G": (f32, (d,)), "LN2_B": (f32, (d,)),
        "W1": (bf, (ffn, d)), "B1": (f32, (ffn,)), "W2": (bf, (d, ffn)), "B2": (f32, (d,)),
    }
    dl = dict(el)
    dl.update({"WQC": (bf, (d, d)), "BQC": (f32, (d,)), "WOC": (bf, (d, d)), "BOC": (f32, (d,)),
               "LN3_G": (f32, (d,)), "LN3_B": (f32, (d,))})
    out = [(n,) + g[n] for n in W_GLOBAL]
    for l in range(L_e):
        out += [(f"enc{l}.{n}",) + el[n] for n in W_ENC_LAYER]
    for l in range(L_d):
        out += [(f"dec{l}.{n}",) + dl[n] for n in W_DEC_LAYER]
    return out


def _alloc(c: Dict, device):
    shapes = slot_shapes(c)
    nb = sum(_pad(_numel(s)) for _, dt, s in shapes if dt == torch.bfloat16)
    nf = sum(_pad(_numel(s)) for _, dt, s in shapes if dt == torch.float32)
    return torch.zeros(nb, dtype=torch.bfloat16, device=device), torch.zeros(nf, dtype=torch.float32, device=device)


def _numel(s):
    n = 1
    for x in s:
        n *= x
    return n


def _pad(n, a=128):
    return (n + a - 1) // a * a  # keeps every slot 256-byte aligned (TMA needs 16)


def _rebind(proto_or_cfg, arena_bf16, arena_f32) -> PackedWeights:
    c = proto_or_cfg.config if isinstance(proto_or_cfg, PackedWeights) else proto_or_cfg
    tensors = []
    ob = of = 0
    for _, dt, s in slot_shapes(c):
        n = _numel(s)
        if dt == torch.bfloat16:
            tensors.append(arena_bf16[ob:ob + n].view(*s))
            ob += _pad(n)
        else:
            tensors.append(arena_f32[of:of + n].view(*s))
            of += _pad(n)
    return PackedWeights(tensors, arena_bf16, arena_f32, c)


def empty_packed(config: Dict, device) -> PackedWeights:
    """Zero-filled weights of the right layout (receive side of the NCCL broadcast)."""
    a, b = _alloc(config, device)
    return _rebind(config, a, b)


def make_config(*, d_model, n_heads, enc_layers, dec_layers, ffn_dim, vocab, n_mels, eos_id, no_timestamps_id,
                alignment_heads, n_audio_ctx=1500, n_text_ctx=448, median_filter_width=7, suppress_tokens=None,
                begin_suppress_tokens=None, max_initial_timestamp_index=None, decoder_start_token_id=None,
                lang_id=None, task_id=None, **extra) -> Dict:
    if d_model != 64 * n_heads:
        raise ValueError("crisperwhisper_b200 requires head_dim == 64 (true for every Whisper size)")
    c = dict(d_model=d_model, n_heads=n_heads, enc_layers=enc_layers, dec_layers=dec_layers, ffn_dim=ffn_dim,
             vocab=vocab, vocab_padded=_pad(vocab, 128), n_mels=n_mels, n_audio_ctx=n_audio_ctx, n_text_ctx=n_text_ctx,
             eos_id=eos_id, no_timestamps_id=no_timestamps_id, alignment_heads=[list(map(int, x)) for x in alignment_heads],
             median_filter_width=median_filter_width, suppress_tokens=list(suppress_tokens or []),
             begin_suppress_tokens=list(begin_suppress_tokens or []),
             max_initial_timestamp_index=max_initial_timestamp_index, decoder_start_token_id=decoder_start_token_id,
             lang_id=lang_id, task_id=task_id)
    c.update(extra)
    return c


def config_from_hf(model, generation_config=None) -> Dict:
    """Read dims from a HF WhisperConfig and the generation fields the logits processors use
    (HF/models/whisper/generation_whisper.py:1774-1812)."""
    hc = model.config
    gc = generation_config if generation_config is not None else model.generation_config
    heads = getattr(gc, "alignment_heads", None)
    if heads is None:
        raise ValueError("generation_config.alignment_heads is required for word timestamps "
                         "(HF raises the same at generation_whisper.py:1348-1352)")
    eos = gc.eos_token_id if gc.eos_token_id is not None else hc.eos_token_id
    if isinstance(eos, (list, tuple)):
        eos = eos[0]
    # prompt-token resolution fields (language / task / forced ids; generate.init_token_template follows HF's
    # _retrieve_init_tokens, generation_whisper.py:1455-1608)
    extra = dict(lang_to_id=dict(getattr(gc, "lang_to_id", None) or {}), task_to_id=dict(getattr(gc, "task_to_id", None) or {}),
                 language=getattr(gc, "language", None), task=getattr(gc, "task", None),
                 is_multilingual=bool(getattr(gc, "is_multilingual", False)),
                 forced_decoder_ids=getattr(gc, "forced_decoder_ids", None) or getattr(hc, "forced_decoder_ids", None))
    return make_config(
        **extra,
        d_model=hc.d_model, n_heads=hc.encoder_attention_heads, enc_layers=hc.encoder_layers,
        dec_layers=hc.decoder_layers, ffn_dim=hc.encoder_ffn_dim, vocab=hc.vocab_size, n_mels=hc.num_mel_bins,
        eos_id=int(eos), no_timestamps_id=int(gc.no_timestamps_token_id), alignment_heads=heads,
        n_audio_ctx=hc.max_source_positions, n_text_ctx=hc.max_target_positions,
        median_filter_width=getattr(hc, "median_filter_width", 7),
        suppress_tokens=getattr(gc, "suppress_tokens", None), begin_suppress_tokens=getattr(gc, "begin_suppress_tokens", None),
        max_initial_timestamp_index=getattr(gc, "max_initial_timestamp_index", None),
        decoder_start_token_id=hc.decoder_start_token_id)


@torch.no_grad()
def pack_state_dict(sd: Dict[str, torch.Tensor], config: Dict, device="cpu") -> PackedWeights:
    """sd: state dict of WhisperForConditionalGeneration (keys `model.encoder...`, `model.decoder...`)."""
    c = config
    d, L_e, L_d = c["d_model"], c["enc_layers"], c["dec_layers"]
    pw = empty_packed(c, "cpu")
    slots = {name: t for (name, _, _), t in zip(slot_shapes(c), pw.tensors)}
    f = lambda k: sd[k].detach().to(torch.float32).cpu()
    pre = "model." if any(k.startswith("model.") for k in sd) else ""
    E, D = pre + "encoder.", pre + "decoder."

    def put(name, val):
        slots[name].copy_(val.to(slots[name].dtype))

    w1 = f(E + "conv1.weight")  # [d, n_mels, 3]
    w1p = torch.zeros(d, 3, MELS_PADDED)
    w1p[:, :, : w1.shape[1]] = w1.permute(0, 2, 1)
    put("CONV1_W", w1p.reshape(d, 3 * MELS_PADDED))
    put("CONV1_B", f(E + "conv1.bias"))
    put("CONV2_W", f(E + "conv2.weight").permute(0, 2, 1).reshape(d, 3 * d))
    put("CONV2_B", f(E + "conv2.bias"))
    put("ENC_POS", f(E + "embed_positions.weight"))
    put("ENC_LNF_G", f(E + "layer_norm.weight"))
    put("ENC_LNF_B", f(E + "layer_norm.bias"))
    emb = f(D + "embed_tokens.weight")
    slots["TOK_EMB"][: emb.shape[0]].copy_(emb.to(torch.bfloat16))
    put("DEC_POS", f(D + "embed_positions.weight"))
    put("DEC_LNF_G", f(D + "layer_norm.weight"))
    put("DEC_LNF_B", f(D + "layer_norm.bias"))

    def attn(prefix, dst_prefix, names):
        q_w, q_b = f(prefix + "q_proj.weight") * 0.125, f(prefix + "q_proj.bias") * 0.125
        k_w = f(prefix + "k_proj.weight")
        v_w, v_b = f(prefix + "v_proj.weight"), f(prefix + "v_proj.bias")
        put(dst_prefix + names[0], torch.cat([q_w, k_w, v_w], 0))
        put(dst_prefix + names[1], torch.cat([q_b, torch.zeros(d), v_b], 0))
        put(dst_prefix + names[2], f(prefix + "out_proj.weight"))
        put(dst_prefix + names[3], f(prefix + "out_proj.bias"))

    def ffn(prefix, dst_prefix):
        put(dst_prefix + "W1", f(prefix + "fc1.weight")); put(dst_prefix + "B1", f(prefix + "fc1.bias"))
        put(dst_prefix + "W2", f(prefix + "fc2.weight")); put(dst_prefix + "B2", f(prefix + "fc2.bias"))

    for l in range(L_e):
        P, Q = f"{E}layers.{l}.", f"enc{l}."
        put(Q + "LN1_G", f(P + "self_attn_layer_norm.weight")); put(Q + "LN1_B", f(P + "self_attn_layer_norm.bias"))
        attn(P + "self_attn.", Q, ["WQKV", "BQKV", "WO", "BO"])
        put(Q + "LN2_G", f(P + "final_layer_norm.weight")); put(Q + "LN2_B", f(P + "final_layer_norm.bias"))
        ffn(P, Q)
    xw, xb = [], []
    for l in range(L_d):
        P, Q = f"{D}layers.{l}.", f"dec{l}."
        put(Q + "LN1_G", f(P + "self_attn_layer_norm.weight")); put(Q + "LN1_B", f(P + "self_attn_layer_norm.bias"))
        attn(P + "self_attn.", Q, ["WQKV", "BQKV", "WO", "BO"])
        put(Q + "LN2_G", f(P + "encoder_attn_layer_norm.weight")); put(Q + "LN2_B", f(P + "encoder_attn_layer_norm.bias"))
        put(Q + "WQC", f(P + "encoder_attn.q_proj.weight") * 0.125); put(Q + "BQC", f(P + "encoder_attn.q_proj.bias") * 0.125)
        put(Q + "WOC", f(P + "encoder_attn.out_proj.weight")); put(Q + "BOC", f(P + "encoder_attn.out_proj.bias"))
        put(Q + "LN3_G", f(P + "final_layer_norm.weight")); put(Q + "LN3_B", f(P + "final_layer_norm.bias"))
        ffn(P, Q)
        xw += [f(P + "encoder_attn.k_proj.weight"), f(P + "encoder_attn.v_proj.weight")]
        xb += [torch.zeros(d), f(P + "encoder_attn.v_proj.bias")]
    put("XKV_W", torch.cat(xw, 0))
    put("XKV_B", torch.cat(xb, 0))
    return pw.to(device) if str(device) != "cpu" else pw


def pack_hf_model(model, generation_config=None, device="cpu") -> PackedWeights:
    return pack_state_dict(model.state_dict(), config_from_hf(model, generation_config), device)


@torch.no_grad()
def synthetic_weights(config: Dict, device, seed: int = 0, std: float = 0.02) -> PackedWeights:
    """Random-init weights of the right architecture generated directly on `device` (no checkpoint is available
    offline — BASELINE.md §2): matrices ~ N(0, std), LayerNorm gamma 1 / beta 0, biases ~ N(0, std), sinusoidal
    encoder positions (modeling_whisper.py:55-64)."""
    import math
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    a, b = _alloc(config, device)
    pw = _rebind(config, a, b)
    for (name, dt, shape), t in zip(slot_shapes(config), pw.tensors):
        base = name.split(".")[-1]
        if base.endswith("_G"):
            t.fill_(1.0)
        elif base.startswith("LN") or base.endswith("LNF_B"):
            t.zero_()
        elif base == "ENC_POS":
            n, ch = shape
            inc = math.log(10000.0) / (ch // 2 - 1)
            inv = torch.exp(-inc * torch.arange(ch // 2, device=device, dtype=torch.float32))
            st = torch.arange(n, device=device, dtype=torch.float32)[:, None] * inv[None, :]
            t.copy_(torch.cat([st.sin(), st.cos()], 1))
        else:
            t.copy_((torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std).to(dt))
    d = config["d_model"]
    # structural zeros the real packing has
    slots = {name: t for (name, _, _), t in zip(slot_shapes(config), pw.tensors)}
    slots["TOK_EMB"][config["vocab"]:].zero_()
    slots["CONV1_W"].view(d, 3, MELS_PADDED)[:, :, config["n_mels"]:].zero_()
    xb = slots["XKV_B"].view(config["dec_layers"], 2, d)
    xb[:, 0].zero_()
    for l in range(config["enc_layers"]):
        slots[f"enc{l}.BQKV"][d:2 * d].zero_()
    for l in range(config["dec_layers"]):
        slots[f"dec{l}.BQKV"][d:2 * d].zero_()
    return pw


def large_v3_config(n_align_heads: int = 20, median_filter_width: int = 7) -> Dict:
    """Whisper large-v3 shape with the token-id layout derived in SURVEY §8c (CrisperWhisper's real
    generation_config is not available offline; alignment heads follow SURVEY §8d cfg 2)."""
    heads = [[l, (7 * l) % 20] for l in range(32 - n_align_heads, 32)]
    return make_config(d_model=1280, n_heads=20, enc_layers=32, dec_layers=32, ffn_dim=5120, vocab=51866, n_mels=128,
                       eos_id=50257, no_timestamps_id=50364, alignment_heads=heads, median_filter_width=median_filter_width,
                       decoder_start_token_id=50258, lang_id=50259, task_id=50360)
