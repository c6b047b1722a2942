"""oracle/hf_harness.py — builds the reference path itself out of the installed `transformers` (5.5.0): a random-init
WhisperForConditionalGeneration, a synthetic byte-level WhisperTokenizer, and the reference's exact
`pipeline("automatic-speech-recognition", ..., return_timestamps="word")` call (REF/transcribe.py:21-31) with greedy
pinned.  TEST INFRASTRUCTURE ONLY: used by tests/, tests/golden/make_golden.py and the
`--impl reference` / cpu_baseline legs of bench.py.  No checkpoint or tokenizer files exist offline, hence synthetic.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch


def synthetic_tokenizer():
    """256 byte symbols, Whisper specials (eot=256, sot=257, languages, ..., notimestamps=363) and the
    1501 timestamp tokens 364..1864."""
    from tokenizers import AddedToken
    from tokenizers.pre_tokenizers import ByteLevel
    from transformers import WhisperTokenizer
    from transformers.models.whisper.tokenization_whisper import LANGUAGES

    vocab = {ch: i for i, ch in enumerate(sorted(ByteLevel.alphabet()))}
    specials = ["<|endoftext|>", "<|startoftranscript|>"] + [f"<|{l}|>" for l in LANGUAGES] + \
               ["<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nospeech|>", "<|notimestamps|>"]
    for s in specials:
        vocab[s] = len(vocab)
    tok = WhisperTokenizer(vocab=vocab, merges=[], language="en", task="transcribe", additional_special_tokens=specials[1:])
    tok.add_tokens([AddedToken("<|%.2f|>" % (i * 0.02), special=False, normalized=False) for i in range(1501)])
    tok.pad_token = "<|endoftext|>"
    return tok


TOK_IDS = dict(eos=256, sot=257, en=258, translate=358, transcribe=359, startoflm=360, startofprev=361, nospeech=362,
               no_timestamps=363, vocab=1865)


def tiny_hf_config(n_mels: int = 128, d_model: int = 128, heads: int = 2, layers: int = 2, ffn: int = 512,
                   vocab: int = TOK_IDS["vocab"], median_filter_width: int = 7):
    from transformers import WhisperConfig
    return WhisperConfig(vocab_size=vocab, num_mel_bins=n_mels, d_model=d_model, encoder_layers=layers, decoder_layers=layers,
                         encoder_attention_heads=heads, decoder_attention_heads=heads, encoder_ffn_dim=ffn,
                         decoder_ffn_dim=ffn, decoder_start_token_id=TOK_IDS["sot"], eos_token_id=TOK_IDS["eos"],
                         pad_token_id=TOK_IDS["eos"], bos_token_id=TOK_IDS["eos"], begin_suppress_tokens=None,
                         suppress_tokens=None, median_filter_width=median_filter_width)


def large_v3_hf_config(median_filter_width: int = 7):
    from transformers import WhisperConfig
    return WhisperConfig(vocab_size=51866, num_mel_bins=128, d_model=1280, encoder_layers=32, decoder_layers=32,
                         encoder_attention_heads=20, decoder_attention_heads=20, encoder_ffn_dim=5120, decoder_ffn_dim=5120,
                         decoder_start_token_id=50258, eos_token_id=50257, pad_token_id=50257, bos_token_id=50257,
                         begin_suppress_tokens=None, suppress_tokens=None, median_filter_width=median_filter_width)


def build_model(hf_config, seed: int = 0, alignment_heads: Optional[List[List[int]]] = None, ids: Optional[Dict] = None,
                logit_scale: float = 1.0, max_new_tokens: Optional[int] = None, suppress_tokens=None,
                begin_suppress_tokens=None, bf16_round: bool = True, pos_scale: float = 1.0, fast_init: bool = False):
    """Random-init model whose parameters are rounded to bf16 and upcast (so the fp32 oracle and the bf16 kernels
    share values, SURVEY §7 'hard parts'), with the generation_config fields Whisper's generate needs."""
    from transformers import WhisperForConditionalGeneration
    torch.manual_seed(seed)
    if fast_init:  # large models: skip HF's slow per-module init, fill the parameters directly (values are irrelevant for timing)
        from transformers.initialization import no_init_weights
        with no_init_weights():
            m = WhisperForConditionalGeneration(hf_config).eval()
        with torch.no_grad():
            for name, p in m.named_parameters():
                if p.dim() >= 2:
                    p.normal_(0.0, 0.02)
                elif name.endswith("layer_norm.weight") or name.endswith("layernorm.weight"):
                    p.fill_(1.0)
                else:
                    p.zero_()
            m.model.encoder.embed_positions.weight.copy_(torch.randn_like(m.model.encoder.embed_positions.weight) * 0.02)
    else:
        m = WhisperForConditionalGeneration(hf_config).eval()
    with torch.no_grad():
        if logit_scale != 1.0:  # enlarge the (tied) embedding to create logit margin between tokens
            m.model.decoder.embed_tokens.weight.mul_(logit_scale)
        if pos_scale != 1.0:  # make the decoder output depend strongly on the position: varied tokens per step
            m.model.decoder.embed_positions.weight.mul_(pos_scale)
        if bf16_round:
            for p in m.parameters():
                p.copy_(p.to(torch.bfloat16).to(torch.float32))
    ids = ids or TOK_IDS
    gc = m.generation_config
    gc.no_timestamps_token_id = ids["no_timestamps"]
    gc.lang_to_id = {"<|en|>": ids["en"]}
    gc.task_to_id = {"transcribe": ids["transcribe"], "translate": ids["translate"]}
    gc.is_multilingual = True
    gc.language = "en"
    gc.task = "transcribe"
    gc.alignment_heads = alignment_heads or [[l, h] for l in range(hf_config.decoder_layers)
                                             for h in range(hf_config.decoder_attention_heads)][-4:]
    gc.suppress_tokens = suppress_tokens
    gc.begin_suppress_tokens = begin_suppress_tokens
    gc.num_beams = 1
    gc.max_length = hf_config.max_target_positions
    if max_new_tokens is not None:
        gc.max_length = None
        gc.max_new_tokens = max_new_tokens
    return m


def build_pipeline(model, tokenizer, batch_size: int = 16, chunk_length_s: int = 30, device: str = "cpu"):
    """The reference's call (REF/transcribe.py:21-31) with greedy pinned on the pipeline object."""
    from transformers import WhisperFeatureExtractor, pipeline
    fe = WhisperFeatureExtractor(feature_size=model.config.num_mel_bins)
    pipe = pipeline("automatic-speech-recognition", model=model, tokenizer=tokenizer, feature_extractor=fe,
                    chunk_length_s=chunk_length_s, batch_size=batch_size, return_timestamps="word", torch_dtype=torch.float32,
                    device=device)
    pipe.generation_config.num_beams = 1
    return pipe


def noise(seed: int, n: int = 480000, scale: float = 0.1) -> np.ndarray:
    """BASELINE.md §2 synthetic audio: chunk i = rng(i).standard_normal(n) * 0.1."""
    return (np.random.default_rng(seed).standard_normal(n) * scale).astype(np.float32)


def speechlike(seed: int, n: int = 480000) -> np.ndarray:
    """Noise bursts with 0.3 s gaps and a few harmonics: exercises the `max - 8` floor of the log-mel."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    x = 0.05 * rng.standard_normal(n)
    for f0 in (110.0, 220.0, 330.0, 1250.0):
        x += 0.08 * np.sin(2 * np.pi * f0 * t + rng.uniform(0, 6.28))
    env = ((t % 1.3) < 1.0).astype(np.float64)
    return (x * env).astype(np.float32)
