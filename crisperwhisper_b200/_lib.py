"""ctypes binding of libcrisper.so (include/crisper.h).  There is NO fallback: if the CUDA library is missing or a
call fails, a RuntimeError is raised — the product path never routes through a CPU implementation."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcrisper.so")

CW_OK = 0
CW_DEC_SUPPRESS_EOS = 1
CW_DEC_NO_TIMESTAMP_RULES = 2
CW_DEC_NO_GRAPH = 4
CW_DEC_PROFILE = 8
CW_DEC_NO_PDL = 16
CW_DEC_NO_MEGA = 32
CW_DEC_NO_SUPPRESS = 64

# weight slot enums (must mirror include/crisper.h)
W_GLOBAL = ["CONV1_W", "CONV1_B", "CONV2_W", "CONV2_B", "ENC_POS", "ENC_LNF_G", "ENC_LNF_B", "XKV_W", "XKV_B",
            "TOK_EMB", "DEC_POS", "DEC_LNF_G", "DEC_LNF_B"]
W_ENC_LAYER = ["LN1_G", "LN1_B", "WQKV", "BQKV", "WO", "BO", "LN2_G", "LN2_B", "W1", "B1", "W2", "B2"]
W_DEC_LAYER = ["LN1_G", "LN1_B", "WQKV", "BQKV", "WO", "BO", "LN2_G", "LN2_B", "WQC", "BQC", "WOC", "BOC",
               "LN3_G", "LN3_B", "W1", "B1", "W2", "B2"]


class ModelDesc(C.Structure):
    _fields_ = [
        ("d_model", C.c_int32), ("n_heads", C.c_int32), ("enc_layers", C.c_int32), ("dec_layers", C.c_int32),
        ("ffn_dim", C.c_int32), ("vocab", C.c_int32), ("vocab_padded", C.c_int32), ("n_mels", C.c_int32),
        ("n_audio_ctx", C.c_int32), ("n_text_ctx", C.c_int32), ("eos_id", C.c_int32), ("no_timestamps_id", C.c_int32),
        ("max_initial_timestamp_index", C.c_int32), ("median_filter_width", C.c_int32), ("n_align_heads", C.c_int32),
        ("align_heads_host", C.POINTER(C.c_int32)), ("n_suppress", C.c_int32), ("suppress_host", C.POINTER(C.c_int32)),
        ("n_begin_suppress", C.c_int32), ("begin_suppress_host", C.POINTER(C.c_int32)),
    ]


EXPORTS = {
    # name: (restype, argtypes)
    "cw_abi_version": (C.c_int, []),
    "cw_init": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "cw_destroy": (None, [C.c_void_p]),
    "cw_last_error": (C.c_char_p, []),
    "cw_load_weights": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.POINTER(ModelDesc)]),
    "cw_logmel_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "cw_logmel": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "cw_encode_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "cw_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "cw_decode_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "cw_decode_greedy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    "cw_decode_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "cw_align_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "cw_align": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                           C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "cw_gemm_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                               C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cw_gemm_bf16_check": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cw_attention_enc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cw_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "cw_decode_cross_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cw_decode_pack_bytes": (C.c_size_t, [C.c_void_p]),
    "cw_decode_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "cw_resample_out_len": (C.c_longlong, [C.c_longlong, C.c_int, C.c_int]),
    "cw_resample_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "cw_resample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p,
                              C.c_size_t, C.c_void_p]),
    "cw_words_from_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_void_p, C.c_int64,
                                       C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "cw_launch_count": (C.c_longlong, [C.c_void_p]),
    "cw_event_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "cw_event_record": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cw_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "cw_event_destroy": (C.c_int, [C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load libcrisper.so (built by crisperwhisper_b200.build / __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m crisperwhisper_b200.build` "
            "(there is no CPU fallback for the CrisperWhisper hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export what include/crisper.h declares
        fn.restype = res
        fn.argtypes = args
    if lib.cw_abi_version() != 2:
        raise RuntimeError("libcrisper.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "libcrisper") -> None:
    if rc != CW_OK:
        msg = load().cw_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (cw_status {rc}): {msg}")
