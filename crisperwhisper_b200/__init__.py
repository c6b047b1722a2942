"""crisperwhisper_b200 — B200-native (sm_100a) inference-and-alignment path with the call surface of
nyrahealth/CrisperWhisper: `pipeline(..., return_timestamps="word")`, `adjust_pauses_for_hf_pipeline_output`,
and a `transcribe.py`-style CLI.  All compute runs in libcrisper.so (include/crisper.h); there is no CPU fallback."""
from .utils import adjust_pauses_for_hf_pipeline_output, timestamps_to_vtt  # noqa: F401

__all__ = ["pipeline", "adjust_pauses_for_hf_pipeline_output", "timestamps_to_vtt", "Engine"]


def __getattr__(name):  # lazy: importing the package must not need torch/CUDA (build() runs on a CPU box)
    if name == "pipeline":
        from .asr_pipeline import pipeline
        return pipeline
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(name)
