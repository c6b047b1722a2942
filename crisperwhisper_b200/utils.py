"""Post-processing helpers with the reference's names and behaviour (REF/utils.py:1-29, REF/app.py:74-82).

`adjust_pauses_for_hf_pipeline_output(pipeline_output, split_threshold=0.12)` redistributes the silence between
adjacent word chunks: a pause `p = next_start - cur_end > 0` is shared evenly, `min(p, split_threshold) / 2` being added
to the end of the current word and subtracted from the start of the next.  Kept from the reference on purpose
(SURVEY §7.1 Q9): only the list is copied, so the chunk dicts of the argument are updated in place; words are
visited left to right, so a word's start has already been moved when its end is considered; results are not rounded.
"""
from __future__ import annotations


def adjust_pauses_for_hf_pipeline_output(pipeline_output, split_threshold=0.12):
    words = list(pipeline_output["chunks"])
    for cur, nxt in zip(words, words[1:]):
        cur_start, cur_end = cur["timestamp"]
        nxt_start, nxt_end = nxt["timestamp"]
        pause = nxt_start - cur_end
        if pause > 0:
            share = (split_threshold if pause > split_threshold else pause) / 2
            cur["timestamp"] = (cur_start, cur_end + share)
            nxt["timestamp"] = (nxt_start - share, nxt_end)
    pipeline_output["chunks"] = words
    return pipeline_output


def _vtt_time(t: float) -> str:
    # h:mm:ss.mmm with unpadded hours; the seconds field is rounded to the millisecond on its own (so 59.9996 prints as
    # "60.000" without carrying into the minutes) — as REF/app.py:79-80 formats it
    return f"{int(t // 3600)}:{int(t // 60 % 60):02d}:{t % 60:06.3f}"


def timestamps_to_vtt(timestamps) -> str:
    """Word chunks [{"text", "timestamp": (start, end)}, ...] -> WebVTT text, one cue per word (REF/app.py:74-82)."""
    cues = ["WEBVTT\n\n"]
    for word in timestamps:
        start, end = word["timestamp"]
        cues.append(f"{_vtt_time(start)} --> {_vtt_time(end)}\n{word['text']}\n\n")
    return "".join(cues)
