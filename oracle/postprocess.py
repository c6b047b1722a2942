"""oracle/postprocess.py — restatement of REF/utils.py:1-29 (adjust_pauses_for_hf_pipeline_output).
TEST INFRASTRUCTURE ONLY (the product's own implementation lives in crisperwhisper_b200/utils.py).

Semantics kept: the chunk list is shallow-copied (REF/utils.py:6) so the dicts — and therefore the caller's
input — are mutated in place; pauses are processed left to right; a pause > 0 is split evenly up to
`split_threshold` (:14-20); floats are not rounded (:23,26)."""


def adjust_pauses(pipeline_output, split_threshold=0.12):
    chunks = pipeline_output["chunks"].copy()
    for i in range(len(chunks) - 1):
        cs, ce = chunks[i]["timestamp"]
        ns, ne = chunks[i + 1]["timestamp"]
        pause = ns - ce
        if pause > 0:
            dist = split_threshold / 2 if pause > split_threshold else pause / 2
            chunks[i]["timestamp"] = (cs, ce + dist)
            chunks[i + 1]["timestamp"] = (ns - dist, ne)
    pipeline_output["chunks"] = chunks
    return pipeline_output
