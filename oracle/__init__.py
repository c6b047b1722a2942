"""CPU oracle for the CrisperWhisper hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``crisperwhisper_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it, and only
as the checker / the thing the GPU path is compared with — never as the thing that is shipped.

The reference repository (nyrahealth/CrisperWhisper @ 4a24d3d) contains no arithmetic of its own: all of it
lives in the un-vendored, un-pinned dependency ``transformers`` (REF/requirements.txt:3; version 5.5.0 is the
one installed in this image and therefore the binding behaviour).  The reference ships no tests, golden
vectors or fixtures for this path ("parity unpinned" at the reference level), so every restatement here is
pinned instead against the outputs of the installed ``transformers`` functions themselves
(tests/test_oracle_pins.py, tests/golden/make_golden.py) — the "outputs of the reference itself run here"
route.

Modules
  logmel.py      numpy restatement of WhisperFeatureExtractor (feature_extraction_whisper.py:135-164,296-337)
  align.py       numpy/C restatement of _median_filter/_dynamic_time_warping/_extract_token_timestamps
                 (generation_whisper.py:43-115,241-381); dtw.c is the C inner loop
  logits.py      numpy restatement of the three Whisper logits processors (logits_process.py:1847-2043)
  whisper_ref.py plain PyTorch fp32 restatement of encoder/decoder forward + greedy loop
                 (modeling_whisper.py:215-796,1081; generation/utils.py:2743-2800)
  postprocess.py restatement of REF/utils.py:1-29 (pause redistribution)
  hf_harness.py  synthetic Whisper model / tokenizer / pipeline factory on top of the installed transformers
                  — the reference arm of bench.py and the generator of tests/golden/
"""
