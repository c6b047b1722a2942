#!/bin/bash
# A/B of prebuilt libcrisper.so variants under _ab/ on one box: cw_align time at cfg-5 sub-batch size
cp crisperwhisper_b200/libcrisper.so /tmp/orig.so
for v in "$@"; do
  echo "=== $v"
  if [ "$v" != "base" ]; then cp _ab/$v.so crisperwhisper_b200/libcrisper.so; else cp /tmp/orig.so crisperwhisper_b200/libcrisper.so; fi
  timeout 120 python tools/align_bench.py --n ${N:-128} --iters 5 2>&1 | tail -1
done
cp /tmp/orig.so crisperwhisper_b200/libcrisper.so
