#!/bin/bash
# A/B of prebuilt libcrisper.so variants under _ab/ on one box: ms/step (T=64) and per-phase timings
cp crisperwhisper_b200/libcrisper.so /tmp/orig.so
for v in "$@"; do
  echo "=== $v"
  if [ "$v" != "base" ]; then cp _ab/$v.so crisperwhisper_b200/libcrisper.so; else cp /tmp/orig.so crisperwhisper_b200/libcrisper.so; fi
  T=${T:-64} ONLY_MEGA=1 timeout 200 python tools/decode_bench.py 2>&1 | grep "mega (default)"
  timeout 200 python tools/mega_debug.py 2>&1 | grep -A26 "CW_MEGA_DEBUG" | tail -26
done
cp /tmp/orig.so crisperwhisper_b200/libcrisper.so
