#!/bin/bash
# A/B of prebuilt libcrisper.so variants under _ab/ on one box: ms/step (T=128 and T=445) and per-phase timings
cp crisperwhisper_b200/libcrisper.so /tmp/orig.so
for v in "$@"; do
  echo "=== $v"
  cp _ab/$v.so crisperwhisper_b200/libcrisper.so
  T=128 ONLY_MEGA=1 timeout 200 python tools/decode_bench.py 2>&1 | grep "mega (default)"
  T=445 ONLY_MEGA=1 timeout 200 python tools/decode_bench.py 2>&1 | grep "mega (default)"
  timeout 200 python tools/mega_debug.py 2>&1 | grep -A12 "CW_MEGA_DEBUG" | tail -12
done
cp /tmp/orig.so crisperwhisper_b200/libcrisper.so
