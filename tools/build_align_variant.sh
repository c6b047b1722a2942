#!/bin/bash
# build a libcrisper.so variant with extra nvcc flags for align.cu:  tools/build_align_variant.sh NAME -DCW_ALIGN_SKIP=7
set -e
name=$1; shift
cd "$(dirname "$0")/.."
O=crisperwhisper_b200/_obj
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c crisperwhisper_b200/csrc/align.cu -o /tmp/align_$name.o
mkdir -p _ab
nvcc -shared -o _ab/$name.so $O/api.o $O/logmel.o /tmp/align_$name.o $O/gemm.o $O/encoder.o $O/resample.o $O/postproc.o $O/decoder.o -gencode arch=compute_100a,code=sm_100a -cudart static
python - <<PY
import ctypes; ctypes.CDLL("_ab/$name.so"); print("built _ab/$name.so (loads)")
PY
