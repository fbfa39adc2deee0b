#!/bin/bash
# Developer aid: GEMM-only builds of the library with A/B macros (gemm.cu + runtime.cu), for tools/gemm_ab.py
set -e
cd "$(dirname "$0")/../.."
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --use_fast_math --expt-relaxed-constexpr -shared -cudart shared"
for v in "cur:" "legacy:-DAVS_GEMM_LEGACY_ROLES" $EXTRA_VARIANTS; do
  name=${v%%:*}; defs=${v#*:}
  nvcc $F $defs -o tools/ab/libgemm_$name.so avsiam_b200/csrc/gemm.cu avsiam_b200/csrc/runtime.cu -Xptxas -v 2> tools/ab/ptxas_$name.log &
done
wait
grep -h "spill" tools/ab/ptxas_*.log | sort | uniq -c
