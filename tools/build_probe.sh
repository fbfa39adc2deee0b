#!/bin/bash
# builds tools/gemm_probe against the in-tree libavsiam_b200.so
set -e
cd "$(dirname "$0")/.."
python -m avsiam_b200.build
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -cudart shared -o tools/gemm_probe tools/gemm_probe.cu \
  -L avsiam_b200 -lavsiam_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../avsiam_b200'
