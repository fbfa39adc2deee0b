set -x
python -m pytest tests/test_kernels_gpu.py -x -q -k attention 2>&1 | tail -3
python tools/attn_bench.py > gpurun_out/attn_plain.log 2>&1 && cat gpurun_out/attn_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 13 -c 3 -o gpurun_out/r01_attn2 python tools/attn_bench.py > gpurun_out/attn_ncu.log 2>&1
