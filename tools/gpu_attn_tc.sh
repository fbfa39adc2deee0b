set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k attention 2>&1 | tail -4
timeout 120 python tools/attn_bench.py 2>&1 | tail -6
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 3 -c 1 -o gpurun_out/r01_attn_tc_bwd python tools/attn_bench.py > gpurun_out/attn_ncu.log 2>&1
