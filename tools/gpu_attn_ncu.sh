set -x
python tools/attn_bench.py > gpurun_out/attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 10 -c 3 -o gpurun_out/r01_attn python tools/attn_bench.py > gpurun_out/attn_ncu.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench2.log 2>&1; tail -c 1500 gpurun_out/r01_bench2.log
