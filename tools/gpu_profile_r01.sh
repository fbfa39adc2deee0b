set -x
python bench.py --steps 3 --warmup 3 --kernel-only --detail > gpurun_out/r01_detail.log 2>&1
python bench.py --steps 1 --warmup 3 --kernel-only > gpurun_out/r01_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 2700 -c 1000 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 1 --warmup 3 --kernel-only > gpurun_out/r01_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 300 -c 6 -o gpurun_out/r01_gemm python bench.py --steps 1 --warmup 3 --kernel-only > gpurun_out/r01_ncu_gemm.log 2>&1
tail -3 gpurun_out/r01_ncu_launches.log gpurun_out/r01_ncu_gemm.log
head -70 gpurun_out/r01_detail.log
