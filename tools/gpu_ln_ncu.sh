set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:layernorm_bwd -s 100 -c 2 -o gpurun_out/r01_ln_bwd python bench.py --steps 1 --warmup 1 --kernel-only > gpurun_out/ln_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"layernorm_fwd|colsum" -s 140 -c 4 -o gpurun_out/r01_ln_fwd_colsum python bench.py --steps 1 --warmup 1 --kernel-only >> gpurun_out/ln_ncu.log 2>&1
tail -2 gpurun_out/ln_ncu.log
