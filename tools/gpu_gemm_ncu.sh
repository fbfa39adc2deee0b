set -x
timeout 120 python tools/loss_parity_report.py 2>&1 | tail -4
timeout 120 python tools/gemm_bench.py 2>&1 | tail -8
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o gpurun_out/r01_gemm_gelu python tools/gemm_bench.py gelu > gpurun_out/gemm_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o gpurun_out/r01_gemm_dgelu python tools/gemm_bench.py dgelu >> gpurun_out/gemm_ncu.log 2>&1
