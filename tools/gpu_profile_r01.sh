# Round-1 evidence run (one B200): tests, bench, launch list, probes, parity report -> gpurun_out/ (copied to profiles/)
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r01_tests.log
timeout 400 python bench.py > gpurun_out/r01_bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --kernel-only --detail > gpurun_out/r01_detail.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 700 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 1 --warmup 2 --kernel-only > gpurun_out/r01_ncu_launches.log 2>&1
timeout 60 ./tools/umma_probe > gpurun_out/r01_umma_probe.log 2>&1
timeout 200 python tools/loss_parity_report.py 2>&1 | grep "^case" > gpurun_out/r01_loss_parity.log
timeout 120 python tools/attn_bench.py > gpurun_out/r01_attn_bench.log 2>&1
timeout 120 python tools/gemm_bench.py > gpurun_out/r01_gemm_bench.log 2>&1
timeout 120 python tools/augment_bench.py > gpurun_out/r01_augment_bench.log 2>&1
timeout 100 python tools/ln_bench.py > gpurun_out/r01_ln_bench.log 2>&1
tail -1 gpurun_out/r01_bench.log | cut -c1-300; cat gpurun_out/r01_tests.log
