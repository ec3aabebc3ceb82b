#!/bin/bash
# round 2, final single-GPU record: GPU suite, bench (both arms), launch list of the bench command, ncu full of the scan kernels
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/f_gpu.log 2>&1
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/f_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/f_tests.log
( time timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err ) 2> gpurun_out/f_bench.time
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
timeout 300 python tools/prof_query.py 4 > gpurun_out/f_prof_query.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-sharded --no-scan-large > gpurun_out/f_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vscan_kernel|probe_exact_smem_kernel' -c 4 -o gpurun_out/f_vscan_short -f python tools/prof_query.py 2 > gpurun_out/f_ncu_query.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vscan_kernel' -s 1 -c 1 -o gpurun_out/f_vscan_8ppl -f python tools/prof_scan_large.py 2048 16 40000000 4096 vector > gpurun_out/f_ncu_scan8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'seed_round_tile_kernel' -s 20 -c 1 -o gpurun_out/f_seed_tile -f python tools/prof_cfg2.py 125000 3 1 > gpurun_out/f_ncu_seed.log 2>&1
tail -3 gpurun_out/f_tests.log; cat gpurun_out/f_bench.time; cat gpurun_out/f_prof_query.log
