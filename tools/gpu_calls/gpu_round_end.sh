#!/bin/bash
# what the round-end record needs, in one box: GPU tests, bench (plain), launch list of the same bench
# command under ncu, one ncu --set full capture of the query kernels and of the long-list scan kernels
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_bench_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fscan_kernel|fselect_kernel|probe_exact_kernel|probe_select_kernel|probe_finalize_kernel' -c 6 -o gpurun_out/prof_query_final -f python tools/prof_query.py 1 > gpurun_out/ncu_query_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pscan16_kernel|fscan_kernel' -c 2 -o gpurun_out/prof_scan_large_final -f python tools/prof_scan_large.py 4096 16 10000000 1024 partition16,query > gpurun_out/ncu_scan_large_final.log 2>&1
ls -la gpurun_out/*final*
