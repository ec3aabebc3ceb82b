#!/bin/bash
# round 2, call N: the whole record again on one box (the container holding the earlier gpurun_out/ was lost):
# GPU suite, bench both arms, launch list of the bench command, ncu --set full of the scan / selection / tc kernels
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/n_gpu.log 2>&1
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/n_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/n_tests.log
( time timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err ) 2> gpurun_out/n_bench.time
( time timeout 600 python bench.py --steps 5 --warmup 3 --no-sharded > gpurun_out/n_bench_nosh.json 2> gpurun_out/n_bench_nosh.err ) 2> gpurun_out/n_bench_nosh.time
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/n_bench_ref.json 2> gpurun_out/n_bench_ref.err
timeout 300 python tools/prof_query.py 4 > gpurun_out/n_prof_query.log 2>&1
timeout 300 python tools/prof_tc.py > gpurun_out/n_prof_tc.log 2>&1
timeout 300 python tools/prof_scan_large.py 8192 16 40000000 4096 query,partition16,vector > gpurun_out/n_scan_32ppl.log 2>&1
timeout 300 python tools/prof_scan_large.py 4096 16 10000000 1024 query,partition16,vector > gpurun_out/n_scan_64ppl.log 2>&1
timeout 300 python tools/prof_scan_large.py 2048 16 40000000 4096 query,vector > gpurun_out/n_scan_8ppl.log 2>&1
timeout 300 python tools/prof_e2e.py 7 > gpurun_out/n_prof_e2e.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-sharded --no-scan-large > gpurun_out/n_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vscan_kernel|fscan_kernel|fselect_kernel|pmerge_kernel|vq_quant_kernel|probe_exact_kernel|probe_select_kernel' -s 14 -c 14 -o gpurun_out/n_query_kernels -f python tools/prof_query.py 3 > gpurun_out/n_ncu_query.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vscan_kernel|pscan16_kernel|fscan_kernel' -c 3 -o gpurun_out/n_scan_large -f python tools/prof_scan_large.py 4096 16 10000000 1024 vector,partition16,query > gpurun_out/n_ncu_scan_large.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_assign_kernel' -s 4 -c 2 -o gpurun_out/n_tc_assign_pq -f python tools/prof_tc.py > gpurun_out/n_ncu_tc.log 2>&1
tail -3 gpurun_out/n_tests.log; cat gpurun_out/n_bench.time gpurun_out/n_bench_nosh.time; cat gpurun_out/n_prof_query.log gpurun_out/n_prof_tc.log gpurun_out/n_scan_*.log
