#!/bin/bash
# round 2, call E: full GPU suite with the vector-lane scan as the default + timings
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/e_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/e_tests.log
FDB_FILTER_STATS=1 timeout 300 python tools/prof_scan_large.py 8192 16 40000000 4096 vector > gpurun_out/e_scan_32ppl.log 2>&1
FDB_FILTER_STATS=1 timeout 300 python tools/prof_scan_large.py 2048 16 40000000 4096 vector > gpurun_out/e_scan_8ppl.log 2>&1
FDB_FILTER_STATS=1 timeout 300 python tools/prof_query.py 4 > gpurun_out/e_prof_query.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e_launches_query.csv python tools/prof_query.py 2 > gpurun_out/e_ncu1.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-sharded > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
FDB_VSCAN_OFF=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-sharded --no-scan-large > gpurun_out/e_bench_off.json 2> gpurun_out/e_bench_off.err
tail -5 gpurun_out/e_tests.log; cat gpurun_out/e_scan_*.log gpurun_out/e_prof_query.log
