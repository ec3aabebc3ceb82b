#!/bin/bash
# round 2, call D: vscan v3 (pre-quantised tables, retries) + library-owned comm
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -k "vector_lane or forced_scan or library_owned" > gpurun_out/d_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/d_tests.log
timeout 300 python tools/prof_scan_large.py 8192 16 40000000 4096 vector > gpurun_out/d_scan_32ppl.log 2>&1
timeout 300 python tools/prof_scan_large.py 2048 16 40000000 4096 vector > gpurun_out/d_scan_8ppl.log 2>&1
FDB_VSCAN_DEFAULT=1 timeout 300 python tools/prof_query.py 4 > gpurun_out/d_prof_query.log 2>&1
FDB_VSCAN_DEFAULT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches_query_vscan.csv python tools/prof_query.py 2 > gpurun_out/d_ncu1.log 2>&1
FDB_VSCAN_DEFAULT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vscan_kernel' -s 1 -c 1 -o gpurun_out/d_vscan_short -f python tools/prof_query.py 2 > gpurun_out/d_ncu2.log 2>&1
FDB_VSCAN_DEFAULT=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-scan-large --no-sharded > gpurun_out/d_bench_vscan.json 2> gpurun_out/d_bench_vscan.err
tail -5 gpurun_out/d_tests.log; cat gpurun_out/d_scan_*.log gpurun_out/d_prof_query.log
