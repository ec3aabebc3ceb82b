#!/bin/bash
# round 2, call P: vscan two-pass cold start: parity, README-shape timing (vector forced vs default), ncu
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "vector_lane or forced_scan or cold_start" > gpurun_out/p_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/p_tests.log
timeout 300 python tools/prof_query.py 4 > gpurun_out/p_prof_query_default.log 2>&1
FDB_FILTER_SCAN=vector FDB_FILTER_STATS=1 timeout 300 python tools/prof_query.py 4 > gpurun_out/p_prof_query_vector.log 2>&1
FDB_FILTER_SCAN=vector FDB_VSCAN_2PASS_MAX=0 timeout 300 python tools/prof_query.py 4 > gpurun_out/p_prof_query_vector_rounds.log 2>&1
FDB_FILTER_SCAN=vector timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/p_launches_query_vscan.csv python tools/prof_query.py 2 > gpurun_out/p_ncu1.log 2>&1
FDB_FILTER_SCAN=vector timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vscan_kernel' -s 1 -c 1 -o gpurun_out/p_vscan_short -f python tools/prof_query.py 2 > gpurun_out/p_ncu2.log 2>&1
timeout 300 python tools/prof_scan_large.py 2048 16 40000000 4096 vector > gpurun_out/p_scan_8ppl.log 2>&1
timeout 300 python tools/prof_scan_large.py 4096 16 10000000 1024 vector > gpurun_out/p_scan_64ppl.log 2>&1
tail -3 gpurun_out/p_tests.log; cat gpurun_out/p_prof_query_*.log gpurun_out/p_scan_*.log
