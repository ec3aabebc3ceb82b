#!/bin/bash
# round 2, call C: vscan v2 (conflict-free layout) + library-owned comm at world 1
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "vector_lane or forced_scan or library_owned" > gpurun_out/c_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c_tests.log
timeout 300 python tools/prof_scan_large.py 8192 16 40000000 4096 vector > gpurun_out/c_scan_32ppl.log 2>&1
timeout 300 python tools/prof_scan_large.py 4096 16 10000000 1024 vector > gpurun_out/c_scan_64ppl.log 2>&1
timeout 300 python tools/prof_scan_large.py 2048 16 40000000 4096 vector > gpurun_out/c_scan_8ppl.log 2>&1
FDB_VSCAN_DEFAULT=1 timeout 300 python tools/prof_query.py 4 > gpurun_out/c_prof_query.log 2>&1
FDB_VSCAN_DEFAULT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c_launches_query_vscan.csv python tools/prof_query.py 2 > gpurun_out/c_ncu1.log 2>&1
FDB_VSCAN_DEFAULT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vscan_kernel' -s 1 -c 1 -o gpurun_out/c_vscan_short -f python tools/prof_query.py 2 > gpurun_out/c_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vscan_kernel' -s 1 -c 1 -o gpurun_out/c_vscan_long -f python tools/prof_scan_large.py 4096 16 10000000 1024 vector > gpurun_out/c_ncu3.log 2>&1
tail -5 gpurun_out/c_tests.log; cat gpurun_out/c_scan_*.log gpurun_out/c_prof_query.log
