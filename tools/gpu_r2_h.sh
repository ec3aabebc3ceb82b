#!/bin/bash
# round 2, call H: tc_assign with problem-major pieces + resident centroids; lagged Lloyd loop; build timing
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "tc_ or lloyd or full_build or golden or batched or seeding or update or reassign or database_builder or live" > gpurun_out/h_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/h_tests.log
timeout 300 python tools/prof_tc.py > gpurun_out/h_prof_tc.log 2>&1
FDB_TC_ROW_MAJOR=1 timeout 300 python tools/prof_tc.py > gpurun_out/h_prof_tc_rowmajor.log 2>&1
FDB_TC_ROW_MAJOR=1 FDB_TC_NO_BRES=1 timeout 300 python tools/prof_tc.py > gpurun_out/h_prof_tc_r01.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_assign_kernel' -s 4 -c 1 -o gpurun_out/h_tc_assign_pq -f python tools/prof_tc.py > gpurun_out/h_ncu_tc.log 2>&1
FDB_BENCH_BUILD_PROFILE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-sharded --no-scan-large > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
tail -3 gpurun_out/h_tests.log; cat gpurun_out/h_prof_tc*.log; tail -3 gpurun_out/h_ncu_tc.log
