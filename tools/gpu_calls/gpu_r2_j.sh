#!/bin/bash
# round 2, call J: tc_assign with the two-group epilogue; host path trace; sharded coarse Lloyd round at 125k rows
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tc_ or lloyd or full_build or golden or batched or seeding or update or reassign or database_builder or live" > gpurun_out/j_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/j_tests.log
timeout 300 python tools/prof_tc.py > gpurun_out/j_prof_tc.log 2>&1
timeout 300 python tools/prof_e2e.py 7 > gpurun_out/j_prof_e2e.log 2>&1
timeout 300 python tools/prof_build2.py 125000 6 > gpurun_out/j_prof_build2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j_launches_build2.csv python tools/prof_build2.py 125000 4 > gpurun_out/j_ncu_build2.log 2>&1
tail -3 gpurun_out/j_tests.log; cat gpurun_out/j_prof_tc.log; cat gpurun_out/j_prof_e2e.log; cat gpurun_out/j_prof_build2.log
