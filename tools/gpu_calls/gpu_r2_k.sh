#!/bin/bash
# round 2, call K: tc_assign timing experiments (no loads / no epilogue), CTA-per-row recheck of overflow rows
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tc_ or lloyd or full_build or golden or batched or seeding or update or reassign or database_builder or live" > gpurun_out/k_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/k_tests.log
for d in 0 1 2 3; do
  echo "FDB_TC_DEBUG=$d" >> gpurun_out/k_prof_tc.log
  FDB_TC_DEBUG=$d timeout 300 python tools/prof_tc.py >> gpurun_out/k_prof_tc.log 2>&1
done
timeout 300 python tools/prof_build2.py 125000 6 > gpurun_out/k_prof_build2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/k_launches_build2.csv python tools/prof_build2.py 125000 4 > gpurun_out/k_ncu_build2.log 2>&1
tail -3 gpurun_out/k_tests.log; cat gpurun_out/k_prof_tc.log; cat gpurun_out/k_prof_build2.log
