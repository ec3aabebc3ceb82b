#!/bin/bash
# round 2, late call: host batches on 2 / 3 / 4 / 6 slice streams x slice sizes
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for S in 2 3 4 6; do
  echo "== FDB_QUERY_STREAMS=$S" >> gpurun_out/i2_prof_e2e.log
  FDB_QUERY_STREAMS=$S SLICES=834,1000,1250,1667,2500 PLANS="25,25,25,25" timeout 300 python tools/prof_e2e.py 9 2>&1 | grep -v "^\[fdb q" | grep -v "scan query" >> gpurun_out/i2_prof_e2e.log
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "page_locked or filter_path_equals" > gpurun_out/i2_tests.log 2>&1
FDB_QUERY_STREAMS=4 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "page_locked or filter_path_equals" >> gpurun_out/i2_tests.log 2>&1
cat gpurun_out/i2_prof_e2e.log; tail -2 gpurun_out/i2_tests.log
