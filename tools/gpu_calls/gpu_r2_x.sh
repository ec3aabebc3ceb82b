#!/bin/bash
# round 2, call X: probe_exact with rows staged in shared memory; vscan on short lists only for large batches; e2e
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -x -m gpu ) > gpurun_out/x_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/x_tests.log
timeout 300 python tools/prof_query.py 4 > gpurun_out/x_prof_query.log 2>&1
FDB_PROBE_EXACT_NO_SMEM=1 timeout 300 python tools/prof_query.py 4 > gpurun_out/x_prof_query_old_probe.log 2>&1
SLICES=2500 PLANS="25,25,25,25;20,30,30,20" timeout 600 python tools/prof_e2e.py 9 > gpurun_out/x_prof_e2e.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x_launches_query.csv python tools/prof_query.py 2 > gpurun_out/x_ncu1.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-sharded --no-scan-large > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
tail -3 gpurun_out/x_tests.log; cat gpurun_out/x_prof_query.log gpurun_out/x_prof_query_old_probe.log; grep -v "^\[fdb" gpurun_out/x_prof_e2e.log
