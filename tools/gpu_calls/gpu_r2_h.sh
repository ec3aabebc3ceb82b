#!/bin/bash
# round 2, late call: pmerge with lane-parallel descriptor fetch; stored-query events; full suite + bench
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/h2_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/h2_tests.log
timeout 300 python tools/prof_query.py 4 > gpurun_out/h2_prof_query.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/h2_launches_query.csv python tools/prof_query.py 2 > gpurun_out/h2_ncu1.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-sharded > gpurun_out/h2_bench.json 2> gpurun_out/h2_bench.err
tail -3 gpurun_out/h2_tests.log; cat gpurun_out/h2_prof_query.log
