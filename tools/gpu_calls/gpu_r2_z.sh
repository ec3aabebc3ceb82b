#!/bin/bash
# round 2, call Z: stored-semantic probes from sparse rows; chunked probe_exact; accumulate with 16 in flight; full suite + full bench
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/z_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/z_tests.log
( time timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err ) 2> gpurun_out/z_bench.time
timeout 300 python tools/prof_query.py 4 > gpurun_out/z_prof_query.log 2>&1
timeout 300 python tools/prof_tc.py > gpurun_out/z_prof_tc.log 2>&1
timeout 300 python tools/prof_cfg2.py 125000 100 2 > gpurun_out/z_cfg2_125k.log 2>&1
tail -3 gpurun_out/z_tests.log; cat gpurun_out/z_bench.time; cat gpurun_out/z_prof_query.log gpurun_out/z_prof_tc.log gpurun_out/z_cfg2_125k.log
