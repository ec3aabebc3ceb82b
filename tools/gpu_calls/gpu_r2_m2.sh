#!/bin/bash
# round 2, late call: device-memory cache (exact-size reuse of freed blocks): full suite, bench, build timing stability
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/m2_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/m2_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/m2_smoke.log 2>&1
( time timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/m2_bench.json 2> gpurun_out/m2_bench.err ) 2> gpurun_out/m2_bench.time
timeout 600 python bench.py --steps 5 --warmup 3 --no-sharded --no-scan-large > gpurun_out/m2_bench2.json 2> gpurun_out/m2_bench2.err
tail -3 gpurun_out/m2_tests.log; tail -1 gpurun_out/m2_smoke.log; cat gpurun_out/m2_bench.time
