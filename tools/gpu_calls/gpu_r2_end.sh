#!/bin/bash
# round 2, end: the GPU suite and smoke on the final tree
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/end_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/end_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/end_smoke.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-sharded --no-scan-large > gpurun_out/end_bench.json 2> gpurun_out/end_bench.err
tail -5 gpurun_out/end_tests.log; tail -1 gpurun_out/end_smoke.log
