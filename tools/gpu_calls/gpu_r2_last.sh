#!/bin/bash
# round 2, last record on the final tree: GPU suite, smoke, bench (both arms)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/l_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/l_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/l_smoke.log 2>&1
( time timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err ) 2> gpurun_out/l_bench.time
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/l_bench_ref.json 2> gpurun_out/l_bench_ref.err
tail -3 gpurun_out/l_tests.log; tail -1 gpurun_out/l_smoke.log; cat gpurun_out/l_bench.time
