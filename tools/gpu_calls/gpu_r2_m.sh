#!/bin/bash
# round 2, call M: full GPU suite + bench after the tc_assign epilogue work
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/m_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/m_tests.log
timeout 300 python tools/prof_tc.py > gpurun_out/m_prof_tc.log 2>&1
timeout 300 python tools/prof_build2.py 125000 6 > gpurun_out/m_prof_build2.log 2>&1
FDB_BENCH_BUILD_PROFILE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-sharded --no-scan-large > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
tail -3 gpurun_out/m_tests.log; cat gpurun_out/m_prof_tc.log gpurun_out/m_prof_build2.log; tail -3 gpurun_out/m_bench.err
