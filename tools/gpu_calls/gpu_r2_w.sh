#!/bin/bash
# round 2, call W: host batch path: early results + pinned staging, tapered slice plans, fscan vs vscan slices
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "page_locked or host or filter_path_equals" > gpurun_out/w_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/w_tests.log
SLICES=2500 timeout 600 python tools/prof_e2e.py 9 > gpurun_out/w_prof_e2e.log 2>&1
tail -3 gpurun_out/w_tests.log; cat gpurun_out/w_prof_e2e.log
