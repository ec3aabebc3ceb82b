#!/bin/bash
# round 2, call Z3: NBestByKey push as one prefix-maximum sweep (WarpNBest): full suite, stored-semantic phases at P = 16384
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/z3_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/z3_tests.log
timeout 300 python tools/prof_phases_p16k.py 128 > gpurun_out/z3_phases_128.log 2>&1
timeout 300 python tools/prof_phases_p16k.py 32 > gpurun_out/z3_phases_32.log 2>&1
timeout 300 python tools/prof_phases_p16k.py 64 > gpurun_out/z3_phases_64.log 2>&1
tail -3 gpurun_out/z3_tests.log; cat gpurun_out/z3_phases_*.log | grep "^mode"
