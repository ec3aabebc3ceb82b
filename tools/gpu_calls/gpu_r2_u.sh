#!/bin/bash
# round 2, call U: tiled seeding, thread per (row, problem); stored / async / attribute tests
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "seeding or full_build or golden or database_builder or cpp or native or stored or example or library_owned" > gpurun_out/u_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/u_tests.log
timeout 300 python tools/prof_cfg2.py 125000 100 3 > gpurun_out/u_cfg2_125k.log 2>&1
timeout 300 python tools/prof_cfg2.py 1000000 100 3 > gpurun_out/u_cfg2_1m.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/u_launches_cfg2_125k.csv python tools/prof_cfg2.py 125000 3 1 > gpurun_out/u_ncu2.log 2>&1
tail -3 gpurun_out/u_tests.log; cat gpurun_out/u_cfg2_*.log
