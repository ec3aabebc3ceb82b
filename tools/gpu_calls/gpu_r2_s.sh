#!/bin/bash
# round 2, call S: tiled k-means++ round for many short problems (configs[2] PQ seeding): parity + phase times
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "seeding or full_build or golden or batched or library_owned or database_builder" > gpurun_out/s_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/s_tests.log
timeout 300 python tools/prof_cfg2.py 125000 100 2 > gpurun_out/s_cfg2_125k.log 2>&1
timeout 300 python tools/prof_cfg2.py 1000000 100 2 > gpurun_out/s_cfg2_1m.log 2>&1
FDB_SEED_NO_TILE=1 timeout 300 python tools/prof_cfg2.py 1000000 100 2 > gpurun_out/s_cfg2_1m_notile.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s_launches_cfg2_125k.csv python tools/prof_cfg2.py 125000 3 1 > gpurun_out/s_ncu2.log 2>&1
tail -3 gpurun_out/s_tests.log; cat gpurun_out/s_cfg2_*.log
