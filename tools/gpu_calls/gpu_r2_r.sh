#!/bin/bash
# round 2, call R: fselect with staged codes + prefetched rows, vscan default on the README shape; configs[2] shard phases
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r_tests.log
timeout 300 python tools/prof_query.py 4 > gpurun_out/r_prof_query.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r_launches_query.csv python tools/prof_query.py 2 > gpurun_out/r_ncu1.log 2>&1
timeout 300 python tools/prof_cfg2.py 125000 100 2 > gpurun_out/r_cfg2_125k.log 2>&1
timeout 300 python tools/prof_cfg2.py 1000000 100 2 > gpurun_out/r_cfg2_1m.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r_launches_cfg2_125k.csv python tools/prof_cfg2.py 125000 3 1 > gpurun_out/r_ncu2.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-sharded --no-scan-large > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
tail -3 gpurun_out/r_tests.log; cat gpurun_out/r_prof_query.log gpurun_out/r_cfg2_*.log
