#!/bin/bash
# round 2, call Z2: accumulate tail batched; where the stored semantic spends its time at P = 16384; e2e slice plans; launch list
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "update or lloyd or full_build or golden or database_builder or library_owned" > gpurun_out/z2_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/z2_tests.log
FDB_FILTER_STATS=1 timeout 300 python tools/prof_phases_p16k.py 128 > gpurun_out/z2_phases_128.log 2>&1
FDB_FILTER_STATS=1 timeout 300 python tools/prof_phases_p16k.py 32 > gpurun_out/z2_phases_32.log 2>&1
SLICES=2500 PLANS="25,25,25,25;28,28,28,16;30,30,25,15;30,30,30,10;22,22,22,22,12" timeout 600 python tools/prof_e2e.py 9 > gpurun_out/z2_prof_e2e.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/z2_launches_build.csv python tools/prof_tc.py > gpurun_out/z2_ncu1.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-sharded --no-scan-large > gpurun_out/z2_bench.json 2> gpurun_out/z2_bench.err
tail -3 gpurun_out/z2_tests.log; cat gpurun_out/z2_phases_*.log | grep -v "^$" | cut -c1-400; grep -v "^\[fdb q" gpurun_out/z2_prof_e2e.log
