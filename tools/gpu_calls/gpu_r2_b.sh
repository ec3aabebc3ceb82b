#!/bin/bash
# round 2, call B: where the vector-lane scan spends its time on short lists (README shape) and on long lists
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export FDB_VSCAN_DEFAULT=1
timeout 300 python tools/prof_query.py 3 > gpurun_out/b_prof_query.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launches_query_vscan.csv python tools/prof_query.py 2 > gpurun_out/b_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vscan_kernel' -s 1 -c 1 -o gpurun_out/b_vscan_short -f python tools/prof_query.py 2 > gpurun_out/b_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vscan_kernel' -s 1 -c 1 -o gpurun_out/b_vscan_long -f python tools/prof_scan_large.py 4096 16 10000000 1024 vector > gpurun_out/b_ncu3.log 2>&1
cat gpurun_out/b_prof_query.log
