#!/bin/bash
# round 2, call O: the vector-lane scan forced on the README shape (short lists): phase times, launch list, ncu full with source
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export FDB_FILTER_SCAN=vector
FDB_FILTER_STATS=1 timeout 300 python tools/prof_query.py 4 > gpurun_out/o_prof_query.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/o_launches_query_vscan.csv python tools/prof_query.py 2 > gpurun_out/o_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vscan_kernel|vq_quant_kernel|pmerge_kernel|pg_items_kernel|fselect_kernel' -s 5 -c 5 -o gpurun_out/o_vscan_short -f python tools/prof_query.py 2 > gpurun_out/o_ncu2.log 2>&1
cat gpurun_out/o_prof_query.log
