#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
rm -f gpurun_out/l_prof_tc.log
for d in 16 56; do
  echo "FDB_TC_DEBUG=$d" >> gpurun_out/l_prof_tc.log
  FDB_TC_DEBUG=$d timeout 300 python tools/prof_tc.py 2>&1 | tail -4 >> gpurun_out/l_prof_tc.log
done
cat gpurun_out/l_prof_tc.log
