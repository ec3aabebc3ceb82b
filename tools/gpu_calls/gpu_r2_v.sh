#!/bin/bash
# round 2, call V: full GPU suite + bench (both arms) on the current tree, launch list of the bench command, ncu full of the query kernels
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/v_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/v_tests.log
( time timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err ) 2> gpurun_out/v_bench.time
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/v_bench_ref.json 2> gpurun_out/v_bench_ref.err
FDB_QUERY_TRACE=1 timeout 300 python tools/prof_e2e.py 7 > gpurun_out/v_prof_e2e.log 2>&1
timeout 300 python tools/prof_query.py 4 > gpurun_out/v_prof_query.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/v_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-sharded --no-scan-large > gpurun_out/v_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vscan_kernel|fselect_kernel|pmerge_kernel|vq_quant_kernel|probe_exact_kernel|probe_select_kernel|probe_finalize_kernel|tc_assign_kernel|split_rows_kernel' -s 9 -c 9 -o gpurun_out/v_query_kernels -f python tools/prof_query.py 2 > gpurun_out/v_ncu_query.log 2>&1
tail -3 gpurun_out/v_tests.log; cat gpurun_out/v_bench.time; cat gpurun_out/v_prof_query.log; tail -12 gpurun_out/v_prof_e2e.log
