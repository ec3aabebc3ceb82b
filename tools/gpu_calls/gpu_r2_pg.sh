#!/bin/bash
# round 2, late call: pageable query batches staged through a page-locked ring by host threads
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nproc > gpurun_out/pg_nproc.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "filter_path_equals or page_locked or vector_lane_scan_equals or partition_major_scan_on_clustered" > gpurun_out/pg_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/pg_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-sharded --no-scan-large > gpurun_out/pg_bench.json 2> gpurun_out/pg_bench.err
for T in 2 4 16; do FDB_STAGE_THREADS=$T timeout 600 python bench.py --steps 5 --warmup 3 --no-sharded --no-scan-large > gpurun_out/pg_bench_t$T.json 2> gpurun_out/pg_bench_t$T.err; done
FDB_QUERY_NO_STAGING=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-sharded --no-scan-large > gpurun_out/pg_bench_nostage.json 2> gpurun_out/pg_bench_nostage.err
tail -2 gpurun_out/pg_tests.log; cat gpurun_out/pg_nproc.log
for f in gpurun_out/pg_bench.json gpurun_out/pg_bench_t2.json gpurun_out/pg_bench_t4.json gpurun_out/pg_bench_t16.json gpurun_out/pg_bench_nostage.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', d['e2e']['other_host_buffers']['pageable'], d['e2e']['ms_per_step'])"; done
