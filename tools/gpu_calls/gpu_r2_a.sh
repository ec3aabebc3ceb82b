#!/bin/bash
# round 2, call A: parity of the vector-lane scan + first timings
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "vector_lane or forced_scan" > gpurun_out/a_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/a_tests.log
timeout 300 python tools/prof_scan_large.py 8192 16 40000000 4096 query,partition16,vector > gpurun_out/a_scan_32ppl.log 2>&1
timeout 300 python tools/prof_scan_large.py 4096 16 10000000 1024 partition16,vector > gpurun_out/a_scan_64ppl.log 2>&1
timeout 300 python tools/prof_scan_large.py 2048 16 40000000 4096 query,vector > gpurun_out/a_scan_8ppl.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-scan-large > gpurun_out/a_bench_default.json 2> gpurun_out/a_bench_default.err
FDB_VSCAN_DEFAULT=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-scan-large > gpurun_out/a_bench_vscan.json 2> gpurun_out/a_bench_vscan.err
tail -3 gpurun_out/a_tests.log; cat gpurun_out/a_scan_*.log | grep -v "^$"
