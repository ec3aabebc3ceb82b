#!/bin/bash
# round 2, very last call: the N = 1 bench line of the final tree (both arms)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err ) 2> gpurun_out/fin_bench.time
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fin_bench_ref.json 2> gpurun_out/fin_bench_ref.err
cat gpurun_out/fin_bench.time
