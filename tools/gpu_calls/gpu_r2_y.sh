#!/bin/bash
# round 2, call Y (8 GPUs): the driver's N = 8 and N = 2 commands on the final tree
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 1 > gpurun_out/y_bench_g8.json 2> gpurun_out/y_bench_g8.err ) 2> gpurun_out/y_bench_g8.time
echo "exit $?" >> gpurun_out/y_bench_g8.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 1 --no-sharded-query > gpurun_out/y_bench_g2.json 2> gpurun_out/y_bench_g2.err ) 2> gpurun_out/y_bench_g2.time
echo "exit $?" >> gpurun_out/y_bench_g2.err
timeout 120 python bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > gpurun_out/y_bench_ref_g8.json 2> gpurun_out/y_bench_ref_g8.err
tail -c 400 gpurun_out/y_bench_g8.err; cat gpurun_out/y_bench_g8.time; tail -c 400 gpurun_out/y_bench_g2.err; cat gpurun_out/y_bench_g2.time
