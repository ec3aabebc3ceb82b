#!/bin/bash
# round 2, final 8-GPU record: the driver's N = 8 command on the final tree + smoke
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 1 > gpurun_out/g8_bench.json 2> gpurun_out/g8_bench.err ) 2> gpurun_out/g8_bench.time
echo "exit $?" >> gpurun_out/g8_bench.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check2.py > gpurun_out/g8_dist_check2.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > gpurun_out/g8_bench_ref.json 2> gpurun_out/g8_bench_ref.err
timeout 200 python __graft_entry__.py smoke > gpurun_out/g8_smoke.log 2>&1
tail -3 gpurun_out/g8_dist_check2.log | cut -c1-300; cat gpurun_out/g8_bench.time; tail -2 gpurun_out/g8_smoke.log
