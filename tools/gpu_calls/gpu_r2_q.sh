#!/bin/bash
# round 2, call Q (8 GPUs): sharded bench at 8 ranks (the driver's N = 8 command, fewer steps), library-owned comm vs oracle
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/q_topo.log 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 1 > gpurun_out/q_bench_g8.json 2> gpurun_out/q_bench_g8.err ) 2> gpurun_out/q_bench_g8.time
echo "exit $?" >> gpurun_out/q_bench_g8.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check2.py > gpurun_out/q_dist_check2.log 2>&1
echo "exit $?" >> gpurun_out/q_dist_check2.log
tail -3 gpurun_out/q_dist_check2.log; tail -c 600 gpurun_out/q_bench_g8.err; cat gpurun_out/q_bench_g8.time
