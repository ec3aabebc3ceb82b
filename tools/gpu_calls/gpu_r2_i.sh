#!/bin/bash
# round 2, call I (8 GPUs): sharded bench at 8 and 4 ranks; library-owned comm vs oracle at 8
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/i_topo.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check2.py > gpurun_out/i_dist_check2.log 2>&1
echo "exit $?" >> gpurun_out/i_dist_check2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 1 > gpurun_out/i_bench_g8.json 2> gpurun_out/i_bench_g8.err
echo "exit $?" >> gpurun_out/i_bench_g8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 1 > gpurun_out/i_bench_g4.json 2> gpurun_out/i_bench_g4.err
echo "exit $?" >> gpurun_out/i_bench_g4.err
tail -3 gpurun_out/i_dist_check2.log; tail -c 600 gpurun_out/i_bench_g8.err; tail -c 600 gpurun_out/i_bench_g4.err
