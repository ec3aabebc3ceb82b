#!/bin/bash
# round 2, late call (2 GPUs): the device-memory cache under the sharded paths
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check2.py > gpurun_out/n2_dist_check2.log 2>&1
echo "exit $?" >> gpurun_out/n2_dist_check2.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 1 > gpurun_out/n2_bench_g2.json 2> gpurun_out/n2_bench_g2.err ) 2> gpurun_out/n2_bench_g2.time
echo "exit $?" >> gpurun_out/n2_bench_g2.err
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "two_gpu or 2_gpu or sharded or library_owned" > gpurun_out/n2_tests.log 2>&1
tail -3 gpurun_out/n2_dist_check2.log | cut -c1-300; tail -c 300 gpurun_out/n2_bench_g2.err; cat gpurun_out/n2_bench_g2.time; tail -2 gpurun_out/n2_tests.log
