#!/bin/bash
# round 2, late call (2 GPUs): the N > 1 bench line after the last bench.py edit
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 1 --no-sharded-query --no-sift --no-single-gpu-base > gpurun_out/n2b_bench_g2.json 2> gpurun_out/n2b_bench_g2.err ) 2> gpurun_out/n2b_bench_g2.time
echo "exit $?" >> gpurun_out/n2b_bench_g2.err
tail -c 300 gpurun_out/n2b_bench_g2.err; cat gpurun_out/n2b_bench_g2.time; python -c "import json; d=json.load(open('gpurun_out/n2b_bench_g2.json')); print(d['value'], d['e2e'])"
