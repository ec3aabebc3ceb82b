#!/bin/bash
# round 2, call G (2 GPUs): full suite, stored-tie fix, tc_assign ncu, N=1 bench with sharded extras
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/g_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/g_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check2.py > gpurun_out/g_dist_check2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_assign_kernel' -s 7 -c 1 -o gpurun_out/g_tc_assign_pq -f python tools/prof_tc.py > gpurun_out/g_ncu_tc.log 2>&1
CUDA_VISIBLE_DEVICES=0 timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench exit $?" >> gpurun_out/g_bench.err
tail -5 gpurun_out/g_tests.log; tail -3 gpurun_out/g_dist_check2.log; tail -c 800 gpurun_out/g_bench.err
