#!/bin/bash
# round 2, call F (2 GPUs): library-owned comm vs oracle, sharded bench; tc_assign with resident centroids
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check2.py > gpurun_out/f_dist_check2.log 2>&1
echo "exit $?" >> gpurun_out/f_dist_check2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 1 > gpurun_out/f_bench_g2.json 2> gpurun_out/f_bench_g2.err
echo "exit $?" >> gpurun_out/f_bench_g2.err
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tc_ or lloyd or full_build or golden" > gpurun_out/f_tests_tc.log 2>&1
echo "exit $?" >> gpurun_out/f_tests_tc.log
timeout 300 python tools/prof_tc.py > gpurun_out/f_prof_tc.log 2>&1
FDB_TC_NO_BRES=1 timeout 300 python tools/prof_tc.py > gpurun_out/f_prof_tc_nobres.log 2>&1
tail -3 gpurun_out/f_dist_check2.log; tail -c 1500 gpurun_out/f_bench_g2.err; tail -3 gpurun_out/f_tests_tc.log; cat gpurun_out/f_prof_tc.log gpurun_out/f_prof_tc_nobres.log
