#!/bin/bash
# round 2, call T: tiled seeding with batched loads; attributes + async mirror on the GPU; C++ example; ncu of the tile kernel
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "seeding or full_build or golden or database_builder or cpp or native or stored or example" > gpurun_out/t_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t_tests.log
timeout 300 python tools/prof_cfg2.py 125000 100 3 > gpurun_out/t_cfg2_125k.log 2>&1
timeout 300 python tools/prof_cfg2.py 1000000 100 3 > gpurun_out/t_cfg2_1m.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t_launches_cfg2_125k.csv python tools/prof_cfg2.py 125000 3 1 > gpurun_out/t_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'seed_round_tile_kernel' -s 20 -c 1 -o gpurun_out/t_seed_tile -f python tools/prof_cfg2.py 125000 3 1 > gpurun_out/t_ncu3.log 2>&1
tail -3 gpurun_out/t_tests.log; cat gpurun_out/t_cfg2_*.log
