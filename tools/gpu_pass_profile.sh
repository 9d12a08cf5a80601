#!/bin/bash
# One-GPU measurement pass: tile-kernel variant sweep, headline bench, ncu full captures, GPU parity tests.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/p_smi.txt 2>&1
date +%s > gpurun_out/p_t0
timeout 300 python tools/tile_sweep.py --sizes ${SWEEP_SIZES:-256,512,128} --out gpurun_out/tile_sweep.txt > gpurun_out/p_sweep.log 2>&1
date +%s > gpurun_out/p_t1
timeout 180 python bench.py > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
date +%s > gpurun_out/p_t2
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_tiles -s 2 -c 1 -f -o gpurun_out/p_prof_default \
    python tools/profile_driver.py > gpurun_out/p_prof_default.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_tiles -s 2 -c 1 -f -o gpurun_out/p_prof_512 \
    python tools/profile_driver.py --cluster-size 512 > gpurun_out/p_prof_512.log 2>&1
date +%s > gpurun_out/p_t3
timeout 330 python -m pytest tests -m gpu -x -q > gpurun_out/p_pytest.log 2>&1
date +%s > gpurun_out/p_t4
tail -3 gpurun_out/p_pytest.log; cat gpurun_out/tile_sweep.txt; tail -c 600 gpurun_out/p_bench.json
