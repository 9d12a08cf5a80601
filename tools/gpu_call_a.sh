#!/bin/bash
# One-GPU measurement pass: tile-kernel variant sweep, headline bench, ncu full captures, GPU parity tests.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
date +%s > gpurun_out/a_t0
timeout 300 python tools/tile_sweep.py --sizes ${SWEEP_SIZES:-256,512,128} --out gpurun_out/tile_sweep.txt > gpurun_out/a_sweep.log 2>&1
date +%s > gpurun_out/a_t1
timeout 180 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
date +%s > gpurun_out/a_t2
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_tiles -s 2 -c 1 -f -o gpurun_out/a_prof_default \
    python tools/profile_driver.py > gpurun_out/a_prof_default.log 2>&1
TETSIM_TILE_TPT=2 TETSIM_TILE_STAGES=3 timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_tiles -s 2 -c 1 -f -o gpurun_out/a_prof_tpt2 \
    python tools/profile_driver.py > gpurun_out/a_prof_tpt2.log 2>&1
date +%s > gpurun_out/a_t3
timeout 330 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1
date +%s > gpurun_out/a_t4
tail -3 gpurun_out/a_pytest.log; cat gpurun_out/tile_sweep.txt; tail -c 600 gpurun_out/a_bench.json
