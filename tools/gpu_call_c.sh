#!/bin/bash
# One-GPU pass: accuracy table, tile-kernel variant sweep, all GPU parity tests, ncu capture of the candidate default.
mkdir -p gpurun_out
timeout 120 python tools/jacobi_errors.py > gpurun_out/c_errors.txt 2>&1
timeout 300 python tools/tile_sweep.py --sizes ${SWEEP_SIZES:-256,512} --out gpurun_out/tile_sweep_c.txt > gpurun_out/c_sweep.log 2>&1
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1
TETSIM_TILE_TPT=2 TETSIM_TILE_STAGES=2 TETSIM_TILE_MINB=4 timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_tiles -s 2 -c 1 -f -o gpurun_out/c_prof_512_2_2_4 \
    python tools/profile_driver.py --cluster-size 512 > gpurun_out/c_prof.log 2>&1
cat gpurun_out/c_errors.txt; tail -5 gpurun_out/c_pytest.log; cat gpurun_out/tile_sweep_c.txt || tail -20 gpurun_out/c_sweep.log
