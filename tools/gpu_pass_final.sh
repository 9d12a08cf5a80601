#!/bin/bash
# One-GPU pass: peer-exchange protocol tests, RPF experiment, all GPU tests, headline bench, ncu launch list of the bench.
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -k single_process > gpurun_out/i_peer.log 2>&1
timeout 120 python tools/tile_sweep.py --sizes 512 --out gpurun_out/tile_sweep_i.txt > gpurun_out/i_sweep.log 2>&1
timeout 200 python -m pytest tests -m gpu -q --deselect tests/test_multigpu_gpu.py > gpurun_out/i_pytest.log 2>&1
timeout 150 python bench.py > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/i_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/i_bench_under_ncu.log 2>&1
tail -15 gpurun_out/i_peer.log; tail -4 gpurun_out/i_pytest.log; cat gpurun_out/tile_sweep_i.txt || tail -20 gpurun_out/i_sweep.log; tail -c 900 gpurun_out/i_bench.json; tail -3 gpurun_out/i_bench.err
