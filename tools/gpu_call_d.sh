#!/bin/bash
# One-GPU pass: peer-exchange protocol test first (bounded), sweep, all GPU tests, headline bench.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -k single_process > gpurun_out/d_peer.log 2>&1
timeout 300 python tools/tile_sweep.py --sizes ${SWEEP_SIZES:-256,512,128} --out gpurun_out/tile_sweep_d.txt > gpurun_out/d_sweep.log 2>&1
timeout 300 python -m pytest tests -m gpu -q --deselect tests/test_multigpu_gpu.py > gpurun_out/d_pytest.log 2>&1
timeout 200 python bench.py > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
tail -25 gpurun_out/d_peer.log; tail -4 gpurun_out/d_pytest.log; cat gpurun_out/tile_sweep_d.txt || tail -20 gpurun_out/d_sweep.log; tail -c 1500 gpurun_out/d_bench.json; tail -3 gpurun_out/d_bench.err
