#!/bin/bash
# One-GPU pass: peer-exchange protocol tests (fused + unfused), sweep of the defaults, GPU tests.
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -k single_process > gpurun_out/f_peer.log 2>&1
timeout 200 python tools/tile_sweep.py --sizes 512,256 --out gpurun_out/tile_sweep_f.txt > gpurun_out/f_sweep.log 2>&1
timeout 300 python -m pytest tests -m gpu -q --deselect tests/test_multigpu_gpu.py > gpurun_out/f_pytest.log 2>&1
tail -25 gpurun_out/f_peer.log; tail -4 gpurun_out/f_pytest.log; cat gpurun_out/tile_sweep_f.txt || tail -20 gpurun_out/f_sweep.log
