#!/bin/bash
# One-GPU pass: tile-kernel variant sweep + GPU parity tests (no profiler).
mkdir -p gpurun_out
timeout 300 python tools/tile_sweep.py --sizes ${SWEEP_SIZES:-256,512} --out gpurun_out/tile_sweep_b.txt > gpurun_out/b_sweep.log 2>&1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1
tail -5 gpurun_out/b_pytest.log; cat gpurun_out/tile_sweep_b.txt || tail -20 gpurun_out/b_sweep.log
