#!/bin/bash
# Round 2, one-GPU pass b: every GPU test (incl. the reference-pinned ones and the bench-configuration test), then bench.py
# with the streamed end-to-end leg and the per-config rates.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest.log 2>&1
tail -5 gpurun_out/r2b_pytest.log
timeout 400 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 6000 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
