#!/bin/bash
# One-GPU kernel iteration pass: Jacobi parity tests, tile sweep (T = 512 and 256 defaults only unless SWEEP_ALL=1), bench, optional ncu.
mkdir -p gpurun_out
TAG=${TAG:-k}
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "jacobi or config4 or peer" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/tile_sweep.py --sizes ${SWEEP_SIZES:-512,256} ${SWEEP_ALL:+--all} --out gpurun_out/${TAG}_tile_sweep.txt > gpurun_out/${TAG}_sweep.log 2>&1
cat gpurun_out/${TAG}_tile_sweep.txt || tail -5 gpurun_out/${TAG}_sweep.log
timeout 300 python bench.py --no-configs --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -1 gpurun_out/${TAG}_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'serial', round(d['e2e']['serial']['value']), 'frac', round(d['roofline']['frac'],3), 'tile ms', round(d['roofline']['ms_per_launch'],4), d['clocks'])" || tail -5 gpurun_out/${TAG}_bench.err
if [ -n "$NCU" ]; then
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_jacobi_tiles|k_jacobi_apply" -s 4 -c 2 -f -o gpurun_out/${TAG}_prof \
      python tools/profile_driver.py --cluster-size 512 --steps 2 > gpurun_out/${TAG}_prof.log 2>&1
fi
