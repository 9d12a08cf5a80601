#!/bin/bash
# Round 2, one-GPU pass c: every GPU test, the full bench line (configs included), the polar workload, ncu of the polar tile kernel.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1
tail -12 gpurun_out/r2c_pytest.log
timeout 500 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -1 gpurun_out/r2c_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'serial', round(d['e2e']['serial']['value']), 'frac', round(d['roofline']['frac'],3), d['clocks'])
for k,v in d['configs'].items():
    if 'substeps_per_s' in v: print(' ', k, round(v['substeps_per_s']), round(v['Mtet_per_s'],1))
" || tail -5 gpurun_out/r2c_bench.err
timeout 300 python bench.py --workload polar --steps 10 > gpurun_out/r2c_polar.json 2> gpurun_out/r2c_polar.err
tail -1 gpurun_out/r2c_polar.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('polar value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'kernel ms', round(d['roofline']['ms_per_launch'],4), 'cpu', d['cpu_baseline']['value'])" || tail -5 gpurun_out/r2c_polar.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_polar_tiles|k_polar_vertex_tiles" -s 4 -c 2 -f -o gpurun_out/r2c_prof_polar \
    python bench.py --workload polar --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2c_prof_polar.log 2>&1
