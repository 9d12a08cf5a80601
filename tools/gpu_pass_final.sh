#!/bin/bash
# Round 2, one-GPU evidence pass: every GPU test, the default bench line, a sustained (>= 2 s timed) line, a jittered-mesh
# line, the polar workload, the ncu launch list of the bench command and ncu --set full captures of the three hot kernels.
mkdir -p gpurun_out
T=${TAG:-r2f}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --steps 600 --no-configs --no-cpu-baseline > gpurun_out/${T}_bench_sustained.json 2> gpurun_out/${T}_bench_sustained.err
timeout 600 python bench.py --jitter 0.2 --no-configs --no-cpu-baseline > gpurun_out/${T}_bench_jitter.json 2> gpurun_out/${T}_bench_jitter.err
timeout 600 python bench.py --workload polar --steps 20 > gpurun_out/${T}_bench_polar.json 2> gpurun_out/${T}_bench_polar.err
for f in bench bench_sustained bench_jitter bench_polar; do
  tail -1 gpurun_out/${T}_$f.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'serial', round(d['e2e']['serial']['value']), 'frac', round(d['roofline']['frac'],3), 'kernel ms', round(d['roofline']['ms_per_launch'],4), d['clocks'])" || tail -3 gpurun_out/${T}_$f.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-sustained > gpurun_out/${T}_bench_under_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_jacobi_tiles|k_jacobi_apply" -s 4 -c 2 -f -o gpurun_out/${T}_prof_tiles \
    python tools/profile_driver.py --cluster-size 512 --steps 2 > gpurun_out/${T}_prof_tiles.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_polar_tiles|k_polar_vertex_tiles" -s 40 -c 2 -f -o gpurun_out/${T}_prof_polar \
    python bench.py --workload polar --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-sustained > gpurun_out/${T}_prof_polar.log 2>&1
