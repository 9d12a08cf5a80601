#!/bin/bash
# First one-GPU pass of round 2: everything round 1 built but could not measure (its GPU budget ran out).
#   1. all GPU tests, then the single-process peer protocol under every experiment mask
#   2. bench.py as is (vertex kernel without the dead velocity store: expect ~-4 us per substep against 3.69 ms/step)
#   3. bench.py with TETSIM_APPLY_INLINE=1 (vertex kernel: partial slots in one record)
#   4. tile sweep (defaults must still read 0.141 ms at T = 512)
#   5. ncu --set full of the tile kernel and of the vertex kernel (both variants)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest.log 2>&1
TETSIM_TEST_PEER_V2="1 2 4 8 16 31" timeout 200 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -k single_process > gpurun_out/r2_peer_masks.log 2>&1
timeout 200 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
TETSIM_APPLY_INLINE=1 timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_inline.json 2> gpurun_out/r2_bench_inline.err
timeout 200 python tools/tile_sweep.py --sizes 512,256 --out gpurun_out/r2_tile_sweep.txt > gpurun_out/r2_sweep.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_jacobi_tiles|k_jacobi_apply" -s 4 -c 2 -f -o gpurun_out/r2_prof \
    python tools/profile_driver.py --cluster-size 512 --steps 2 > gpurun_out/r2_prof.log 2>&1
TETSIM_APPLY_INLINE=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_apply -s 2 -c 1 -f -o gpurun_out/r2_prof_inline \
    python tools/profile_driver.py --cluster-size 512 --steps 2 > gpurun_out/r2_prof_inline.log 2>&1
tail -3 gpurun_out/r2_pytest.log; tail -3 gpurun_out/r2_peer_masks.log
for f in gpurun_out/r2_bench.json gpurun_out/r2_bench_inline.json; do
  tail -1 $f | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3))" || tail -3 ${f%.json}.err
done
cat gpurun_out/r2_tile_sweep.txt
