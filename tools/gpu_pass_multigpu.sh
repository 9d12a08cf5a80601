#!/bin/bash
# N-GPU pass (gpurun --gpus N): real cudaIpc peer-memory exchange check, then bench at N for every exchange mode.
N=${NGPU:-2}
mkdir -p gpurun_out
for X in peer; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N \
      tools/multigpu_check.py --exchange $X --cluster-size ${CHECK_T:-256} 2>&1 | grep -E "world|FAIL|rror|Traceback" | sed "s/^/$X /"
done
for X in ${EXCHANGES:-peer halo allreduce}; do
  for T in ${TILES:-512}; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N \
        bench.py --gpus $N --steps 30 --exchange $X --cluster-size $T > gpurun_out/e_${X}_${N}_$T.log 2>&1
    tail -1 gpurun_out/e_${X}_${N}_$T.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$X T=$T', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'bverts', d['config']['boundary_verts'], 'btiles', d['config'].get('boundary_tiles_rank0'), 'tiles', d['config']['clusters_rank0'], 'tile ms', round(d['roofline']['ms_per_launch'],4))" || tail -5 gpurun_out/e_${X}_${N}_$T.log
  done
done
