#!/bin/bash
# N-GPU pass (gpurun --gpus N -- 'NGPU=N bash tools/gpu_pass_multigpu.sh'): correctness of the exchanges through real
# cudaIpc / NCCL, then bench.py at N for every exchange mode (and every TETSIM_PEER_V2 experiment mask for the peer
# exchange), then one traced run per mode (TETSIM_TRACE=1: per-launch intervals on every rank).
#   EXCHANGES="peer halo allreduce"  TILES="512 256"  PEER_V2S="0 2 4 8 16 31"  TRACE=1  CHECKS="peer"
N=${NGPU:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for X in ${CHECKS:-peer}; do
  timeout 200 bash -c "$(declare -f run); N=$N; run 2954$N tools/multigpu_check.py --exchange $X --cluster-size ${CHECK_T:-256}" 2>&1 \
      | grep -E "world|FAIL|rror|Traceback" | sed "s/^/$X /"
done
for X in ${EXCHANGES:-peer halo allreduce}; do
  for T in ${TILES:-512}; do
    if [ $X = peer ]; then VS=${PEER_V2S:-0}; else VS=0; fi
    for V in $VS; do
      L=gpurun_out/mg_${X}_v${V}_${N}_$T.log
      TETSIM_PEER_V2=$V timeout 300 bash -c "$(declare -f run); N=$N; run 2955$N bench.py --gpus $N --steps 30 --exchange $X --cluster-size $T" > $L 2>&1
      tail -1 $L | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$X v2=$V T=$T', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'bverts', d['config']['boundary_verts'], 'btiles', d['config'].get('boundary_tiles_rank0'), 'tiles', d['config']['clusters_rank0'], 'tile ms', round(d['roofline']['ms_per_launch'],4))" || tail -5 $L
    done
    if [ -n "$TRACE" ]; then
      TETSIM_TRACE=1 timeout 300 bash -c "$(declare -f run); N=$N; run 2956$N bench.py --gpus $N --steps 6 --warmup 1 --no-e2e --exchange $X --cluster-size $T" 2>&1 \
          | grep -A12 "tetsim trace, rank 0 " | head -14 | sed "s/^/$X T=$T /"
    fi
  done
done
