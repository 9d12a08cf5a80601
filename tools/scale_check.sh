# usage: GPUS_CHECK="2 4" GPUS_BENCH="1 2 4" EXCHANGES="allreduce halo" bash tools/scale_check.sh
for X in ${EXCHANGES:-allreduce}; do
for N in $GPUS_CHECK; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N tools/multigpu_check.py --exchange $X 2>&1 | grep -E "world|FAIL|rror" | sed "s/^/$X /"; done
for N in $GPUS_BENCH; do
 if [ $N = 1 ]; then timeout 400 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/scale_${X}_$N.log 2>&1
 else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 30 --exchange $X > gpurun_out/scale_${X}_$N.log 2>&1; fi
 tail -1 gpurun_out/scale_${X}_$N.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$X', d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['config']['boundary_verts'], d['config'].get('boundary_tiles_rank0'), d['config']['clusters_rank0'], round(d['roofline']['ms_per_launch'],4))" || tail -5 gpurun_out/scale_${X}_$N.log
done
done
