#!/bin/bash
# Final N-GPU pass (gpurun --gpus N -- 'NGPU=N bash tools/gpu_pass_final_multigpu.sh'): the multi-GPU tests (N >= 2), the
# exchange check through real cudaIpc, and bench.py launched exactly as the driver launches it.
N=${NGPU:-2}
T=${TAG:-r2f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q > gpurun_out/${T}_pytest_${N}gpu.log 2>&1
tail -3 gpurun_out/${T}_pytest_${N}gpu.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N tools/multigpu_check.py --exchange peer --cluster-size 256 2>&1 | grep -E "world|FAIL|rror|Traceback"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
tail -1 gpurun_out/${T}_bench_${N}gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'sustained', round(d['sustained']['value']), 'parity', d.get('parity'), 'c5', d['configs'].get('C5_784x_gs_exact_fast_sharded'), 'tile ms', round(d['roofline']['ms_per_launch'],4), d['clocks'])" || tail -5 gpurun_out/${T}_bench_${N}gpu.err
