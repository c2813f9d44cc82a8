#!/bin/bash
# usage: bash tools/gpu_n.sh <N> <tag>    -- multi-GPU checks: dp test (N>=2), bench at N, timeline at N
N=$1; tag=$2
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dp.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
PYTHONFAULTHANDLER=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_n$N.json'))
print({k:d[k] for k in ('n_gpus','ms_per_step','value','gpu_launches','loss')}, 'e2e', d['e2e']['value'])
print('parity', d['parity'])
print('roofline', {k:d['roofline'][k] for k in ('launch_ms','fwd_kernel_ms','step_frac_of_peak')})
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tools/dp_timeline.py 32768 2>/dev/null | grep -v "^\*\|OMP\|Warn\|_warn" | cut -c1-140 > gpurun_out/${tag}_timeline$N.log
head -3 gpurun_out/${tag}_timeline$N.log
