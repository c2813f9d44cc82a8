#!/bin/bash
# usage: bash tools/gpu_run.sh <tag> "<pytest -k expression or empty>" [extra commands...]   -- one gpurun call's worth of work
tag=$1; sel=$2; shift 2
mkdir -p gpurun_out
if [ -n "$sel" ]; then
  python -m pytest tests -m gpu -q -p no:cacheprovider -k "$sel" 2>&1 | tail -${TAILN:-40} > gpurun_out/${tag}_tests.log
  tail -${TAILN:-40} gpurun_out/${tag}_tests.log
fi
for c in "$@"; do echo "+ $c"; eval "$c"; done
