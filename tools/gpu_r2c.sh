python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench1.json 2> gpurun_out/r2c_bench1.err; tail -3 gpurun_out/r2c_bench1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench1.json'))
print({k:d[k] for k in ('ms_per_step','value','gpu_launches','loss')}, d['e2e']['value'])
r=d['roofline']; print({k:r[k] for k in ('frac','launch_ms','fwd_kernel_ms','step_frac_of_peak')})
print(d['extra']['cfg2']['ms_per_step'], d['extra']['cfg2']['step_frac_of_peak'], d['extra']['cfg4_zeroshot']['roofline']['frac'])
PY
