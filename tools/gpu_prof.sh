mkdir -p gpurun_out
python tools/nce_time.py 32768 512 > gpurun_out/r2q_nce_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:nce_fwd2 -s 2 -c 1 -o gpurun_out/r2q_prof_fwd2 python tools/nce_time.py 32768 512 > gpurun_out/r2q_ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nce_bwdc -s 1 -c 1 -o gpurun_out/r2q_prof_bwdc python tools/nce_time.py 32768 512 > gpurun_out/r2q_ncu_bwd.log 2>&1
python bench.py --steps 2 --warmup 3 --eager --no-extras --no-cpu-baseline > gpurun_out/r2q_bench_eager.json 2> gpurun_out/r2q_bench_eager.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-extras --no-cpu-baseline > gpurun_out/r2q_ncu_launch.log 2>&1
ncu --set full --clock-control none -k regex:"gemm_bf16_kernel|bce_heads_mma|layernorm_bwd" -s 30 -c 12 -o gpurun_out/r2q_prof_small python bench.py --steps 2 --warmup 3 --eager --no-extras --no-cpu-baseline > gpurun_out/r2q_ncu_small.log 2>&1
grep -v Warn gpurun_out/r2q_nce_plain.log; tail -2 gpurun_out/r2q_ncu_bwd.log | cut -c1-120; wc -l gpurun_out/r2q_launches.csv
