set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r2a_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench1.json 2> gpurun_out/r2a_bench1.err
tail -5 gpurun_out/r2a_bench1.err
python bench.py --workload zeroshot --steps 5 --warmup 3 > gpurun_out/r2a_zs.json 2> gpurun_out/r2a_zs.err && \
ncu --set full --clock-control none --import-source on -k regex:zeroshot_kernel -s 3 -c 1 -o gpurun_out/r2a_prof_zs python bench.py --workload zeroshot --steps 3 --warmup 3 > gpurun_out/r2a_ncu_zs.log 2>&1
cat gpurun_out/r2a_tests.log | tail -30
