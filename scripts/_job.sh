python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/r02/j15_bench.json 2> gpurun_out/r02/j15_bench.err; tail -3 gpurun_out/r02/j15_bench.err
DEISM_B200_GRAPHS=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/j15_bench_nograph.json 2> gpurun_out/r02/j15_bench_nograph.err; tail -3 gpurun_out/r02/j15_bench_nograph.err
