mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu-baseline --no-extras --no-configs 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c3', round(d['ms_per_step'],3), d['kernel_ms_per_launch'], d['parity_rel_l2_vs_reference'], d['roofline']['frac'])"
python bench.py --workload c3 --general-impedance --steps 6 --warmup 3 --no-cpu-baseline --no-extras --no-configs 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c3 general', round(d['ms_per_step'],3), d['kernel_ms_per_launch'])"
