mkdir -p gpurun_out/r02
python -m pytest tests/test_gpu_shoebox.py -x -q -m gpu -k "org or c3 or mix or c5 or frequency_dependent" 2>&1 | tail -3
for w in c3; do
python bench.py --workload $w --no-configs --steps 10 --warmup 3 > gpurun_out/r02/u_$w.json 2> gpurun_out/r02/u_$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02/u_$w.json').read().strip().split('\n')[-1])
print('$w', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d.get('kernel_ms_per_launch'), d['roofline']['frac'], d.get('parity_rel_l2_vs_reference'))
PY
done
