mkdir -p gpurun_out/r02
for w in c3 c5 c2; do
python bench.py --workload $w --no-configs --no-cpu-baseline --no-extras --steps 10 --warmup 3 > gpurun_out/r02/w.json 2> gpurun_out/r02/w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02/w.json').read().strip().split('\n')[-1])
print('$w', 'ms', round(d['ms_per_step'],3), d.get('kernel_ms_per_launch'), d.get('parity_rel_l2_vs_reference'))
PY
done
