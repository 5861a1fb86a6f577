mkdir -p gpurun_out/r02
DEISM_DEBUG=1 python bench.py --workload c3 --steps 4 --warmup 2 --no-cpu-baseline --no-extras --no-configs 2>gpurun_out/r02/j21.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c3', round(d['ms_per_step'],3), d['kernel_ms_per_launch'])"
grep deism gpurun_out/r02/j21.err | sort | uniq -c
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02/j21_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-configs > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02/j21_launches_c3.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size') if 'Grid Size' in hdr else None
for r in rows[-12:]: print(r[ki][:60], r[vi], r[gi] if gi is not None else '')
PY
