#!/bin/bash
# A/B of library variants (variants/lib<NAME>.so) on the per-term wall-factor path:
#   scripts/ab_general.sh NAME...      C5 and C2-monopole with --general-impedance
cp deism_b200/libdeism_b200.so /tmp/lib_keep.so
for v in "$@"; do
  cp variants/lib$v.so deism_b200/libdeism_b200.so
  for w in c5 c2mono c3; do
    python bench.py --workload $w --general-impedance --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v $w general ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d.get('kernel_ms_per_launch'))"
  done
done
cp /tmp/lib_keep.so deism_b200/libdeism_b200.so
