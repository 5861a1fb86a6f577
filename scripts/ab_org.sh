#!/bin/bash
# A/B timing of library variants (variants/lib<NAME>.so) on the ORG workload C3: scripts/ab_org.sh NAME...
# ("cur" = the library in the tree)
export DEISM_B200_SKIP_HASH=1
cp deism_b200/libdeism_b200.so /tmp/lib_keep.so
for v in "$@"; do
  if [ "$v" != "cur" ]; then cp variants/lib$v.so deism_b200/libdeism_b200.so; else cp /tmp/lib_keep.so deism_b200/libdeism_b200.so; fi
  echo "== $v"
  python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-configs 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C3 ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['roofline']['kernel'], d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'], d.get('kernel_ms_per_launch'), 'parity', d.get('parity_rel_l2_vs_reference'))"
done
cp /tmp/lib_keep.so deism_b200/libdeism_b200.so
