#!/bin/bash
# A/B timing of library variants under variants/lib<NAME>.so on the same box: scripts/ab_lc.sh NAME...
cp deism_b200/libdeism_b200.so /tmp/lib_keep.so
for v in "$@"; do
  name=${v%%:*}; env_kv=""
  [[ "$v" == *:* ]] && env_kv=${v#*:}
  cp variants/lib$name.so deism_b200/libdeism_b200.so
  echo "== $v"
  env $env_kv python scripts/quick_lc_perf.py 2>&1 | grep -v Warn
done
cp /tmp/lib_keep.so deism_b200/libdeism_b200.so
