#!/bin/bash
# occupancy / register-budget variants of the sweep kernel (prebuilt under csrc/_variants), C3 bench each.
# usage: bench_sweep_variants.sh variant[:NF] ...
L=online_convex_optimization_b200/csrc/liboco_b200.so
cp $L /tmp/lib_orig.so
for spec in orig "$@"; do
  v=${spec%%:*}; nf=""; [[ "$spec" == *:* ]] && nf=${spec##*:}
  if [ "$v" = orig ]; then cp /tmp/lib_orig.so $L; else cp online_convex_optimization_b200/csrc/_variants/lib_$v.so $L; fi
  echo "== $spec"
  OCO_SWEEP_NF=$nf timeout 300 python bench.py --no-cpu --no-extras --steps 4 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
cp /tmp/lib_orig.so $L
