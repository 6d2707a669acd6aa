#!/bin/bash
# tuning sweep for the wide-d TMA kernel (lanes per trajectory at d = 64)
for g in 8 16 32; do
  for mode in linear two_pass; do
    echo "== G=$g $mode"
    OCO_WIDE_G=$g timeout 300 python bench.py --workload c4 --trials ${TRIALS:-9472} --comparator $mode --steps 5 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['ms_per_step'])"
  done
done
