#!/bin/bash
# same-box A/B of two builds of the engine on the default bench line: tools/diag/ab_lib.sh other.so [bench args]
O=$1; shift
for r in 1 2 3; do
  for L in "" "$O"; do
    B2C_LIB=$L python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-sub "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('${L:-product}', round(d['value']/1e6,2), round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['kernel_ms_all'].items()})"
  done
done
