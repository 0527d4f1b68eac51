#!/bin/bash
# A/B of engine builds on the same GPU box: bench kernel times per variant; usage: tools/ab_bench.sh MODE lib1 lib2 ... ("default" = product build)
MODE=${1:-preset}; shift
for L in "$@"; do
  if [ "$L" = default ]; then unset B2C_LIB; else export B2C_LIB=$PWD/$L; fi
  echo "== $L"
  python bench.py --mode $MODE --steps 20 --warmup 5 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); st=json.loads(d['roofline']['note'].split('step: ')[-1])
print('poses/s %.0f  e2e %.0f' % (d['value'], d['e2e']['value']), {k: round(v,3) for k,v in d['roofline']['kernel_ms_all'].items()}, st)"
done
