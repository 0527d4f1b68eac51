#!/bin/bash
# A/B of runtime switches on the same GPU box; usage: tools/ab_env.sh MODE "VAR=val ..." "VAR=val ..." ...  ("-" = no overrides)
MODE=${1:-preset}; shift
for E in "$@"; do
  echo "== $E"
  if [ "$E" = "-" ]; then E=""; fi
  env $E python bench.py --mode $MODE --steps 20 --warmup 5 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); st=json.loads(d['roofline']['note'].split('step: ')[-1])
print('poses/s %.0f  e2e %.0f' % (d['value'], d['e2e']['value']), {k: round(v,3) for k,v in d['roofline']['kernel_ms_all'].items()}, 'pbox %d ptri %d' % (st['point_box_tests'], st['point_tri_tests']))"
done
