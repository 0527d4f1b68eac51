#!/bin/bash
# A/B of engine variants on the GPU box: parity of the distance tests + LIVE bench kernel times per variant
# usage: tools/ab_live.sh [mode] lib1.so lib2.so ...   ("default" = product build)
MODE=${1:-live}; shift
for L in "$@"; do
  if [ "$L" = default ]; then unset B2C_LIB; else export B2C_LIB=$PWD/$L; fi
  echo "== $L"
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
  python bench.py --mode $MODE --steps 5 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); st=json.loads(d['roofline']['note'].split('traversal work per step: ')[1])
print('poses/s %.0f' % d['value'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_all'].items()}, 'dbox %d dtri %d' % (st['dist_box_tests'], st['dist_tri_tests']))"
done
