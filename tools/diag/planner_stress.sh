#!/bin/bash
# repeat the host planner's GPU check to catch rare host-memory faults: tools/diag/planner_stress.sh [runs]
python - <<'PY'
import sys, pathlib
sys.path.insert(0, '.')
from tests.test_host_planner import _dump_scene
p = pathlib.Path('/tmp/planner_scene'); p.mkdir(exist_ok=True)
print(_dump_scene(p))
PY
fail=0
for i in $(seq ${1:-30}); do
  MALLOC_CHECK_=3 MALLOC_PERTURB_=165 graspit_b200/build/b200planner_check gpu /tmp/planner_scene/scene.txt > /tmp/planner_out.txt 2> /tmp/planner_err.txt
  rc=$?
  if [ $rc -ne 0 ]; then fail=$((fail+1)); echo "run $i rc=$rc"; head -40 /tmp/planner_err.txt; tail -5 /tmp/planner_out.txt; fi
done
echo "failures: $fail of ${1:-30}"
