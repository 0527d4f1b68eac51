#!/bin/bash
# Runs on the GPU box (under gpurun): GPU test suite, both bench modes, then the ncu launch list and one
# full capture of the hot kernels for MODE (preset|live).  Outputs land in gpurun_out/ with TAG in their names.
#   tools/gpu_profile.sh TAG [MODE] [skip-tests]
TAG=${1:-r01}; MODE=${2:-live}; SKIPT=${3:-}
mkdir -p gpurun_out
if [ -z "$SKIPT" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_$TAG.log
fi
python bench.py --mode preset > gpurun_out/bench_${TAG}_preset.json 2> gpurun_out/bench_${TAG}_preset.err; echo "bench preset rc=$?"
python bench.py --mode live --steps 8 --warmup 3 > gpurun_out/bench_${TAG}_live.json 2> gpurun_out/bench_${TAG}_live.err; echo "bench live rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; echo "bench reference rc=$?"
CMD="python bench.py --mode $MODE --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_$MODE.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'dist_kernel|eval_kernel|energy_kernel' -s 9 -c 3 -f -o gpurun_out/prof_${TAG}_$MODE $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu rc=$?"
cat gpurun_out/bench_${TAG}_preset.json gpurun_out/bench_${TAG}_live.json gpurun_out/bench_${TAG}_reference.json | cut -c1-1500
