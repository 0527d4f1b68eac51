#!/bin/bash
# Runs on the GPU box (under gpurun): GPU test suite, both bench modes + the reference arm, then the ncu launch list
# and full captures of the hot kernels in both modes.  Outputs land in gpurun_out/ with TAG in their names.
#   tools/gpu_profile.sh TAG [skip-tests]
TAG=${1:-r01}; SKIPT=${2:-}
mkdir -p gpurun_out
if [ -z "$SKIPT" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; grep -v "Collision found" gpurun_out/tests_$TAG.log | tail -3
fi
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; echo "bench reference rc=$?"
python bench.py --mode preset > gpurun_out/bench_${TAG}_preset.json 2> gpurun_out/bench_${TAG}_preset.err; echo "bench preset rc=$?"
python bench.py --mode live --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_live.json 2> gpurun_out/bench_${TAG}_live.err; echo "bench live rc=$?"
P="python bench.py --mode preset --steps 2 --warmup 3 --no-cpu-baseline"
L="python bench.py --mode live --steps 2 --warmup 3 --no-cpu-baseline"
$P > gpurun_out/plain_${TAG}_preset.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_preset.csv $P > gpurun_out/ncu_launches_${TAG}_preset.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fk_kernel|broad_kernel|narrow_kernel|energy_kernel' -s 16 -c 4 -f -o gpurun_out/prof_${TAG}_preset $P > gpurun_out/ncu_full_${TAG}_preset.log 2>&1
echo "ncu preset rc=$?"
$L > gpurun_out/plain_${TAG}_live.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_live.csv $L > gpurun_out/ncu_launches_${TAG}_live.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'dist_kernel|energy_kernel' -s 8 -c 2 -f -o gpurun_out/prof_${TAG}_live $L > gpurun_out/ncu_full_${TAG}_live.log 2>&1
echo "ncu live rc=$?"
cat gpurun_out/bench_${TAG}_preset.json gpurun_out/bench_${TAG}_live.json gpurun_out/bench_${TAG}_reference.json | cut -c1-1200
