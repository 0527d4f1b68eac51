#!/bin/bash
# Runs on the GPU box (under gpurun): the driver-style bench lines, then -- each only after the same command has exited 0
# without ncu -- the ncu launch lists and one full capture of the hot kernels of the three workloads.
#   tools/gpu_profile.sh TAG [tests]
TAG=${1:-r02}; TESTS=${2:-}
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; grep -v "Collision found" gpurun_out/tests_$TAG.log | tail -3
fi
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; echo "bench reference rc=$?"
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
NS="--no-cpu-baseline --no-sub"
P="python bench.py --steps 2 --warmup 3 $NS"
L="python bench.py --mode live --steps 2 --warmup 3 $NS"
C="python bench.py --workload config5 --steps 1 --warmup 3 $NS"
$P > gpurun_out/plain_${TAG}_preset.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_preset.csv $P > gpurun_out/ncu_launches_${TAG}_preset.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fk_kernel|broad_kernel|narrow_kernel|energy_kernel' -s 12 -c 4 -f -o gpurun_out/prof_${TAG}_preset $P > gpurun_out/ncu_full_${TAG}_preset.log 2>&1
echo "ncu preset rc=$?"
$L > gpurun_out/plain_${TAG}_live.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_live.csv $L > gpurun_out/ncu_launches_${TAG}_live.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'dist_kernel|dist_refine_kernel' -s 6 -c 2 -f -o gpurun_out/prof_${TAG}_live $L > gpurun_out/ncu_full_${TAG}_live.log 2>&1
echo "ncu live rc=$?"
$C > gpurun_out/plain_${TAG}_config5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_config5.csv $C > gpurun_out/ncu_launches_${TAG}_config5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'narrow_kernel|energy_kernel' -s 6 -c 2 -f -o gpurun_out/prof_${TAG}_config5 $C > gpurun_out/ncu_full_${TAG}_config5.log 2>&1
echo "ncu config5 rc=$?"
cut -c1-600 gpurun_out/bench_${TAG}.json gpurun_out/bench_${TAG}_reference.json
