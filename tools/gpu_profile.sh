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
# config 3 (batched contacts + region batch): tool line with the phase timing, launch list, one full capture per kernel
T="python tools/contacts_bench.py 16384"
$T > gpurun_out/bench_${TAG}_contacts.json 2> gpurun_out/bench_${TAG}_contacts.err &&
B2C_TIMING=1 $T 2>&1 | grep "contacts_batch:\|region_batch:" | tail -2 >> gpurun_out/bench_${TAG}_contacts.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}_config3.csv $T > gpurun_out/ncu_launches_${TAG}_config3.log 2>&1
# (the warm-up call of the tool runs cull + contact kernel twice -- its candidate buffer starts from an estimate -- hence -s 4)
ncu --set full --clock-control none --import-source on -k regex:'contact_kernel|contact_cull_kernel' -s 4 -c 2 -f -o gpurun_out/prof_${TAG}_contact $T > gpurun_out/ncu_full_${TAG}_contact.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'contact_order_kernel' -s 8 -c 1 -f -o gpurun_out/prof_${TAG}_order $T > gpurun_out/ncu_full_${TAG}_order.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'region_batch_kernel|region_keep_kernel|region_write_kernel' -s 3 -c 3 -f -o gpurun_out/prof_${TAG}_region $T > gpurun_out/ncu_full_${TAG}_region.log 2>&1
echo "ncu config3 rc=$?"
