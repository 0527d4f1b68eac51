#!/bin/bash
# one full ncu capture of kernels matching a regex from a short bench run; usage: tools/ncu_one.sh TAG MODE KERNEL_REGEX [skip] [count]
TAG=$1; MODE=$2; K=$3; SKIP=${4:-3}; CNT=${5:-1}
CMD="python bench.py --mode $MODE --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
