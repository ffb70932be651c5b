#!/bin/bash
# One ncu --set full capture of a kernel of a bench workload, after the same command exited 0 without ncu
# (B200_PROFILING.md).  Usage (under gpurun): profiles/tools/ncu_capture.sh <workload> <kernel regex> <tag> [frames] [skip] [count]
set -u
WL=$1; K=$2; TAG=$3; FR=${4:-16}; SKIP=${5:-40}; CNT=${6:-2}
CMD="python bench.py --workload $WL --frames $FR --steps 1 --warmup 1 --no-e2e --no-cpu --no-also"
$CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err && \
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu_capture $TAG rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
