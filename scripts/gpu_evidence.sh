#!/bin/bash
# Evidence for the default bench command: the bench line itself, the ncu launch list of the SAME command and one
# `ncu --set full` capture of the dominant kernel at the benchmarked size (DRAM traffic per launch).
# Usage (on the GPU box, from the repo root):  bash scripts/gpu_evidence.sh [tag]
TAG=${1:-run}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -2 $OUT/bench_default.err
cut -c1-900 $OUT/bench_default.json
# launch list of the same command (fewer steps), only after it exited 0 without ncu
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/plain_default.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 600 --csv \
    --log-file $OUT/launches_default_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:BucketAccumulate -s 2 -c 1 \
    -o $OUT/prof_acc24_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 $OUT/ncu_full.log
ncu -i $OUT/prof_acc24_$TAG.ncu-rep --page raw --csv > $OUT/prof_acc24_${TAG}_raw.csv 2>/dev/null
ls -la $OUT | grep -E "$TAG|default"
