#!/bin/bash
# Round 2, GPU call 26: where do the fronts run? (lab trace build) + the no-overlap build beside it
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
B="--steps 3 --warmup 1 --no-cpu-baseline --no-prove --no-oneshot"
timeout 600 python bench.py $B --log-n 24 --lib $LAB/pipe_trace.so > $OUT/trace_bench.json 2> $OUT/trace_err.txt; echo rc=$?; grep "pipe trace" $OUT/trace_err.txt | tail -8 | tee $OUT/pipe_trace.txt
timeout 600 python bench.py $B --log-n 24 --lib $LAB/no_front_overlap.so > $OUT/noov_bench.json 2> $OUT/noov_err.txt; echo rc=$?; tail -5 $OUT/noov_err.txt; cut -c1-300 $OUT/noov_bench.json
