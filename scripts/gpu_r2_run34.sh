#!/bin/bash
# Round 2, GPU call 34: G1 ReduceLevel in two loops / at 6 blocks per SM, per-rank configuration of the 8-GPU run and 2^24
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
for v in std g1_red_two_loops g1_red_two_loops_mb6 g1_red_mb6; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g1 2^21 c20: "; timeout 300 python $T --group g1 --log-n 21 --precompute-bits 20 --steps 10 $L 2>&1 | tail -1 | cut -c1-420
done | tee $OUT/lab_g1_red_two_loops.txt
for v in std g1_red_two_loops_mb6; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g1 2^24: "; timeout 300 python $T --group g1 --log-n 24 --steps 4 $L 2>&1 | tail -1 | cut -c1-420
done | tee -a $OUT/lab_g1_red_two_loops.txt
