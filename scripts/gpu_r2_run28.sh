#!/bin/bash
# Round 2, GPU call 28: launch shapes of the accumulate kernels re-swept after the multi-product multiplications
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
for v in std g1_mb5 g1_mb7 g1_b128_mb3 g1_b32_mb12 std; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g1 2^24: "; timeout 300 python $T --group g1 --log-n 24 --steps 4 $L 2>&1 | tail -1 | cut -c1-420
done | tee $OUT/lab_launch_shapes_final.txt
for v in std g2_mb2_b128 g2_b32; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g2 2^20: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 $L 2>&1 | tail -1 | cut -c1-420
done | tee -a $OUT/lab_launch_shapes_final.txt
