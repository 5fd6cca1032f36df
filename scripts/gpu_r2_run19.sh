#!/bin/bash
# Round 2, GPU call 19: G2 accumulate with the Fq2 multiplication as a call (code 10x smaller) -- A/B
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
for v in std g2_acc_fq2_calls g2_acc_fq2_calls_mb6; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g2 2^20: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 $L 2>&1 | tail -1 | cut -c1-420
done | tee $OUT/lab_g2_acc_fq2_calls.txt
