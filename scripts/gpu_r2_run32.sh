#!/bin/bash
# Round 2, GPU call 32: G2 ReduceLevel in two loops (suffix sums parked in place, then their sum) -- parity and A/B
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "g2 or adversarial or golden" > $OUT/pytest_run32.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_run32.log
for v in std g2_red_one_loop std g2_red_one_loop; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g2 2^20: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 $L 2>&1 | tail -1 | cut -c1-420
done | tee $OUT/lab_g2_red_two_loops.txt
echo -n "std g2 2^20 u64: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 --bits 64 --precompute-bits 17 2>&1 | tail -1 | cut -c1-420 | tee -a $OUT/lab_g2_red_two_loops.txt
