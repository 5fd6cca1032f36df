#!/bin/bash
# Round 2, GPU call 22: y3 of every group addition as one multi-product multiplication (mul_diff) -- suite, A/B, bench
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_run22.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_run22.log
for v in std g1_acc_no_mul_diff std g1_acc_no_mul_diff; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g1 2^24: "; timeout 300 python $T --group g1 --log-n 24 --steps 5 $L 2>&1 | tail -1 | cut -c1-420
done | tee $OUT/lab_mul_diff.txt
for v in std g2_acc_no_mul_diff; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g2 2^20: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 $L 2>&1 | tail -1 | cut -c1-420
done | tee -a $OUT/lab_mul_diff.txt
timeout 900 python bench.py > $OUT/bench_run22.json 2> $OUT/bench_run22.err; echo "bench rc=$?"; cut -c1-900 $OUT/bench_run22.json
timeout 600 python zero-knowledge-proofs_b200/tools/bench_setup.py --log-n 22 --steps 3 2>&1 | cut -c1-260 | tee $OUT/setup_run22.jsonl
