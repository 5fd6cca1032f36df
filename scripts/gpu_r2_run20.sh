#!/bin/bash
# Round 2, GPU call 20: G2 bucket accumulation on lane pairs (pair_g2.cuh) -- parity, then A/B against the one-thread kernel
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "g2 or adversarial or golden or exceptional or field or faithful or chunked or fixed_base" > $OUT/pytest_g2_pair.log 2>&1; echo "pytest g2 rc=$?"; tail -3 $OUT/pytest_g2_pair.log
for v in std g2_acc_thread g2_acc_thread_dual g2_pair_mb4 g2_pair_b128_mb3 g2_pair_b32_mb10; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g2 2^20: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 $L 2>&1 | tail -1 | cut -c1-420
done | tee $OUT/lab_g2_pair.txt
echo -n "std g2 2^20 u64: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 --bits 64 2>&1 | tail -1 | cut -c1-420 | tee -a $OUT/lab_g2_pair.txt
