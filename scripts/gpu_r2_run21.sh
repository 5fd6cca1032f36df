#!/bin/bash
# Round 2, GPU call 21: two-product Fq2 multiplication as the default -- whole GPU suite, then per-kernel A/B against Karatsuba
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_run21.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_run21.log
for v in std g2_acc_karatsuba g2_red_karatsuba g2_red_inline_dual g2_fb_karatsuba g2_pre_karatsuba g2_pair; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g2 2^20: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 $L 2>&1 | tail -1 | cut -c1-420
done | tee $OUT/lab_g2_dual.txt
echo -n "std g2 2^20 u64: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 --bits 64 2>&1 | tail -1 | cut -c1-420 | tee -a $OUT/lab_g2_dual.txt
timeout 600 python zero-knowledge-proofs_b200/tools/bench_prove.py --log-n 20 --steps 5 --no-cpu 2>&1 | grep gpu_ms | cut -c1-300 | tee $OUT/prove_run21.txt
