#!/bin/bash
# Round 2, GPU call 23: per-unit choice of the Fq2 multiplication form (cold units back to Karatsuba), G2 reduction occupancy A/B
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "g2 or adversarial or golden or exceptional or field or fixed_base" > $OUT/pytest_run23.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_run23.log
for v in std g2_red_mb6 g2_red_mb8 g2_red_mb12; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g2 2^20: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 $L 2>&1 | tail -1 | cut -c1-420
done | tee $OUT/lab_g2_red_occupancy.txt
timeout 600 python zero-knowledge-proofs_b200/tools/bench_setup.py --log-n 22 --steps 3 --groups g2 2>&1 | cut -c1-200 | tee $OUT/setup_run23.jsonl
timeout 600 python zero-knowledge-proofs_b200/tools/bench_setup.py --log-n 22 --steps 3 --groups g2 --lib $LAB/g2_fb_inline.so 2>&1 | cut -c1-200 | tee -a $OUT/setup_run23.jsonl
