#!/bin/bash
# Round 2, GPU call 36: G2 tile levels with 8 entries x 32 elements per tile (128 threads, 254 registers) instead of 4 x 64
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
for v in std g2_tile_k8 std g2_tile_k8; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g2 2^20: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 $L 2>&1 | tail -1 | cut -c1-420
done | tee $OUT/lab_g2_tile_k8.txt
echo -n "g2_tile_k8 g2 2^20 u64 c17: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 --bits 64 --precompute-bits 17 --lib $LAB/g2_tile_k8.so 2>&1 | tail -1 | cut -c1-420 | tee -a $OUT/lab_g2_tile_k8.txt
