#!/bin/bash
# Round 2, last GPU call: GPU suite + smoke on the final build (G2 tiles 8 x 32), then tile-shape A/B for G1 and K = 16 for G2
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/final3_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/final3_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/final3_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/final3_smoke.log
for v in std g1_tile_k8 std g1_tile_k8; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g1 2^21 c20: "; timeout 300 python $T --group g1 --log-n 21 --precompute-bits 20 --steps 10 $L 2>&1 | tail -1 | cut -c1-420
done | tee $OUT/lab_tile_shapes.txt
for v in std g2_tile_k16; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v g2 2^20: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 $L 2>&1 | tail -1 | cut -c1-420
done | tee -a $OUT/lab_tile_shapes.txt
