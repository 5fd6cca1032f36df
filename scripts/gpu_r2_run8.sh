#!/bin/bash
# Round 2, GPU call 8: prove schedule with the MSM tails on high-priority streams (A/B against the same build without),
# prove parity tests, and the ncu --set full capture of the FULL-SIZE accumulate launch at 2^24.
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
timeout 900 python -m pytest tests/test_gpu_prove.py -m gpu -x -q > $OUT/pytest_prove.log 2>&1; echo "pytest prove rc=$?"; tail -3 $OUT/pytest_prove.log
for v in std prove_no_split_tail std prove_no_split_tail; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo "== $v"; timeout 600 python zero-knowledge-proofs_b200/tools/bench_prove.py --log-n 20 --steps 5 --no-cpu $L 2>&1 | tail -2 | cut -c1-260
done | tee $OUT/lab_prove_split_tail.txt
echo "== ncu --set full: BucketAccumulate<Fq> at 2^24, the full-size launch (second launch of the run)"
B="--steps 1 --warmup 1 --no-cpu-baseline --no-prove --no-oneshot"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:BucketAccumulate -s 1 -c 1 \
    -o $OUT/r02_acc_g1_2p24_full python bench.py $B > $OUT/ncu_acc.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_acc.log
