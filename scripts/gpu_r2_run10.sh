#!/bin/bash
# Round 2, GPU call 10: the re-ordered prove schedule (sort stages first, chained accumulations, scalar multiplications of
# pi_A / pi_B' under later accumulations): parity, timeline, timing; then the whole suite.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_prove.py -m gpu -x -q > $OUT/pytest_prove.log 2>&1; echo "pytest prove rc=$?"; tail -3 $OUT/pytest_prove.log
timeout 600 python zero-knowledge-proofs_b200/tools/bench_prove.py --log-n 20 --steps 5 --no-cpu --timeline 2>&1 | grep -E "timeline|gpu_ms" | cut -c1-200 | tee $OUT/prove_timeline.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python zero-knowledge-proofs_b200/tools/bench_r1cs.py 2>&1 | tail -1 | cut -c1-500 | tee $OUT/r1cs.txt
