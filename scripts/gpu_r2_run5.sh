#!/bin/bash
# Round 2, GPU call 5: fused-stage NTT (parity + timings), then the default bench.
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
for n in 12 16 20 22; do timeout 600 python zero-knowledge-proofs_b200/tools/bench_quotient.py --log-n $n --steps 3 2>&1 | tail -1; done | tee $OUT/quotient.jsonl
timeout 600 python zero-knowledge-proofs_b200/tools/bench_r1cs.py 2>&1 | tail -4 | cut -c1-600 | tee $OUT/r1cs.txt
