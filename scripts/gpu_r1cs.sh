#!/bin/bash
# sparse R1CS path on the GPU: parity tests, then setup + prove for synthetic circuits
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
timeout 600 python zero-knowledge-proofs_b200/tools/bench_r1cs.py --log-m 12 --check > $OUT/r1cs12.json 2> $OUT/r1cs12.err; echo "r1cs 2^12 rc=$?"; tail -3 $OUT/r1cs12.err; cat $OUT/r1cs12.json
timeout 900 python zero-knowledge-proofs_b200/tools/bench_r1cs.py --log-m 16 --check > $OUT/r1cs16.json 2> $OUT/r1cs16.err; echo "r1cs 2^16 rc=$?"; tail -3 $OUT/r1cs16.err; cat $OUT/r1cs16.json
timeout 1500 python zero-knowledge-proofs_b200/tools/bench_r1cs.py --log-m 20 --check > $OUT/r1cs20.json 2> $OUT/r1cs20.err; echo "r1cs 2^20 rc=$?"; tail -3 $OUT/r1cs20.err; cat $OUT/r1cs20.json
