#!/bin/bash
# Round 2, GPU call 18: G2 bucket reduction with the Fq2 multiplication as the call boundary (instead of the Fq one)
OUT=gpurun_out; mkdir -p $OUT
T=zero-knowledge-proofs_b200/tools/bench_stages.py
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "g2 or adversarial or golden" > $OUT/pytest_g2.log 2>&1; echo "pytest g2 rc=$?"; tail -2 $OUT/pytest_g2.log
for i in 1 2; do echo -n "g2 2^20: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 2>&1 | tail -1 | cut -c1-420; done | tee $OUT/g2_stages.txt
echo -n "g2 2^20 u64: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 --bits 64 2>&1 | tail -1 | cut -c1-420 | tee -a $OUT/g2_stages.txt
timeout 600 python zero-knowledge-proofs_b200/tools/bench_prove.py --log-n 20 --steps 5 --no-cpu 2>&1 | grep gpu_ms | cut -c1-200 | tee $OUT/prove.txt
