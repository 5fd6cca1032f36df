#!/bin/bash
# Round 2, GPU call 27: tables for a promised scalar width (g16_pk_precompute_bits) -- prove parity, window choices, prove timing
OUT=gpurun_out; mkdir -p $OUT
T=zero-knowledge-proofs_b200/tools/bench_stages.py
timeout 1500 python -m pytest tests/test_gpu_prove.py -m gpu -x -q > $OUT/pytest_prove_run27.log 2>&1; echo "pytest prove rc=$?"; tail -3 $OUT/pytest_prove_run27.log
for c in 18 19 20; do echo -n "g2 2^20 full c=$c: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 --precompute-bits $c 2>&1 | tail -1 | cut -c1-420; done | tee $OUT/lab_window_choice.txt
for c in 16 17 18 20; do echo -n "g2 2^20 u64 c=$c: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 --bits 64 --precompute-bits $c 2>&1 | tail -1 | cut -c1-420; done | tee -a $OUT/lab_window_choice.txt
for c in 16 17 18 20; do echo -n "g1 2^20 u64 c=$c: "; timeout 300 python $T --group g1 --log-n 20 --steps 5 --bits 64 --precompute-bits $c 2>&1 | tail -1 | cut -c1-420; done | tee -a $OUT/lab_window_choice.txt
timeout 600 python zero-knowledge-proofs_b200/tools/bench_prove.py --log-n 20 --steps 5 --no-cpu 2>&1 | grep gpu_ms | cut -c1-300 | tee $OUT/prove_run27.txt
timeout 600 python zero-knowledge-proofs_b200/tools/bench_prove.py --log-n 20 --steps 5 --no-cpu --no-precompute-bits 2>&1 | grep gpu_ms | cut -c1-300 | tee -a $OUT/prove_run27.txt
