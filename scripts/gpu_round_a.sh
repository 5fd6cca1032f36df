#!/bin/bash
# full GPU parity suite, default bench, config 5 on one GPU (2^20 with the exponent check, 2^24 with it too)
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -2 $OUT/bench_default.err; cut -c1-600 $OUT/bench_default.json
timeout 600 python zero-knowledge-proofs_b200/tools/bench_config5.py --gpus 1 --log-n 20 --check-exponent > $OUT/config5_g1_2p20.jsonl 2> $OUT/config5_g1_2p20.err; echo "config5 2^20 rc=$?"; tail -2 $OUT/config5_g1_2p20.err; cut -c1-420 $OUT/config5_g1_2p20.jsonl
timeout 1500 python zero-knowledge-proofs_b200/tools/bench_config5.py --gpus 1 --log-n 24 --steps 2 --check-exponent > $OUT/config5_g1_2p24.jsonl 2> $OUT/config5_g1_2p24.err; echo "config5 2^24 rc=$?"; tail -2 $OUT/config5_g1_2p24.err; cut -c1-420 $OUT/config5_g1_2p24.jsonl
