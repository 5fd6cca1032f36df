#!/bin/bash
# Round 2, first GPU call: parity suite, smoke, default bench (with prove sub-record), reference arm.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
nproc; free -g | sed -n 2p
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -5 $OUT/bench_default.err; cat $OUT/bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference rc=$?"; tail -3 $OUT/bench_reference.err; cat $OUT/bench_reference.json
