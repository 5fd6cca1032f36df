#!/bin/bash
# Round 2, GPU call 9: per-lane stage timeline of the prove schedule (where do the 50 ms go?)
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python zero-knowledge-proofs_b200/tools/bench_prove.py --log-n 20 --steps 3 --no-cpu --timeline 2>&1 | grep -E "timeline|gpu_ms" | cut -c1-200 | tee $OUT/prove_timeline.txt
