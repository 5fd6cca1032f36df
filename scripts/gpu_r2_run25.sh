#!/bin/bash
# Round 2, GPU call 25: stages 1-3 of range k+1 under the accumulation of range k (front stream) -- suite, A/B, sweep
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_run25.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_run25.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-prove --no-oneshot"
for v in std no_front_overlap; do
  for ln in 24 22; do
    L=""; [ $v != std ] && L="--lib $LAB/$v.so"
    echo -n "$v 2^$ln: "; timeout 600 python bench.py $B --log-n $ln $L 2>$OUT/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d.get(k) for k in ('ms_per_step','value','e2e','stage_ms')}))"
  done
done | tee $OUT/lab_front_overlap.txt
for ln in 20 21 22; do echo -n "front_pipe_min_20 2^$ln: "; timeout 600 python bench.py $B --log-n $ln --lib $LAB/front_pipe_min_20.so 2>$OUT/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d.get(k) for k in ('ms_per_step','value','e2e','stage_ms')}))"; done | tee -a $OUT/lab_front_overlap.txt
for ln in 20 21 26; do echo -n "std 2^$ln: "; timeout 600 python bench.py $B --log-n $ln 2>$OUT/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d.get(k) for k in ('ms_per_step','value','e2e','stage_ms')}))"; done | tee -a $OUT/lab_front_overlap.txt
