#!/bin/bash
# parity + stage timings after a change to the reduction tree
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
for n in 16 20 21 24; do
  timeout 400 python bench.py --log-n $n --steps 5 --warmup 3 --no-cpu-baseline > $OUT/q$n.json 2> $OUT/q$n.err; echo "bench 2^$n rc=$?"; tail -2 $OUT/q$n.err
  python - $OUT/q$n.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["workload"], round(d["ms_per_step"], 3), "ms | e2e", round(d["e2e"]["ms_per_step"], 3), "|", {k: round(v, 3) for k, v in d["stage_ms"].items()}, "c", d["config"]["window_bits"], "launches", d["gpu_launches"])
except Exception as e:
    print("no result", e)
PY
done
for g in g2; do for n in 16 20; do timeout 300 python zero-knowledge-proofs_b200/tools/bench_stages.py --group $g --log-n $n 2>&1 | tail -1; done; done
timeout 900 python zero-knowledge-proofs_b200/tools/bench_prove.py --log-n 20 --steps 3 --no-cpu 2>&1 | tail -2 | cut -c1-300
