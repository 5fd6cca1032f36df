#!/bin/bash
# quick GPU check: parity tests, bench 2^16/2^20/2^24, G1/G2 stage timings, prove
OUT=gpurun_out; mkdir -p $OUT
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["workload"], "gpus", d["n_gpus"], round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e6, 2), "Mpts/s | e2e",
          round(d["e2e"]["ms_per_step"], 3), "ms |", {k: round(v, 3) for k, v in d["stage_ms"].items()}, "| frac",
          d["roofline"].get("whole_step_frac"), "c", d["config"]["window_bits"], "launches", d["gpu_launches"])
except Exception as e:
    print("no result in", sys.argv[1], e)
PY
}
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
for n in 16 20 24; do
  timeout 400 python bench.py --log-n $n --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench$n.json 2> $OUT/bench$n.err; echo "bench 2^$n rc=$?"; tail -2 $OUT/bench$n.err; summ $OUT/bench$n.json
done
timeout 400 python bench.py --log-n 24 --steps 5 --warmup 3 --no-cpu-baseline --no-precompute > $OUT/bench24_plain.json 2> $OUT/bench24_plain.err; summ $OUT/bench24_plain.json
for g in g1 g2; do for n in 16 20; do timeout 300 python zero-knowledge-proofs_b200/tools/bench_stages.py --group $g --log-n $n 2>&1 | tail -1; done; done
timeout 300 python zero-knowledge-proofs_b200/tools/bench_stages.py --group g2 --log-n 20 --no-precompute 2>&1 | tail -1
timeout 300 python zero-knowledge-proofs_b200/tools/bench_stages.py --group g2 --log-n 20 --bits 64 2>&1 | tail -1
timeout 900 python zero-knowledge-proofs_b200/tools/bench_prove.py --log-n 20 --steps 3 --no-cpu > $OUT/prove20.json 2> $OUT/prove20.err; echo "prove rc=$?"; cat $OUT/prove20.json; tail -3 $OUT/prove20.err
