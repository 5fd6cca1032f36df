#!/bin/bash
# size sweep of the headline benchmark on one GPU (north_star: 2^16 ... 2^26) + parity tests
OUT=gpurun_out; mkdir -p $OUT
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["workload"], "gpus", d["n_gpus"], round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e6, 2), "Mpts/s | e2e",
          round(d["e2e"]["ms_per_step"], 3), "ms |", {k: round(v, 3) for k, v in d["stage_ms"].items()}, "| frac",
          d["roofline"].get("whole_step_frac"), "c", d["config"]["window_bits"], "|", d["config"]["bases"][:60])
except Exception as e:
    print("no result in", sys.argv[1], e)
PY
}
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
for n in 16 18 20 22 24 26; do
  timeout 900 python bench.py --log-n $n --steps 5 --warmup 3 --no-cpu-baseline > $OUT/sweep$n.json 2> $OUT/sweep$n.err; echo "bench 2^$n rc=$?"; tail -2 $OUT/sweep$n.err | cut -c1-300; summ $OUT/sweep$n.json
done
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "default rc=$?"; summ $OUT/bench_default.json
