#!/bin/bash
# two-pass partitioned scatter: exact large-size parity, then one-pass vs two-pass at 2^22 / 2^24 / 2^26 (+ plain bases)
OUT=gpurun_out; mkdir -p $OUT
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["workload"], round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e6, 2), "Mpts/s | e2e",
          round(d["e2e"]["ms_per_step"], 3), "ms |", {k: round(v, 3) for k, v in d["stage_ms"].items()}, "c", d["config"]["window_bits"], "launches", d["gpu_launches"])
except Exception as e:
    print("no result in", sys.argv[1], e)
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "discrete_log or golden or adversarial" > $OUT/pytest_scatter.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_scatter.log
for n in 22 24 26; do for mode in two one; do
  if [ $mode = one ]; then export G16_SCATTER_ONE_PASS=1; else unset G16_SCATTER_ONE_PASS; fi
  timeout 600 python bench.py --log-n $n --steps 4 --warmup 3 --no-cpu-baseline > $OUT/scatter_${mode}_$n.json 2> $OUT/scatter_${mode}_$n.err; echo "2^$n $mode-pass rc=$?"; tail -1 $OUT/scatter_${mode}_$n.err | cut -c1-200; summ $OUT/scatter_${mode}_$n.json
done; done
unset G16_SCATTER_ONE_PASS
timeout 600 python bench.py --log-n 24 --steps 4 --warmup 3 --no-cpu-baseline --no-precompute > $OUT/scatter_two_24_plain.json 2> $OUT/scatter_two_24_plain.err; summ $OUT/scatter_two_24_plain.json
