#!/bin/bash
# affine accumulation, run 14: L2 fetch granularity hint x rounds, then the per-launch list
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
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "affine or golden or wire" > $OUT/pytest_affine.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_affine.log
for f in 0 32 64; do for r in 0 3; do
  if [ $f = 0 ]; then unset G16_L2_FETCH; else export G16_L2_FETCH=$f; fi
  G16_AFFINE_ROUNDS=$r timeout 400 python bench.py --log-n 24 --steps 4 --warmup 3 --no-cpu-baseline > $OUT/aff24_f${f}_r$r.json 2> $OUT/aff24_f${f}_r$r.err; echo "2^24 fetch=$f rounds=$r rc=$?"; tail -2 $OUT/aff24_f${f}_r$r.err; summ $OUT/aff24_f${f}_r$r.json
done; done
export G16_L2_FETCH=32
LOGN=24 R=3 bash scripts/gpu_affine_launches.sh 2>&1 | cut -c1-330
