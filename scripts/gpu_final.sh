#!/bin/bash
# final state of the round on one GPU: full parity suite, smoke, default bench, size sweep, launch list, prove
OUT=gpurun_out; mkdir -p $OUT
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["workload"], round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e6, 2), "Mpts/s | e2e",
          round(d["e2e"]["ms_per_step"], 3), "ms |", {k: round(v, 3) for k, v in d["stage_ms"].items()}, "| frac",
          round(d["roofline"].get("whole_step_frac") or 0, 3), "c", d["config"]["window_bits"])
except Exception as e:
    print("no result in", sys.argv[1], e)
PY
}
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; summ $OUT/bench_default.json
for n in 16 18 20 21 22 26; do
  timeout 900 python bench.py --log-n $n --steps 5 --warmup 3 --no-cpu-baseline > $OUT/sweep$n.json 2> $OUT/sweep$n.err; summ $OUT/sweep$n.json
done
timeout 600 python zero-knowledge-proofs_b200/tools/bench_prove.py --log-n 20 --steps 3 --no-cpu > $OUT/prove20.json 2> $OUT/prove20.err; echo "prove rc=$?"; cut -c1-200 $OUT/prove20.json
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/plain_default.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 400 --csv --log-file $OUT/launches_default_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_default.log 2>&1; echo "ncu rc=$?"
