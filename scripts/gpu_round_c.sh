#!/bin/bash
# run 16 (one GPU): size sweep of the headline bench, config 5 at 2^24 with pinned host buffers, launch list of the default bench
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
for n in 16 18 20 22 24 26; do
  timeout 900 python bench.py --log-n $n --steps 5 --warmup 3 --no-cpu-baseline > $OUT/sweep$n.json 2> $OUT/sweep$n.err; echo "bench 2^$n rc=$?"; tail -1 $OUT/sweep$n.err | cut -c1-200; summ $OUT/sweep$n.json
done
timeout 900 python zero-knowledge-proofs_b200/tools/bench_config5.py --gpus 1 --log-n 24 --steps 3 > $OUT/config5_g1_2p24_pinned.jsonl 2> $OUT/config5_g1_2p24_pinned.err; echo "config5 rc=$?"; tail -2 $OUT/config5_g1_2p24_pinned.err; cut -c1-300 $OUT/config5_g1_2p24_pinned.jsonl
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/plain_default.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 400 --csv --log-file $OUT/launches_default_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_default.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm rc=$?"; cut -c1-400 $OUT/bench_reference.json
