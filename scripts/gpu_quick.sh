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
# per-level timing of the reduction tree at 2^24 (only after the plain run above exited 0)
python bench.py --log-n 24 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/plain24.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"ReduceLevel|tile_reduce|WindowCombine|chunk_merge|ItemScatter|ItemCount|DigitDecompose|Scatter" -c 60 --csv --log-file $OUT/launches24_tail.csv \
    python bench.py --log-n 24 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_tail.log 2>&1; echo "ncu tail rc=$?"
python - <<'PY'
import csv, re
try:
    lines = [l for l in open("gpurun_out/launches24_tail.csv") if not l.startswith("==")]
    for row in csv.DictReader(lines):
        m = re.search(r"(ReduceLevel|tile_reduce_kernel|WindowCombine|chunk_merge_kernel|ItemScatter|ItemCount|DigitDecompose|ScatterRanked|ScatterByWindow)", row["Kernel Name"])
        print(row["ID"], m.group(1) if m else row["Kernel Name"][:30], row["Grid Size"], row["Block Size"], row["Metric Value"], row["Metric Unit"])
except Exception as e:
    print("no launch list", e)
PY
