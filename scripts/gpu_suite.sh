#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench sweep, prove benchmark, ncu capture of the hot kernel.
# Usage (from the repo root, on the GPU box):  bash scripts/gpu_suite.sh [tag]
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["workload"], "gpus", d["n_gpus"], round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e6, 2), "Mpts/s | e2e",
          round(d["e2e"]["ms_per_step"], 3), "ms |", {k: round(v, 3) for k, v in d["stage_ms"].items()}, "| frac",
          d["roofline"].get("whole_step_frac"), "c", d["config"]["window_bits"], "launches", d["gpu_launches"], d.get("clocks"))
except Exception as e:
    print("no result in", sys.argv[1], e)
PY
}
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
for n in 16 20 24; do
  timeout 400 python bench.py --log-n $n --steps 5 --warmup 3 > $OUT/bench$n.json 2> $OUT/bench$n.err; echo "bench 2^$n rc=$?"; tail -2 $OUT/bench$n.err; summ $OUT/bench$n.json
done
G16_ATOMIC_SCATTER=1 timeout 400 python bench.py --log-n 24 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench24_atomic_scatter.json 2> $OUT/bench24_as.err; summ $OUT/bench24_atomic_scatter.json
timeout 400 python bench.py --log-n 24 --steps 5 --warmup 3 --no-cpu-baseline --no-precompute > $OUT/bench24_plain.json 2> $OUT/bench24_plain.err; summ $OUT/bench24_plain.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference rc=$?"; cut -c1-400 $OUT/bench_reference.json
timeout 900 python zero-knowledge-proofs_b200/tools/bench_prove.py --log-n 20 --steps 3 > $OUT/prove20.json 2> $OUT/prove20.err; echo "prove rc=$?"; cat $OUT/prove20.json; tail -3 $OUT/prove20.err
# ncu: only after the same command has exited 0 without ncu
python bench.py --log-n 20 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/plain20.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:BucketAccumulate -s 1 -c 1 \
    -o $OUT/prof_acc20_$TAG python bench.py --log-n 20 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/ncu_full.log
ls -la $OUT | head -40
