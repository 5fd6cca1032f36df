#!/bin/bash
# per-launch durations of the affine accumulation kernels at 2^LOGN (ncu launch list; shares, not absolutes)
OUT=gpurun_out; mkdir -p $OUT
LOGN=${LOGN:-24}; R=${R:-3}
G16_AFFINE_ROUNDS=$R python bench.py --log-n $LOGN --steps 1 --warmup 1 --no-cpu-baseline > $OUT/aff_launch_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/aff_launch_plain.log; exit 1; }
G16_AFFINE_ROUNDS=$R timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --kernel-name-base demangled -k regex:"Affine|BatchInverse|BucketAccumulate" -c 24 --csv --log-file $OUT/aff_launches_2p${LOGN}_r$R.csv \
   python bench.py --log-n $LOGN --steps 1 --warmup 1 --no-cpu-baseline > $OUT/aff_launch_ncu.log 2>&1; echo "ncu rc=$?"
python - $OUT/aff_launches_2p${LOGN}_r$R.csv <<'PY'
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = {}
for row in csv.DictReader(lines):
    m = re.search(r"(AffinePhase1|AffinePhase2|AffineTail|BatchInverse|BucketAccumulate)", row["Kernel Name"])
    rows.setdefault((row["ID"], m.group(1) if m else "?"), {})[row["Metric Name"]] = row["Metric Value"] + " " + row["Metric Unit"]
for (i, k), v in rows.items():
    print(i, k, v)
PY
