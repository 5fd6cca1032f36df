#!/bin/bash
# Round 2, GPU call 4: safegcd inversion on the device (parity suite), A/B of the item floor and of the reduction split,
# per-kernel launch list of a G2 MSM, fixed-base rates with the new inversion.
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== policy A/B (G1)"
for v in std item_floor8 item_floor32 red_14_15 red_14_14 red_16_15 red_13_13; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  for cfg in "16 0" "18 0" "20 0" "21 20"; do
    set -- $cfg
    echo -n "$v 2^$1 c=$2: "; timeout 200 python $T --group g1 --log-n $1 --precompute-bits $2 --steps 10 $L 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms'],3), d['plan'], d['stage_ms'])"
  done
done | tee $OUT/lab_policy_g1.txt
echo "== fixed base"
timeout 600 python zero-knowledge-proofs_b200/tools/bench_setup.py --log-n 22 --steps 3 2>&1 | tee $OUT/setup_fixed_base_2p22.jsonl | cut -c1-330
echo "== G2 launch list"
python $T --group g2 --log-n 20 --steps 1 > $OUT/g2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 200 --csv --log-file $OUT/launches_g2_2p20.csv \
    python $T --group g2 --log-n 20 --steps 1 > $OUT/ncu_g2_launches.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, re
try:
    lines = [l for l in open("gpurun_out/launches_g2_2p20.csv") if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    for row in rows[-20:]:
        m = re.search(r"(ReduceLevel|tile_reduce_kernel|WindowCombine|chunk_merge_kernel|ChunkMergeSerial|item_scatter|item_count|DigitDecompose|ScatterRanked|ScatterFinal|scatter_partition|BucketAccumulate|scan_\w+|PartialCombine|partial_combine\w+|Precompute\w+|Fb\w+)", row["Kernel Name"])
        print(row["ID"], m.group(1) if m else row["Kernel Name"][:30], row["Grid Size"], row["Block Size"], row["Metric Value"], row["Metric Unit"])
except Exception as e:
    print("no launch list", e)
PY
