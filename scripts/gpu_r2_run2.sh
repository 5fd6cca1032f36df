#!/bin/bash
# Round 2, GPU call 2: A/B of launch shapes for the accumulate kernels, per-kernel launch list at the per-rank size of
# the 8-GPU run, ncu --set full of the G2 accumulate kernel.
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
echo "== A: G1 2^24"
for v in std g1_mb4 g1_b64_mb6 g1_b64_mb8; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v: "; timeout 300 python $T --group g1 --log-n 24 --steps 5 $L 2>&1 | tail -1
done | tee $OUT/lab_g1_2p24.txt
echo "== B: G2 2^20"
for v in std g2_mb3 g2_b64_mb4 g2_b64_mb6; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 $L 2>&1 | tail -1
done | tee $OUT/lab_g2_2p20.txt
echo "== E: small sizes"
for n in 16 18 20 21; do echo -n "g1 2^$n: "; timeout 200 python $T --group g1 --log-n $n --steps 10 2>&1 | tail -1; done | tee $OUT/stages_small.txt
echo "== C: per-rank configuration of the 8-GPU run (2^21 pairs, c = 20)"
timeout 300 python bench.py --log-n 21 --precompute-bits 20 --steps 5 --warmup 3 --no-cpu-baseline --no-prove --no-oneshot > $OUT/bench_2p21_c20.json 2> $OUT/bench_2p21_c20.err; echo "rc=$?"; cut -c1-1200 $OUT/bench_2p21_c20.json
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 120 --csv --log-file $OUT/launches_2p21_c20.csv \
    python bench.py --log-n 21 --precompute-bits 20 --steps 2 --warmup 1 --no-cpu-baseline --no-prove --no-oneshot > $OUT/ncu_2p21.log 2>&1; echo "ncu launches rc=$?"
python - <<'PY'
import csv, re
try:
    lines = [l for l in open("gpurun_out/launches_2p21_c20.csv") if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    for row in rows[-22:]:
        m = re.search(r"(ReduceLevel|tile_reduce_kernel|WindowCombine|chunk_merge_kernel|ChunkMergeSerial|item_scatter|item_count|DigitDecompose|ScatterRanked|ScatterFinal|scatter_partition|BucketAccumulate|scan_\w+|PartialCombine|partial_combine\w+)", row["Kernel Name"])
        print(row["ID"], m.group(1) if m else row["Kernel Name"][:30], row["Grid Size"], row["Block Size"], row["Metric Value"], row["Metric Unit"])
except Exception as e:
    print("no launch list", e)
PY
echo "== D: ncu --set full, G2 accumulate at 2^20"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:BucketAccumulate -s 2 -c 1 \
    -o $OUT/r02_g2_acc_2p20 python $T --group g2 --log-n 20 --steps 1 > $OUT/ncu_g2_full.log 2>&1; echo "ncu full rc=$?"; tail -2 $OUT/ncu_g2_full.log
ls -la $OUT | tail -12
