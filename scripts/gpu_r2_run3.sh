#!/bin/bash
# Round 2, GPU call 3: parity suite on the new fixed-base (16-bit signed windows) and quad G2 reduction, fixed-base rates,
# A/B of the remaining launch shapes, G2 window sweep, default bench.
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== fixed base"
timeout 600 python zero-knowledge-proofs_b200/tools/bench_setup.py --log-n 22 --steps 3 2>&1 | tee $OUT/setup_fixed_base_2p22.jsonl | cut -c1-400
echo "== A: G1 2^24"
for v in std g1_b32_mb12 g1_b64_mb7 g1_b96_mb4; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v: "; timeout 300 python $T --group g1 --log-n 24 --steps 5 $L 2>&1 | tail -1
done | tee $OUT/lab_g1_2p24_b.txt
echo "== B: G2 2^20"
for v in std g2_b32; do
  L=""; [ $v != std ] && L="--lib $LAB/$v.so"
  echo -n "$v: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 $L 2>&1 | tail -1
done | tee $OUT/lab_g2_2p20_b.txt
for c in 17 18 19; do echo -n "g2 c=$c: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 --precompute-bits $c 2>&1 | tail -1; done | tee -a $OUT/lab_g2_2p20_b.txt
echo -n "g2 u64 scalars: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 --bits 64 2>&1 | tail -1 | tee -a $OUT/lab_g2_2p20_b.txt
echo "== default bench"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -5 $OUT/bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value ms", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], d["stage_ms"], "prove", {k: d["prove"].get(k) for k in ("full_width", "ref_faithful_u64", "error")}, d["roofline"].get("executed"))
PY
