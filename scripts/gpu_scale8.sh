#!/bin/bash
# strong scaling of the headline bench at 2^24 over 8 / 4 / 2 ranks of one box (N = 1 comes from the one-GPU runs)
OUT=gpurun_out; mkdir -p $OUT
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["workload"], d["scaling"], "gpus", d["n_gpus"], round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e6, 2), "Mpts/s | e2e",
          round(d["e2e"]["ms_per_step"], 3), "ms |", {k: round(v, 3) for k, v in d["stage_ms"].items()}, "c", d["config"]["window_bits"])
except Exception as e:
    print("no result in", sys.argv[1], e)
PY
}
for g in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29600 + g)) \
      bench.py --gpus $g --steps 5 --warmup 3 > $OUT/scale_g$g.json 2> $OUT/scale_g$g.err; echo "gpus=$g rc=$?"; summ $OUT/scale_g$g.json
done
timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/scale_g1.json 2> $OUT/scale_g1.err; summ $OUT/scale_g1.json
