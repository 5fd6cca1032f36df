#!/bin/bash
# 1 -> 8 GPU scaling of the headline benchmark (strong scaling at 2^24, BASELINE config 4) on one box.
# Usage: gpurun --gpus 8 -- bash scripts/gpu_scale.sh
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
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
timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 > $OUT/scale_g1.json 2> $OUT/scale_g1.err; summ $OUT/scale_g1.json
for g in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29600 + g)) \
      bench.py --gpus $g --steps 5 --warmup 3 > $OUT/scale_g$g.json 2> $OUT/scale_g$g.err; echo "gpus=$g rc=$?"; tail -2 $OUT/scale_g$g.err | cut -c1-300; summ $OUT/scale_g$g.json
done
# weak scaling (2^24 per GPU) at 8 GPUs, and a multi-device single-process context check
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 \
    bench.py --gpus 8 --steps 3 --warmup 3 --scaling weak > $OUT/scale_weak_g8.json 2> $OUT/scale_weak_g8.err; summ $OUT/scale_weak_g8.json
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_device" > $OUT/pytest_multidev.log 2>&1; tail -2 $OUT/pytest_multidev.log
