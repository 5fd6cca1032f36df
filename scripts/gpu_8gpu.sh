#!/bin/bash
# one 8-GPU box: BASELINE config 5 (setup + prove at 2^24 on a multi-device context) and config 4 (strong scaling of
# the G1 MSM at 2^24 over 8 / 4 / 2 ranks, NCCL all-gather of the partial sums).  Usage: gpurun --gpus 8 -- bash scripts/gpu_8gpu.sh
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
timeout 900 python zero-knowledge-proofs_b200/tools/bench_config5.py --gpus 8 --log-n 24 --steps 3 > $OUT/config5_g8_2p24.jsonl 2> $OUT/config5_g8_2p24.err; echo "config5 8 GPUs rc=$?"; tail -2 $OUT/config5_g8_2p24.err; cut -c1-330 $OUT/config5_g8_2p24.jsonl
for g in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29600 + g)) \
      bench.py --gpus $g --steps 5 --warmup 3 > $OUT/scale_g$g.json 2> $OUT/scale_g$g.err; echo "gpus=$g rc=$?"; tail -2 $OUT/scale_g$g.err | cut -c1-300; summ $OUT/scale_g$g.json
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_device" > $OUT/pytest_multidev.log 2>&1; tail -2 $OUT/pytest_multidev.log
