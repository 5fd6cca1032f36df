#!/bin/bash
# tuning run: split of the bucket reduction between thread levels and tile levels
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunked or golden or multi_device" 2>&1 | tail -2
for n in ${SIZES:-16 18 21}; do for g in ${GROUPS_LOG2:-12 13 14 15}; do for t in ${TILES_LOG2:-12 13 14 15}; do
  [ $t -gt $g ] && continue
  G16_REDUCE_GROUPS_LOG2=$g G16_TILE_MAX_LOG2=$t timeout 300 python bench.py --log-n $n --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('2^$n groups=$g tile=$t', round(d['ms_per_step'],3), 'reduce', round(d['stage_ms']['reduce'],3), 'combine', round(d['stage_ms']['combine'],3))"
done; done; done
