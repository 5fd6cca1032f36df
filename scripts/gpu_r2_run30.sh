#!/bin/bash
# Round 2, GPU call 30: work-item limit sweep at small and medium sizes (g16_ctx_set_item_max)
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python zero-knowledge-proofs_b200/tools/sweep_item_max.py --group g1 --log-n 16 17 18 19 20 21 2>&1 | tail -8 | tee $OUT/sweep_item_max_g1.jsonl
timeout 300 python zero-knowledge-proofs_b200/tools/sweep_item_max.py --group g2 --log-n 16 18 20 --item-max 0 8 11 16 22 32 64 --steps 5 2>&1 | tail -4 | tee $OUT/sweep_item_max_g2.jsonl
