#!/bin/bash
# Round 2, GPU call 35: ncu --set full of the two block-cooperative reduction levels of one G2 MSM (latency tail)
OUT=gpurun_out; mkdir -p $OUT
T=zero-knowledge-proofs_b200/tools/bench_stages.py
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:tile_reduce -c 2 \
    -o $OUT/r02_run35_tile_reduce_g2_2p20 python $T --group g2 --log-n 20 --steps 1 > $OUT/ncu_tile.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_tile.log
ls -la $OUT/r02_run35_tile_reduce_g2_2p20.ncu-rep
