#!/bin/bash
# Round 2, GPU call 24: ncu evidence of the shipped build after the multi-product multiplications: launch list of the
# default bench, --set full captures of BucketAccumulate<Fq> (2^24, full-size launch) and BucketAccumulate<Fq2> (2^20).
OUT=gpurun_out; mkdir -p $OUT
B="--steps 2 --warmup 1 --no-cpu-baseline --no-prove --no-oneshot"
python bench.py $B > $OUT/bench_plain.json 2> $OUT/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 400 --csv --log-file $OUT/r02_run24_ncu_launches_default_bench_2p24.csv \
    python bench.py $B > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?"
B1="--steps 1 --warmup 1 --no-cpu-baseline --no-prove --no-oneshot"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:BucketAccumulate -s 1 -c 1 \
    -o $OUT/r02_run24_acc_g1_2p24 python bench.py $B1 > $OUT/ncu_acc.log 2>&1; echo "g1 rc=$?"; tail -2 $OUT/ncu_acc.log
T=zero-knowledge-proofs_b200/tools/bench_stages.py
timeout 300 python $T --group g2 --log-n 20 --steps 2 > $OUT/g2_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:BucketAccumulate -s 1 -c 1 \
    -o $OUT/r02_run24_acc_g2_2p20 python $T --group g2 --log-n 20 --steps 2 > $OUT/ncu_acc_g2.log 2>&1; echo "g2 rc=$?"; tail -2 $OUT/ncu_acc_g2.log
ls -la $OUT/*.ncu-rep
