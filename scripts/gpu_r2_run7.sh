#!/bin/bash
# Round 2, GPU call 7: prefetch A/B for the accumulate kernels, then the ncu evidence of the shipped build: launch list of
# the default bench, --set full captures of the hot kernel (2^24), the fused NTT pass and the fixed-base kernel.
OUT=gpurun_out; mkdir -p $OUT
LAB=zero-knowledge-proofs_b200/lib/lab
T=zero-knowledge-proofs_b200/tools/bench_stages.py
echo "== prefetch A/B"
for v in std g1_prefetch; do L=""; [ $v != std ] && L="--lib $LAB/$v.so"; echo -n "$v g1 2^24: "; timeout 300 python $T --group g1 --log-n 24 --steps 5 $L 2>&1 | tail -1 | cut -c1-420; done | tee $OUT/lab_prefetch.txt
for v in std g1_prefetch; do L=""; [ $v != std ] && L="--lib $LAB/$v.so"; echo -n "$v g1 2^21 c20: "; timeout 300 python $T --group g1 --log-n 21 --precompute-bits 20 --steps 10 $L 2>&1 | tail -1 | cut -c1-420; done | tee -a $OUT/lab_prefetch.txt
for v in std g2_prefetch; do L=""; [ $v != std ] && L="--lib $LAB/$v.so"; echo -n "$v g2 2^20: "; timeout 300 python $T --group g2 --log-n 20 --steps 5 $L 2>&1 | tail -1 | cut -c1-420; done | tee -a $OUT/lab_prefetch.txt
echo "== launch list of the default bench (after the same command exited 0 without ncu)"
B="--steps 2 --warmup 1 --no-cpu-baseline --no-prove --no-oneshot"
python bench.py $B > $OUT/bench_plain.json 2> $OUT/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 400 --csv --log-file $OUT/r02_ncu_launches_default_bench_2p24.csv \
    python bench.py $B > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?"
echo "== ncu --set full: BucketAccumulate<Fq> at 2^24"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:BucketAccumulate -s 3 -c 1 \
    -o $OUT/r02_acc_g1_2p24 python bench.py $B > $OUT/ncu_acc.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_acc.log
echo "== ncu --set full: ntt_fused_kernel at 2^20 (one middle pass, batch 3)"
timeout 600 python zero-knowledge-proofs_b200/tools/bench_quotient.py --log-n 20 --steps 1 > $OUT/q_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:ntt_fused -s 4 -c 1 \
    -o $OUT/r02_ntt_fused_2p20 python zero-knowledge-proofs_b200/tools/bench_quotient.py --log-n 20 --steps 1 > $OUT/ncu_ntt.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_ntt.log
echo "== ncu --set full: FbMul<Fq> at 2^22"
timeout 600 python zero-knowledge-proofs_b200/tools/bench_setup.py --log-n 20 --steps 1 > $OUT/fb_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:FbMul -s 3 -c 1 \
    -o $OUT/r02_fbmul_g1_2p20 python zero-knowledge-proofs_b200/tools/bench_setup.py --log-n 20 --steps 1 > $OUT/ncu_fb.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_fb.log
ls -la $OUT/*.ncu-rep
