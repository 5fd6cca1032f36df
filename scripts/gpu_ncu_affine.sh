#!/bin/bash
# ncu --set full of the affine accumulation kernel (one launch) at 2^LOGN
OUT=gpurun_out; mkdir -p $OUT
LOGN=${LOGN:-22}; R=${R:-2}
G16_AFFINE_ROUNDS=$R python bench.py --log-n $LOGN --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_aff_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/ncu_aff_plain.log; exit 1; }
G16_AFFINE_ROUNDS=$R timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"accumulate_affine" -c 1 -o $OUT/aff_full_2p${LOGN}_r$R -f \
   python bench.py --log-n $LOGN --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_aff.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/ncu_aff.log
ncu -i $OUT/aff_full_2p${LOGN}_r$R.ncu-rep --page raw --csv > $OUT/aff_full_2p${LOGN}_r${R}_raw.csv 2>/dev/null
ncu -i $OUT/aff_full_2p${LOGN}_r$R.ncu-rep --page details > $OUT/aff_full_2p${LOGN}_r${R}_details.txt 2>/dev/null
ls -la $OUT | tail -5
