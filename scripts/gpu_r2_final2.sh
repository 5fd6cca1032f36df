#!/bin/bash
# Round 2, last GPU call: GPU suite + smoke on the final build, launch list of one G2 MSM, prove timing (tool and bench record)
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/final2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/final2_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/final2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/final2_smoke.log
T=zero-knowledge-proofs_b200/tools/bench_stages.py
timeout 300 python $T --group g2 --log-n 20 --steps 2 > $OUT/final2_g2_plain.log 2>&1 && tail -1 $OUT/final2_g2_plain.log | cut -c1-420 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 80 --csv --log-file $OUT/r02_run33_ncu_launches_g2_2p20_final.csv \
    python $T --group g2 --log-n 20 --steps 2 > $OUT/final2_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 python zero-knowledge-proofs_b200/tools/bench_prove.py --log-n 20 --steps 5 --no-cpu 2>&1 | grep gpu_ms | cut -c1-200 | tee $OUT/final2_prove.txt
