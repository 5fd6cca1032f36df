#!/bin/bash
# Round 2, last GPU call: the driver's sequence on the final build (GPU suite, smoke, default bench), then compute-sanitizer
# memcheck over a small MSM / prove selection.
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/final_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/final_smoke.log
timeout 1200 python bench.py > $OUT/final_bench_default.json 2> $OUT/final_bench_default.err; echo "bench rc=$?"; cut -c1-700 $OUT/final_bench_default.json
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or exceptional or window_sweep or adversarial" > $OUT/final_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $OUT/final_memcheck.log
