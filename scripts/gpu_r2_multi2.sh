#!/bin/bash
# Round 2, second 8-GPU call (final build): the driver's N = 8 command (full bench line incl. BASELINE config 5), the
# throughput matrix at N = 8 and N = 4 / 2 / 1 side by side on disjoint GPUs of the same box.
OUT=gpurun_out; mkdir -p $OUT
run() {  # run <N> <port> <visible devices> <args...>
  N=$1; PORT=$2; DEV=$3; shift 3
  if [ $N -eq 1 ]; then CUDA_VISIBLE_DEVICES=$DEV python bench.py --gpus 1 "$@"
  else CUDA_VISIBLE_DEVICES=$DEV python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N "$@"; fi
}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== N=8 full bench (driver command)"
timeout 1500 bash -c "$(declare -f run); run 8 29511 0,1,2,3,4,5,6,7 --steps 10 --warmup 3" > $OUT/final_bench_n8.json 2> $OUT/final_bench_n8.err; echo "rc=$?"; tail -3 $OUT/final_bench_n8.err | cut -c1-300; tail -1 $OUT/final_bench_n8.json | cut -c1-1200
SW="--steps 5 --warmup 3 --no-cpu-baseline --no-prove --no-oneshot"
echo "== sweep N=8"
for L in 16 18 20 22 24 26; do
  timeout 600 bash -c "$(declare -f run); run 8 $((29520 + L)) 0,1,2,3,4,5,6,7 --log-n $L $SW" 2> $OUT/final_sweep_n8_$L.err | tail -1 > $OUT/final_sweep_n8_2p$L.json; echo "N=8 2^$L rc=$?"
done
echo "== sweep N=4 | N=2 | N=1 side by side"
for L in 16 18 20 22 24 26; do
  (timeout 600 bash -c "$(declare -f run); run 4 $((29600 + L)) 0,1,2,3 --log-n $L $SW" 2> $OUT/final_sweep_n4_$L.err | tail -1 > $OUT/final_sweep_n4_2p$L.json) &
  (timeout 600 bash -c "$(declare -f run); run 2 $((29700 + L)) 4,5 --log-n $L $SW" 2> $OUT/final_sweep_n2_$L.err | tail -1 > $OUT/final_sweep_n2_2p$L.json) &
  (timeout 600 bash -c "$(declare -f run); run 1 0 6 --log-n $L $SW" 2> $OUT/final_sweep_n1_$L.err | tail -1 > $OUT/final_sweep_n1_2p$L.json) &
  wait
  echo "2^$L done"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_sweep_n*_2p*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "N", d["n_gpus"], d["config"]["workload"], round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e6, 1), "M/s e2e", round(d["e2e"]["ms_per_step"], 3), "c", d["config"]["window_bits"], "frac", round(d["roofline"].get("whole_step_frac", 0), 3))
    except Exception as e:
        print(f, "no result", e)
PY
