#!/usr/bin/env python
"""Headline benchmark: G1 MSM points/s at 2^24 on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n 24] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...        (one rank per GPU, NCCL)

A "step" is one full G1 MSM (digits -> sort -> bucket accumulate -> reduce -> combine -> affine) over
N = 2^log_n synthetic (scalar, point) pairs: scalars uniform in [0, r) from SplitMix64(0x5eed0000+log_n),
points k_i*G with k_i from SplitMix64(0xba5e0000+log_n) built on the GPU by the fixed-base kernel
(SURVEY.md 8d).  With N GPUs the pairs are partitioned by index range (strong scaling, BASELINE
config 4); each rank reduces its range to one projective partial sum, the partials are exchanged with
one NCCL all-gather (192 bytes per rank) and folded on every rank.

  value : pairs/s, scalars already resident in HBM when the timed region starts
  e2e   : pairs/s through the C ABI with scalars in pinned HOST memory: H2D of the scalars and D2H of
          the affine result inside the timed region, every step
  roofline : the dominant kernel (bucket accumulation) against the measured IMAD peak
  cpu_baseline : the C port of the reference's CPU path (ark msm_bigint_wnaf) on a bounded sample
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "zero-knowledge-proofs_b200")
for p in (PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

# algorithmic work per (scalar, point) pair fixed by SURVEY.md 8(d): 16 windows x (8M + 2S) = 160 Fq
# multiplications, 600 32-bit IMAD issues each
FQ_MUL_PER_PAIR = 160
IMAD_PER_FQ_MUL = 600


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--no-precompute", action="store_true",
                    help="skip the one-time table of multiples 2^(c w) P for the resident bases")
    ap.add_argument("--precompute-bits", type=int, default=0)
    ap.add_argument("--cpu-sample-log-n", type=int, default=None,
                    help="log2 of the bounded CPU sample (default 17 for cpu_baseline, 19 for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2] or [r for (_, r) in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def imad_peak():
    """Run the integer-pipe microbenchmark (lib/imad_peak) once; returns its JSON."""
    exe = os.path.join(PKG, "lib", "imad_peak")
    try:
        out = subprocess.run([exe, "1.0"], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}


def cpu_baseline(oracle, pts, inf, sc, threads):
    t = time.perf_counter()
    oracle.g1_msm(pts, inf, sc, threads=threads)
    return time.perf_counter() - t


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The reference
    is Rust (no toolchain here), so this is the C port in oracle/ (cpu_baseline.kind = "port"), all host
    threads (windows in parallel = ark's `parallel` feature), each step one MSM over a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import bls12_381 as bls
    import cpu_oracle as oracle
    oracle.build()
    log_s = min(args.cpu_sample_log_n or 19, args.log_n)
    n = 1 << log_s
    threads = oracle.max_threads()
    g1 = np.array(bls.g1_to_mont(bls.G1_GEN)[0], dtype=np.uint64)
    k = oracle.gen_scalars(0xba5e0000 + args.log_n, n, 64)   # small multiples keep input generation cheap
    pts, inf = oracle.g1_fixed_base_mul(g1, k, threads=threads)
    sc = oracle.gen_scalars(0x5eed0000 + args.log_n, n)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_baseline(oracle, pts, inf, sc, threads)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_baseline(oracle, pts, inf, sc, threads)
    ms = t / args.steps * 1e3
    value = n / (ms * 1e-3)
    sample = f"one G1 MSM over the first 2^{log_s} pairs of the 2^{args.log_n} workload per step"
    line = {
        "impl": "reference", "metric": "g1_msm_points_per_sec", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u32x12 (381-bit Fq)",
        "data": "synthetic",
        "config": {"workload": f"g1_msm_2^{args.log_n}", "log_n": args.log_n, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import bls12_381 as bls
    import groth16_cuda

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if rank == 0 and not os.path.exists(groth16_cuda.DEFAULT_LIB):
        # normally prebuilt by __graft_entry__.build(); never rebuilt here when present (the GPU box runs
        # the library that was built and tested with the snapshot)
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build_library()
    if world > 1:
        dist.barrier()
    lib = groth16_cuda.load_library()
    lib.g16_launch_count.restype = ctypes.c_ulonglong
    lib.g16_ctx_enable_stage_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.g16_ctx_last_stage_ms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]

    peak = imad_peak() if rank == 0 else None

    ctx = groth16_cuda.Context([local])
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    if args.window_bits:
        ctx.set_window_bits(args.window_bits)

    # ---- workload: this rank's index range of the global arrays ------------------------------------
    n_total = (1 << args.log_n) * (world if args.scaling == "weak" else 1)
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    n_loc = hi - lo
    import cpu_oracle as oracle
    oracle.build()
    g1 = np.array(bls.g1_to_mont(bls.G1_GEN)[0], dtype=np.uint64)
    # generate the full deterministic streams and keep this rank's slice (cheap: ~50 ns per scalar)
    k_all = oracle.gen_scalars(0xba5e0000 + args.log_n, n_total)[lo:hi]
    s_all = oracle.gen_scalars(0x5eed0000 + args.log_n, n_total)[lo:hi]
    d_k = torch.from_numpy(k_all.view(np.int64)).to(dev)
    h_s = torch.from_numpy(np.ascontiguousarray(s_all).view(np.int64)).pin_memory()
    d_s = h_s.to(dev)
    d_pts = torch.empty((n_loc, 24), dtype=torch.int32, device=dev)
    ctx.fixed_base_mul_device("g1", g1, d_k.data_ptr(), n_loc, d_pts.data_ptr())
    torch.cuda.synchronize()
    del d_k
    bases = ctx.bases_from_device("g1", d_pts.data_ptr(), n_loc, keepalive=d_pts)
    pre_bits, pre_ms = 0, 0.0
    if not args.no_precompute:
        # one-time preprocessing of the resident CRS-style bases (outside the timed region, reported)
        t_pre = time.perf_counter()
        free_b, _total_b = torch.cuda.mem_get_info(dev)
        # the table may take half of the free HBM (19 GB at 2^24, 77 GB at 2^26); if nothing fits the
        # bases stay plain and the engine uses per-window bucket sets
        pre_bits = bases.precompute(args.precompute_bits, int(free_b * 0.5))
        torch.cuda.synchronize()
        pre_ms = (time.perf_counter() - t_pre) * 1e3

    partial = torch.zeros(48, dtype=torch.int32, device=dev)
    gathered = torch.zeros(48 * world, dtype=torch.int32, device=dev)
    out = torch.zeros(25, dtype=torch.int32, device=dev)

    from groth16_cuda.dist import msm_sharded

    def step_device(scalars_ptr):
        # local pipeline -> (N > 1: NCCL all-gather of the 192-byte partials -> fold) -> affine result
        msm_sharded(ctx, "g1", bases, scalars_ptr, n_loc, partial, gathered, out, world)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value") ---------------------------------------------------------
    lib.g16_ctx_enable_stage_timing(ctx.handle, 1)
    for _ in range(args.warmup):
        step_device(d_s.data_ptr())
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.g16_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = np.zeros(6, dtype=np.float64)
    t_wall0 = time.time()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device(d_s.data_ptr())
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    launches = (lib.g16_launch_count() - launches0) // max(1, args.steps)
    ms = e0.elapsed_time(e1) / args.steps
    st = (ctypes.c_float * 6)()
    plan = (ctypes.c_uint * 3)()
    if lib.g16_ctx_last_stage_ms(ctx.handle, st, plan) == 0:
        stage_ms = np.array(list(st))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    lib.g16_ctx_enable_stage_timing(ctx.handle, 0)
    result_dev = out.cpu().numpy().copy()

    # ---- end to end through the C ABI with host scalars ("e2e") -------------------------------------
    def step_e2e():
        # the reference-facing C-ABI call with HOST buffers: the library copies this step's scalars from
        # pinned host memory (chunked, overlapped with the pipeline) and leaves the result on the device
        if world == 1:
            ctx.msm_async("g1", bases, h_s.data_ptr(), n_loc, out.data_ptr(), 0)
        else:
            ctx.msm_async("g1", bases, h_s.data_ptr(), n_loc, 0, partial.data_ptr())
            dist.all_gather_into_tensor(gathered, partial)
            ctx.combine_partials_device("g1", gathered.data_ptr(), world, out.data_ptr())
        return out.cpu()                               # D2H of the affine result (synchronises)

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    assert (res.numpy() == result_dev).all(), "e2e result differs from the device-resident result"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) -------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        m = min(n_loc, 1 << (args.cpu_sample_log_n or 17))
        pts_h = d_pts[:m].cpu().numpy().view(np.uint32).view(np.uint64).reshape(m, 12)
        inf_h = (~pts_h.any(axis=1)).astype(np.uint8)
        sc_h = np.ascontiguousarray(s_all[:m])
        dt1 = cpu_baseline(oracle, pts_h, inf_h, sc_h, 1)
        th = oracle.max_threads()
        dtn = cpu_baseline(oracle, pts_h, inf_h, sc_h, th)
        cpu = {"value": m / dt1, "unit": "points/s", "cores": 1, "kind": "port",
               "sample": f"first 2^{int(np.log2(m))} pairs of the workload, C port of ark-ec 0.4.2 msm_bigint_wnaf "
                         f"(the reference runs it single-threaded)",
               "all_cores": {"value": m / dtn, "cores": th}}
        # parity of the sample: GPU vs oracle on the same prefix
        got, ginf = ctx.g1_msm(bases, sc_h)
        exp, einf = oracle.g1_msm(pts_h, inf_h, sc_h, threads=th)
        assert ginf == einf and (got == exp).all(), "GPU result differs from the CPU oracle on the sample"

    value = n_total / (ms * 1e-3)
    e2e_value = n_total / (e2e_ms * 1e-3)
    acc_ms = float(stage_ms[3]) if stage_ms[3] > 0 else None
    imad_peak_per_s = (peak or {}).get("imad_per_s")
    alg_imad = n_loc * FQ_MUL_PER_PAIR * IMAD_PER_FQ_MUL
    roof = {"bound": "imad", "unit": "TIMAD/s", "peak": imad_peak_per_s / 1e12 if imad_peak_per_s else None,
            "peak_source": "measured live by lib/imad_peak (mad.lo.u32, all SMs)",
            "achieved": alg_imad / (acc_ms * 1e-3) / 1e12 if acc_ms else None,
            "kernel": "BucketAccumulate<Fq>", "kernel_ms": acc_ms, "traffic": None,
            "algorithmic": f"{FQ_MUL_PER_PAIR} Fq-mul/pair x {IMAD_PER_FQ_MUL} IMAD (SURVEY.md 8d)"}
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture of
    # this exact configuration (profiles/r01_ncu_traffic.json); null for configurations never captured
    try:
        for cap in json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))["captures"]:
            if (cap["log_n"], cap["window_bits"], cap["windows"]) == (args.log_n, int(plan[0]), int(plan[1])) and world == 1:
                roof["traffic"] = cap["dram_bytes_read"] + cap["dram_bytes_write"]
                roof["traffic_unit"] = "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)"
                roof["algorithmic_bytes"] = cap["algorithmic_bytes"]
    except (OSError, KeyError, ValueError):
        pass
    if roof["achieved"] and roof["peak"]:
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["whole_step_frac"] = value / world * FQ_MUL_PER_PAIR * IMAD_PER_FQ_MUL / imad_peak_per_s
    if acc_ms and (peak or {}).get("fq_mul_per_s"):
        # what the kernel actually executes: one XYZZ mixed addition (8M + 2S) per non-zero digit, against the
        # measured throughput of the engine's own Fq multiplication (lib/imad_peak)
        ex = float(n_loc) * float(plan[1]) * 10.0 / (acc_ms * 1e-3)
        roof["executed"] = {"fq_mul_per_pair": int(plan[1]) * 10, "fq_mul_per_s": ex,
                            "fq_mul_peak_per_s": peak["fq_mul_per_s"], "frac_of_fq_mul_peak": ex / peak["fq_mul_per_s"]}
    if acc_ms:
        pt_bytes = float(n_loc) * float(plan[1]) * 96.0
        roof["point_stream_hbm"] = {"achieved_GBps": pt_bytes / (acc_ms * 1e-3) / 1e9,
                                    "peak_GBps": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
                                    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0}
    line = {
        "metric": "g1_msm_points_per_sec", "value": value, "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "u32x12 (381-bit Fq)", "data": "synthetic",
        "config": {"workload": f"g1_msm_2^{args.log_n}", "log_n": args.log_n, "pairs_total": n_total,
                   "pairs_per_gpu": n_loc, "parallelism": f"index-range x{world}",
                   "window_bits": int(plan[0]), "windows": int(plan[1]), "l2": "inputs_exceed_l2",
                   "bases": "resident, precomputed multiples 2^(c w) P (one-time %.0f ms, c=%d)" % (pre_ms, pre_bits)
                            if pre_bits else "resident, plain"},
        "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(n_loc * 32), "d2h_bytes_per_step": 100},
        "gpu_launches": int(launches),
        "stage_ms": {k: float(v) for k, v in zip(["count", "scan", "scatter", "accumulate", "reduce", "combine"], stage_ms)},
        "roofline": roof, "imad_microbench": peak, "cpu_baseline": cpu, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
