#!/usr/bin/env python
"""Headline benchmark: G1 MSM points/s at 2^24 on B200 + Groth16 prove ms (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n 24] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...        (one rank per GPU, NCCL)

A "step" is one full G1 MSM (digits -> sort -> bucket accumulate -> reduce -> combine -> affine) over
N = 2^log_n synthetic (scalar, point) pairs: scalars uniform in [0, r) from SplitMix64(0x5eed0000+log_n),
points k_i*G with k_i from SplitMix64(0xba5e0000+log_n) built on the GPU by the fixed-base kernel
(SURVEY.md 8d).  With N GPUs the pairs are partitioned by index range (strong scaling, BASELINE
config 4); each rank reduces its range to one projective partial sum, the partials are exchanged with
one NCCL all-gather (192 bytes per rank) and folded on every rank.

  value    : pairs/s, scalars already resident in HBM when the timed region starts
  e2e      : pairs/s through the C ABI with scalars in pinned HOST memory: H2D of the scalars and D2H of
             the affine result inside the timed region, every step
  oneshot  : (N = 1) the reference seam as it stands, `Prover::multi_scalar_mult_g1(&scalars, &points)`:
             bases AND scalars uploaded from the host on every call (g16_g1_msm_oneshot)
  prove    : the second half of the metric -- `g16_prove` (4 x G1 + 1 x G2 MSM schedule of Prover::prove) through
             the C ABI with host buffers on a synthetic 2^20 key (BASELINE config 3; N = 8: config 5, setup + prove
             at 2^24 on a multi-device context), checked exactly in the exponent, CPU five-MSM baseline beside it
  roofline : the dominant kernel (bucket accumulation) against the measured IMAD peak
  cpu_baseline : the C port of the reference's CPU path (ark msm_bigint_wnaf) on a bounded sample
The full-size result is verified OUTSIDE the timed region through the discrete-log identity
sum s_i (k_i G) = (sum s_i k_i mod r) G (exact, any size).
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
FQ_MUL_PER_MADD = (8 * 300 + 444) / 300.0   # executed: 8 M + (2 products, 1 reduction), in units of one 300-MAD multiplication
IMAD_PER_FQ_MUL = 600
SEED_POINTS, SEED_SCALARS = 0xba5e0000, 0x5eed0000
L2_BYTES = 126 << 20


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
                    help="log2 of the bounded CPU sample (default: 17 for cpu_baseline; --impl reference sizes its "
                         "sample from a probe so that warmup + steps finish in about two minutes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prove", action="store_true", help="skip the Groth16 prove sub-record")
    ap.add_argument("--no-oneshot", action="store_true", help="skip the one-shot seam measurement (N = 1)")
    ap.add_argument("--prove-log-n", type=int, default=0, help="size of the prove sub-record (default 20; 24 at N = 8)")
    ap.add_argument("--prove-steps", type=int, default=5)
    ap.add_argument("--lib", default=None, help="an A/B build of the library (tools/lab_build.py) instead of lib/libg16cuda.so; "
                                                "the line then carries \"lab_build\": <path>")
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


def source_hash() -> str:
    """The same digest build.py bakes into g16_version(): a stale prebuilt library is detected, not trusted."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("g16_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.source_hash()


def cpu_msm_seconds(oracle, pts, inf, sc, threads):
    t = time.perf_counter()
    oracle.g1_msm(pts, inf, sc, threads=threads)
    return time.perf_counter() - t


def cpu_workload_prefix(oracle, log_n, m, threads):
    """First m pairs of the bench workload, built on the CPU: the SAME points (k_i G, full-width k_i from
    SplitMix64(SEED_POINTS + log_n)) and scalars the GPU arm uses."""
    import bls12_381 as bls
    g1 = np.array(bls.g1_to_mont(bls.G1_GEN)[0], dtype=np.uint64)
    k = oracle.gen_scalars(SEED_POINTS + log_n, m)
    pts, inf = oracle.g1_fixed_base_mul(g1, k, threads=threads)
    sc = oracle.gen_scalars(SEED_SCALARS + log_n, m)
    return pts, inf, sc


def cpu_prove_sample(oracle, log_m, threads):
    """The five MSMs of Prover::prove (crates/groth16-core/src/lib.rs:179,197,220,255,264) on the CPU over a synthetic
    key of 2^log_m variables (C port of ark's Pippenger, `threads` host threads).  Returns seconds."""
    import bls12_381 as bls
    import prove_model as pm
    g1 = np.array(bls.g1_to_mont(bls.G1_GEN)[0], dtype=np.uint64)
    g2 = np.array(bls.g2_to_mont(bls.G2_GEN)[0], dtype=np.uint64)
    m = 1 << log_m
    k = pm.synthetic_key_exponents(m, 0xc9000, 1)
    pk = {"num_public": 1}
    for name in ("a_g1", "b_g1", "ic_g1", "h_g1"):
        pk[name], pk[name + "_inf"] = oracle.g1_fixed_base_mul(g1, k[name], threads=threads)
    pk["b_g2"], pk["b_g2_inf"] = oracle.g2_fixed_base_mul(g2, k["b_g2"], threads=threads)
    for name in ("alpha_g1", "beta_g1", "delta_g1"):
        pk[name] = oracle.g1_fixed_base_mul(g1, k[name][None])[0][0]
    for name in ("beta_g2", "delta_g2"):
        pk[name] = oracle.g2_fixed_base_mul(g2, k[name][None])[0][0]
    w = oracle.gen_scalars(0xc9100, m); w[0] = pm.ONE
    h = oracle.gen_scalars(0xc9200, m - 1)
    r, s = oracle.gen_scalars(0xc9300, 2)
    t = time.perf_counter()
    proof = pm.five_msms_cpu(pk, w, h, r, s, threads=threads)
    dt = time.perf_counter() - t
    assert pm.proofs_equal(proof, pm.proof_in_exponent(k, 1, w, h, r, s, (g1, g2), threads=threads))
    return dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The reference
    is Rust (no toolchain here), so this is the C port in oracle/ (cpu_baseline.kind = "port").  All host threads
    this process may run on (os.sched_getaffinity -- NOT the OpenMP environment, so the figure is the same however
    bench.py was launched), windows in parallel = ark's `parallel` feature; each step one MSM over a bounded prefix
    of the SAME 2^log_n workload the GPU arm runs, sized from a short probe so that warmup + steps take about two
    minutes.  The reference itself runs this path single-threaded: that figure is reported too."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cpu_oracle as oracle
    oracle.build()
    threads = oracle.max_threads()
    warm = max(1, min(args.warmup, 1))
    # probe at 2^15, then the largest power of two that keeps (warm + steps) MSMs within ~120 s
    pts, inf, sc = cpu_workload_prefix(oracle, args.log_n, 1 << 15, threads)
    cpu_msm_seconds(oracle, pts, inf, sc, threads)
    rate = (1 << 15) / cpu_msm_seconds(oracle, pts, inf, sc, threads)
    single = (1 << 15) / cpu_msm_seconds(oracle, pts, inf, sc, 1)
    if args.cpu_sample_log_n:
        log_s = min(args.cpu_sample_log_n, args.log_n)
    else:
        log_s = 15
        while log_s < min(args.log_n, 22) and (warm + args.steps) * (2 << log_s) / (1.3 * rate) < 120.0:
            log_s += 1
    n = 1 << log_s
    if log_s != 15:
        pts, inf, sc = cpu_workload_prefix(oracle, args.log_n, n, threads)
    for _ in range(warm):
        cpu_msm_seconds(oracle, pts, inf, sc, threads)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_msm_seconds(oracle, pts, inf, sc, threads)
    ms = t / args.steps * 1e3
    value = n / (ms * 1e-3)
    c_s, c_f = oracle.msm_window(n), oracle.msm_window(1 << args.log_n)
    sample = (f"one G1 MSM over the first 2^{log_s} pairs of the 2^{args.log_n} workload per step (same points and scalars "
              f"as the GPU arm); ark's window rule gives c = {c_s} ({-(-255 // c_s)} windows) at the sample size and c = {c_f} "
              f"({-(-255 // c_f)} windows) at 2^{args.log_n}, so the full-size CPU rate is about "
              f"{(-(-255 // c_s)) / (-(-255 // c_f)):.2f}x the sampled one")
    prove_log_m = 16
    prove_s = cpu_prove_sample(oracle, prove_log_m, threads)
    line = {
        "impl": "reference", "metric": "g1_msm_points_per_sec", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u32x12 (381-bit Fq)",
        "data": "synthetic",
        "config": {"workload": f"g1_msm_2^{args.log_n}", "log_n": args.log_n, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port", "sample": sample,
                         "single_thread": {"value": single, "cores": 1, "sample": "first 2^15 pairs; what the reference "
                                           "itself runs (no `parallel` feature, SURVEY.md 0.5)"}},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "prove": {"metric": "groth16_prove_ms", "sample": f"five MSMs of Prover::prove on a synthetic 2^{prove_log_m} key, "
                  f"{threads} threads", "ms_sample": prove_s * 1e3,
                  "ms_extrapolated_2^20": prove_s * 1e3 * (1 << (20 - prove_log_m)),
                  "extrapolation": "linear in the key size (conservative: larger MSMs cost ~10 % less per pair)"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


LAB_LIB = None   # --lib


def ensure_library(rank, world, dist):
    """Load lib/libg16cuda.so; rebuild it first when it is missing or was built from other sources (the digest of
    csrc/ + include/ is baked into g16_version())."""
    import groth16_cuda
    want = source_hash()
    if LAB_LIB:   # A/B build: use as is
        groth16_cuda.DEFAULT_LIB = LAB_LIB
        lib = groth16_cuda.load_library()
        return lib, lib.g16_version().decode(), want, False
    if rank == 0:
        stale = not os.path.exists(groth16_cuda.DEFAULT_LIB)
        if not stale:
            probe = subprocess.run([sys.executable, "-c",
                                    "import ctypes,sys; l=ctypes.CDLL(sys.argv[1]); l.g16_version.restype=ctypes.c_char_p; "
                                    "print(l.g16_version().decode())", groth16_cuda.DEFAULT_LIB], capture_output=True, text=True)
            stale = want not in probe.stdout
        if stale:
            sys.path.insert(0, ROOT)
            import __graft_entry__
            __graft_entry__.build_library()
    if world > 1:
        dist.barrier()
    lib = groth16_cuda.load_library()
    version = lib.g16_version().decode()
    return lib, version, want, want in version


def prove_record(args, oracle, devices, log_n, steps, with_setup):
    """Groth16 prove through the reference-facing C ABI (g16_prove) with HOST buffers on `devices` (one process; a
    multi-device context shards every key array by index range).  Synthetic ProvingKey-shaped key of 2^log_n variables
    built by the engine's own fixed-base path (= the group work of CRS::generate_from_qap, timed as `setup` when
    with_setup), resident + precomputed; proofs verified exactly in the exponent at full size."""
    import torch
    import bls12_381 as bls
    import groth16_cuda
    import prove_model as pm
    n = 1 << log_n
    g1 = np.array(bls.g1_to_mont(bls.G1_GEN)[0], dtype=np.uint64)
    g2 = np.array(bls.g2_to_mont(bls.G2_GEN)[0], dtype=np.uint64)
    th = oracle.max_threads()

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.int64) if a.dtype == np.uint64 else np.ascontiguousarray(a))
        return t.pin_memory().numpy().view(a.dtype)

    ctx = groth16_cuda.Context(list(devices))
    try:
        k = pm.synthetic_key_exponents(n, 0xc0de00, 1)
        names = ("a_g1", "b_g1", "ic_g1", "h_g1", "b_g2")
        ks = {nm: pin(k[nm]) for nm in names}
        outs = {nm: (pin(np.zeros((ks[nm].shape[0], 24 if nm == "b_g2" else 12), dtype=np.uint64)),
                     pin(np.zeros(ks[nm].shape[0], dtype=np.uint8))) for nm in names}
        ctx.fixed_base_mul_g1(g1, ks["a_g1"][:4096])      # tables + workspaces
        ctx.fixed_base_mul_g2(g2, ks["b_g2"][:4096])
        t0 = time.perf_counter()
        pk = {"num_public": 1}
        for nm in names:
            f = ctx.fixed_base_mul_g2 if nm == "b_g2" else ctx.fixed_base_mul_g1
            pk[nm], pk[nm + "_inf"] = f(g2 if nm == "b_g2" else g1, ks[nm], out=outs[nm])
        setup_ms = (time.perf_counter() - t0) * 1e3
        for nm in ("alpha_g1", "beta_g1", "delta_g1"):
            pk[nm] = oracle.g1_fixed_base_mul(g1, k[nm][None])[0][0]
        for nm in ("beta_g2", "delta_g2"):
            pk[nm] = oracle.g2_fixed_base_mul(g2, k[nm][None])[0][0]
        idx = [0, 1, n // 3, n - 3]
        for nm in names:       # spot check of the generated key against the oracle
            f = oracle.g2_fixed_base_mul if nm == "b_g2" else oracle.g1_fixed_base_mul
            exp, _ = f(g2 if nm == "b_g2" else g1, k[nm][idx])
            assert (pk[nm][idx] == exp).all(), f"setup array {nm} differs from the oracle"
        t0 = time.perf_counter()
        dev_pk = ctx.pk_upload(pk)
        upload_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        ctx.pk_precompute(dev_pk)
        pre_s = time.perf_counter() - t0
        r, s = oracle.gen_scalars(0xaabb, 2)
        out = {"metric": "groth16_prove_ms", "n_gpus": len(devices), "steps": steps,
               "config": {"workload": f"groth16_prove_2^{log_n}" + ("_with_setup" if with_setup else ""),
                          "key": f"synthetic ProvingKey, N = n = 2^{log_n} variables / H coefficients, 1 public input, "
                                 f"resident + precomputed (upload {upload_s:.2f} s, precompute {pre_s:.2f} s, one-time)",
                          "msms": "4 x G1 + 1 x G2 (+ ad-hoc terms)", "context": f"one process, {len(devices)} device(s)"},
               "through": "g16_prove: assignment + H coefficients in pinned host memory in, proof out",
               "h2d_bytes_per_step": (2 * n - 1) * 32, "d2h_bytes_per_step": 51 * 4}
        if with_setup:
            out["setup"] = {"ms": setup_ms, "what": f"4 x 2^{log_n} G1 + 2^{log_n} G2 fixed-base multiplications "
                            "(full-width scalars), host scalars in, host points out, pinned", "points_per_s": 5 * n / (setup_ms * 1e-3)}
        for dist_name, bits in (("full_width", 255), ("ref_faithful_u64", 64)):
            if bits == 64:   # tables for the scalars the reference's truncation yields (g16_pk_precompute_bits: a hint)
                t0 = time.perf_counter()
                ctx.pk_precompute(dev_pk, scalar_bits=64)
                out["config"]["key"] += f"; ref_faithful_u64: tables rebuilt for 64-bit scalars ({time.perf_counter() - t0:.2f} s)"
            w = oracle.gen_scalars(0x1000 + bits, n, bits)
            w[0] = pm.ONE
            w = pin(w)
            h = pin(oracle.gen_scalars(0x2000 + bits, n - 1, bits))
            for _ in range(2):
                proof = ctx.prove(dev_pk, w, h, r, s)
            t0 = time.perf_counter()
            for _ in range(steps):
                proof = ctx.prove(dev_pk, w, h, r, s)
            ms = (time.perf_counter() - t0) / steps * 1e3
            ok = pm.proofs_equal(proof, pm.proof_in_exponent(k, 1, w, h, r, s, (g1, g2), threads=th))
            assert ok, f"2^{log_n} proof ({dist_name}) differs from the exact computation in the exponent"
            out[dist_name] = {"ms": ms, "bit_exact_in_exponent": True}
        out["value"] = out["full_width"]["ms"]
        out["unit"] = "ms"
        out["scalars"] = ("full_width: uniform < r (value); ref_faithful_u64: uniform < 2^64, what the reference's "
                          "truncation (lib.rs:156-161) produces")
        dev_pk.free()
    finally:
        ctx.close()
    if not args.no_cpu_baseline:
        log_m = 16
        cpu_s = cpu_prove_sample(oracle, log_m, th)
        out["cpu_baseline"] = {"kind": "port", "cores": th, "ms_sample": cpu_s * 1e3,
                               "sample": f"the same five MSMs on a synthetic 2^{log_m} key (C port of ark-ec 0.4.2 msm)",
                               "ms_extrapolated": cpu_s * 1e3 * (1 << (log_n - log_m)),
                               "extrapolation": "linear in the key size (conservative: larger MSMs cost ~10 % less per pair)"}
    return out


def main():
    args = parse_args()
    global LAB_LIB
    LAB_LIB = os.path.abspath(args.lib) if args.lib else None
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
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")   # host-side waits that must not occupy the GPUs

    lib, version, src_hash, lib_fresh = ensure_library(rank, world, dist)
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
    # the full deterministic streams (cheap: ~50 ns per scalar); every rank keeps its slice, rank 0 the whole for the check
    k_full = oracle.gen_scalars(SEED_POINTS + args.log_n, n_total)
    s_full = oracle.gen_scalars(SEED_SCALARS + args.log_n, n_total)
    k_all, s_all = k_full[lo:hi], s_full[lo:hi]
    expected = None
    if rank == 0:
        exp_xy, exp_inf = oracle.g1_fixed_base_mul(g1, oracle.dot_mod_r(s_full, k_full)[None])
        expected = (exp_xy[0], int(exp_inf[0]))
    del k_full, s_full
    d_k = torch.from_numpy(np.ascontiguousarray(k_all).view(np.int64)).to(dev)
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

    # working set of one step: scalars + the base table it gathers from; below ~2 x L2 the cache is flushed between
    # timed steps (each step then gets its own event pair), otherwise the inputs themselves exceed L2
    nwin_guess = -(-256 // pre_bits) if pre_bits else 1
    flush = n_loc * (32 + 96 * nwin_guess) < 2 * L2_BYTES
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush else None

    # ---- device-resident timing ("value") ---------------------------------------------------------
    lib.g16_ctx_enable_stage_timing(ctx.handle, 1)
    for _ in range(args.warmup):
        step_device(d_s.data_ptr())
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.g16_launch_count()
    stage_ms = np.zeros(6, dtype=np.float64)
    t_wall0 = time.time()
    barrier()
    if not flush:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_device(d_s.data_ptr())
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / args.steps
    else:
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in pairs:
            flush_buf.fill_(1)                         # evicts L2 (256 MiB write), outside the event pair
            a.record(stream)
            step_device(d_s.data_ptr())
            b.record(stream)
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in pairs) / args.steps
    t_wall1 = time.time()
    launches = (lib.g16_launch_count() - launches0) // max(1, args.steps)
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
    if rank == 0:
        # full-size parity, outside the timed region: sum s_i (k_i G) == (sum s_i k_i mod r) G
        got = result_dev.view(np.uint32)
        assert int(got[24]) == expected[1] and (got[:24].view(np.uint64) == expected[0]).all(), \
            f"GPU result at 2^{args.log_n} differs from the exact value (discrete-log identity)"

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

    # ---- the reference seam as it stands: bases + scalars from the host on every call (N = 1) --------
    oneshot = None
    if world == 1 and not args.no_oneshot:
        try:
            pts_h = d_pts.cpu().numpy().view(np.uint32).view(np.uint64).reshape(n_loc, 12)
            sc_h = np.ascontiguousarray(s_all)
            ctx.multi_scalar_mult_g1(sc_h, pts_h)                      # warm-up
            reps = 2
            t0 = time.perf_counter()
            for _ in range(reps):
                got_xy, got_inf = ctx.multi_scalar_mult_g1(sc_h, pts_h)
            one_ms = (time.perf_counter() - t0) / reps * 1e3
            assert got_inf == expected[1] and (got_xy == expected[0]).all(), "one-shot result differs"
            oneshot = {"value": n_total / (one_ms * 1e-3), "unit": "points/s", "ms_per_step": one_ms, "steps": reps,
                       "call": "g16_g1_msm_oneshot = Prover::multi_scalar_mult_g1(&scalars, &points) unchanged "
                               "(crates/groth16-core/src/lib.rs:275-286): pageable host bases + scalars uploaded per call, "
                               "no precomputed table",
                       "h2d_bytes_per_step": int(n_loc * (96 + 32)), "d2h_bytes_per_step": 97}
            del pts_h
        except Exception as e:  # pragma: no cover
            oneshot = {"error": repr(e)}

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) -------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        m = min(n_loc, 1 << (args.cpu_sample_log_n or 17))
        pts_s = d_pts[:m].cpu().numpy().view(np.uint32).view(np.uint64).reshape(m, 12)
        inf_s = (~pts_s.any(axis=1)).astype(np.uint8)
        sc_s = np.ascontiguousarray(s_all[:m])
        th = oracle.max_threads()
        dt1 = cpu_msm_seconds(oracle, pts_s, inf_s, sc_s, 1)
        dtn = cpu_msm_seconds(oracle, pts_s, inf_s, sc_s, th)
        cpu = {"value": m / dt1, "unit": "points/s", "cores": 1, "kind": "port",
               "sample": f"first 2^{int(np.log2(m))} pairs of the workload, C port of ark-ec 0.4.2 msm_bigint_wnaf "
                         f"(the reference runs it single-threaded)",
               "all_cores": {"value": m / dtn, "cores": th}}
        # parity of the sample: GPU vs oracle on the same prefix
        got, ginf = ctx.g1_msm(bases, sc_s)
        exp, einf = oracle.g1_msm(pts_s, inf_s, sc_s, threads=th)
        assert ginf == einf and (got == exp).all(), "GPU result differs from the CPU oracle on the sample"

    # ---- Groth16 prove (second half of the metric) -------------------------------------------------
    # release the MSM workload first: the prove key and its tables need the memory at 2^24
    bases.free()
    ctx.close()
    del d_pts, d_s, bases
    torch.cuda.empty_cache()
    prove = None
    if not args.no_prove and args.scaling == "strong":
        if world > 1:
            dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                plog = args.prove_log_n or (24 if world == 8 else 20)
                prove = prove_record(args, oracle, list(range(world)), plog, args.prove_steps, with_setup=(world == 8))
            except Exception as e:  # pragma: no cover
                prove = {"error": repr(e)}
        if world > 1:
            dist.barrier(group=cpu_group)                # the other ranks wait on the host, their GPUs stay free

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

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
    # this exact configuration (profiles/*ncu_traffic.json); null for configurations never captured
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            for cap in json.load(open(os.path.join(ROOT, "profiles", name)))["captures"]:
                if (cap["log_n"], cap["window_bits"], cap["windows"]) == (args.log_n, int(plan[0]), int(plan[1])) and world == 1:
                    roof["traffic"] = cap["dram_bytes_read"] + cap["dram_bytes_write"]
                    roof["traffic_unit"] = "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)"
                    roof["traffic_source"] = "profiles/" + name
                    roof["algorithmic_bytes"] = cap["algorithmic_bytes"]
        except (OSError, KeyError, ValueError):
            continue
        if roof["traffic"]:
            break
    if roof["achieved"] and roof["peak"]:
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["whole_step_frac"] = value / world * FQ_MUL_PER_PAIR * IMAD_PER_FQ_MUL / imad_peak_per_s
        roof["note"] = ("frac > 1 is possible: the accounting charges the canonical 16 windows per pair, the resident "
                        "precomputed table lets the kernel run ceil(256/c) windows (see `executed`); the table costs "
                        f"{pre_ms:.0f} ms once and {-(-256 // max(pre_bits, 1)) * n_loc * 96 / 1e9:.1f} GB per GPU, outside the timed region")
    if acc_ms and (peak or {}).get("fq_mul_per_s"):
        # what the kernel actually executes: one XYZZ mixed addition per non-zero digit = 8 multiplications + one
        # two-product multiplication for y3 (two products, one reduction: 444 of the 600 wide MADs of two
        # multiplications) = 9.48 multiplication equivalents, against the measured throughput of the engine's own Fq
        # multiplication (lib/imad_peak)
        ex = float(n_loc) * float(plan[1]) * FQ_MUL_PER_MADD / (acc_ms * 1e-3)
        roof["executed"] = {"fq_mul_per_pair": round(int(plan[1]) * FQ_MUL_PER_MADD, 2), "fq_mul_per_s": ex,
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
                   "window_bits": int(plan[0]), "windows": int(plan[1]),
                   "l2": "flushed between timed steps (256 MiB write, outside the per-step event pairs)" if flush else "inputs_exceed_l2",
                   "bases": "resident, precomputed multiples 2^(c w) P (one-time %.0f ms, c=%d)" % (pre_ms, pre_bits)
                            if pre_bits else "resident, plain",
                   "parity": f"full 2^{args.log_n} result == (sum s_i k_i mod r) G, checked outside the timed region"},
        "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(n_loc * 32), "d2h_bytes_per_step": 100},
        "gpu_launches": int(launches),
        "stage_ms": {k: float(v) for k, v in zip(["count", "scan", "scatter", "accumulate", "reduce", "combine"], stage_ms)},
        "roofline": roof, "imad_microbench": peak, "cpu_baseline": cpu, "oneshot_seam": oneshot, "prove": prove,
        "library": {"version": version, "source_hash": src_hash, "matches_source": bool(lib_fresh), "lab_build": LAB_LIB},
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
