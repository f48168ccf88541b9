#!/usr/bin/env python
"""bench.py — headline benchmark of the GP-posterior + acquisition hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

metric  : acquisition candidate-points/sec, ExpectedImprovement on an exact GPR with N=4096 training
          points, fp64 (BASELINE.json `metric`); workload = the synthetic "headline" config of
          SURVEY.md §8d (Ackley-10D data, Matern52, build_gpr defaults, candidates ~ U[0,1]^10).
step    : one pass of predict + EI + first-max argmax over one batch of M candidates per GPU
          (weak scaling: every rank owns its own shard of M candidates; one NCCL all-gather of the
          per-rank (value, global index) pair picks the winner — SURVEY.md §8e).
value   : whole-job candidates/s with the candidates already resident in HBM.
e2e     : the same metric through the C-ABI call with HOST (pinned) buffers: H2D of the candidates and
          D2H of all M acquisition values + the best pair are inside the timed region.
roofline: dominant kernel = triangular DMMA GEMM (A = Linv·K*): fp64 tensor-pipe bound; achieved =
          N^2 flop per candidate x candidates per launch / average launch duration from CUDA events
          recorded on the library's stream around every launch in the timed region; peak = cuBLAS DGEMM
          measured in this process (MEASURED_PEAKS.json carries no fp64 figure).  The HBM view the
          north-star asks for is reported alongside (`hbm`).
--impl reference : the CPU restatement of the reference's path (oracle/gp_oracle.py; TensorFlow/GPflow are
          not installable here, so there is no baseline/_ref) timed on the host cores with all threads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN = 4096
DIM = 10
M_PER_GPU = 1_212_416  # 32 chunks of 37,888 candidates (= 2 full waves of 148 CTAs x 128 candidates)
METRIC = "acquisition candidate-points/sec (EI on GPR N=4096 fp64)"
UNIT = "candidates/s"


def synth_problem():
    """SURVEY.md §8d headline config: X ~ U[0,1]^10 seed 0, y = Ackley-10, Matern52, lengthscale
    0.2*sqrt(D), variance Var(y), mean mean(y), noise Var(y)/100."""
    from trieste_b200.objectives import ackley

    rng = np.random.default_rng(0)
    X = rng.uniform(size=(N_TRAIN, DIM))
    y = ackley(X)
    return X, y


def clocks_sampler(stop_evt, out):
    """Sample nvidia-smi clocks + throttle reasons during the timed region."""
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    dev = os.environ.get("LOCAL_RANK", "0")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", dev],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    out["proc"] = p
    lines = []

    def reader():
        for ln in p.stdout:
            lines.append(ln.strip())

    t = threading.Thread(target=reader, daemon=True)
    t.start()
    stop_evt.wait()
    p.terminate()
    try:
        p.wait(timeout=5)
    except Exception:
        p.kill()
    t.join(timeout=2)
    out["lines"] = lines


def summarise_clocks(lines):
    sm, mx, reasons = [], [], set()
    for ln in lines or []:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 7:
            continue
        try:
            sm.append(float(f[0]))
            mx.append(float(f[1]))
        except ValueError:
            continue
        for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
            if v.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def dgemm_peak_tflops():
    """Measured fp64 GEMM peak on this GPU (cuBLAS via torch.matmul, 6144^3, best of 5)."""
    import torch

    n = 6144
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n**3 / (best * 1e-3) / 1e12


_ORACLE_MODEL = None


def _oracle_model():
    """Oracle model of the headline config, built once (outside every timed region)."""
    global _ORACLE_MODEL
    if _ORACLE_MODEL is None:
        from oracle import gp_oracle as o  # cpu_baseline / reference leg: the checker timed beside the product

        X, y = synth_problem()
        var = float(np.var(y))
        om = o.build_model("matern52", X, y, var, np.full(DIM, 0.2 * math.sqrt(DIM)), var / 100.0, float(np.mean(y)))
        _ORACLE_MODEL = (o, om, o.ei_eta(om))
    return _ORACLE_MODEL


def cpu_reference_rate(sample_target_s=12.0, chunk=8192, seed=1):
    """Oracle (CPU restatement of the reference path: predict + EI + argmax) on a bounded sample of the
    same workload, all host threads (NumPy/SciPy BLAS).  Returns (candidates/s, candidates, seconds)."""
    o, om, eta = _oracle_model()
    try:  # torchrun exports OMP_NUM_THREADS=1: give the BLAS pool every host core back
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass
    rng = np.random.default_rng(seed)
    Xc = rng.uniform(size=(chunk, DIM))
    t0 = time.perf_counter()
    o.expected_improvement_at(om, Xc, eta, chunk=chunk)
    dt1 = time.perf_counter() - t0
    nchunks = max(1, min(64, int(sample_target_s / max(dt1, 1e-3))))
    Xc = rng.uniform(size=(chunk * nchunks, DIM))
    t0 = time.perf_counter()
    ei = o.expected_improvement_at(om, Xc, eta, chunk=chunk)
    int(np.argmax(ei))
    dt = time.perf_counter() - t0
    return (chunk * nchunks) / dt, chunk * nchunks, dt


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path is TensorFlow/GPflow, which cannot
    be installed here (no wheels, no network) -> the oracle port is timed instead, on rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    _oracle_model()
    per_step, total = [], 0
    for i in range(args.warmup + args.steps):
        rate, n, dt = cpu_reference_rate(sample_target_s=4.0 if i >= args.warmup else 1.0, seed=1 + i)
        if i >= args.warmup:
            per_step.append(dt)
            total += n
    value = total / sum(per_step)
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(per_step)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"headline: EI on GPR N={N_TRAIN} D={DIM} Matern52 fp64, Ackley-10 synthetic (SURVEY.md §8d)",
                   "candidates_per_step": total // max(1, args.steps)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{total} candidates over {args.steps} steps (NumPy/SciPy oracle, chunks of 8192, all host "
                                   "threads; TensorFlow/GPflow not installable -> CPU restatement)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--candidates", type=int, default=M_PER_GPU, help="candidates per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--engine", default="int8", choices=["int8", "fp64"],
                    help="variance GEMM engine: int8 = fp64-accurate Ozaki split on the INT8 tensor cores (default), fp64 = native DMMA")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    g.build()
    import trieste_b200 as tb
    from trieste_b200 import _lib
    from trieste_b200.acquisition import ExpectedImprovement

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- model (replicated on every rank; once-per-step precompute is outside the timed region) ----
    X, y = synth_problem()
    ds = tb.Dataset(X, y)
    spec = tb.build_gpr(ds, tb.Box([0.0] * DIM, [1.0] * DIM))
    model = tb.GaussianProcessRegression(spec, device=local_rank)
    model.set_engine(args.engine)
    fn = ExpectedImprovement().prepare_acquisition_function(model, ds)
    lib = _lib.lib()
    h = model.handle

    M = args.candidates
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1 + rank)
    xc_dev = torch.rand(M, DIM, dtype=torch.float64, device="cuda", generator=gen)  # Box.sample semantics
    vals_dev = torch.empty(M, dtype=torch.float64, device="cuda")
    xc_host = torch.empty(M, DIM, dtype=torch.float64).pin_memory()
    xc_host.copy_(xc_dev.cpu())
    vals_host = torch.empty(M, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    sp = C.c_void_p()
    _lib.check(lib.tb_gp_stream(h, C.byref(sp)))
    stream = torch.cuda.ExternalStream(sp.value)
    best_v = C.c_double()
    best_i = C.c_int64()

    def exchange(v, i):
        """single collective of the path: all-gather of (value, global index); first-max wins."""
        if world == 1:
            return v, i
        t = torch.tensor([v, float(rank * M + i)], dtype=torch.float64, device="cuda")
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        pairs = [(float(o[0]), int(o[1])) for o in out]
        bv, bi = pairs[0]
        for pv, pi in pairs[1:]:
            if pv > bv or (pv == bv and pi < bi):
                bv, bi = pv, pi
        return bv, bi

    def step_device():
        flush.zero_()
        _lib.check(lib.tb_acq_argmax(h, _lib.ACQ_EI, fn.eta, xc_dev.data_ptr(), M, vals_dev.data_ptr(),
                                     C.byref(best_v), C.byref(best_i)))
        return exchange(best_v.value, best_i.value)

    def step_host():
        flush.zero_()
        _lib.check(lib.tb_acq_argmax(h, _lib.ACQ_EI, fn.eta, xc_host.data_ptr(), M, vals_host.data_ptr(),
                                     C.byref(best_v), C.byref(best_i)))
        return exchange(best_v.value, best_i.value)

    def timed(step_fn, steps, profile=False):
        for _ in range(args.warmup):
            step_fn()
        if profile:
            lib.tb_gp_profile(h, 1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.tb_launch_count()
        e0.record(stream)
        for _ in range(steps):
            res = step_fn()
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.tb_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, launches, res

    # ---- clocks during the timed region ----
    stop_evt, clk = threading.Event(), {}
    th = threading.Thread(target=clocks_sampler, args=(stop_evt, clk), daemon=True)
    if rank == 0:
        th.start()
        time.sleep(0.3)

    ms_dev, launches, res = timed(step_device, args.steps, profile=True)
    tg_ms, tg_n, tg_fl = C.c_double(), C.c_int64(), C.c_double()
    lib.tb_gp_profile_read(h, C.byref(tg_ms), C.byref(tg_n), C.byref(tg_fl))
    lib.tb_gp_profile(h, 0)
    ms_e2e, _, _ = timed(step_host, args.steps)

    if rank == 0:
        stop_evt.set()
        th.join(timeout=10)
    clocks = summarise_clocks(clk.get("lines"))

    value = world * M * args.steps / (ms_dev * 1e-3)
    e2e_value = world * M * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        dgemm_tf = dgemm_peak_tflops()
        fp64_eq_tf = tg_fl.value / (tg_ms.value * 1e-3) / 1e12 if tg_ms.value > 0 else 0.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        nb = N_TRAIN // 128
        cand_per_launch = tg_fl.value / max(tg_n.value, 1) / (N_TRAIN**2)
        avg_launch_s = tg_ms.value * 1e-3 / max(tg_n.value, 1)
        if args.engine == "int8":
            # 21 exact int8 digit products per fp64 product: algorithmic int8 ops = 21 N^2 / candidate.
            # int8 dense rate = 2x the bf16 dense rate on this part (4.5 vs 2.25 POP/s nominal): peak = 2 x the measured
            # cuBLAS bf16 figure of MEASURED_PEAKS.json (sustained: the kernel is timed inside a long step)
            ops_per_cand = 21.0 * N_TRAIN**2
            achieved = ops_per_cand * cand_per_launch / avg_launch_s / 1e12 if avg_launch_s > 0 else 0.0
            bf16 = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
            peak = 2.0 * bf16
            roof_unit = "TOP/s"
            kernel_name = "oz::trigemm_i8_kernel (tcgen05 kind::i8, TMEM accumulators)"
            peak_src = ("of measured: 2 x bf16_tflops_sustained of MEASURED_PEAKS.json (int8 dense = 2 x bf16 dense on B200); "
                        "tools/i8_umma_test.cu measured 3838 TOP/s burst for the same 128x128 SS MMA shape")
            # HBM: the K* digit tiles (6 B / element) are written once and, with the L2-friendly CTA order, read ~once
            bytes_per_cand = 2.0 * 6.0 * N_TRAIN
            traffic_key = "trigemm_i8_dram_bytes_per_launch"
        else:
            achieved, peak, roof_unit = fp64_eq_tf, dgemm_tf, "TFLOP/s"
            kernel_name = "trigemm_kernel<false, EPI_SUMSQ> (fp64 DMMA)"
            peak_src = "of measured: cuBLAS DGEMM 6144^3 in this process (MEASURED_PEAKS.json has no fp64 figure)"
            bytes_per_cand = 8.0 * N_TRAIN * (nb + 1) / 2.0
            traffic_key = "trigemm_dram_bytes_per_launch"
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(traffic_key)
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, n, dt = cpu_reference_rate()
            cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"{n} candidates of the same workload in {dt:.1f} s (NumPy/SciPy oracle, chunks of 8192, all host threads)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "engine": ("int8: fp64 operands split error-free into 6 base-256 int8 digits, 21 exact digit GEMMs on tcgen05 "
                           "kind::i8 with int32 TMEM accumulators, fp64 recombination; parity to the fp64 oracle at 1e-9 sigma_f^2"
                           if args.engine == "int8" else "fp64: native DMMA triangular GEMM"),
                "workload": f"headline: EI on GPR N={N_TRAIN} D={DIM} Matern52 fp64, Ackley-10 synthetic (SURVEY.md §8d)",
                "candidates_per_gpu_per_step": M, "parallelism": f"candidate-sharded x{world}, 1 NCCL all-gather/step",
                "l2": "256 MiB L2 flush between timed iterations; per-chunk Ks scratch (1.2 GB) also exceeds L2",
            },
            "roofline": {
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": roof_unit,
                "frac": achieved / peak if peak > 0 else None, "traffic": traffic,
                "kernel": kernel_name, "launches_timed": tg_n.value,
                "avg_launch_ms": avg_launch_s * 1e3, "candidates_per_launch": cand_per_launch,
                "peak_source": peak_src,
                # transparency: the same achieved figure against the two other candidates for an int8 denominator
                "frac_of_nominal_int8_4500": (achieved / 4500.0) if args.engine == "int8" else None,
                "frac_of_measured_int8_burst_3838": (achieved / 3838.0) if args.engine == "int8" else None,
                "fp64_equivalent_tflops": fp64_eq_tf, "fp64_dgemm_peak_tflops": dgemm_tf,
                "hbm": {"algorithmic_bytes_per_candidate": bytes_per_cand,
                        "achieved_gbs": bytes_per_cand * cand_per_launch / avg_launch_s / 1e9 if avg_launch_s > 0 else None,
                        "peak_gbs": hbm_peak,
                        "frac": (bytes_per_cand * cand_per_launch / avg_launch_s / 1e9) / hbm_peak if avg_launch_s > 0 else None,
                        "note": "tensor-pipe bound (N^2 multiply-adds per candidate); HBM fraction reported as the north-star asks"},
            },
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * M * DIM * 8, "d2h_bytes_per_step": world * (M * 8 + 16),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "best": {"value": res[0], "global_index": res[1]},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
