#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): SO(5) Matern-5/2 EigenbasisSumKernel covariance, 16384 x 16384, 100 irreps, fp64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 16384] [--workload so5|so3|su4]

A step = one full covariance evaluation through the public API (`kernel(x, y)`: two embedding launches + the
fused pairwise kernel), inputs resident in HBM.  N > 1 (torchrun): the x rows are sharded across ranks, no
collective on the data path; the total problem is fixed (strong scaling, as BASELINE.json quotes it).
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference algorithm
(oracle/lsk_oracle.py; the reference itself is pure torch-CPU Python and cannot run the full size) on a bounded
block of the same workload, on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_ENTRY = {"so5": 756.0, "so3": 98.0, "su4": 800.0}       # algorithmic work per covariance entry
KERNEL_NAME = {"so5": "lsk::so5::stair_kernel", "so3": "lsk::pairwise_kernel<CHEB1>", "su4": "lsk::pairwise_kernel<HORNER>"}
EXECUTED_FLOP_PER_ENTRY = {"so5": 2.0 * (44 + 99 + 4)}
WORKLOADS = {
    "so5": dict(group="SO", n=5, order=100, measure=dict(kind="matern", dim=10, lengthscale=1.0, nu=2.5),
                name="SO(5) Matern(nu=5/2) EigenbasisSumKernel, 100 irreps, fp64"),
    "so3": dict(group="SO", n=3, order=20, measure=dict(kind="heat", dim=3, lengthscale=1.0),
                name="SO(3) heat EigenbasisSumKernel, 20 irreps, fp64"),
    "su4": dict(group="SU", n=4, order=20, measure=dict(kind="heat", dim=15, lengthscale=1.0),
                name="SU(4) heat EigenbasisSumKernel, 20 irreps, complex128 characters"),
}


def fp64_peak():
    """Measured DFMA peak of this pool's B200 (tools/fp64_peak.cu, profiles/fp64_peaks_r01.json); MEASURED_PEAKS.json
    carries no fp64 figure.  Fallback: nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz."""
    path = os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["dfma_tflops"]), "measured (tools/fp64_peak.cu on this pool, profiles/fp64_peaks_r01.json)"
    except Exception:
        return 37.2, "nominal fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        top = sm[len(sm) // 2:] if sm else []          # samples under load = upper half
        return {"sm_mhz": (top[len(top) // 2] if top else None), "sm_max_mhz": (max(mx) if mx else None),
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_block(workload, nx, ny, seed=0):
    """One bounded block of the workload through the CPU port of the reference algorithm; returns (seconds, entries)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lsk_oracle as orc
    w = WORKLOADS[workload]
    # all host cores (torchrun exports OMP_NUM_THREADS=1, which would handicap the CPU arm)
    if torch.get_num_threads() < (os.cpu_count() or 1):
        torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(seed)
    og = orc.Group(w["group"], w["n"], w["order"])
    om = orc.Measure(w["measure"]["kind"], w["measure"]["dim"], w["measure"]["lengthscale"], w["measure"].get("nu"))
    x, y = og.rand(nx), og.rand(ny)
    og.chi(og.irreps[0][0], og.pairwise_embed(x[:1], y[:1]))      # table construction outside the timed region
    for irrep in og.irreps:
        og._table(irrep[0])
    t0 = time.perf_counter()
    with torch.no_grad():
        orc.eigensum_unnormalised(og, om, x, y)
    return time.perf_counter() - t0, nx * ny


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)       # torchrun exports OMP_NUM_THREADS=1; the CPU arm may use every host core
    block = {"so5": (16, 16), "so3": (256, 256), "su4": (96, 96)}[args.workload]
    if args.workload == "so5":                       # ~50 entries/s: keep the whole run near two minutes for any --steps
        side = max(6, min(16, int((5000.0 / max(args.steps, 1)) ** 0.5)))
        block = (side, side)
    for _ in range(min(args.warmup, 1)):
        oracle_block(args.workload, block[0] // 2, block[1] // 2)
    total_t, total_e = 0.0, 0
    for s in range(args.steps):
        t, e = oracle_block(args.workload, block[0], block[1], seed=s)
        total_t += t
        total_e += e
    value = total_e / total_t
    w = WORKLOADS[args.workload]
    cores = torch.get_num_threads()
    sample = "%dx%d block of the %dx%d covariance per step (the reference materialises O(N*M*n^2) temporaries and cannot run the full size)" % (
        block[0], block[1], args.size, args.size)
    line = {"impl": "reference", "metric": "kernel entries/sec", "value": value, "unit": "entries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "entry_irreps_per_s": value * w["order"],
            "config": {"workload": "%s, %dx%d random Haar points" % (w["name"], args.size, args.size), "irreps": w["order"]},
            "cpu_baseline": {"value": value, "unit": "entries/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from liestationarykernels_b200 import _lib, ops
    from liestationarykernels_b200.spaces import SO, SU
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel, spectral_weights
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w = WORKLOADS[args.workload]
    N = M = args.size
    space = (SO if w["group"] == "SO" else SU)(w["n"], order=w["order"])
    ms = w["measure"]
    measure = MaternSpectralMeasure(ms["dim"], ms["lengthscale"], ms["nu"]) if ms["kind"] == "matern" else SqExpSpectralMeasure(ms["dim"], ms["lengthscale"])
    torch.manual_seed(0)
    kern = EigenbasisSumKernel(measure, space)
    kern.eval()
    x_all, y = space.rand(N), space.rand(M)              # every rank draws the same points (same seed)
    lo, hi = rank * N // world, (rank + 1) * N // world   # row shard of this rank; y and the tables are replicated
    if args.shard_of > 1 and world == 1:                  # development aid: this GPU does the work of rank 0 of `shard_of` ranks
        lo, hi = 0, N // args.shard_of
    x = x_all[lo:hi].contiguous()
    out = torch.empty((hi - lo, M), dtype=torch.float64, device="cuda")

    def step():
        with torch.no_grad():
            kern(x, y, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    launches = _lib.launches() - launches0
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = ms_total.item() / args.steps
    job_entries = (hi - lo) * M if (args.shard_of > 1 and world == 1) else N * M
    value = job_entries / (ms_step * 1e-3)

    # ---- dominant kernel alone (roofline): the fused pairwise kernel on this rank's shard ----
    plan = space.plan
    ep = kern._epilogue(True)
    ea, eb = ops.embed(plan, x, "a"), ops.embed(plan, y, "b")
    for _ in range(3):
        ops.pairwise_poly(ea, eb, hi - lo, M, plan.kg_total, ep, out=out)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        ops.pairwise_poly(ea, eb, hi - lo, M, plan.kg_total, ep, out=out)
    k1.record()
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / args.steps

    # ---- end to end through the public API with host buffers ----
    xh, yh = x.cpu().pin_memory(), y.cpu().pin_memory()
    oh = torch.empty((hi - lo, M), dtype=torch.float64).pin_memory()
    chk = torch.empty(1, dtype=torch.float64).pin_memory()

    def e2e_step(full):
        with torch.no_grad():
            xd, yd = xh.to("cuda", non_blocking=True), yh.to("cuda", non_blocking=True)
            k = kern(xd, yd, out=out)
            if full:
                oh.copy_(k, non_blocking=True)
            else:
                chk.copy_(k.sum().reshape(1), non_blocking=True)

    res = {}
    for full in (True, False):
        e2e_step(full)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(2, min(args.steps, 5))
        t0.record()
        for _ in range(n_e2e):
            e2e_step(full)
        t1.record()
        barrier()
        t = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[full] = job_entries / (t.item() / n_e2e * 1e-3)

    clocks = sampler.stop() if rank == 0 else None      # sampled over the timed loops above (device-resident and e2e)
    if rank == 0:
        peak, peak_src = fp64_peak()
        flop = FLOP_PER_ENTRY[args.workload]
        achieved = (hi - lo) * M * flop / (kern_ms * 1e-3) / 1e12
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                hbm_peak = float(json.load(fh)["hbm_gbs"])
        except Exception:
            hbm_peak = 6650.0
        store_gbs = (hi - lo) * M * 8 / (kern_ms * 1e-3) / 1e9
        traffic = None
        try:    # dram__bytes_read + dram__bytes_write of one launch from the committed ncu capture (same workload, N=1)
            with open(os.path.join(ROOT, "profiles", "pairwise_traffic_r01.json")) as fh:
                tj = json.load(fh)
            if args.workload == "so5" and args.size == 16384 and world == 1:
                traffic = tj["traffic_bytes_per_launch"]
        except Exception:
            traffic = None
        # cpu baseline: bounded block of the same workload through the oracle port on the host cores (N=1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            blk = {"so5": (24, 24), "so3": (384, 384), "su4": (128, 128)}[args.workload]
            t, e = oracle_block(args.workload, blk[0], blk[1])
            cpu = {"value": e / t, "unit": "entries/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": "%dx%d block of the same workload, %.1f s of CPU work (oracle/lsk_oracle.py: bmm + LAPACK eigvals + per-monomial loop)" % (blk[0], blk[1], t)}
        line = {
            "metric": "kernel entries/sec", "value": value, "unit": "entries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "entry_irreps_per_s": value * w["order"],
            "config": {"workload": "%s, %dx%d random Haar points" % (w["name"], N, M), "irreps": w["order"],
                       "sharding": "x rows split over %d rank(s); y and coefficient tables replicated; no collective" % world,
                       "cache": "output (%.2f GB per step) is far larger than L2 and rewritten every step; the 2x%.1f MB inputs are L2-resident by nature" % ((hi - lo) * M * 8 / 1e9, N * w["n"] ** 2 * (16 if w["group"] == "SU" else 8) / 1e6),
                       "polynomial_form": {2: "monomial staircase (fast)", 3: "Weyl-orbit Chebyshev (robust)", 1: "Clenshaw", 8: "nested-Horner tape"}.get(ep.kind, str(ep.kind))},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": KERNEL_NAME[args.workload], "kernel_ms": kern_ms,
                         "flop_per_entry_algorithmic": flop, "peak_source": peak_src,
                         "executed": None if args.workload not in EXECUTED_FLOP_PER_ENTRY else {
                             "flop_per_entry": EXECUTED_FLOP_PER_ENTRY[args.workload],
                             "tflops": (hi - lo) * M * EXECUTED_FLOP_PER_ENTRY[args.workload] / (kern_ms * 1e-3) / 1e12,
                             "frac_of_peak": (hi - lo) * M * EXECUTED_FLOP_PER_ENTRY[args.workload] / (kern_ms * 1e-3) / 1e12 / peak,
                             "note": "what the implemented formulation issues per entry: 44 MAC on the DMMA path (tr g: 25 + 3 pad, Spin(5) rotor form: 16) + 99 Horner DFMA + 4 variable-map"},
                         "secondary_hbm": {"achieved": store_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": store_gbs / hbm_peak}},
            "e2e": {"value": res[True], "unit": "entries/s", "h2d_bytes_per_step": int(xh.numel() * xh.element_size() + yh.numel() * yh.element_size()),
                    "d2h_bytes_per_step": int(oh.numel() * 8), "note": "full covariance copied back to pinned host memory every step"},
            "e2e_device_consumer": {"value": res[False], "unit": "entries/s", "d2h_bytes_per_step": 8,
                                    "note": "covariance stays in HBM for the next GP stage; only its checksum is read back"},
            "gpu_launches": launches, "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--workload", default="so5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard-of", type=int, default=1, help="development: time the per-rank work of an N-rank run on one GPU (value is then NOT a whole-job number)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
