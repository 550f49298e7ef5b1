#!/usr/bin/env python
"""Benchmarks of the BASELINE.json configs.  Default = the headline: C2, SO(5) Matern-5/2 EigenbasisSumKernel covariance,
16384 x 16384, 100 irreps, fp64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4|c5] [--no-configs]

Workloads (a step = one pass of the public API over the whole synthetic input, inputs resident in HBM):
  c1  SO(3) heat EigenbasisSumKernel, 1024 x 1024, 20 irreps            step = kernel(x, y)
  c2  SO(5) Matern-5/2 EigenbasisSumKernel, 16384 x 16384, 100 irreps   step = kernel(x, y)                      [default]
  c3  SU(4) heat EigenbasisSumKernel, 16384 x 16384, 20 irreps          step = kernel(x, y)
  c4  V(5,3) RandomPhaseKernel, N = 32768, P = 2048, A = 30, 10 irreps  step = feature map + Gram Phi Phi^T  (+ prior sample timed beside it)
  c5  H^3 / SPD(3) RandomFourierFeatureKernel, N = 65536, m = 8192      step = feature map + Gram Psi Psi^T  (+ prior sample timed beside it)
N > 1 (torchrun, one rank per GPU, NCCL): c1-c3 shard the x rows (no collective); c4 / c5 shard the upper triangle of the Gram by
balanced row chunks (no collective, `distributed.sharded_symmetric_gram`) and the prior sample by phases / spherical functions with
ONE all-reduce of N doubles inside the timed region.  The total problem is fixed: strong scaling, as BASELINE.json quotes it.

Prints ONE JSON line (rank 0).  The default (c2) line also carries `configs`: the other four workloads measured in the same
process with fewer steps, each with its own roofline block.  `--impl reference` times the CPU restatement of the reference
algorithm (oracle/lsk_oracle.py; the reference is pure torch-CPU Python and cannot hold any of these sizes) on a bounded sample
of the same workload on the host cores.
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

ALIASES = {"so3": "c1", "so5": "c2", "su4": "c3"}
EIG = {
    "c1": dict(group="SO", n=3, order=20, size=1024, measure=dict(kind="heat", dim=3, lengthscale=1.0), flop=98.0,
               name="C1: SO(3) heat EigenbasisSumKernel, 20 irreps, fp64", kernel="lsk::pairwise_kernel<CHEB1>",
               ncu="profiles/c1_ncu_summary_r02.json"),
    "c2": dict(group="SO", n=5, order=100, size=16384, measure=dict(kind="matern", dim=10, lengthscale=1.0, nu=2.5), flop=756.0,
               name="C2: SO(5) Matern(nu=5/2) EigenbasisSumKernel, 100 irreps, fp64", kernel="lsk::so5::stair_kernel",
               ncu="profiles/so5_ncu_summary_r02.json"),
    "c3": dict(group="SU", n=4, order=20, size=16384, measure=dict(kind="heat", dim=15, lengthscale=1.0), flop=800.0,
               name="C3: SU(4) heat EigenbasisSumKernel, 20 irreps, complex128 characters", kernel="lsk::su4::tile_kernel",
               ncu="profiles/su4_ncu_summary_r02.json"),
}
C4 = dict(N=32768, P=2048, A=30, L=10, name="C4: Stiefel V(5,3) RandomPhaseKernel + RandomPhaseApproximation, Matern-5/2, 10 irreps, A=30")
C5 = dict(N=65536, m=8192, name="C5: H^3 and SPD(3) RandomFourierFeatureKernel + RandomFourierApproximation, Matern-5/2")


# ======================================================================================================================
# helpers
# ======================================================================================================================
def fp64_peak():
    """Measured DFMA peak of this pool's B200 (tools/fp64_peak.cu -> profiles/fp64_peaks_r01.json); MEASURED_PEAKS.json
    carries no fp64 figure.  Fallback: nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz."""
    try:
        with open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")) as fh:
            return float(json.load(fh)["dfma_tflops"]), "measured DFMA peak (tools/fp64_peak.cu on this pool, profiles/fp64_peaks_r01.json)"
    except Exception:
        return 37.2, "nominal fallback (148 SM x 64 DFMA/clk x 2 x 1.965 GHz)"


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6550.0, "fallback of B200_PROFILING.md"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],     # the recipe samples at 200 ms; 20 ms queries perturbed 0.3 ms steps on rank 0's GPU
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


class Ctx:
    """Process-group plumbing: one rank per GPU, barrier + synchronize around timed regions, max over ranks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.numa = self._bind_to_gpu_numa_node() if self.world > 1 else None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self._flush = None

    def _bind_to_gpu_numa_node(self):
        """Multi-rank runs: keep this rank's threads (and, by first touch, its pinned host buffers) on the NUMA node its GPU
        hangs off, so that the end-to-end copies do not cross the socket interconnect.  Best effort: returns the node or None."""
        try:
            prop = self.torch.cuda.get_device_properties(self.local)
            path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
            with open(path) as fh:
                node = int(fh.read().strip())
            if node < 0:
                return None
            with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
                cpus = set()
                for part in fh.read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = cpus & os.sched_getaffinity(0)
            if not allowed:
                return None
            os.sched_setaffinity(0, allowed)
            return node
        except Exception:
            return None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, value):
        t = self.torch.tensor([value], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def flush_l2(self):
        """write a 256 MB buffer (2x the 126 MB L2) so that the next step starts from HBM"""
        if self._flush is None:
            self._flush = self.torch.empty(256 << 20, dtype=self.torch.uint8, device="cuda")
        self._flush.fill_(1)

    def time_steps(self, fn, steps, warmup, flush=False):
        """ms per step, max over ranks.  Without `flush` the K steps run back to back between one pair of events (their
        working set exceeds L2); with it every step is bracketed by its own events and L2 is flushed in between."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        import gc
        gc_was_on = gc.isenabled()
        gc.disable()             # as `timeit` does: a collection pause inside a 0.3 ms step (8-GPU shards) is 1-2 ms on one rank
        try:
            return self._timed(fn, steps, flush)
        finally:
            if gc_was_on:
                gc.enable()

    def _timed(self, fn, steps, flush):
        torch = self.torch
        if not flush:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            self.barrier()
            total = e0.elapsed_time(e1)
        else:
            evs = []
            for _ in range(steps):
                self.flush_l2()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                evs.append((a, b))
            self.barrier()
            total = sum(a.elapsed_time(b) for a, b in evs)
        return self.max_over_ranks(total) / steps

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def make_measure(spec, variance=1.0):
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure, SqExpSpectralMeasure
    if spec["kind"] == "matern":
        return MaternSpectralMeasure(spec["dim"], spec["lengthscale"], spec["nu"], variance)
    return SqExpSpectralMeasure(spec["dim"], spec["lengthscale"], variance)


def executed_flop_per_entry(tag, plan, ep):
    """What the implemented formulation issues on the fp64 pipe per covariance entry (1 MAC = 2 flop): the bilinear forms on the
    DMMA path (4 multiply-adds per k-group, padding included) + the polynomial (one DFMA per coefficient / tape step) + the
    affine variable map.  C2 is the count read off the SASS of `so5::stair_kernel` (44 MAC + 102 DFMA + scale)."""
    if tag == "c2":
        return 2.0 * (44 + 103), "44 MAC on the DMMA path (tr g: 25 + 3 pad, Spin(5) rotor form: 16) + 99 Horner DFMA + 4 variable-map (SASS count)"
    if tag == "c3":
        return 2.0 * (84 + 20), "84 MAC on the DMMA path (21 k-groups) + 13 DFMA + 4 DMUL + 3 DADD per entry (SASS of su4::tile_kernel<4>: variable map, 27-step tape of which 10 are register moves, scale)"
    mac = 4.0 * plan.kg_total
    poly = float(ep.nterms) * (2.0 if tag == "c1" else 1.0)        # Clenshaw: two pipe operations per term
    vmap = float(ep.nvars * (1 + ep.nforms))
    return 2.0 * (mac + poly + vmap), "%d MAC on the DMMA path (k-groups x 4) + %d polynomial steps%s + %d variable-map" % (
        mac, ep.nterms, " x 2 (Clenshaw)" if tag == "c1" else "", vmap)


# ======================================================================================================================
# C1 / C2 / C3: EigenbasisSumKernel on a compact Lie group
# ======================================================================================================================
def run_eigensum(ctx, tag, steps, warmup, size=None, with_e2e=True, shard_of=1):
    torch = ctx.torch
    from liestationarykernels_b200 import _lib, ops
    from liestationarykernels_b200.spaces import SO, SU
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel
    w = EIG[tag]
    N = M = size or w["size"]
    space = (SO if w["group"] == "SO" else SU)(w["n"], order=w["order"])
    torch.manual_seed(0)
    kern = EigenbasisSumKernel(make_measure(w["measure"]), space)
    kern.eval()
    kern.assume_static = bool(os.environ.get("LSK_BENCH_ASSUME_STATIC"))     # development: skip the per-call hyper-parameter verification
    x_all, y = space.rand(N), space.rand(M)                      # every rank draws the same points (same seed)
    lo, hi = ctx.rank * N // ctx.world, (ctx.rank + 1) * N // ctx.world
    if shard_of > 1 and ctx.world == 1:                          # development aid: the work of rank 0 of `shard_of` ranks
        lo, hi = 0, N // shard_of
    x = x_all[lo:hi].contiguous()
    out = torch.empty((hi - lo, M), dtype=torch.float64, device="cuda")
    small = (hi - lo) * M * 8 < (256 << 20)                      # output smaller than 2x L2: flush between steps

    def step():
        with torch.no_grad():
            kern(x, y, out=out)

    launches0 = _lib.launches()
    ms_step = ctx.time_steps(step, steps, max(warmup, 3), flush=small)
    launches = (_lib.launches() - launches0) * steps // (steps + max(warmup, 3))
    job_entries = (hi - lo) * M if (shard_of > 1 and ctx.world == 1) else N * M
    value = job_entries / (ms_step * 1e-3)

    # ---- dominant kernel alone (roofline): the fused pairwise kernel on this rank's shard ----
    plan = space.plan
    ep = kern._epilogue(True)
    ea, eb = ops.embed(plan, x, "a"), ops.embed(plan, y, "b")
    kern_ms = ctx.time_steps(lambda: ops.pairwise_poly(ea, eb, hi - lo, M, plan.kg_total, ep, out=out), steps, 3, flush=small)

    res = {"value": value, "unit": "entries/s", "ms_per_step": ms_step, "entry_irreps_per_s": value * w["order"], "launches": launches}
    if with_e2e:
        xh, yh = x.cpu().pin_memory(), y.cpu().pin_memory()
        oh = torch.empty((hi - lo, M), dtype=torch.float64).pin_memory()
        chk = torch.empty(1, dtype=torch.float64).pin_memory()

        def e2e_step(full):
            with torch.no_grad():
                if full:       # host buffers in, host buffer out: row blocks computed while the previous block's copy is in flight
                    kern.evaluate_to_host(xh, yh, out=oh)
                else:
                    xd, yd = xh.to("cuda", non_blocking=True), yh.to("cuda", non_blocking=True)
                    chk.copy_(kern(xd, yd, out=out).sum().reshape(1), non_blocking=True)

        n_e2e = max(2, min(steps, 5))
        t_full = ctx.time_steps(lambda: e2e_step(True), n_e2e, 1)
        t_chk = ctx.time_steps(lambda: e2e_step(False), n_e2e, 1)
        res["e2e"] = {"value": job_entries / (t_full * 1e-3), "unit": "entries/s",
                      "h2d_bytes_per_step": int(xh.numel() * xh.element_size() + yh.numel() * yh.element_size()),
                      "d2h_bytes_per_step": int(oh.numel() * 8), "note": "EigenbasisSumKernel.evaluate_to_host: pinned host inputs, full covariance written to pinned host memory every step (row blocks, copy overlapped with compute)"}
        res["e2e_device_consumer"] = {"value": job_entries / (t_chk * 1e-3), "unit": "entries/s", "d2h_bytes_per_step": 8,
                                      "note": "covariance stays in HBM for the next GP stage; only its checksum is read back"}
    if ctx.rank == 0:
        peak, peak_src = fp64_peak()
        hpeak, hsrc = hbm_peak()
        entries = (hi - lo) * M
        ex_flop, ex_note = executed_flop_per_entry(tag, plan, ep)
        executed = entries * ex_flop / (kern_ms * 1e-3) / 1e12
        algorithmic = entries * w["flop"] / (kern_ms * 1e-3) / 1e12
        store = entries * 8 / (kern_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        if tag == "c2" and N == 16384 and ctx.world == 1:
            try:    # dram__bytes_read + dram__bytes_write of one launch: STATIC, from the committed ncu capture of this workload
                with open(os.path.join(ROOT, "profiles", "pairwise_traffic_r02.json")) as fh:
                    traffic = json.load(fh)["traffic_bytes_per_launch"]
                traffic_src = "static: profiles/pairwise_traffic_r02.json (ncu --set full capture of the same launch, not re-measured in this run)"
            except Exception:
                pass
        res["roofline"] = {
            "bound": "fp64", "unit": "TFLOP/s", "peak": peak, "peak_source": peak_src,
            "achieved": executed, "frac": executed / peak,
            "frac_definition": "executed fp64 flop (what the kernel issues on the shared DFMA/DMMA datapath) / measured peak; ncu's pipe-busy figure is in the file named by `ncu`",
            "flop_per_entry_executed": ex_flop, "executed_note": ex_note,
            "flop_per_entry_algorithmic": w["flop"], "achieved_vs_survey_count": algorithmic, "frac_vs_survey_count": algorithmic / peak,
            "survey_note": "SURVEY.md 8(d) counts the Chebyshev/orbit formulation; the implemented invariant-polynomial form needs fewer flop, so this ratio may exceed 1",
            "traffic": traffic, "traffic_source": traffic_src, "kernel": w["kernel"], "kernel_ms": kern_ms, "ncu": w["ncu"],
            "secondary_hbm": {"achieved": store, "peak": hpeak, "unit": "GB/s", "frac": store / hpeak, "peak_source": hsrc}}
        res["config"] = {
            "workload": "%s, %dx%d random Haar points" % (w["name"], N, M), "irreps": w["order"],
            "sharding": "x rows split over %d rank(s); y and coefficient tables replicated; no collective" % ctx.world,
            "cache": ("L2 flushed (256 MB write) between timed steps, each step timed by its own event pair: the %.1f MB output fits in L2" % ((hi - lo) * M * 8 / 1e6)) if small else
                     ("output (%.2f GB per step) is far larger than L2 and rewritten every step; the 2x%.1f MB inputs are L2-resident by nature" % ((hi - lo) * M * 8 / 1e9, N * w["n"] ** 2 * (16 if w["group"] == "SU" else 8) / 1e6)),
            "polynomial_form": {2: "monomial staircase (fast)", 3: "Weyl-orbit Chebyshev (robust)", 1: "Clenshaw", 8: "nested-Horner tape"}.get(ep.kind, str(ep.kind))}
    return res


# ======================================================================================================================
# C4 / C5: feature-map kernels (Gram on the fp64 tensor cores) and prior samplers
# ======================================================================================================================
def _gram_roofline(ctx, n, k_inner, gram_ms, symmetric, kernel_note):
    peak, peak_src = fp64_peak()
    full = 2.0 * n * n * k_inner
    executed = full / 2 if symmetric else full
    tf = executed / (gram_ms * 1e-3) / 1e12                       # whole-job flop / max-over-ranks time = aggregate rate
    return {"bound": "fp64 tensor (DMMA)", "unit": "TFLOP/s", "peak": peak * ctx.world, "peak_source": peak_src + (" x %d GPUs" % ctx.world if ctx.world > 1 else ""),
            "achieved": tf, "frac": tf / (peak * ctx.world),
            "frac_definition": "flop of the tiles actually computed (upper triangle only in symmetric mode) / (time x peak)",
            "flop_algorithmic": full, "achieved_vs_survey_count": full / (gram_ms * 1e-3) / 1e12,
            "frac_vs_survey_count": full / (gram_ms * 1e-3) / 1e12 / (peak * ctx.world),
            "survey_note": "SURVEY.md 8(d) counts 2 N M K for the full matrix; Phi Phi^T is symmetric and only the upper triangle of tiles is computed (and mirrored at N=1)",
            "traffic": None, "kernel": "lsk::pairwise_kernel<LINEAR>", "kernel_ms": gram_ms, "kernel_note": kernel_note,
            "ncu": "profiles/gram_ncu_summary_r02.json"}


def _sharded_gram_buffers(ctx, n):
    from liestationarykernels_b200 import distributed as D
    torch = ctx.torch
    if ctx.world == 1:
        return [torch.empty((n, n), dtype=torch.float64, device="cuda")]
    return [torch.empty((hi - lo, n - lo), dtype=torch.float64, device="cuda") for lo, hi in D.triangular_chunks(n, ctx.world, ctx.rank)]


def run_c4(ctx, steps, warmup, with_e2e=True):
    torch = ctx.torch
    from liestationarykernels_b200 import _lib, distributed as D
    from liestationarykernels_b200.prior_approximation import RandomPhaseApproximation
    from liestationarykernels_b200.spaces import Stiefel
    from liestationarykernels_b200.spectral_kernel import EigenbasisSumKernel, RandomPhaseKernel
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    N, P, A, L = C4["N"], C4["P"], C4["A"], C4["L"]
    torch.manual_seed(0)                                           # every rank draws the same points, phases and weights
    space = Stiefel(5, 3, order=L, average_order=A)
    meas = MaternSpectralMeasure(9, 1.0, 2.5)
    kern = RandomPhaseKernel(meas, space, phase_order=P).eval()
    esum = EigenbasisSumKernel(meas, space).eval()
    sampler = D.PhaseShardedSampler(RandomPhaseApproximation(esum, phase_order=P))
    x = space.rand(N).contiguous()
    bufs = _sharded_gram_buffers(ctx, N)
    scale = float(torch.abs(meas.variance[0]).item() / float(kern.normalizer))
    state = {}

    def features():
        with torch.no_grad():
            state["phi"] = kern._features(x, kern._row_scales(), panel=True)

    def gram():
        with torch.no_grad():
            D.sharded_symmetric_gram(state["phi"], scale, buffers=bufs)

    def step():
        features()
        gram()

    def sample():
        with torch.no_grad():
            state["f"] = sampler(x)

    launches0 = _lib.launches()
    ms_step = ctx.time_steps(step, steps, warmup)
    launches = (_lib.launches() - launches0) * steps // (steps + warmup)
    feat_ms = ctx.time_steps(features, max(steps, 2), 1)
    gram_ms = ctx.time_steps(gram, steps, 1)
    samp_ms = ctx.time_steps(sample, max(steps, 10), 3)         # ms-scale call: enough repetitions to average out NCCL's lazy set-up
    res = {"value": N * N / (ms_step * 1e-3), "unit": "entries/s", "ms_per_step": ms_step, "launches": launches,
           "features_ms": feat_ms, "feature_pairs_per_s": N * P / (feat_ms * 1e-3), "gram_ms": gram_ms,
           "sampler": {"ms": samp_ms, "points_per_s": N / (samp_ms * 1e-3),
                       "what": "RandomPhaseApproximation(x): %d phases per rank, fused feature kernel + reduction%s" % (
                           P // ctx.world, ", NCCL all-reduce of %d doubles inside the timed region" % N if ctx.world > 1 else "")}}
    if with_e2e:
        xh = x.cpu().pin_memory()
        back = torch.empty(N + 1, dtype=torch.float64).pin_memory()

        def e2e_step():
            with torch.no_grad():
                xd = xh.to("cuda", non_blocking=True)
                phi = kern._features(xd, kern._row_scales(), panel=True)
                blocks = D.sharded_symmetric_gram(phi, scale, buffers=bufs)
                f = sampler(xd)
                back[:N].copy_(f, non_blocking=True)
                back[N:].copy_(sum(b.sum() for _, _, b in blocks).reshape(1), non_blocking=True)

        t = ctx.time_steps(e2e_step, max(2, min(steps, 3)), 1)
        res["e2e"] = {"value": N * N / (t * 1e-3), "unit": "entries/s", "h2d_bytes_per_step": int(xh.numel() * 8), "d2h_bytes_per_step": int(back.numel() * 8),
                      "note": "host points in; feature map + Gram + prior sample; the sample f[N] and a checksum of the Gram come back (the %.1f GB Gram stays in HBM for the GP step)" % (N * N * 8 / 1e9)}
    if ctx.rank == 0:
        res["roofline"] = _gram_roofline(ctx, N, L * P, gram_ms, True, "K = L*P = %d; %s" % (
            L * P, "one symmetric-mode launch (triangular tile schedule, mirrored)" if ctx.world == 1 else
            "per rank: 2 row chunks x (symmetric diagonal block + general rectangle), upper triangle only"))
        res["roofline_features"] = {"kernel": "lsk::homog5_stair_kernel", "kernel_ms": feat_ms, "ncu": "profiles/homog5_ncu_summary_r02.json",
                                    "flop_algorithmic": 3.8e11, "achieved_vs_survey_count": 3.8e11 / (feat_ms * 1e-3) / 1e12,
                                    "note": "SURVEY.md 8(d): ~5.6 kflop per (point, phase); every rank computes the whole map"}
        res["config"] = {"workload": "%s; N=%d points, P=%d phases" % (C4["name"], N, P),
                         "sharding": ("single GPU: one symmetric-mode Gram launch" if ctx.world == 1 else
                                      "upper triangle of the Gram in 2x%d balanced row chunks, feature map recomputed on every rank, no collective; sampler phase-sharded + all-reduce" % ctx.world),
                         "cache": "feature map 5.4 GB and Gram 8.6 GB per step: far larger than L2"}
    del bufs, state
    torch.cuda.empty_cache()
    return res


def run_c5(ctx, steps, warmup, which=("h3", "spd3"), with_e2e=True):
    torch = ctx.torch
    from liestationarykernels_b200 import _lib, distributed as D
    from liestationarykernels_b200.prior_approximation import RandomFourierApproximation
    from liestationarykernels_b200.spaces import HyperbolicSpace, SymmetricPositiveDefiniteMatrices
    from liestationarykernels_b200.spectral_kernel import RandomFourierFeatureKernel, _rff_scale
    from liestationarykernels_b200.spectral_measure import MaternSpectralMeasure
    N, m = C5["N"], C5["m"]
    out = {}
    for name in which:
        torch.manual_seed(0)
        if name == "h3":
            space, meas = HyperbolicSpace(3, order=m), MaternSpectralMeasure(3, 1.0, 2.5)
        else:
            space, meas = SymmetricPositiveDefiniteMatrices(3, order=m), MaternSpectralMeasure(6, 1.0, 2.5)
        kern = RandomFourierFeatureKernel(meas, space).eval()
        sampler = D.FeatureShardedSampler(RandomFourierApproximation(kern))
        x = space.rand(N)
        bufs = _sharded_gram_buffers(ctx, N)
        s = _rff_scale(kern, True)
        state = {}

        def features():
            with torch.no_grad():
                state["psi"] = space._features(space.to_group(x), space.lb_eigenspaces, s, "panel")

        def gram():
            with torch.no_grad():
                D.sharded_symmetric_gram(state["psi"], 1.0, buffers=bufs)

        def step():
            features()
            gram()

        def sample():
            with torch.no_grad():
                state["f"] = sampler(x)

        launches0 = _lib.launches()
        ms_step = ctx.time_steps(step, steps, warmup)
        launches = (_lib.launches() - launches0) * steps // (steps + warmup)
        feat_ms = ctx.time_steps(features, max(steps, 2), 1)
        gram_ms = ctx.time_steps(gram, steps, 1)
        samp_ms = ctx.time_steps(sample, max(steps, 10), 3)         # ms-scale call: enough repetitions to average out NCCL's lazy set-up
        r = {"value": N * N / (ms_step * 1e-3), "unit": "entries/s", "ms_per_step": ms_step, "launches": launches,
             "features_ms": feat_ms, "features_per_s": N * m / (feat_ms * 1e-3), "gram_ms": gram_ms,
             "sampler": {"ms": samp_ms, "points_per_s": N / (samp_ms * 1e-3),
                         "what": "RandomFourierApproximation(x): %d spherical functions per rank, feature map fused away%s" % (
                             m // ctx.world, ", NCCL all-reduce of %d doubles inside the timed region" % N if ctx.world > 1 else "")}}
        if with_e2e:
            xh = x.cpu().pin_memory()
            back = torch.empty(N + 1, dtype=torch.float64).pin_memory()

            def e2e_step():
                with torch.no_grad():
                    xd = xh.to("cuda", non_blocking=True)
                    psi = space._features(space.to_group(xd), space.lb_eigenspaces, s, "panel")
                    blocks = D.sharded_symmetric_gram(psi, 1.0, buffers=bufs)
                    f = sampler(xd)
                    back[:N].copy_(f, non_blocking=True)
                    back[N:].copy_(sum(b.sum() for _, _, b in blocks).reshape(1), non_blocking=True)

            t = ctx.time_steps(e2e_step, 2, 1)
            r["e2e"] = {"value": N * N / (t * 1e-3), "unit": "entries/s", "h2d_bytes_per_step": int(xh.numel() * 8), "d2h_bytes_per_step": int(back.numel() * 8),
                        "note": "host points in; feature map + Gram + prior sample; the sample f[N] and a checksum of the Gram come back (the %.0f GB Gram stays in HBM)" % (N * N * 8 / 1e9)}
        if ctx.rank == 0:
            r["roofline"] = _gram_roofline(ctx, N, 2 * m, gram_ms, True, "K = 2m = %d; %s" % (
                2 * m, "one symmetric-mode launch, 34 GB output" if ctx.world == 1 else "per rank: 2 row chunks x (symmetric diagonal block + general rectangle)"))
            r["roofline_features"] = {"kernel": "lsk::hyp_features_kernel" if name == "h3" else "lsk::spd_features_kernel", "kernel_ms": feat_ms,
                                      "features_per_s": N * m / (feat_ms * 1e-3)}
        out[name] = r
        del bufs, state
        torch.cuda.empty_cache()
    first = out[which[0]]
    res = dict(first)
    if len(which) > 1:
        res["spd3"] = out["spd3"]
    if ctx.rank == 0:
        res["config"] = {"workload": "%s; N=%d points, m=%d spherical functions (headline value: %s; both spaces reported)" % (C5["name"], N, m, which[0]),
                         "sharding": ("single GPU: one symmetric-mode Gram launch" if ctx.world == 1 else
                                      "upper triangle of the Gram in 2x%d balanced row chunks, feature map recomputed on every rank, no collective; sampler feature-sharded + all-reduce" % ctx.world),
                         "cache": "feature map 8.6 GB and Gram 34 GB per step: far larger than L2"}
    return res


# ======================================================================================================================
# CPU arm: the oracle port on bounded samples (reference algorithm: materialised pair grid, LAPACK eigvals, per-monomial loop)
# ======================================================================================================================
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lsk_oracle as orc
    return orc


def _all_threads():
    import torch
    if torch.get_num_threads() < (os.cpu_count() or 1):       # torchrun exports OMP_NUM_THREADS=1, which would handicap the CPU arm
        torch.set_num_threads(os.cpu_count() or 1)
    return torch.get_num_threads()


def oracle_eigensum_block(tag, nx, ny, seed=0, device=None):
    """One bounded block of an EigenbasisSumKernel workload through the port; returns (seconds, entries).  With `device='cuda'`
    the same port runs on CUDA tensors (ATen kernels on the GPU: what the reference's stock code path does on a GPU box, its
    `device = 'cuda' if available` at spectral_kernel.py:11-12)."""
    import torch
    orc = _oracle()
    w = EIG[tag]
    torch.manual_seed(seed)
    og = orc.Group(w["group"], w["n"], w["order"])
    ms = w["measure"]
    om = orc.Measure(ms["kind"], ms["dim"], ms["lengthscale"], ms.get("nu"))
    x, y = og.rand(nx), og.rand(ny)
    for irrep in og.irreps:                                     # table construction outside the timed region
        og._table(irrep[0])
    if device is None:
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.eigensum_unnormalised(og, om, x, y)
        return time.perf_counter() - t0, nx * ny
    prev = torch.get_default_device()
    try:
        torch.set_default_device(device)
        og.id = og.id.to(device)
        og._tables = {k: (c.to(device), m.to(device)) for k, (c, m) in og._tables.items()}
        om.lengthscale, om.variance = om.lengthscale.to(device), om.variance.to(device)
        if getattr(om, "nu", None) is not None:
            om.nu = om.nu.to(device)
        x, y = x.to(device), y.to(device)
        with torch.no_grad():
            orc.eigensum_unnormalised(og, om, x[:2], y[:2])    # warm-up (library handles, MAGMA init)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            orc.eigensum_unnormalised(og, om, x, y)
            torch.cuda.synchronize()
        return time.perf_counter() - t0, nx * ny
    finally:
        torch.set_default_device(prev)


def oracle_gram_workload(tag, n_points, gemm_rows):
    """Bounded sample of C4 / C5 through the port: the feature map of `n_points` points against ALL phases / spherical functions
    plus a `gemm_rows`^2 block of the Gram at the full inner dimension (torch CPU DGEMM).  Returns (seconds of CPU work,
    extrapolated whole-job entries/s, description): T_full = N t_point + N^2 t_entry."""
    import torch
    orc = _oracle()
    torch.manual_seed(0)
    if tag == "c4":
        N, K = C4["N"], C4["L"] * C4["P"]
        sp = orc.Homogeneous("stiefel", 5, 3, order=C4["L"], average_order=C4["A"])
        om = orc.Measure("matern", 9, 1.0, 2.5)
        x, phases = sp.rand(n_points), sp.rand(C4["P"])
        for irrep in sp.g.irreps:
            sp.g._table(irrep[0])
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.random_phase_embedding(sp, om, phases, x, torch.sqrt)
        t_feat = time.perf_counter() - t0
    else:
        N, K = C5["N"], 2 * C5["m"]
        sp = orc.Hyperbolic(3, order=C5["m"])
        om = orc.Measure("matern", 3, 1.0, 2.5)
        x = sp.rand(n_points)
        lmd, shift = sp.generate(om)
        t0 = time.perf_counter()
        with torch.no_grad():
            sp.features(sp.to_group(x), lmd, shift)
        t_feat = time.perf_counter() - t0
    a = torch.randn(gemm_rows, K, dtype=torch.float64)
    t0 = time.perf_counter()
    a @ a.t()
    t_gemm = time.perf_counter() - t0
    t_point, t_entry = t_feat / n_points, t_gemm / gemm_rows ** 2
    total = N * t_point + N * N * t_entry
    desc = "feature map of %d points x all %d columns (%.1f s) + a %dx%d block of the Gram at K=%d (%.1f s, torch CPU DGEMM %.2f TFLOP/s); extrapolated T = N t_point + N^2 t_entry = %.0f s" % (
        n_points, K, t_feat, gemm_rows, gemm_rows, K, t_gemm, 2.0 * gemm_rows ** 2 * K / t_gemm / 1e12, total)
    return t_feat + t_gemm, N * N / total, desc


def cpu_baseline(tag, budget="bench"):
    import torch
    cores = _all_threads()
    if tag in EIG:
        blk = {"c2": (24, 24), "c1": (384, 384), "c3": (128, 128)}[tag]
        t, e = oracle_eigensum_block(tag, blk[0], blk[1])
        return {"value": e / t, "unit": "entries/s", "cores": cores, "kind": "port",
                "sample": "%dx%d block of the same workload, %.1f s of CPU work (oracle/lsk_oracle.py: bmm + LAPACK eigvals + per-monomial loop)" % (blk[0], blk[1], t)}
    t, v, desc = oracle_gram_workload(tag, 6 if tag == "c4" else 4096, 2048 if tag == "c4" else 6144)
    return {"value": v, "unit": "entries/s", "cores": cores, "kind": "port", "sample": desc}


def gpu_port_baseline(tag):
    """The port with CUDA tensors (ATen on the GPU), bounded block: the reference's own stock path on a GPU box."""
    blk = {"c2": (96, 96), "c1": (1024, 1024), "c3": (192, 192)}[tag]
    try:
        t, e = oracle_eigensum_block(tag, blk[0], blk[1], device="cuda")
        return {"value": e / t, "unit": "entries/s", "kind": "port on CUDA tensors (ATen kernels; torch.linalg.eigvals + per-monomial loop on the GPU)",
                "sample": "%dx%d block, %.2f s" % (blk[0], blk[1], t)}
    except Exception as exc:       # e.g. no GPU eigvals backend in the torch build
        return {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:200])}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = _all_threads()
    tag = args.workload
    total_t, values = 0.0, []
    if tag in EIG:
        w = EIG[tag]
        block = {"c2": (16, 16), "c1": (256, 256), "c3": (96, 96)}[tag]
        if tag == "c2":                       # ~50 entries/s: keep the whole run near two minutes for any --steps
            side = max(6, min(16, int((5000.0 / max(args.steps, 1)) ** 0.5)))
            block = (side, side)
        for _ in range(min(args.warmup, 1)):
            oracle_eigensum_block(tag, block[0] // 2, block[1] // 2)
        total_e = 0
        for s in range(args.steps):
            t, e = oracle_eigensum_block(tag, block[0], block[1], seed=s)
            total_t += t
            total_e += e
        value = total_e / total_t
        size = args.size or w["size"]
        sample = "%dx%d block of the %dx%d covariance per step (the reference materialises O(N*M*n^2) temporaries and cannot run the full size)" % (
            block[0], block[1], size, size)
        config = {"workload": "%s, %dx%d random Haar points" % (w["name"], size, size), "irreps": w["order"]}
        extra = {"entry_irreps_per_s": value * w["order"]}
    else:
        steps = max(1, min(args.steps, 4))
        for s in range(steps):
            t, v, sample = oracle_gram_workload(tag, 3 if tag == "c4" else 2048, 1536 if tag == "c4" else 4096)
            total_t += t
            values.append(v)
        value = sum(values) / len(values)
        config = {"workload": (C4 if tag == "c4" else C5)["name"]}
        extra = {"steps_run": steps}
    line = {"impl": "reference", "metric": "kernel entries/sec", "value": value, "unit": "entries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": value, "unit": "entries/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line.update(extra)
    print(json.dumps(line), flush=True)


# ======================================================================================================================
def run_workload(ctx, tag, steps, warmup, args, primary):
    if tag in EIG:
        return run_eigensum(ctx, tag, steps, warmup, size=args.size if primary else None, with_e2e=True, shard_of=args.shard_of if primary else 1)
    if tag == "c4":
        return run_c4(ctx, steps, warmup)
    return run_c5(ctx, steps, warmup)


def run_ours(args):
    ctx = Ctx()
    torch = ctx.torch
    tag = args.workload
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0:
        sampler.start()
    steps, warmup = args.steps, max(args.warmup, 3)
    if tag in ("c4", "c5"):
        steps, warmup = max(1, min(args.steps, 5 if tag == "c4" else 3)), 3
    res = run_workload(ctx, tag, steps, warmup, args, primary=True)
    clocks = sampler.stop() if ctx.rank == 0 else None         # sampled over the timed loops of the primary workload
    configs = None
    if tag == "c2" and not args.no_configs and args.shard_of == 1 and args.size in (None, 16384):
        configs = {}
        for other in ("c1", "c3", "c4", "c5"):
            try:
                k = {"c1": min(args.steps, 20), "c3": min(args.steps, 10), "c4": 3, "c5": 2}[other]
                r = run_workload(ctx, other, k, 3, args, primary=False)       # 3 warm-up steps: the caching allocator has both feature-map buffers by then
                r["steps"] = k
                configs[other] = r
            except Exception as exc:                             # a secondary config must not take the headline line down
                configs[other] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
                torch.cuda.empty_cache()
    if ctx.rank == 0:
        line = {"metric": "kernel entries/sec", "value": res.pop("value"), "unit": res.pop("unit"), "n_gpus": ctx.world, "steps": steps,
                "warmup": warmup, "ms_per_step": res.pop("ms_per_step"), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": res.pop("config"), "roofline": res.pop("roofline"),
                "e2e": res.pop("e2e"), "gpu_launches": res.pop("launches"), "clocks": clocks}
        line.update(res)
        if ctx.world > 1:
            line["host_affinity"] = {"rank0_numa_node": ctx.numa, "note": "each rank pinned to the CPUs of its GPU's NUMA node (best effort; null = not available)"}
        if ctx.world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(tag)
            if tag in EIG:
                line["gpu_port_baseline"] = gpu_port_baseline(tag)
        if configs is not None:
            line["configs"] = configs
        print(json.dumps(line), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=None, help="c1-c3: override the number of points per side")
    ap.add_argument("--workload", default="c2", choices=sorted(list(EIG) + ["c4", "c5"] + list(ALIASES)))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="c2 only: skip the `configs` block (the other four workloads)")
    ap.add_argument("--shard-of", type=int, default=1, help="development: time the per-rank work of an N-rank run on one GPU (value is then NOT a whole-job number)")
    args = ap.parse_args()
    args.workload = ALIASES.get(args.workload, args.workload)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
