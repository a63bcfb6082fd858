#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its config 2: omnidirectional experimental variogram
(semivariance + covariance + pair counts, the ExperimentalVariogram default) of 1M synthetic points,
80 lags (step 500, max_range 40500), pair tiles sharded across the ranks.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU restatement of the reference on the host cores

One step = one full variogram pass over the field with the points already resident in HBM
(spatial sort + pair kernel + reduction [+ one NCCL all-reduce of the per-lag sums when N > 1]).
`value` = unordered point pairs of the input, N(N-1)/2, per second (the reference evaluates every
pair; this implementation skips tile pairs farther apart than the last lag, `pairs_evaluated` says how
many distances were really computed, and only those count as achieved FLOPs in `roofline`).
`e2e` = the same pass through the host-buffer entry point (pinned host points in, per-lag sums out).
The kriging half of the metric (kriged points/s) is measured after the timed region and reported
under `extra` (it does not enter `value`).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS = 1_000_000
EXTENT = 100_000.0
STEP, MAX_RANGE = 500.0, 40500.0          # np.arange(500, 40500, 500) -> 80 lags
FLOP_PER_PAIR = 13                        # semivariance + covariance per evaluated pair
METRIC = "variogram_point_pairs_per_s"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--no-extra", action="store_true", help="skip the kriging measurements")
    return ap.parse_args()


def workload_config(n, world):
    return {"workload": f"omnidirectional experimental variogram (semivariance+covariance+counts), {n} synthetic "
                        f"DEM-like points on {EXTENT/1000:.0f} km, 80 lags (step 500, max_range 40500) — BASELINE config 2",
            "points": n, "lags": 80, "parallelism": f"pair-tile shards x{world} + 1 all-reduce" if world > 1 else "single GPU",
            "l2": "256 MB buffer written between timed steps (inputs are 24 MB, smaller than L2)"}


# ---------------------------------------------------------------------------------------------------
# CPU legs (oracle port, test infrastructure used as the measured CPU baseline)
# ---------------------------------------------------------------------------------------------------
def cpu_variogram_sample(pts, lags, target_s=12.0):
    """Time the C restatement of the reference (all host threads) on the first R rows of the pair
    triangle of the SAME field; R is sized from a short probe so the sample takes ~target_s."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.set_threads(os.cpu_count() or 1)         # torchrun exports OMP_NUM_THREADS=1
    n = len(pts)
    probe_rows = 64
    t0 = time.perf_counter()
    oracle.variogram_sums(pts, lags, rows=(0, probe_rows))
    dt = time.perf_counter() - t0
    rate = probe_rows * (n - 1 - probe_rows / 2) / max(dt, 1e-6)
    rows = int(min(n - 1, max(probe_rows, target_s * rate / n)))
    t0 = time.perf_counter()
    oracle.variogram_sums(pts, lags, rows=(0, rows))
    dt = time.perf_counter() - t0
    pairs = rows * (n - 1) - rows * (rows - 1) // 2
    return pairs / dt, oracle.num_threads(), f"rows [0, {rows}) of the {n}-point pair triangle = {pairs:.3e} pairs in {dt:.1f} s"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pyinterpolate_b200.synthetic import dem_like_field
    pts = dem_like_field(args.points, extent=EXTENT, seed=2)
    lags = np.arange(STEP, MAX_RANGE, STEP)
    rates, sample, cores = [], "", 1
    per_step = max(2.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        r, cores, sample = cpu_variogram_sample(pts, lags, target_s=per_step)
        if i >= args.warmup:
            rates.append(r)
    value = float(np.mean(rates))
    n = args.points
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (n * (n - 1) / 2) / value,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample + " per step; ms_per_step is the extrapolation to the full triangle"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.gpu_index = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)], stdout=self.file,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for row in self.file.read().splitlines():
            f = [c.strip() for c in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.file.name)
        if sm:
            loaded = [c for c, p in zip(sm, power) if p > 0.5 * max(power)] or sm
            out = {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                   "power_w_max": max(power), "samples": len(sm)}
        return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            pass
    return {}


def profile_traffic(kernel):
    """dram bytes per launch from the committed ncu --set full summary, if there is one."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(kernel)
        except Exception:
            pass
    return None


def run_b200(args):
    import torch
    import torch.distributed as dist
    from pyinterpolate_b200 import _lib
    from pyinterpolate_b200.distributed import allreduce_variogram_sums
    from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything native libraries print there (NCCL's version banner)
    # goes to stderr while the run is in progress
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback (use --impl reference for the CPU leg)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.points
    pts = dem_like_field(n, extent=EXTENT, seed=2)
    lags = np.arange(STEP, MAX_RANGE, STEP)
    pinned = torch.empty((n, 3), dtype=torch.float64).pin_memory()
    pinned.copy_(torch.from_numpy(pts))
    d_pts = pinned.to(dev, non_blocking=True)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    total_pairs = n * (n - 1) / 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(device_resident=True):
        """One variogram pass; returns the (all-reduced) sums and this rank's pair-kernel ms."""
        flush.fill_(1)                                                  # evict the points from L2
        if device_resident:
            sums = _lib.variogram_device(d_pts, lags, shard=(rank, world), want_pairs=True)
        else:
            h = _lib.variogram_host(pinned.numpy(), lags, shard=(rank, world))
            sums = {k: (torch.from_numpy(np.ascontiguousarray(v)).to(dev) if isinstance(v, np.ndarray) else v)
                    for k, v in h.items()}
        if world > 1:
            sums = allreduce_variogram_sums(sums, device=dev)
        return sums

    fp64_peak = _lib.fp64_peak_tflops()

    # ---- device-resident timing: K steps between barriers, CUDA events, max over ranks -------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, evaluated = [], 0
    ev0.record()
    for _ in range(args.steps):
        sums = step()
        ms, launches = _lib.last_kernel_ms(_lib.KERNEL_PAIR)
        kernel_ms.append(ms)
        evaluated = sums["pairs_evaluated"]
    ev1.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms, float(np.mean(kernel_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, pair_kernel_ms = float(t[0]), float(t[1])
    ms_per_step = elapsed_ms / args.steps
    value = total_pairs / (ms_per_step * 1e-3)

    # ---- end to end: pinned host buffers through the host entry point ------------------------------
    for _ in range(2):
        step(device_resident=False)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        step(device_resident=False)
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = total_pairs / float(e2e_s[0])
    out_bytes = 80 * (3 * 8 + 8) + 8

    # ---- roofline of the dominant kernel (pair kernel) --------------------------------------------
    evaluated_per_rank = evaluated / world                              # evaluated is the all-reduced total
    achieved = FLOP_PER_PAIR * evaluated_per_rank / (pair_kernel_ms * 1e-3) / 1e12
    roofline = {"bound": "fp64", "kernel": "pair_ring_kernel<16,512>", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak,
                "peak_source": "DFMA-saturation probe run in this process (pib_fp64_peak_tflops); MEASURED_PEAKS.json "
                               "has no FP64 figure (its HBM %.0f GB/s / bf16 figures do not bound this kernel)"
                               % measured_peaks().get("hbm_gbs", float("nan")),
                "flop_per_pair": FLOP_PER_PAIR, "pairs_evaluated_per_launch": evaluated_per_rank,
                "kernel_ms_per_launch": pair_kernel_ms, "kernel_share_of_step": pair_kernel_ms / ms_per_step,
                "traffic": profile_traffic("pair_ring_kernel")}

    extra = {"pairs_evaluated": evaluated, "evaluated_pairs_per_s": evaluated / (ms_per_step * 1e-3),
             "fp64_peak_tflops_probe": fp64_peak}
    cpu_baseline = None
    if rank == 0:
        v, cores, sample = cpu_variogram_sample(pts, lags)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    # ---- BASELINE config 3 (directional, 4 ellipse directions, tolerance 0.2, 200k points), rank 0 only ----
    if not args.no_extra and rank == 0:
        from pyinterpolate_b200 import core
        p3 = torch.from_numpy(dem_like_field(200_000, extent=EXTENT, seed=3)).to(dev)
        W = np.stack([core.define_whitening_matrix(a, 0.2) for a in (90, 0, 45, 135)]).reshape(-1, 4)
        lower = core.ellipse_lower_limits(lags)
        r3 = _lib.variogram_device(p3, lags, lower=lower, W=W, want_pairs=True)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            r3 = _lib.variogram_device(p3, lags, lower=lower, W=W, want_pairs=True)
        c1.record()
        torch.cuda.synchronize()
        ms3 = c0.elapsed_time(c1) / 3
        extra["directional_config3"] = {"points": 200_000, "directions": 4, "tolerance": 0.2, "ms": ms3,
                                        "direction_pairs_per_s": 4 * (200_000 * 199_999 / 2) / (ms3 * 1e-3),
                                        "pairs_evaluated": r3["pairs_evaluated"]}

    # ---- kriging half of the metric (outside the timed region; every rank kriges its own locations) ----
    if not args.no_extra:
        var = float(np.var(pts[:, 2]))
        for name, k, m, model, params in (
                ("ok_k32", 32, 1_000_000, "linear", [0.0, var, 8000.0]),
                ("ok_k64", 64, 1_000_000, "spherical", [1.0, var, 20000.0]),
                ("sk_k64", 64, 1_000_000, "spherical", [1.0, var, 20000.0])):
            unk = torch.from_numpy(uniform_locations(m, extent=EXTENT, seed=5 + rank)).to(dev)
            mode = _lib.KRIGE_SIMPLE if name.startswith("sk") else _lib.KRIGE_ORDINARY
            call = lambda: _lib.krige_device(d_pts, unk, _lib.MODEL_IDS[model], np.array([params]), mode, k,
                                             process_mean=float(pts[:, 2].mean()))
            call()
            barrier()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            for _ in range(3):
                call()
            k1.record()
            barrier()
            tk = torch.tensor([k0.elapsed_time(k1) / 3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tk, op=dist.ReduceOp.MAX)
            knn_ms, _ = _lib.last_kernel_ms(_lib.KERNEL_KNN)
            solve_ms, _ = _lib.last_kernel_ms(_lib.KERNEL_SOLVE)
            dim = k + 1 if mode == _lib.KRIGE_ORDINARY else k
            solve_flop = (2 * dim ** 3 / 3 + 2 * dim ** 2) * m
            extra[f"kriged_points_per_s_{name}"] = world * m / (float(tk) * 1e-3)
            extra[f"kriging_{name}"] = {"known": n, "unknown_per_gpu": m, "model": model, "ms": float(tk),
                                        "knn_kernel_ms": knn_ms, "solve_kernel_ms": solve_ms,
                                        "solve_fp64_tflops": solve_flop / (solve_ms * 1e-3) / 1e12,
                                        "solve_frac_of_fp64_peak": solve_flop / (solve_ms * 1e-3) / 1e12 / fp64_peak,
                                        "scaling": "weak"}

    # kriging end to end through the public Python call (host arrays in, host array out), rank 0 only
    if rank == 0 and not args.no_extra:
        import pyinterpolate_b200 as pib
        var = float(np.var(pts[:, 2]))
        unk_h = uniform_locations(1_000_000, extent=EXTENT, seed=5)
        mdl = {"variogram_model_type": "linear", "nugget": 0.0, "sill": var, "rang": 8000.0}
        pib.ordinary_kriging(mdl, pts, unk_h[:1000], no_neighbors=32)
        t0 = time.perf_counter()
        res = pib.ordinary_kriging(mdl, pts, unk_h, no_neighbors=32)
        dt = time.perf_counter() - t0
        extra["kriged_points_per_s_ok_k32_public_api_host_buffers"] = {
            "value": len(unk_h) / dt, "h2d_bytes": int(pts.nbytes + unk_h.nbytes), "d2h_bytes": int(res.nbytes + len(unk_h))}

    # CPU restatement of the kriging path on the host cores (bounded sample), for context next to the GPU numbers
    if rank == 0 and not args.no_extra:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle
        oracle.set_threads(os.cpu_count() or 1)
        var = float(np.var(pts[:, 2]))
        sample = uniform_locations(64 * max(1, oracle.num_threads()), extent=EXTENT, seed=99)
        t0 = time.perf_counter()
        oracle.krige(pts, sample, "linear", 0.0, var, 8000.0, 32, "ok")
        dt = time.perf_counter() - t0
        extra["cpu_port_kriged_points_per_s_ok_k32"] = {"value": len(sample) / dt, "cores": oracle.num_threads(),
                                                         "sample": f"{len(sample)} locations against {n} known points in {dt:.1f} s"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(n, world),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 24), "d2h_bytes_per_step": out_bytes},
                "gpu_launches": 8 * args.steps,
                "gpu_launches_note": "per step: bbox_partial, curve_key, gather_soa, tile_box, group_sum, item_cost, pair_ring, "
                                     "reduce_partials (+ CUB radix-sort passes and memsets, not counted)",
                "roofline": roofline, "cpu_baseline": cpu_baseline, "extra": extra}
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
