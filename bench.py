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
The kriging half of the metric (kriged points/s) is measured after the timed region: the top-level `kriging` block
is BASELINE config 4 (1M known points, 10M unknown locations, no_neighbors=64, ordinary + simple kriging) through the
public call with host arrays, the locations strong-scaled over the ranks, with its own roofline / e2e / cpu_baseline;
device-resident rates per kernel sit under `extra` (none of it enters `value`).
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
COUNT_CHECKSUM_CONFIG2 = 172371810095     # sum over the 80 lags of the unordered pair counts of the seed-2 field (1 GPU = oracle)
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
    ap.add_argument("--krige-points", type=int, default=10_000_000, help="unknown locations of the config-4 kriging block")
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


def reference_python_config1():
    """The UNMODIFIED reference (oracle/_ref, a git-ignored copy made by oracle/make_ref.py; or /root/reference in the
    build container) on BASELINE config 1 in full, nothing extrapolated: ExperimentalVariogram(step 500, max_range 40000)
    on 10k points, then ordinary_kriging(no_neighbors=32) of a bounded sample of locations."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_import
        ref = ref_import.import_reference()
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}
    from pyinterpolate_b200.synthetic import dem_like_field, uniform_locations
    pts = dem_like_field(10_000, extent=EXTENT, seed=0)
    t0 = time.perf_counter()
    ev = ref.ExperimentalVariogram(pts, step_size=500, max_range=40000)
    dt_v = time.perf_counter() - t0
    n = len(pts)
    out = {"config": "BASELINE config 1: 10k points, step 500, max_range 40000 (79 lags), semivariance + covariance",
           "variogram_s": dt_v, "pairs_per_s": (n * (n - 1) / 2) / dt_v, "cores": os.cpu_count(),
           "lags": int(len(ev.lags))}
    try:
        model = ref.TheoreticalVariogram()
        model.from_dict({"variogram_model_type": "linear", "nugget": 0.0, "sill": float(np.var(pts[:, 2])), "rang": 8000.0})
        unk = uniform_locations(300, extent=EXTENT, seed=1)
        t0 = time.perf_counter()
        ref.ordinary_kriging(theoretical_model=model, known_locations=pts, unknown_locations=unk, no_neighbors=32,
                             progress_bar=False)
        dt_k = time.perf_counter() - t0
        out.update({"kriging_sample_locations": len(unk), "kriging_s": dt_k, "kriged_points_per_s_ok_k32": len(unk) / dt_k})
    except Exception as e:  # noqa: BLE001
        out["kriging_unavailable"] = f"{type(e).__name__}: {e}"
    return out


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
                             "sample": sample + " per step; ms_per_step is the extrapolation to the full triangle",
                             "reference_python": reference_python_config1()},
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

    def step(device_resident=True, want_pairs=False):
        """One variogram pass; returns the (all-reduced) sums.  ``want_pairs`` reads the evaluated-pair counter back (a host
        synchronisation: only used outside the timed region)."""
        flush.fill_(1)                                                  # evict the points from L2
        if device_resident:
            sums = _lib.variogram_device(d_pts, lags, shard=(rank, world), want_pairs=want_pairs)
        else:
            h = _lib.variogram_host(pinned.numpy(), lags, shard=(rank, world))
            sums = {k: (torch.from_numpy(np.ascontiguousarray(v)).to(dev) if isinstance(v, np.ndarray) else v)
                    for k, v in h.items()}
        if world > 1:
            sums = allreduce_variogram_sums(sums, device=dev)
        return sums

    fp64_peak = _lib.fp64_peak_tflops()

    # ---- device-resident timing: K steps between barriers, CUDA events, max over ranks -------------
    sampler = ClockSampler(local_rank)                 # started ahead of the warm-up: nvidia-smi needs ~0.3 s to deliver its
    sampler.start()                                    # first sample and an 8-GPU timed region is shorter than that
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the K steps are enqueued back to back, nothing is read back between them: the library keeps its CUDA-event records of the
    # pair kernel across the calls (pib_timing_accumulate) and they are read once, after the timed region
    _lib.timing_accumulate(True)
    ev0.record()
    for _ in range(args.steps):
        sums = step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    pair_ms_sum, pair_launches = _lib.last_kernel_ms(_lib.KERNEL_PAIR)
    assert pair_launches == args.steps, (pair_launches, args.steps)
    _lib.timing_accumulate(False)
    evaluated = step(want_pairs=True)["pairs_evaluated"]        # this rank's evaluated pairs (untimed extra pass)
    t = torch.tensor([elapsed_ms, pair_ms_sum / pair_launches], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, pair_kernel_ms = float(t[0]), float(t[1])
    ms_per_step = elapsed_ms / args.steps
    # every rank holds the reduced histogram: its pair-count checksum must be the single-GPU one
    count_checksum = int(sums["count"].sum().item())
    if n == N_POINTS:
        assert count_checksum == COUNT_CHECKSUM_CONFIG2, f"rank {rank}: pair-count checksum {count_checksum} != {COUNT_CHECKSUM_CONFIG2}"
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
    roofline = {"bound": "fp64", "kernel": _lib.last_kernel_name(_lib.KERNEL_PAIR), "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak,
                "peak_source": "DFMA-saturation probe run in this process (pib_fp64_peak_tflops); MEASURED_PEAKS.json "
                               "has no FP64 figure (its HBM %.0f GB/s / bf16 figures do not bound this kernel)"
                               % measured_peaks().get("hbm_gbs", float("nan")),
                "flop_per_pair": FLOP_PER_PAIR, "pairs_evaluated_per_launch": evaluated_per_rank,
                "flop_note": "algorithmic fp64 operations per pair of the reference formulation (SURVEY section 8d: 5 for the "
                             "distance, 8 for the three sums), not instructions executed: with 'fp32 screening' in the kernel name "
                             "the lag of a pair is decided in packed single precision (FADD2 / FMUL2 / FFMA2) and confirmed in fp64 "
                             "only for the undecided groups, so the FP64 pipe itself is busy well below this fraction "
                             "(profiles/r2_pair_ring6_f32_ncu_full.md)",
                "kernel_ms_per_launch": pair_kernel_ms, "kernel_share_of_step": pair_kernel_ms / ms_per_step,
                "traffic": profile_traffic("pair_ring_kernel")}

    extra = {"pairs_evaluated": evaluated, "evaluated_pairs_per_s": evaluated / (ms_per_step * 1e-3),
             "pair_count_checksum": count_checksum,
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

    # ---- kriging half of the metric (outside the variogram's timed region) ------------------------------------------
    kriging = None
    if not args.no_extra:
        import pyinterpolate_b200 as pib
        var, mean = float(np.var(pts[:, 2])), float(pts[:, 2].mean())
        # (1) device-resident rates per kernel: every rank kriges 1M locations of its own (weak scaling)
        for name, k, m, model, params, mode in (
                ("ok_k32", 32, 1_000_000, "linear", [0.0, var, 8000.0], _lib.KRIGE_ORDINARY),
                ("ok_k64", 64, 1_000_000, "spherical", [1.0, var, 20000.0], _lib.KRIGE_ORDINARY),
                ("sk_k64", 64, 1_000_000, "spherical", [1.0, var, 20000.0], _lib.KRIGE_SIMPLE),
                ("ok_sk_k64", 64, 1_000_000, "spherical", [1.0, var, 20000.0], _lib.KRIGE_BOTH)):
            unk = torch.from_numpy(uniform_locations(m, extent=EXTENT, seed=5 + rank)).to(dev)
            call = lambda: _lib.krige_device(d_pts, unk, _lib.MODEL_IDS[model], np.array([params]), mode, k, process_mean=mean)
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
            # LU 2n^3/3 + 2n^2 per system, n = k + 1 (ordinary) or k (simple); both = the sum of the two
            flop_ok, flop_sk = 2 * (k + 1) ** 3 / 3 + 2 * (k + 1) ** 2, 2 * k ** 3 / 3 + 2 * k ** 2
            solve_flop = m * (flop_ok if mode == _lib.KRIGE_ORDINARY else flop_sk if mode == _lib.KRIGE_SIMPLE else flop_ok + flop_sk)
            knn_bytes = m * (16 + 24 * k + 32)                                # algorithmic bytes per query (section 8d)
            extra[f"kriged_points_per_s_{name}"] = world * m / (float(tk) * 1e-3)
            extra[f"kriging_{name}"] = {"known": n, "unknown_per_gpu": m, "model": model, "ms": float(tk),
                                        "knn_kernel_ms": knn_ms, "solve_kernel_ms": solve_ms,
                                        "solve_kernel": _lib.last_kernel_name(_lib.KERNEL_SOLVE),
                                        "solve_fp64_tflops": solve_flop / (solve_ms * 1e-3) / 1e12,
                                        "solve_frac_of_fp64_peak": solve_flop / (solve_ms * 1e-3) / 1e12 / fp64_peak,
                                        "knn_algorithmic_gbs": knn_bytes / (knn_ms * 1e-3) / 1e9,
                                        "scaling": "weak"}
        del unk

        # (2) BASELINE config 4 through the PUBLIC call with host arrays: 1M known (seed 4), 10M unknown (seed 5),
        # no_neighbors = 64, spherical {nugget 1, sill var, rang 20000}, ordinary + simple kriging from one pass.
        # Strong scaling: the locations are split over the ranks (no data-path collective); time = max over ranks.
        known4 = dem_like_field(1_000_000, extent=EXTENT, seed=4)
        m4 = args.krige_points
        unk4 = uniform_locations(m4, extent=EXTENT, seed=5)
        b4, e4 = (rank * m4) // world, ((rank + 1) * m4) // world
        mdl4 = {"variogram_model_type": "spherical", "nugget": 1.0, "sill": float(np.var(known4[:, 2])), "rang": 20000.0}
        mean4 = float(known4[:, 2].mean())
        # warm-up: both lanes of the chunk pipeline and their pinned staging at full size (first-call costs, like context creation)
        pib.ordinary_and_simple_kriging(mdl4, known4, unk4[b4:min(e4, b4 + 1_200_000)], process_mean=mean4, no_neighbors=64)
        times = []
        for _ in range(2):
            barrier()
            t0 = time.perf_counter()
            ok4, sk4 = pib.ordinary_and_simple_kriging(mdl4, known4, unk4[b4:e4], process_mean=mean4, no_neighbors=64)
            torch.cuda.synchronize()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            times.append(float(tt))
        dt4 = min(times)
        knn_ms4, _ = _lib.last_kernel_ms(_lib.KERNEL_KNN)
        solve_ms4, _ = _lib.last_kernel_ms(_lib.KERNEL_SOLVE)
        ms4 = torch.tensor([knn_ms4, solve_ms4], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms4, op=dist.ReduceOp.MAX)
        knn_ms4, solve_ms4 = float(ms4[0]), float(ms4[1])
        m_rank = e4 - b4
        flop4 = m_rank * ((2 * 65 ** 3 / 3 + 2 * 65 ** 2) + (2 * 64 ** 3 / 3 + 2 * 64 ** 2))
        finite = bool(np.isfinite(ok4[:, 0]).all() and np.isfinite(sk4[:, 0]).all())
        kriging = {"metric": "kriged_points_per_s", "value": m4 / dt4, "unit": "points/s (each location: ordinary AND simple kriging)",
                   "n_gpus": world, "scaling": "strong", "seconds": dt4, "all_finite": finite,
                   "config": {"workload": "BASELINE config 4: ordinary + simple kriging of %d unknown locations from 1M known points, "
                                          "no_neighbors=64, spherical model (nugget 1, sill var(z), rang 20000); public call "
                                          "ordinary_and_simple_kriging with host arrays, locations split over the ranks" % m4,
                              "known": 1_000_000, "unknown": m4, "no_neighbors": 64},
                   "e2e": {"value": m4 / dt4, "unit": "points/s",
                           "h2d_bytes_per_step": int(known4.nbytes * world + unk4.nbytes),
                           "d2h_bytes_per_step": int(2 * m4 * 4 * 8 + 2 * m4)},
                   "roofline": {"solve": {"bound": "fp64", "kernel": _lib.last_kernel_name(_lib.KERNEL_SOLVE),
                                          "achieved": flop4 / (solve_ms4 * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                                          "frac": flop4 / (solve_ms4 * 1e-3) / 1e12 / fp64_peak, "kernel_ms_per_rank": solve_ms4,
                                          "flop_per_location": flop4 / m_rank},
                                "knn": {"bound": "hbm", "achieved": m_rank * (16 + 24 * 64 + 32) / (knn_ms4 * 1e-3) / 1e9,
                                        "peak": measured_peaks().get("hbm_gbs"), "unit": "GB/s (algorithmic bytes: 16 + 24 k + 32 per location)",
                                        "frac": m_rank * (16 + 24 * 64 + 32) / (knn_ms4 * 1e-3) / 1e9 / (measured_peaks().get("hbm_gbs") or float("nan")),
                                        "kernel_ms_per_rank": knn_ms4},
                                "kernel_ms_note": "kernel times are summed over the chunks of the call; the two lanes of the host pipeline "
                                                  "run their kernels concurrently, so the sums can exceed the call's wall time (the "
                                                  "device-resident rates per kernel are under extra.kriging_*)"}}
        del ok4, sk4
        # the k = 32 public call of config 1's shape (ordinary kriging, linear model), same strong-scaled split
        mdl1 = {"variogram_model_type": "linear", "nugget": 0.0, "sill": var, "rang": 8000.0}
        m1 = m4                                              # the same 10M locations as config 4, split over the ranks
        b1, e1 = (rank * m1) // world, ((rank + 1) * m1) // world
        pib.ordinary_kriging(mdl1, pts, unk4[b1:min(e1, b1 + 2_200_000)], no_neighbors=32)
        t1s = []
        for _ in range(2):                                   # best of two, like config 4 above
            barrier()
            t0 = time.perf_counter()
            res1 = pib.ordinary_kriging(mdl1, pts, unk4[b1:e1], no_neighbors=32)
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t1s.append(float(tt))
        tt = min(t1s)
        kriging["ok_k32_public_call"] = {"value": m1 / float(tt), "unit": "points/s", "unknown": m1, "known": n, "seconds": float(tt),
                                         "h2d_bytes": int(pts.nbytes * world + unk4[:m1].nbytes), "d2h_bytes": int(m1 * 4 * 8 + m1)}
        del res1

        # CPU restatement of the kriging path on the host cores (bounded sample), rank 0 only
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle
            oracle.set_threads(os.cpu_count() or 1)
            sample = unk4[: 48 * max(1, oracle.num_threads())]
            t0 = time.perf_counter()
            oracle.krige(known4, sample, "spherical", 1.0, mdl4["sill"], 20000.0, 64, "ok")
            oracle.krige(known4, sample, "spherical", 1.0, mdl4["sill"], 20000.0, 64, "sk", process_mean=mean4)
            dtc = time.perf_counter() - t0
            kriging["cpu_baseline"] = {"value": len(sample) / dtc, "unit": "points/s (ordinary + simple)", "cores": oracle.num_threads(),
                                       "kind": "port", "sample": f"{len(sample)} locations against 1M known points, k=64, in {dtc:.1f} s"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(n, world),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 24), "d2h_bytes_per_step": out_bytes},
                "gpu_launches": 15 * args.steps,
                "gpu_launches_note": "per step: bbox_partial, bbox_final, curve_key, gather_soa, tile_box, mean_partial, mean_final, "
                                     "group_sum, tile_f32, item_cost, pair_ring6, reduce_partials, cancellation_guard, pair_kernel + "
                                     "reduce_partials (conditional re-run: both return at once unless the guard fired) "
                                     "(+ CUB radix-sort passes and memsets, not counted)",
                "roofline": roofline, "cpu_baseline": cpu_baseline, "kriging": kriging, "extra": extra}
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
