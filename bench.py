#!/usr/bin/env python
"""bench.py — rays/s of TraceNonSequential on BASELINE.json's configurations.

Headline (`value`, `e2e`, `roofline`, `cpu_baseline`): configs[1], DaviesCotton.C — 88-facet Davies-Cotton reflector, 9 field
angles 0..4 deg, each ARayShooter::Square(400 nm, 14 m, n=3334) = 11 115 556 rays, 100 040 004 rays per step.
One "step" = one TraceNonSequential pass over all nine batches plus the on-device PSF reducers (histogram + moments + D80 per
field angle).

  value    : rays/s, inputs resident in HBM (rays generated on device by rbg_shoot before the timed region)
  e2e      : rays/s through the C ABI with HOST (pinned) buffers; H2D of the 64 B/ray inputs and D2H of the 68 B/ray results are
             inside the timed region
  roofline : the navigation kernel k_nav (dominant): algorithmic 132 B/ray (64 in + 68 out) over the CUDA-event time of the
             bounce kernels of a pass, against the measured HBM peak
  cpu_baseline : the CPU oracle (reference algorithm restated; ROOT is unavailable) on all host threads, on a sample of
             windows spread evenly over the whole beam, with the daughter-box trees that play TGeoVoxelFinder's role
             ("port+voxels"); the walk over all daughters ("port") is reported next to it
  configs  : the same figures for every BASELINE config (1-5) at its BASELINE size on one GPU, each with its own bounce-kernel
             time, roofline fraction, CPU row and host-buffer (e2e) rate
  cfg5_strong (torchrun, or 1 GPU): the 1e9-ray HexWinstonCone run of configs[4], rays [rank * 1e9/N, (rank+1) * 1e9/N) per
             rank by shard_range, global ray ids and seeds, generated / traced / reduced on the device batch by batch

Multi-GPU (torchrun): rays shard across ranks (each rank traces its own field-angle sweep, geometry replicated), no data-path
collective; the PSF histograms/moments are all-reduced with NCCL at the end of each step.  The headline scales weakly (per-GPU
work fixed), cfg5_strong strongly (total work fixed).

`--impl reference` times the reference arm: the reference's own implementation cannot be built here (needs CERN ROOT), so it is
the oracle port on all host threads (cpu_baseline.kind = "port+voxels"), on the same spread-out sample of the same workload.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_SIDE = 3334
ANGLES = [0.5 * i for i in range(9)]
BYTES_IN, BYTES_OUT = 64, 68
FP64_PEAK_TFLOPS = 34.19  # DFMA peak measured on this pool (profiles/r1j_fp64_peak.json, profiles/fp64_peak.cu)

# the five BASELINE configs at their BASELINE sizes on one GPU: (cfg, field angle, rays, builder kwargs, beam side for cfg 5)
CONFIGS = {
    1: dict(name="SimpleParabolicTelescope.C", theta=0.0, n=1000 * 1000, kw={}, what="on-axis Square(400 nm, 5 m, n=1000)"),
    2: dict(name="DaviesCotton.C", theta=2.0, n=N_SIDE * N_SIDE, kw={}, what="one field angle (2 deg) of the headline sweep, Square(400 nm, 14 m, n=3334)"),
    3: dict(name="SchwarzschildCouder.C", theta=0.0, n=10000 * 10000, kw={}, what="on-axis Square(400 nm, 20 m, n=10000)"),
    4: dict(name="SchmidtCassegrain.C", theta=0.0, n=10_000_000, kw={}, what="RandomCircle(12.5 in), 300-700 nm, Fresnel + absorption on"),
    5: dict(name="HexWinstonCone.C", theta=20.0, n=125_000_000, kw=dict(rings=10), side=84.0,
            what="331 cells, multilayer-coated walls (direct TMM), RandomSquare at 20 deg: one GPU's share (1/8) of the 1e9-ray run"),
}
CFG5_TOTAL = 1_000_000_000
CFG5_BATCH = 125_000_000


def make_clock_reader():
    """-> function returning one row [sm MHz, max sm MHz, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap].  In-process NVML
    (nvidia_ml_py), initialised HERE, before the timed region: nvmlInit and every `nvidia-smi` process take the driver's global
    lock for ~0.1 s, which showed up as 80 vs 115 ms per step in a 0.4 s timed region.  nvidia-smi is the fallback."""
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    try:
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(dev)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = (0x8, 0x40, 0x20, 0x4)  # HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap

        def read():
            r = int(get_reasons(h))
            return [str(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), str(mx)] + ["Active" if r & b else "Not Active" for b in bits]
        read()
        read.inline = True  # cheap enough to call from the timing thread between steps
        return read
    except Exception:
        pass
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def read_smi():
        txt = subprocess.run(["nvidia-smi", "-i", str(dev), "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in txt.strip().split(",")]
        return f if len(f) >= 6 else None
    read_smi.inline = False
    return read_smi


_AFFINITY = {}


def all_host_cpus():
    """undo pin_to_gpu_numa_node for the CPU legs: the oracle rows use every core of the host"""
    if "full" in _AFFINITY:
        os.sched_setaffinity(0, _AFFINITY["full"])


def gpu_side_cpus():
    if "gpu" in _AFFINITY:
        os.sched_setaffinity(0, _AFFINITY["gpu"])


def pin_to_gpu_numa_node(dev):
    """Bind this rank's threads to the CPUs next to its GPU (NVML's ideal affinity) before any pinned host buffer exists: the
    buffers of the host-buffer path are then first touched, and the copy pipeline driven, on the GPU's own NUMA node.  Returns
    the number of CPUs in the mask, or None when NVML (or the call) is unavailable."""
    if os.environ.get("RB_BENCH_NO_AFFINITY"):
        return None
    try:
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(dev)
        words = N.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        full = os.sched_getaffinity(0)
        cpus &= full
        if not cpus:
            return None
        _AFFINITY.setdefault("full", full)
        _AFFINITY["gpu"] = cpus
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def sample_clocks(stop, out, read):
    """background sampler, used only with the nvidia-smi fallback.  NVML queries from a second thread contend for the driver lock
    with the launching thread: measured on B200, steps of 78.5 ms became 110-770 ms now and then.  The NVML reader is therefore
    called by the timing thread itself after each step is queued (the GPU is still working on it: the sample is under load)."""
    if read.inline:
        return
    while not stop.is_set():
        try:
            row = read()
            if row:
                out.append(row)
        except Exception:
            pass
        stop.wait(0.1)


def clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None, "reasons": reasons, "samples": len(samples)}


# ------------------------------------------------------------------------------------------------ CPU rows (the oracle as checker / baseline only)
def load_cpu_oracle():
    """the oracle built for this host's cores (-march=native, built here on first use) if the compiler is around, else the
    portable build that travelled with the repo"""
    import helpers as H
    try:
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "native"], check=True, timeout=240, capture_output=True)
        native = os.path.join(ROOT, "oracle", "_build", "liboracle_native.so")
        if os.path.exists(native):
            H.ORACLE_SO = native
            return H.load_oracle(), "-O3 -march=native"
    except Exception:
        pass
    return H.load_oracle(), "-O3"


def spread_sample(oracle, H, beam, n_total, n_sample, windows=64):
    """`windows` windows of consecutive rays at evenly spaced offsets over the whole beam (not its busiest rows)"""
    import numpy as np
    windows = max(1, min(windows, n_sample // 64 or 1))
    w = max(1, n_sample // windows)
    parts = []
    for k in range(windows):
        first = int((k + 0.5) * n_total / windows) - w // 2
        first = max(0, min(n_total - w, first))
        parts.append(H.make_rays(oracle, beam, first, w).inp)
    return np.concatenate(parts, axis=1)


def cpu_rows(oracle, H, export, beam, n_total, opts, threads, seconds, flags):
    """-> (cpu_baseline dict with voxels, dict of the full daughter walk): rays/s of the oracle on `threads` host threads on a
    spread-out sample sized for about `seconds` of work each"""
    all_host_cpus()
    try:
        import numpy as np
        rows = {}
        for vox in (1, 0):
            oracle.orc_set_voxels(vox)
            try:
                probe = H.Rays(spread_sample(oracle, H, beam, n_total, 4096 * 4).T)
                t0 = time.perf_counter()
                H.trace_with(oracle.orc_trace, export, probe, opts, nthreads=threads)
                rate = probe.n / (time.perf_counter() - t0)
                ns = int(min(8e6, max(20000, rate * (seconds if vox else seconds / 3))))
                rays = H.Rays(spread_sample(oracle, H, beam, n_total, ns).T)
                t0 = time.perf_counter()
                H.trace_with(oracle.orc_trace, export, rays, opts, nthreads=threads)
                dt = time.perf_counter() - t0
                rows[vox] = {"value": rays.n / dt, "unit": "rays/s", "cores": threads, "per_thread": rays.n / dt / threads,
                             "kind": "port+voxels" if vox else "port",
                             "sample": "%d rays in 64 windows spread evenly over the beam, %.1f s; reference algorithm restated in C++ (oracle/, %s), ROOT unavailable%s"
                                       % (rays.n, dt, flags, "; daughter-box trees in TGeoVoxelFinder's role, results bit-identical to the full walk" if vox else
                                          "; every daughter of a volume examined at every step")}
            finally:
                oracle.orc_set_voxels(1)
        return rows[1], rows[0]
    finally:
        gpu_side_cpus()


def run_reference(args):
    """reference arm: CPU oracle (port of the reference algorithm) with all host threads, same workload, spread-out sample"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import helpers as H
    import numpy as np
    from robast_b200 import configs
    oracle, flags = load_cpu_oracle()
    mgr, _k = configs.davies_cotton()
    export = mgr.ExportScene()
    threads = os.cpu_count() or 1
    opts = H.opts(disable_fresnel=1)
    n = N_SIDE * N_SIDE
    probe = H.Rays(spread_sample(oracle, H, configs.beam(2, 2.0, n_side=N_SIDE), n, 65536).T)
    t0 = time.perf_counter()
    H.trace_with(oracle.orc_trace, export, probe, opts, nthreads=threads)
    rate = probe.n / (time.perf_counter() - t0)
    per_angle = max(4096, int(rate * 4.0 / len(ANGLES)))  # ~4 s per step
    batches = [H.Rays(spread_sample(oracle, H, configs.beam(2, th, n_side=N_SIDE), n, per_angle).T) for th in ANGLES]
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for b in batches:
            H.trace_with(oracle.orc_trace, export, b, opts, nthreads=threads)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    nrays = sum(b.n for b in batches)
    value = nrays * args.steps / total
    sample = "%d rays per field angle (64 windows spread evenly over each 3334^2 grid) x 9 angles per step; oracle with daughter-box trees, %s" % (batches[0].n, flags)
    print(json.dumps({
        "impl": "reference", "metric": "rays/s", "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "DaviesCotton.C (BASELINE configs[1]): 88 hex facets + camera + masts, 9 field angles 0-4 deg, Square(400 nm, 14 m, n=3334) each; bounded CPU sample",
                   "rays_per_step": nrays},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "per_thread": value / threads, "kind": "port+voxels", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference itself needs CERN ROOT (absent); this is the restated reference algorithm (oracle/) on all host threads",
    }))


# ------------------------------------------------------------------------------------------------ device helpers
class Batch:
    """device-resident SoA ray batch + its rbg_rays view"""

    def __init__(self, R, torch, dev, n):
        self.n = n
        self.inp = torch.empty((8, n), dtype=torch.float64, device=dev)
        self.out = torch.empty((7, n), dtype=torch.float64, device=dev)
        self.iout = torch.empty((3, n), dtype=torch.int32, device=dev)
        r = R.rbg_rays()
        r.n, r.on_device = n, 1
        for i, key in enumerate(["x", "y", "z", "t", "dx", "dy", "dz", "lambda_"]):
            setattr(r, key, self.inp[i].data_ptr())
        for i, key in enumerate(["ox", "oy", "oz", "ot", "odx", "ody", "odz"]):
            setattr(r, key, self.out[i].data_ptr())
        for i, key in enumerate(["status", "last_node", "npoints"]):
            setattr(r, key, self.iout[i].data_ptr())
        self.struct = r


def profile_reset(R):
    bm, bn, cm, cn = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
    R.rbg_profile_read(C.byref(bm), C.byref(bn), C.byref(cm), C.byref(cn))
    return bm.value, bn.value, cm.value, cn.value


def roofline_of(n_rays_traced, bounce_ms, hbm_peak, flops_per_ray=None):
    """algorithmic 132 B/ray over the summed CUDA-event time of the bounce kernels"""
    if bounce_ms <= 0:
        return None
    achieved = (BYTES_IN + BYTES_OUT) * n_rays_traced / (bounce_ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
         "algorithmic_bytes_per_ray": BYTES_IN + BYTES_OUT, "bounce_kernel_ms": bounce_ms}
    if flops_per_ray:
        tf = flops_per_ray * n_rays_traced / (bounce_ms * 1e-3) / 1e12
        r["fp64"] = {"algorithmic_kflop_per_ray": flops_per_ray / 1e3, "achieved_tflops": tf, "peak_tflops": FP64_PEAK_TFLOPS, "frac": tf / FP64_PEAK_TFLOPS,
                     "note": "SURVEY.md 8d flop estimate x rays / bounce-kernel time; the measured pipe utilisation is in profiles/ (ncu)"}
    return r


KFLOP = {1: 0.6e3, 2: 2.5e3, 3: 6e3, 4: 10e3, 5: 9e3}  # SURVEY.md §8d algorithmic flop estimates per ray


def bench_one_config(R, torch, H, configs, dev, local, stream, cfg, reps, hbm_peak, oracle, flags, cpu_seconds, e2e_cap):
    """device-resident rays/s, bounce-kernel roofline, host-buffer rate and CPU rows of one BASELINE config on this GPU"""
    c = CONFIGS[cfg]
    mgr, _keep = configs.BUILDERS[cfg](**c["kw"])
    export = mgr.ExportScene()
    scene = C.c_void_p()
    R.check(R.rbg_scene_create(export.desc_ptr(), local, C.byref(scene)))
    n = c["n"]
    nside = int(round(n ** 0.5)) if cfg <= 3 else c.get("side")
    beam = configs.beam(cfg, c["theta"], n_side=nside)
    b = Batch(R, torch, dev, n)
    d = H.shoot_desc(beam)
    R.check(R.rbg_shoot(C.byref(d), 0, n, *[b.inp[i].data_ptr() for i in range(8)], local, stream))
    opts = H.opts(disable_fresnel=1 if cfg == 2 else 0, seed=20180601 if cfg != 5 else 20110306)
    for _ in range(2):
        R.check(R.rbg_trace(scene, C.byref(opts), C.byref(b.struct), stream))
    torch.cuda.synchronize()
    R.rbg_profile_enable(1)
    profile_reset(R)
    l0 = R.rbg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        R.check(R.rbg_trace(scene, C.byref(opts), C.byref(b.struct), stream))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bm, bn, cm, cn = profile_reset(R)
    R.rbg_profile_enable(0)
    launches = (R.rbg_launch_count() - l0) // reps
    counts = torch.bincount(b.iout[0].to(torch.int64), minlength=6).cpu().numpy().tolist()
    res = {"workload": "%s: %s" % (c["name"], c["what"]), "rays": n, "value": n / (ms * 1e-3), "unit": "rays/s", "ms_per_trace": ms,
           "kernel_variant": R.rbg_scene_kernel_variant(scene).decode(), "gpu_launches_per_trace": int(launches),
           "bounce_kernel_ms_per_trace": bm / reps, "compaction_and_sort_ms_per_trace": cm / reps,
           "roofline": roofline_of(n, bm / reps, hbm_peak, KFLOP[cfg]), "status_counts": counts, "mean_npoints": float(b.iout[2].float().mean().item())}
    # host-buffer rate (pinned), bounded size
    ne = min(n, e2e_cap)
    if ne > 0:
        hin = torch.empty((8, ne), dtype=torch.float64).pin_memory()
        hin.copy_(b.inp[:, :ne].cpu())
        hout = torch.empty((7, ne), dtype=torch.float64).pin_memory()
        hiout = torch.empty((3, ne), dtype=torch.int32).pin_memory()
        r = R.rbg_rays()
        r.n, r.on_device = ne, 0
        for i, key in enumerate(["x", "y", "z", "t", "dx", "dy", "dz", "lambda_"]):
            setattr(r, key, hin[i].data_ptr())
        for i, key in enumerate(["ox", "oy", "oz", "ot", "odx", "ody", "odz"]):
            setattr(r, key, hout[i].data_ptr())
        for i, key in enumerate(["status", "last_node", "npoints"]):
            setattr(r, key, hiout[i].data_ptr())
        R.check(R.rbg_trace(scene, C.byref(opts), C.byref(r), None))
        t0 = time.perf_counter()
        for _ in range(2):
            R.check(R.rbg_trace(scene, C.byref(opts), C.byref(r), None))
        dt = (time.perf_counter() - t0) / 2
        same = bool((hiout[0] == b.iout[0, :ne].cpu()).all().item())
        res["e2e"] = {"value": ne / dt, "unit": "rays/s", "rays": ne, "h2d_bytes_per_trace": BYTES_IN * ne, "d2h_bytes_per_trace": BYTES_OUT * ne,
                      "statuses_equal_device_path": same}
        del hin, hout, hiout
    # the MINUIT-loop regime (tutorials/Optimize.C, optimize_multilayer.C): 1000-ray calls with host arrays, one after another
    nl = 1000
    small = H.Rays(b.inp[:, :: max(1, n // nl)][:, :nl].cpu().numpy().T)
    rs = small.struct()
    for _ in range(5):
        R.check(R.rbg_trace(scene, C.byref(opts), C.byref(rs), None))
    t0 = time.perf_counter()
    for _ in range(200):
        R.check(R.rbg_trace(scene, C.byref(opts), C.byref(rs), None))
    res["latency_us_1k_rays"] = (time.perf_counter() - t0) / 200 * 1e6
    if cfg == 5:
        # the same array with the coating's reflectance read from the table of AMultilayer::PreCalculateCoherentTMM
        # (include/AMultilayer.h:243-262: 801 wavelengths x 90 angles), the macro author's other option
        mgr2, _keep2 = configs.BUILDERS[5](precalc=True, **c["kw"])
        ex2 = mgr2.ExportScene()
        scene2 = C.c_void_p()
        R.check(R.rbg_scene_create(ex2.desc_ptr(), local, C.byref(scene2)))
        for _ in range(2):
            R.check(R.rbg_trace(scene2, C.byref(opts), C.byref(b.struct), stream))
        e0.record()
        for _ in range(reps):
            R.check(R.rbg_trace(scene2, C.byref(opts), C.byref(b.struct), stream))
        e1.record()
        torch.cuda.synchronize()
        res["value_with_precalculated_tmm_table"] = n / (e0.elapsed_time(e1) / reps * 1e-3)
        R.rbg_scene_destroy(scene2)
    del b
    torch.cuda.empty_cache()
    if oracle is not None:
        vox, brute = cpu_rows(oracle, H, export, beam, n, opts, os.cpu_count() or 1, cpu_seconds, flags)
        res["cpu_baseline"] = vox
        res["cpu_baseline_full_walk"] = brute
    R.rbg_scene_destroy(scene)
    return res


def bench_cfg5_strong(R, torch, H, configs, sharding, dev, local, stream, rank, world, dist, steps):
    """configs[4] as BASELINE states it: 1e9 RandomSquare rays over the ranks (contiguous ranges, global ray ids), generated on
    the device, traced and reduced to the PMT-plane histogram + status counters batch by batch; nothing but the reducers leaves
    the GPU.  Strong scaling: the total is fixed, each rank takes 1e9 / N rays."""
    c = CONFIGS[5]
    mgr, _keep = configs.BUILDERS[5](**c["kw"])
    export = mgr.ExportScene()
    scene = C.c_void_p()
    R.check(R.rbg_scene_create(export.desc_ptr(), local, C.byref(scene)))
    beam = configs.beam(5, c["theta"], n_side=c["side"])
    d = H.shoot_desc(beam)
    lo, hi = sharding.shard_range(CFG5_TOTAL, rank, world)
    nb = min(CFG5_BATCH, hi - lo)
    b = Batch(R, torch, dev, nb)
    hist = torch.zeros(200 * 200, dtype=torch.int64, device=dev)
    mom = torch.zeros(8, dtype=torch.float64, device=dev)
    cnt = torch.zeros(6, dtype=torch.int64, device=dev)
    opts = H.opts(seed=20110306)

    def step():
        hist.zero_(); mom.zero_(); cnt.zero_()
        first = lo
        while first < hi:
            m = min(nb, hi - first)
            R.check(R.rbg_shoot(C.byref(d), first, m, *[b.inp[i].data_ptr() for i in range(8)], local, stream))
            b.struct.n = m
            opts.ray_id_offset = first
            R.check(R.rbg_trace(scene, C.byref(opts), C.byref(b.struct), stream))
            R.check(R.rbg_hist2d(m, b.out[0].data_ptr(), b.out[1].data_ptr(), b.iout[0].data_ptr(), R.RBG_FOCUSED, 200, -45., 45., 200, -45., 45., hist.data_ptr(), local, stream))
            R.check(R.rbg_moments(m, b.out[0].data_ptr(), b.out[1].data_ptr(), b.out[3].data_ptr(), b.iout[0].data_ptr(), R.RBG_FOCUSED, mom.data_ptr(), cnt.data_ptr(), local, stream))
            first += m
        if dist is not None:
            dist.all_reduce(hist)
            dist.all_reduce(mom)
            dist.all_reduce(cnt)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item()) / steps
    counts = cnt.cpu().numpy().tolist()
    res = {"workload": "HexWinstonCone.C (BASELINE configs[4]): 1e9 RandomSquare rays at 20 deg on 331 multilayer-coated cells, generated, traced and reduced on device",
           "rays_total": CFG5_TOTAL, "rays_per_gpu": hi - lo, "batch": nb, "n_gpus": world, "scaling": "strong", "value": CFG5_TOTAL / (ms * 1e-3), "unit": "rays/s",
           "ms_per_step": ms, "steps": steps, "status_counts": counts, "focused_fraction": counts[R.RBG_FOCUSED] / float(max(1, sum(counts))),
           "histogram_entries": int(hist.sum().item())}
    R.rbg_scene_destroy(scene)
    del b
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--steps-per-launch", type=int, default=int(os.environ.get("RB_STEPS_PER_LAUNCH", "0")))
    ap.add_argument("--n-side", type=int, default=N_SIDE, help="grid side per field angle (profiling runs only; the bench metric uses 3334)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-threads", type=int, default=int(os.environ.get("RB_E2E_THREADS", "1")), help="TraceNonSequential calls in flight on the host-buffer path")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config block (configs 1-5) and the cfg5 strong-scaling run")
    ap.add_argument("--latency", action="store_true", help="also measure the small-batch latency per config")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import robast_b200 as R
    from robast_b200 import configs, sharding
    import helpers as H

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if R.rbg_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the tracer has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    mgr, _keep = configs.davies_cotton()
    export = mgr.ExportScene()
    scene = C.c_void_p()
    R.check(R.rbg_scene_create(export.desc_ptr(), local, C.byref(scene)))
    nside = args.n_side
    n = nside * nside
    nang = len(ANGLES)
    stream = torch.cuda.current_stream().cuda_stream
    # this rank's field angles: the 0..4 deg sweep refined by the rank (weak scaling, geometry replicated)
    my_angles = [a + 0.5 * rank / world for a in ANGLES]
    inp = torch.empty((nang, 8, n), dtype=torch.float64, device=dev)
    out = torch.empty((7, n), dtype=torch.float64, device=dev)
    iout = torch.empty((3, n), dtype=torch.int32, device=dev)
    for k, th in enumerate(my_angles):
        d = H.shoot_desc(configs.beam(2, th, n_side=nside))
        R.check(R.rbg_shoot(C.byref(d), 0, n, *[inp[k, i].data_ptr() for i in range(8)], local, stream))
    # the reducers are double-buffered: the D80 search of a step (a sequential search, a few SMs) runs on a side stream while
    # the next step's rays are traced
    hist2 = [torch.zeros((nang, 200 * 200), dtype=torch.int64, device=dev) for _ in range(2)]
    hstats2 = [torch.zeros((nang, 5), dtype=torch.float64, device=dev) for _ in range(2)]
    d80_2 = [torch.zeros((nang, 3), dtype=torch.float64, device=dev) for _ in range(2)]
    d80_all = torch.zeros((world * nang, 3), dtype=torch.float64, device=dev)
    mom2 = [torch.zeros((nang, 8), dtype=torch.float64, device=dev) for _ in range(2)]
    cnt2 = [torch.zeros((nang, 6), dtype=torch.int64, device=dev) for _ in range(2)]
    hsum2 = [torch.zeros_like(hist2[0]) for _ in range(2)]
    side = torch.cuda.Stream(device=dev)
    side_done = [None, None]
    main_stream = torch.cuda.current_stream()
    state = {"k": 0}
    opts = H.opts(disable_fresnel=1, steps_per_launch=args.steps_per_launch, seed=20180601)

    def rays_struct(k):
        r = R.rbg_rays()
        r.n, r.on_device = n, 1
        for i, key in enumerate(["x", "y", "z", "t", "dx", "dy", "dz", "lambda_"]):
            setattr(r, key, inp[k, i].data_ptr())
        for i, key in enumerate(["ox", "oy", "oz", "ot", "odx", "ody", "odz"]):
            setattr(r, key, out[i].data_ptr())
        for i, key in enumerate(["status", "last_node", "npoints"]):
            setattr(r, key, iout[i].data_ptr())
        return r

    structs = [rays_struct(k) for k in range(nang)]

    def step():
        b = state["k"] & 1
        state["k"] += 1
        hist, hstats, d80, mom, cnt = hist2[b], hstats2[b], d80_2[b], mom2[b], cnt2[b]
        state["last"] = b
        if side_done[b] is not None:
            main_stream.wait_event(side_done[b])  # the search that read this buffer pair two steps ago
        hist.zero_()
        hstats.zero_()
        mom.zero_()
        cnt.zero_()
        for k in range(nang):
            opts.ray_id_offset = (rank * nang + k) * n
            R.check(R.rbg_trace(scene, C.byref(opts), C.byref(structs[k]), stream))
            # PSF histogram in image-centred coordinates (window of DaviesCotton.C:210, cm): every angle shares one binning,
            # so the histograms of all ranks stack and the nine D80 searches run as one launch
            cx = 1600. * math.tan(math.radians(my_angles[k]))
            R.check(R.rbg_hist2d_stats(n, out[0].data_ptr(), out[1].data_ptr(), iout[0].data_ptr(), R.RBG_FOCUSED, cx, 0., 200, -4., 10., 200, -7., 7.,
                                       hist[k].data_ptr(), hstats[k].data_ptr(), local, stream))
            R.check(R.rbg_moments(n, out[0].data_ptr(), out[1].data_ptr(), out[3].data_ptr(), iout[0].data_ptr(), R.RBG_FOCUSED, mom[k].data_ptr(), cnt[k].data_ptr(), local, stream))
        # D80 of each of this rank's field angles (AGeoUtil::ContainmentRadius) on the side stream, then the cross-rank reductions
        ev = torch.cuda.Event()
        ev.record(main_stream)
        side.wait_event(ev)
        R.check(R.rbg_containment_radius(nang, hist.data_ptr(), 200, -4., 10., 200, -7., 7., hstats.data_ptr(), 0.8, d80.data_ptr(), local, side.cuda_stream))
        if dist is not None:
            # the cross-rank reductions ride on the side stream as well (every reducer output is double-buffered): the next
            # step's traces do not wait for the other ranks
            with torch.cuda.stream(side):
                dist.all_gather_into_tensor(d80_all, d80)
                dist.all_reduce(mom)
                dist.all_reduce(cnt)
                hsum2[b].copy_(hist)
                dist.all_reduce(hsum2[b])
        side_done[b] = torch.cuda.Event()
        side_done[b].record(side)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clock_reader = make_clock_reader()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    R.rbg_profile_enable(1)
    profile_reset(R)
    stop, samples = threading.Event(), []
    th_clock = threading.Thread(target=sample_clocks, args=(stop, samples, clock_reader), daemon=True)
    th_clock.start()
    launches0 = R.rbg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    marks = []
    for _ in range(args.steps):
        step()
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()
    for ev_side in side_done:  # the timed region ends when the side stream's reducers and collectives of the last steps are done too
        if ev_side is not None:
            main_stream.wait_event(ev_side)
    e1.record()
    # rbg_trace only queues work, so the steps above are all in the stream long before the GPU is through the first one: the
    # clocks are read now, while it works on them, and no NVML call sits between two steps' launches.
    if clock_reader.inline:
        while True:
            try:
                samples.append(clock_reader())
            except Exception:
                pass
            if e1.query() or len(samples) >= 200:
                break
            time.sleep(0.02)
    barrier()
    stop.set()
    th_clock.join()
    ms = e0.elapsed_time(e1)
    step_ms = [round(a.elapsed_time(b), 3) for a, b in zip([e0] + marks[:-1], marks)]
    launches = R.rbg_launch_count() - launches0
    bm, bn, cm_, cn = profile_reset(R)
    R.rbg_profile_enable(0)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    rays_per_step = n * nang * world
    value = rays_per_step * args.steps / (ms * 1e-3)
    cnt = cnt2[state["last"]]
    counts = cnt.sum(0).cpu().numpy().tolist()
    focused_frac = counts[R.RBG_FOCUSED] / float(sum(counts))

    # ---- e2e: host (pinned) buffers through the same C-ABI call
    e2e = None
    if not args.no_e2e:
        # The user-facing call with HOST buffers: AOpticsManager::TraceNonSequential's path, rbg_trace(on_device = 0), H2D and D2H
        # inside.  The field angles are independent TraceNonSequential calls; --e2e-threads of them are in flight at a time, each
        # from its own host thread on its own scene handle and output buffers (the C ABI is thread-safe across handles), so the
        # pipeline fill and drain of one call overlap the steady state of another.
        T = max(1, min(args.e2e_threads, nang))
        hin = torch.empty((nang, 8, n), dtype=torch.float64).pin_memory()
        hin.copy_(inp.cpu())
        hout = [torch.empty((7, n), dtype=torch.float64).pin_memory() for _ in range(T)]
        hiout = [torch.empty((3, n), dtype=torch.int32).pin_memory() for _ in range(T)]
        scenes_e2e = [scene]
        for _ in range(1, T):
            h = C.c_void_p()
            R.check(R.rbg_scene_create(export.desc_ptr(), local, C.byref(h)))
            scenes_e2e.append(h)

        def host_struct(k, t):
            r = R.rbg_rays()
            r.n, r.on_device = n, 0
            for i, key in enumerate(["x", "y", "z", "t", "dx", "dy", "dz", "lambda_"]):
                setattr(r, key, hin[k, i].data_ptr())
            for i, key in enumerate(["ox", "oy", "oz", "ot", "odx", "ody", "odz"]):
                setattr(r, key, hout[t][i].data_ptr())
            for i, key in enumerate(["status", "last_node", "npoints"]):
                setattr(r, key, hiout[t][i].data_ptr())
            return r

        hs = [host_struct(k, k % T) for k in range(nang)]
        host_counts = torch.zeros(6, dtype=torch.int64)
        errors = []

        def e2e_worker(t, count):
            try:
                o = H.opts(disable_fresnel=1, steps_per_launch=args.steps_per_launch, seed=20180601)
                for k in range(t, nang, T):
                    o.ray_id_offset = (rank * nang + k) * n
                    R.check(R.rbg_trace(scenes_e2e[t], C.byref(o), C.byref(hs[k]), None))
                    if count:
                        c = torch.bincount(hiout[t][0].to(torch.int64), minlength=6)
                        with count_lock:
                            host_counts.add_(c)
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        count_lock = threading.Lock()

        def e2e_step(count=False):
            if T == 1:
                e2e_worker(0, count)
            else:
                ths = [threading.Thread(target=e2e_worker, args=(t, count)) for t in range(T)]
                for th in ths:
                    th.start()
                for th in ths:
                    th.join()
            if errors:
                raise errors[0]

        e2e_step()
        barrier()
        ksteps = min(args.steps, 3)
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        # untimed verification: the host-buffer path returns the same status counts as the device-resident path
        e2e_step(count=True)
        e2e = {"value": rays_per_step * ksteps / float(dt.item()), "unit": "rays/s", "h2d_bytes_per_step": BYTES_IN * n * nang, "d2h_bytes_per_step": BYTES_OUT * n * nang,
               "steps": ksteps, "calls_in_flight": T, "status_counts_match_device_path": bool((host_counts == cnt.sum(0).cpu()).all().item()) if world == 1 else None}
        for h in scenes_e2e[1:]:
            R.rbg_scene_destroy(h)
        del hin, hout, hiout

    variant = R.rbg_scene_kernel_variant(scene).decode()
    R.rbg_scene_destroy(scene)
    del inp, out, iout
    torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N=1): the oracle on all host threads, bounded spread-out sample of the same workload
    cpu = cpu_full = None
    oracle = flags = None
    if rank == 0 and world == 1 and not args.no_cpu:
        oracle, flags = load_cpu_oracle()
        cpu, cpu_full = cpu_rows(oracle, H, export, configs.beam(2, 2.0, n_side=N_SIDE), N_SIDE * N_SIDE, H.opts(disable_fresnel=1), os.cpu_count() or 1, 12.0, flags)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    passes = args.steps * nang
    roof = roofline_of(n * passes, bm, hbm_peak, KFLOP[2]) or {}
    roof.update({"kernel": "k_nav<%s> + k_shade<%s> (the two halves of the wavefront bounce; the first k_nav also locates the start points, k_trace finishes the tail)" % (variant, variant),
                 "peak_source": "MEASURED_PEAKS.json (of measured)" if "hbm_gbs" in peaks else "of fallback",
                 "algorithmic_bytes_per_pass": (BYTES_IN + BYTES_OUT) * n, "bounce_ms_per_pass": bm / passes, "bounce_launches_per_pass": bn / passes,
                 "kernel_ms_per_launch": bm / max(1, bn), "kernel_launches": bn, "kernel_share_of_step": bm / ms if ms else None,
                 "compact": {"ms_per_launch": cm_ / max(1, cn), "launches": cn, "share_of_step": cm_ / ms if ms else None},
                 "note": "achieved = 132 B/ray x rays of a pass / summed CUDA-event time of that pass's bounce launches. The bounce kernels are bound by "
                         "instruction issue / latency, not by HBM: see roofline.fp64 and profiles/ (ncu)"})
    try:  # DRAM bytes per pass from the ncu capture committed with this code (stamped with the commit it was taken at)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic_k_nav.json")))
        roof["traffic"] = tr.get("dram_bytes_per_pass")
        roof["traffic_source"] = tr.get("source")
    except Exception:
        pass

    # ---- every BASELINE config on one GPU (rank 0 reports; under torchrun every rank runs them so the GPUs stay in step)
    per_config = strong = None
    if not args.no_configs:
        per_config = {}
        for cfg in (1, 2, 3, 4, 5):
            per_config["cfg%d" % cfg] = bench_one_config(R, torch, H, configs, dev, local, stream, cfg, 3, hbm_peak,
                                                         oracle if (rank == 0 and world == 1) else None, flags, 4.0, 30_000_000 if not args.no_e2e else 0)
        strong = bench_cfg5_strong(R, torch, H, configs, sharding, dev, local, stream, rank, world, dist, 2 if world < 4 else 3)

    if rank == 0:
        print(json.dumps({
            "metric": "rays/s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "DaviesCotton.C (BASELINE configs[1]): 88 hex facets + camera + masts, 9 field angles 0-4 deg, Square(400 nm, 14 m, n=3334) each",
                       "rays_per_step_per_gpu": n * nang, "rays_per_step": rays_per_step, "l2_policy": "inputs (711 MB per batch) larger than L2, no flush",
                       "cpus_bound_to_gpu_numa_node": numa, "steps_per_launch": args.steps_per_launch, "parallelism": "rays sharded over %d GPU(s), geometry replicated" % world},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "cpu_baseline_full_walk": cpu_full, "clocks": clocks_summary(samples),
            "step_ms": step_ms,
            "check": {"focused_fraction": focused_frac, "status_counts": counts, "d80_cm_by_angle": [round(v, 4) for v in d80_2[(state["k"] - 1) & 1][:, 0].cpu().numpy().tolist()]},
            "configs": per_config, "cfg5_strong": strong,
        }))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
