#!/usr/bin/env python
"""bench.py — rays/s of TraceNonSequential on BASELINE.json's headline configuration.

Workload (config.workload): configs[1], DaviesCotton.C — 88-facet Davies-Cotton reflector, 9 field
angles 0..4 deg, each ARayShooter::Square(400 nm, 14 m, n=3334) = 11 115 556 rays, 100 040 004 rays
per step.  One "step" = one TraceNonSequential pass over all nine batches plus the
on-device PSF reducers (histogram + moments per field angle).

  value : rays/s, inputs resident in HBM (rays generated on device by rbg_shoot before the timed region)
  e2e   : rays/s through the C ABI with HOST (pinned) buffers; H2D of the 64 B/ray inputs and D2H of the
          68 B/ray results are inside the timed region
  roofline : k_trace<1>, algorithmic 132 B/ray (64 in + 68 out) / CUDA-event launch time vs measured HBM peak
  cpu_baseline : the CPU oracle (reference algorithm restated; ROOT is unavailable) on all host threads,
          bounded sample of the same workload
Multi-GPU (torchrun): rays shard across ranks (each rank traces its own field-angle sweep, geometry
replicated), no data-path collective; the PSF histograms/moments are all-reduced with NCCL at the end
of each step.  Scaling is weak (per-GPU work fixed).

`--impl reference` times the reference arm: the reference's own implementation cannot be built here
(needs CERN ROOT), so it is the oracle port on all host threads (cpu_baseline.kind = "port").
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


def oracle_rate(oracle, H, export, params, n, threads, opts):
    rays = H.make_rays(oracle, params, 0, n)
    t0 = time.perf_counter()
    H.trace_with(oracle.orc_trace, export, rays, opts, nthreads=threads)
    return n / (time.perf_counter() - t0)


def run_reference(args):
    """reference arm: CPU oracle (port of the reference algorithm) with all host threads"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import helpers as H
    from robast_b200 import configs
    oracle = H.load_oracle()
    mgr, _k = configs.davies_cotton()
    export = mgr.ExportScene()
    threads = os.cpu_count() or 1
    opts = H.opts(disable_fresnel=1)
    rate = oracle_rate(oracle, H, export, configs.beam(2, 0.0, n_side=N_SIDE), 20000, threads, opts)
    per_angle = max(2000, int(rate * 4.0 / len(ANGLES)))  # ~4 s per step
    batches = [H.make_rays(oracle, configs.beam(2, th, n_side=N_SIDE), (N_SIDE * N_SIDE) // 2 - per_angle // 2, per_angle) for th in ANGLES]
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for b in batches:
            H.trace_with(oracle.orc_trace, export, b, opts, nthreads=threads)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    nrays = per_angle * len(ANGLES)
    value = nrays * args.steps / total
    sample = "%d rays per field angle (central rows of each 3334^2 grid) x 9 angles per step" % per_angle
    print(json.dumps({
        "impl": "reference", "metric": "rays/s", "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "DaviesCotton.C 88 hex facets, 9 field angles 0-4 deg, Square(400 nm, 14 m, n=3334) each; bounded CPU sample", "rays_per_step": nrays},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference itself needs CERN ROOT (absent); this is the restated reference algorithm (oracle/) on all host threads",
    }))


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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import robast_b200 as R
    from robast_b200 import configs
    import helpers as H

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if R.rbg_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the tracer has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
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
    hist = torch.zeros((nang, 200 * 200), dtype=torch.int64, device=dev)
    hstats = torch.zeros((nang, 5), dtype=torch.float64, device=dev)
    d80 = torch.zeros((nang, 3), dtype=torch.float64, device=dev)
    d80_all = torch.zeros((world * nang, 3), dtype=torch.float64, device=dev)
    mom = torch.zeros((nang, 8), dtype=torch.float64, device=dev)
    cnt = torch.zeros((nang, 6), dtype=torch.int64, device=dev)
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
        # D80 of each of this rank's field angles (AGeoUtil::ContainmentRadius), then the cross-rank reductions
        R.check(R.rbg_containment_radius(nang, hist.data_ptr(), 200, -4., 10., 200, -7., 7., hstats.data_ptr(), 0.8, d80.data_ptr(), local, stream))
        if dist is not None:
            dist.all_reduce(hist)
            dist.all_reduce(mom)
            dist.all_reduce(cnt)
            dist.all_gather_into_tensor(d80_all, d80)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clock_reader = make_clock_reader()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    R.rbg_profile_enable(1)
    bm, bn, cm_, cn = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
    R.rbg_profile_read(C.byref(bm), C.byref(bn), C.byref(cm_), C.byref(cn))  # reset
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
        if clock_reader.inline:
            try:
                samples.append(clock_reader())
            except Exception:
                pass
    e1.record()
    barrier()
    stop.set()
    th_clock.join()
    ms = e0.elapsed_time(e1)
    step_ms = [round(a.elapsed_time(b), 3) for a, b in zip([e0] + marks[:-1], marks)]
    launches = R.rbg_launch_count() - launches0
    R.rbg_profile_read(C.byref(bm), C.byref(bn), C.byref(cm_), C.byref(cn))
    R.rbg_profile_enable(0)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    rays_per_step = n * nang * world
    value = rays_per_step * args.steps / (ms * 1e-3)
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

    # ---- CPU baseline (rank 0, N=1): the oracle on all host threads, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        oracle = H.load_oracle()
        threads = os.cpu_count() or 1
        o = H.opts(disable_fresnel=1)
        rate = oracle_rate(oracle, H, export, configs.beam(2, 0.0, n_side=N_SIDE), 20000, threads, o)
        ns = int(min(4e6, max(20000, rate * 15.0)))
        first = n // 2 - ns // 2
        rays = H.make_rays(oracle, configs.beam(2, 2.0, n_side=N_SIDE), first, ns)
        t0 = time.perf_counter()
        H.trace_with(oracle.orc_trace, export, rays, o, nthreads=threads)
        dtc = time.perf_counter() - t0
        cpu = {"value": ns / dtc, "unit": "rays/s", "cores": threads, "kind": "port",
               "sample": "%d rays (central rows of the 2.0 deg 3334^2 grid), %.1f s; reference algorithm restated in C++ (oracle/), ROOT unavailable" % (ns, dtc)}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    kernel_ms = bm.value / max(1, bn.value)
    # one bounce launch processes the live rays of one field angle; per TraceNonSequential pass the algorithmic bytes are
    # 132 B/ray (64 in + 68 out), and the bounce kernel runs `launches per pass` times per pass
    passes = args.steps * nang
    achieved = (BYTES_IN + BYTES_OUT) * n * passes / (bm.value * 1e-3) / 1e9 if bm.value > 0 else None
    variant = R.rbg_scene_kernel_variant(scene).decode()
    roof = {"bound": "hbm", "kernel": "k_step<%s> (wavefront bounce kernel; k_trace<%s> finishes the tail)" % (variant, variant), "achieved": achieved, "peak": hbm_peak,
            "unit": "GB/s", "frac": achieved / hbm_peak if achieved else None,
            "traffic": None, "peak_source": "MEASURED_PEAKS.json (of measured)" if "hbm_gbs" in peaks else "of fallback", "algorithmic_bytes_per_ray": BYTES_IN + BYTES_OUT,
            "algorithmic_bytes_per_pass": (BYTES_IN + BYTES_OUT) * n, "bounce_ms_per_pass": bm.value / passes, "bounce_launches_per_pass": bn.value / passes,
            "kernel_ms_per_launch": kernel_ms, "kernel_launches": bn.value, "kernel_share_of_step": bm.value / ms if ms else None,
            "compact": {"ms_per_launch": cm_.value / max(1, cn.value), "launches": cn.value, "share_of_step": cm_.value / ms if ms else None},
            "note": "achieved = 132 B/ray x rays of a pass / summed CUDA-event time of that pass's bounce launches. The bounce kernel is bound by FP64/"
                    "instruction issue, not HBM: see roofline.fp64 and profiles/"}
    for fn, key in (("traffic_k_step.json", "traffic"),):
        try:
            roof[key] = json.load(open(os.path.join(ROOT, "profiles", fn))).get("dram_bytes_per_pass")
        except Exception:
            pass
    try:
        fp = json.load(open(os.path.join(ROOT, "profiles", "fp64_roofline.json")))
        roof["fp64"] = fp
    except Exception:
        pass

    if rank == 0:
        print(json.dumps({
            "metric": "rays/s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "DaviesCotton.C (BASELINE configs[1]): 88 hex facets + camera + masts, 9 field angles 0-4 deg, Square(400 nm, 14 m, n=3334) each",
                       "rays_per_step_per_gpu": n * nang, "rays_per_step": rays_per_step, "l2_policy": "inputs (711 MB per batch) larger than L2, no flush",
                       "steps_per_launch": args.steps_per_launch, "parallelism": "rays sharded over %d GPU(s), geometry replicated" % world},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clocks_summary(samples), "step_ms": step_ms,
            "check": {"focused_fraction": focused_frac, "status_counts": counts, "d80_cm_by_angle": [round(v, 4) for v in d80[:, 0].cpu().numpy().tolist()]},
        }))
    R.rbg_scene_destroy(scene)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
