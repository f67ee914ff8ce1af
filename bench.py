#!/usr/bin/env python
"""bench.py -- NDT scan-matches/sec (HDL-64 scans vs a 1 M-point map) on N B200s of one node.

Workload (BASELINE.json config 4, "batched offline map-matching"): every rank holds the replicated
1 M-point synthetic map (SetInputTarget, NDT res 1.0) and FRAMES synthetic HDL-64 frames (default 4000
per GPU, weak scaling), each down-sampled by VoxelFilter(1.3 m) as the reference does before ScanMatch
(front_end.cpp:106-107, matching.cpp:187-191), with initial guess = truth o perturbation
(U[-0.5,0.5] m, U[-2,2] deg, seed 4000).  One STEP = one pass of the hot path over the whole batch:
NDTRegistration::ScanMatch (step 0.1, eps 0.01, iter 30) for every frame.

  value   matches/s with the filtered sources + guesses resident in HBM (b2ndt_align_batch_device),
          timed with CUDA events on the launching stream, L2 flushed between steps, max over ranks.
  e2e     the same metric through the reference-facing C-ABI call with HOST buffers
          (b2ndt_align_batch: pack -> H2D -> kernel -> D2H of poses + results inside the timed region).
  roofline  the dominant kernel (ndt_match_kernel): algorithmic bytes / event time / measured HBM peak.
  cpu_baseline  the CPU oracle (restatement of pcl::NDT 1.7; PCL is not installable here) on a bounded sample.

`--impl reference` times that CPU implementation alone on all host threads (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ndt_scan_matches_per_sec"
UNIT = "matches/s"
NDT = dict(res=1.0, step_size=0.1, trans_eps=0.01, max_iter=30)
FRAME_LEAF = 1.3


def f32(x):
    return float(np.float32(x))


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML is polled from a thread every ~5 ms with
    wall-clock stamps, and the samples between mark_begin() / mark_end() (the timed steps) are the ones reported;
    without NVML bindings the nvidia-smi loop (-lms 20) of the profiling recipe is used instead."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,"
         "utilization.gpu")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = None
        self.p = None
        self.nv = None
        self.samples = []           # (t, sm_mhz, reasons bitmask)
        self.windows = []           # [t_begin, t_end] of every timed region
        self._stop = False
        self.thread = None
        self.sm_max = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.gpu
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _poll(self):
        nv, h = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop:
            try:
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get_reasons(h))))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            self.nv = self._nvml_handle()
            self.sm_max = float(self.nv[0].nvmlDeviceGetMaxClockInfo(self.nv[1], self.nv[0].NVML_CLOCK_SM))
            import threading
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark_begin(self):
        self.windows.append([time.perf_counter(), None])

    def mark_end(self):
        if self.windows:
            self.windows[-1][1] = time.perf_counter()

    def _stop_nvml(self):
        self._stop = True
        self.thread.join(timeout=2)
        nv = self.nv[0]
        names = (("hw_slowdown", "HwSlowdown"), ("hw_thermal_slowdown", "HwThermalSlowdown"),
                 ("sw_thermal_slowdown", "SwThermalSlowdown"), ("sw_power_cap", "SwPowerCap"))
        masks = {}
        for key, suffix in names:
            for prefix in ("nvmlClocksEventReason", "nvmlClocksThrottleReason"):
                if hasattr(nv, prefix + suffix):
                    masks[key] = int(getattr(nv, prefix + suffix)); break
        inside = [x for x in self.samples if any(b is not None and e is not None and b <= x[0] <= e for b, e in self.windows)]
        use = inside if inside else self.samples
        out = {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": len(self.samples), "source": "nvml"}
        if use:
            bits = 0
            for x in use:
                bits |= x[2]
            out.update(sm_mhz=float(np.median([x[1] for x in use])), reasons=sorted(k for k, m in masks.items() if bits & m),
                       samples_under_load=len(inside))
        return out

    def stop(self):
        if self.nv is not None and self.thread is not None:
            try:
                return self._stop_nvml()
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": len(self.samples), "source": "nvml"}
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvidia-smi"}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, load = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 10:
                continue
            try:
                s_, m_, u_ = float(c[1]), float(c[2]), float(c[9])
            except ValueError:
                continue
            sm.append(s_); mx.append(m_)
            if u_ >= 50.0:
                load.append(s_)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            under = load if load else sm
            out.update(sm_mhz=float(np.median(under)), sm_max_mhz=float(np.max(mx)), reasons=sorted(reasons), samples=len(sm),
                       samples_under_load=len(load))
        return out


def build_workload(frames, seed_base, rank, world, map_points, nthreads, gen_filter):
    """-> scene, map (N,4), list of filtered sources, truth poses (frames,6), guesses (frames,4,4)."""
    from lidar_slam_b200 import synth
    scene = synth.Scene(leg=500.0)
    target = scene.make_map(map_points, 2.0)
    plen = scene.path_length
    # frame k of rank r sits at arclength s along the drive; ranks interleave so every rank covers the whole map
    gidx = np.arange(frames) * world + rank
    s = 5.0 + (plen - 10.0) * (gidx + 0.5) / (frames * world)
    truth = np.stack([scene.path_pose(v) for v in s])
    rng = np.random.default_rng(4000 + rank)
    pert = np.concatenate([rng.uniform(-0.5, 0.5, (frames, 3)), np.deg2rad(rng.uniform(-2.0, 2.0, (frames, 3)))], axis=1)
    guesses = np.stack([synth.pose6_to_matrix(truth[k] + pert[k]).astype(np.float32) for k in range(frames)])
    sources, raw_pts = [], 0
    chunk = 256
    for c0 in range(0, frames, chunk):
        ids = seed_base + gidx[c0:c0 + chunk]
        raws = scene.scans(ids, truth[c0:c0 + chunk], nthreads=nthreads)
        raw_pts += sum(len(r) for r in raws)
        sources += gen_filter(raws)
    return scene, target, sources, truth, guesses, raw_pts


def run_reference(args, rank, world):
    """CPU arm: the oracle (port of pcl::NDT 1.7 as the reference calls it) on all host threads."""
    if rank != 0:
        return
    from oracle import oracle as O
    O.build(ref=False)
    cores = os.cpu_count() or 1
    per_step = max(cores * 2, 16)
    frames = per_step
    t0 = time.time()

    def cpu_filter(raws):
        return [O.voxel_filter(r, FRAME_LEAF, FRAME_LEAF, FRAME_LEAF)[0] for r in raws]

    scene, target, sources, truth, guesses, _ = build_workload(frames, 0x5EED0000, 0, 1, args.map_points, cores, cpu_filter)
    grid = O.Grid(target, NDT["res"])
    prm = O.params(res=NDT["res"], step_size=f32(NDT["step_size"]), trans_eps=f32(NDT["trans_eps"]), max_iter=NDT["max_iter"])
    setup_s = time.time() - t0

    def one(k):
        return O.align(grid, prm, sources[k], guesses[k])["iterations"]

    def step():
        with ThreadPoolExecutor(max_workers=cores) as ex:
            return list(ex.map(one, range(frames)))

    for _ in range(args.warmup):
        step()
    t1 = time.perf_counter()
    its = []
    for _ in range(args.steps):
        its += step()
    dt = time.perf_counter() - t1
    value = frames * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config4 batched scan-to-map NDT: HDL-64 frames (VoxelFilter 1.3) vs %d-pt map, res 1.0 step 0.1 eps 0.01 iter 30"
                   % len(target), "frames_per_step": frames, "mean_iterations": float(np.mean(its)),
                   "mean_source_points": float(np.mean([len(s) for s in sources])), "setup_s": setup_s},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d frames per step, one oracle align per thread over %d threads" % (frames, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=4000, help="frames (matches) per GPU per step")
    ap.add_argument("--map-points", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample", type=int, default=48, help="frames of the CPU baseline sample (rank 0, N=1)")
    ap.add_argument("--batch-cluster", type=int, default=1, help="CTAs per match in batch mode")
    ap.add_argument("--profile-range", action="store_true",
                    help="bracket the timed device-resident steps with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from lidar_slam_b200 import capi
    from lidar_slam_b200.registration import DeviceCloud, NDTRegistration, VoxelFilter

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the registration path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly ONE line (the JSON): anything a library prints to fd 1 meanwhile (NCCL prints its
    # version at communicator creation when NCCL_DEBUG=VERSION) goes to stderr; fd 1 is restored for the result
    sys.stdout.flush()
    saved_stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # a dedicated (non-default) stream: the library launches on it and the CUDA events are recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0

    # ---------------- setup (untimed): map, frames, VoxelFilter on the GPU, device-resident batch -------------
    vf = VoxelFilter(FRAME_LEAF, FRAME_LEAF, FRAME_LEAF, device=local_rank)
    filt_stats = {"raw_pts": 0, "sec": 0.0, "launches": 0}

    def gpu_filter(raws):
        outs = []
        for r in raws:
            t = time.perf_counter()
            l0 = capi.launches()
            outs.append(vf.Filter(r)[1])
            filt_stats["sec"] += time.perf_counter() - t
            filt_stats["launches"] += capi.launches() - l0
            filt_stats["raw_pts"] += len(r)
        return outs

    nthreads = max(1, (os.cpu_count() or 1) // world)
    t_setup = time.time()
    scene, target, sources, truth, guesses, raw_pts = build_workload(args.frames, 0x5EED0000, rank, world, args.map_points,
                                                                     nthreads, gpu_filter)
    B = len(sources)
    reg = NDTRegistration(NDT["res"], NDT["step_size"], NDT["trans_eps"], NDT["max_iter"], device=local_rank)
    reg.SetCluster(16, args.batch_cluster)      # single ScanMatch: up to 16 CTAs per match (the library picks by size)
    t_tgt = time.perf_counter()
    reg.SetInputTarget(target)
    set_target_ms = 1e3 * (time.perf_counter() - t_tgt)
    t_tgt = time.perf_counter()
    reg.SetInputTarget(target)
    set_target_ms = min(set_target_ms, 1e3 * (time.perf_counter() - t_tgt))
    info = reg.TargetInfo()

    # the two other subsystems of the path with everything resident in HBM (device clouds): VoxelFilter of one raw
    # scan and the target-grid build of the map; algorithmic bytes per SURVEY 8(d)
    def timed_ms(fn, reps=10):
        fn()
        ts = []
        for _ in range(reps):
            t = time.perf_counter(); fn(); ts.append(1e3 * (time.perf_counter() - t))
        return float(np.median(ts))

    raw0 = scene.scan(0x5EED0000 + 17, truth[0])
    d_raw, d_filt, d_map = DeviceCloud(raw0, device=local_rank), DeviceCloud(device=local_rank), DeviceCloud(target, device=local_rank)
    vf_ms = timed_ms(lambda: vf.FilterCloud(d_raw, d_filt))
    tg_ms = timed_ms(lambda: reg.SetInputTargetCloud(d_map), reps=5)
    peak_hbm = measured_peaks()[0]
    vf_bytes = 16.0 * len(raw0) + 16.0 * len(d_filt)
    tg_bytes = 16.0 * len(target) + 80.0 * info["n_leaves"]
    resident = {
        "voxel_filter_scan": {"n_in": len(raw0), "n_out": len(d_filt), "ms": vf_ms, "algorithmic_bytes": vf_bytes,
                              "GBps": vf_bytes / vf_ms / 1e6, "frac_of_hbm_peak": vf_bytes / vf_ms / 1e6 / peak_hbm,
                              "note": "one call = ~17 launches + one 32-byte D2H of the count; latency-bound at this size"},
        "set_target_map": {"n_points": len(target), "voxels": info["n_leaves"], "ms": tg_ms, "algorithmic_bytes": tg_bytes,
                           "GBps": tg_bytes / tg_ms / 1e6, "frac_of_hbm_peak": tg_bytes / tg_ms / 1e6 / peak_hbm},
    }
    reg.SetInputTarget(target)       # back to the host-path target (identical grid)
    del d_raw, d_filt, d_map

    cat = np.ascontiguousarray(np.concatenate(sources, axis=0))
    offsets = np.zeros(B + 1, np.uint32)
    offsets[1:] = np.cumsum([len(s) for s in sources])
    n_total = int(offsets[-1])
    g_cm = np.ascontiguousarray(guesses.transpose(0, 2, 1).reshape(B, 16))
    d_src = torch.from_numpy(cat).to(dev)
    d_off = torch.from_numpy(offsets.astype(np.int64)).to(dev).to(torch.int32)   # same bits as uint32
    d_guess = torch.from_numpy(g_cm).to(dev)
    d_pose = torch.zeros((B, 16), dtype=torch.float32, device=dev)
    d_res = torch.zeros((B, capi.RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # > 126 MB L2
    setup_s = time.time() - t_setup

    reg.SetStream(sptr)

    def step_device():
        reg.ScanMatchBatchDevice(d_src.data_ptr(), n_total, d_off.data_ptr(), B, d_guess.data_ptr(), d_pose.data_ptr(), d_res.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing -------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.5)                        # nvidia-smi needs a moment before its first sample
    for _ in range(args.warmup):
        flush.fill_(1.0)
        step_device()
    barrier()
    launches0 = capi.launches()
    evs = []
    barrier()
    if args.profile_range:
        torch.cuda.profiler.start()
    sampler.mark_begin()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(0.0)                      # L2 flush between timed steps (outside the event pair)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    wall_s = time.perf_counter() - wall0
    sampler.mark_end()
    if args.profile_range:
        torch.cuda.profiler.stop()
    gpu_launches = capi.launches() - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = float(np.sum(step_ms))
    res = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=capi.RESULT_DTYPE)
    poses_dev = d_pose.cpu().numpy().reshape(B, 4, 4).transpose(0, 2, 1)

    # ---------------- end to end through the host-buffer C ABI -------------------------------------------
    reg.SetStream(None)
    h2d = n_total * 16 + B * 64 + (B + 1) * 4
    d2h = B * (64 + capi.RESULT_DTYPE.itemsize)
    import ctypes as C
    L = capi.lib()
    # the caller's cloud lives in page-locked host memory (as a ROS/driver ring buffer would): the library
    # then DMAs straight from it
    cat_pin = torch.from_numpy(cat).pin_memory()
    cat = cat_pin.numpy()
    out_pose = np.zeros((B, 16), np.float32)
    out_res = np.zeros(B, capi.RESULT_DTYPE)

    def step_e2e():
        capi.check(L.b2ndt_align_batch(reg._h, cat.ctypes.data, n_total, 16, 12, offsets.ctypes.data_as(C.POINTER(C.c_uint32)), B,
                                       g_cm.ctypes.data_as(C.POINTER(C.c_float)), out_pose.ctypes.data_as(C.POINTER(C.c_float)),
                                       out_res.ctypes.data))

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    sampler.mark_begin()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop()                # the samples inside the device-resident and the end-to-end timed regions
    assert np.array_equal(out_res["iterations"], res["iterations"]), "host and device batch paths disagree"

    # single-match latency through b2ndt_align (p50 over a sample of frames, cluster of 8 CTAs)
    lat = []
    for k in range(3):                          # warm-up: first launch of the cluster kernel loads its module
        reg.ScanMatch(sources[k], guesses[k], want_cloud=False)
    for k in range(0, B, max(1, B // 64)):
        t = time.perf_counter()
        reg.ScanMatch(sources[k], guesses[k], want_cloud=False)
        lat.append(1e3 * (time.perf_counter() - t))

    # ---------------- aggregate over ranks ----------------------------------------------------------------
    t_dev = torch.tensor([dev_ms, e2e_s * 1e3, wall_s * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(B), float(gpu_launches), float(res["passes"].sum()), float(res["pairs"].sum()), float(n_total)],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    max_dev_ms, max_e2e_ms, max_wall_ms = [float(v) for v in t_dev.cpu()]
    all_B = float(tot[0].item())

    if rank == 0:
        peak, peak_src = measured_peaks()
        # roofline of ndt_batch_kernel (one launch per step on this rank): SURVEY 8(d) algorithmic bytes
        from oracle import oracle as O
        O.build(ref=False)
        grid = O.Grid(target, NDT["res"])
        rho = []
        for k in range(0, B, max(1, B // 8)):
            # V_touched of SURVEY 8(d): distinct voxels returned by any neighbour query of one pass (final pose)
            tr = O.transform_points(poses_dev[k], sources[k][:, :3])
            touched_all = set()
            for q in tr:
                touched_all.update(grid.radius_search(q)[0].tolist())
            rho.append(len(touched_all) / max(1, len(sources[k])))
        rho = float(np.mean(rho))
        passes = res["passes"].astype(np.float64)
        npts = np.diff(offsets.astype(np.int64)).astype(np.float64)
        alg_bytes = float(np.sum(passes * (16.0 * npts + 80.0 * rho * npts + 224.0)) + 64.0 * B)
        launch_s = (dev_ms / args.steps) * 1e-3
        achieved = alg_bytes / launch_s / 1e9
        # DRAM traffic of one launch from the committed ncu --set full capture (same command, same frames)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("frames_per_launch") == B:
                traffic = float(tj["dram_bytes_read"] + tj["dram_bytes_write"])
        except Exception:
            pass
        # FP64 work actually issued: ~300 flops per (point, voxel) pair in the Q/P/M formulation (150 DFMA-class
        # operations, csrc/b2_ndt.cu ndt_pair) + ~30 per point and pass for the float transform / cell lookup
        flops = float(np.sum(passes * 30.0 * npts) + 300.0 * res["pairs"].sum())

        # CPU baseline: the oracle, single thread, bounded sample of the same frames
        prm = O.params(res=NDT["res"], step_size=f32(NDT["step_size"]), trans_eps=f32(NDT["trans_eps"]), max_iter=NDT["max_iter"])
        sample = list(range(0, B, max(1, B // max(1, args.cpu_sample))))[:args.cpu_sample]
        cpu_value, parity = None, {}
        if world == 1:
            t0 = time.perf_counter()
            refs = [O.align(grid, prm, sources[k], guesses[k]) for k in sample]
            cpu_s = time.perf_counter() - t0
            cpu_value = len(sample) / cpu_s
            dt = max(float(np.max(np.abs(poses_dev[k][:3, 3] - refs[i]["pose"][:3, 3]))) for i, k in enumerate(sample))
            dr = max(float(np.max(np.abs(res["p"][k][3:] - refs[i]["p"][3:]))) for i, k in enumerate(sample))
            it_eq = all(int(res["iterations"][k]) == refs[i]["iterations"] for i, k in enumerate(sample))
            parity = {"checked": len(sample), "max_dt_m": dt, "max_dr_rad": dr, "iterations_equal": it_eq}

        value = all_B * args.steps / (max_dev_ms * 1e-3)
        e2e_value = all_B * args.steps / (max_e2e_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "config4 batched scan-to-map NDT: %d HDL-64 frames/GPU (VoxelFilter 1.3 m -> mean %.0f pts) vs %d-pt map "
                            "(V=%d voxels, %d searchable), res 1.0 step 0.1 eps 0.01 iter 30, guesses = truth+U[0.5m,2deg]"
                            % (B, n_total / B, info["n_points"], info["n_leaves"], info["n_tree"]),
                "frames_per_gpu": B, "l2": "256 MB flush between timed steps", "batch_cluster": args.batch_cluster,
                "mean_iterations": float(res["iterations"].mean()), "converged_frac": float(res["converged"].mean()),
                "mean_passes": float(res["passes"].mean()), "pairs_per_pass_per_point": float(res["pairs"].sum() / np.sum(passes * npts)),
                "set_target_ms": set_target_ms, "setup_s": setup_s, "device_resident": resident,
                "voxel_filter": {"raw_pts_per_frame": filt_stats["raw_pts"] / max(1, B), "ms_per_frame_host_api": 1e3 * filt_stats["sec"] / max(1, B)},
                "single_match_ms": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)), "n": len(lat)},
                "step_ms": step_ms, "wall_ms_per_step_incl_flush": max_wall_ms / args.steps,
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(tot[1].item()),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "ndt_batch_kernel", "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "v_touched_per_source_point": rho, "fp64_gflops_est": flops / launch_s / 1e9,
                         "note": "working set is L2-resident (DRAM traffic << algorithmic bytes): bound by the L1 data pipe (LSU wavefronts of the gathers + shared-memory rings) and instruction latency, see profiles/README.md"},
            "cpu_baseline": ({"value": cpu_value, "unit": UNIT, "cores": 1, "kind": "port",
                              "sample": "%d of the %d frames, oracle align, 1 thread" % (len(sample), B)} if cpu_value else None),
            "parity": parity,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
