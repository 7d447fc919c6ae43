#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 render loop (BASELINE.json metric: Mrays/s, avenger Whitted 1080p).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one frame = one pass of the hot path (pg1/simpleguidx11.cpp:95-118) over every pixel of the frame.
Workload (N=1 default) = BASELINE.json configs[1]: avenger (stand-in mesh, the reference's OBJ is absent) 1920x1080,
Whitted, depth cut-off 10, 1 spp un-jittered.  For N>1 the same frame is cut into 32x8 tiles dealt round-robin to
the ranks (strong scaling: total work fixed); every rank's kernels store its tiles straight into rank 0's frame over
NVLink (fallback: NCCL gather); completion is one counter per slot in rank 0's memory that every rank's finished frame adds 1
to, waited for with a stream memory operation (no collective per step).  The library picks the scheduler by the size of a
rank's share (pgrt.h, scheduler 3): hybrid from 400 k samples, the fused frame kernel below (8 GPUs).

value  : rays/s of K whole frames, scene resident in HBM: the Producer loop renders frames forever
         (pg1/simpleguidx11.cpp:95-125), so the run is one continuous pipeline (--inflight frames in flight per GPU, own
         stream each, pgrt_render*_begin / pgrt_render_end; L2 flushed before every step on that step's stream, a fill of
         1.125 x L2) cut into consecutive windows of exactly K steps; a CUDA event on the consumer stream closes every
         window, the slowest rank counts per window, and the MEAN of the windows between the first (which fills the pipeline) and the last
         (which drains it) is reported: frames complete in bursts of --inflight, windows alternate between two values, a median would be arbitrary.
         Enough windows run to cover --min-seconds (1 s) of device time, with the clocks sampled throughout.
e2e    : the same metric through the host-buffer C-ABI call every step (pgrt_set_camera + pgrt_render_begin into pinned
         host memory + pgrt_render_end), wall clock, float frame (e2e) and 8-bit frame (e2e_rgba8).  N>1: the frames live
         in host memory shared by all ranks and every rank's tiles cross its own PCIe link (dist.ShardedRenderer mode "host").
clocks : SM clock and throttle reasons sampled through NVML every 2 ms inside the timed region.
--impl reference : the CPU restatement of the reference's loop (oracle/; the reference itself cannot be built here)
                   on all host cores, same config, same metric.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

# one hardware channel per frame-slot stream (the driver's default of 8 makes waiting streams hold back their channel-mates);
# must be in the environment before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s (prim+secondary), avenger Whitted 1080p"
UNIT = "Mrays/s"


def workload(name: str):
    """Scene + params + config description of a named BASELINE.json config."""
    from pgi_raytracing_b200 import scenes

    if name == "c1":
        sc = scenes.avenger_proxy()
        p = dict(sampling_width=1, jitter=0, aperture=0.0, max_depth=7)
        desc = "C1 avenger(stand-in) 640x480 Whitted depth7 1spp"
    elif name == "c2":
        sc = scenes.avenger_proxy()
        sc.camera = scenes.Camera(1920, 1080, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)
        p = dict(sampling_width=1, jitter=0, aperture=0.0, max_depth=10)
        desc = "C2 avenger(stand-in) 1920x1080 Whitted depth10 1spp"
    elif name == "c3":
        sc = scenes.avenger_proxy()
        sc.camera = scenes.Camera(3840, 2160, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)
        p = dict(sampling_width=8, jitter=1, aperture=5.0, focal_distance=200.0, max_depth=7, seed=1)
        desc = "C3 avenger(stand-in) 3840x2160 thin-lens f200 a5 64spp depth7"
    elif name == "c4":
        sc = scenes.palm_grove(n_palms=600)
        p = dict(sampling_width=1, jitter=0, aperture=0.0, max_depth=7, shader_mode=1)
        desc = "C4 palm grove (stand-in for PalmTrees) 1920x1080 Lambert 1spp"
    elif name == "c5":
        sc = scenes.triangle_soup(10_000_000, seed=1)
        p = dict(sampling_width=2, jitter=1, aperture=0.0, max_depth=7, seed=1, shader_mode=1)
        desc = "C5 soup 10M tris 3840x2160 Lambert 4spp"
    else:
        raise SystemExit(f"unknown workload {name}")
    return sc, p, desc


def algorithmic_bytes_per_ray(n_tris: int) -> float:
    """SURVEY.md section 8(d) contract figure: 32 (ray in) + 16 (hit out) + 80*D + 4*48, D = ceil(log8(N/4))."""
    d = max(1, math.ceil(math.log(max(n_tris, 8) / 4.0, 8)))
    return 48.0 + 80.0 * d + 192.0


def flop_per_ray(n_tris: int) -> float:
    """8 slab tests x 24 flop per node + 4 Moeller-Trumbore x 45: F_ray = 192*D + 180."""
    d = max(1, math.ceil(math.log(max(n_tris, 8) / 4.0, 8)))
    return 192.0 * d + 180.0


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  In-process NVML every
    2 ms (a C2 step is 0.4 ms: the whole timed region of a default run is shorter than one `nvidia-smi -lms 100` tick);
    `nvidia-smi` polling is the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index: int, uuid=None):
        self.index, self.uuid, self.rows, self.proc = index, uuid, [], None
        self.nvml, self.handle, self.samples, self.reasons, self.max_mhz = None, None, [], set(), None
        self._stop = threading.Event(); self._thread = None

    def _nvml_open(self):
        import pynvml
        pynvml.nvmlInit()
        h = None
        if self.uuid is not None:
            for u in (f"GPU-{self.uuid}", str(self.uuid)):
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(u.encode() if isinstance(u, str) else u); break
                except Exception:
                    h = None
        if h is None:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        self.nvml, self.handle = pynvml, h
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

    def _nvml_loop(self):
        n, h = self.nvml, self.handle
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                mask = int(reasons_fn(h))
                for name, bit in self.BITS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                break
            self._stop.wait(0.002)

    def start(self):
        try:
            self._nvml_open()
            self._thread = threading.Thread(target=self._nvml_loop, daemon=True); self._thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.nvml is not None:
            self._stop.set(); self._thread.join(timeout=1.0)
            sm = self.samples
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml, 2 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def host_threads() -> int:
    """Threads the CPU arm may use: the cores this process may run on, whatever OMP_NUM_THREADS says (torch.distributed.run
    exports OMP_NUM_THREADS=1 to every rank, which would otherwise turn the all-cores baseline into a one-core one)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference(args):
    """CPU arm: the oracle restatement of the reference loop on all host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle.oracle import Oracle, make_params

    sc, p, desc = workload(args.workload)
    orc = Oracle(sc)
    cores = host_threads()
    W, H = sc.camera.width, sc.camera.height
    # bounded sample: a centred horizontal band of the frame sized to keep each step around a second
    S = p["sampling_width"] ** 2
    rows = H if W * H * S <= 4_000_000 else max(8, int(4_000_000 / (W * S)) // 8 * 8)
    pr = make_params(**p)
    # ... and the whole run around a minute and a half whatever the host: one untimed calibration step, then the band
    # shrinks if (warmup + steps) of it would take longer
    y0 = (H - rows) // 2
    t0 = time.perf_counter()
    orc.render(pr, want_ids=False, threads=cores, region=(0, y0, W, y0 + rows))
    t_step = time.perf_counter() - t0
    budget = 90.0
    if t_step * (args.steps + args.warmup) > budget:
        rows = max(8, int(rows * budget / (t_step * (args.steps + args.warmup))) // 8 * 8)
    y0 = (H - rows) // 2
    region = (0, y0, W, y0 + rows)
    for _ in range(args.warmup):
        orc.render(pr, want_ids=False, threads=cores, region=region)
    t0 = time.perf_counter(); rays = 0
    for _ in range(args.steps):
        _, _, _, st = orc.render(pr, want_ids=False, threads=cores, region=region)
        rays += st["total"]
    dt = time.perf_counter() - t0
    v = rays / dt / 1e6
    sample = f"rows {y0}..{y0 + rows} of {H} ({rows * W} px x {S} spp) per step, {cores} host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "note": "CPU restatement of the reference loop (oracle/pg_oracle.cpp) with its own SAH BVH in place of Embree (binary not vendored); "
                                             "bit-identical to the reference's own pg1/*.cpp object code above the rtcIntersect1 boundary (oracle/_ref, tests/test_oracle_vs_ref.py)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(sc, p, budget_s=12.0):
    """Oracle timed on the host cores on a bounded sample of the same workload."""
    from oracle.oracle import Oracle, make_params

    orc = Oracle(sc)
    cores = host_threads()
    W, H = sc.camera.width, sc.camera.height
    S = p["sampling_width"] ** 2
    rows = H if W * H * S <= 4_000_000 else max(8, int(4_000_000 / (W * S)) // 8 * 8)
    y0 = (H - rows) // 2
    region = (0, y0, W, y0 + rows)
    pr = make_params(**p)
    orc.render(pr, want_ids=False, threads=cores, region=region)
    t0 = time.perf_counter(); rays = 0; n = 0
    while True:
        _, _, _, st = orc.render(pr, want_ids=False, threads=cores, region=region)
        rays += st["total"]; n += 1
        if time.perf_counter() - t0 > budget_s or n >= 1000:
            break
    dt = time.perf_counter() - t0
    # one-thread figure (the reference as shipped renders on one thread, pg1/simpleguidx11.cpp:104) on a thin band
    band = (0, H // 2 - 4, W, H // 2 + 4)
    t1 = time.perf_counter()
    _, _, _, st1 = orc.render(pr, want_ids=False, threads=1, region=band)
    dt1 = time.perf_counter() - t1
    return {"value": rays / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} x rows {y0}..{y0 + rows} of {H} at {W} px x {S} spp, {cores} host threads; single-thread {st1['total'] / dt1 / 1e6:.3f} Mrays/s on 8 rows"}


def ncu_traffic(workload_name: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed summary of the
    `ncu --set full` capture of this build (profiles/r2_ncu_traffic.json, written by tools/ncu_summary.py); None if absent."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(workload_name)
    except Exception:
        return None


def frame_hash(t) -> str:
    """NaN-safe content hash of a frame tensor (the bit patterns, not the values: NaN pixels of negative Phong sums count)."""
    import hashlib
    return hashlib.sha256(t.contiguous().cpu().numpy().tobytes()).hexdigest()[:16]


def run_ours(args):
    import torch
    import torch.distributed as dist

    from pgi_raytracing_b200 import raytracer_for, default_params
    from pgi_raytracing_b200.dist import ShardedRenderer

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render loop has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sc, p, desc = workload(args.workload)
    rt = raytracer_for(sc, device=local)
    params = default_params(**p)
    # frames in flight: the Producer loop renders frames forever (pg1/simpleguidx11.cpp:95-125); a frame whose secondary-ray
    # chains are still running leaves most of the GPU to the next one
    shard_samples = sc.camera.width * sc.camera.height * p.get("sampling_width", 1) ** 2 / world
    auto_depth = 16 if shard_samples >= 4e5 else 32     # profiles/r2_scale_probe_n8.txt, r2_sweep_shard_depth.txt, r2_sched_by_shard.txt
    depth = max(1, min(args.inflight if args.inflight > 0 else auto_depth, 32))
    K, W = args.steps, max(args.warmup, 3)
    sr = ShardedRenderer(rt, rank, world, dev, depth=depth)
    FLUSH_BYTES = int(torch.cuda.get_device_properties(dev).L2_cache_size * 1.125) // 4096 * 4096   # a fill 12.5 % larger than L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def l2_flush(k):
        def f(stream):
            rt.flush_l2(k % depth, FLUSH_BYTES, k & 0xFF)   # a fill larger than L2 on that frame's own stream
        return f

    host_issue = [0.0, 0]

    def run_windows(renderer, n_windows, k_per_window):
        """n_windows x k_per_window frames as ONE continuous pipeline (`depth` in flight, every frame preceded by an L2 flush on
        its own stream), bracketed by barrier + synchronize; a CUDA event on the consumer stream marks the completion of the
        last frame of every window.  Returns (device ms per window, rays of this rank, host seconds between the brackets)."""
        comm = torch.cuda.current_stream()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(n_windows + 1)]
        n = n_windows * k_per_window
        barrier()
        h_start = time.perf_counter()
        marks[0].record(comm)
        for s in renderer.slot_streams:
            s.wait_event(marks[0])
        rays = 0
        for k in range(n):
            if k >= depth:
                rays += renderer.end(k - depth)["total"]
            h0 = time.perf_counter()
            renderer.begin(k, params, before=l2_flush(k))      # leaves the consumer stream waiting for frame k
            host_issue[0] += time.perf_counter() - h0; host_issue[1] += 1
            if (k + 1) % k_per_window == 0:
                marks[(k + 1) // k_per_window].record(comm)
        for k in range(max(0, n - depth), n):
            rays += renderer.end(k)["total"]
        barrier()
        h_stop = time.perf_counter()
        return [marks[i].elapsed_time(marks[i + 1]) for i in range(n_windows)], rays, h_stop - h_start

    # warm-up: every slot allocates its queues and captures its frame graph on first use; also the estimate that sizes the run
    win, _, _ = run_windows(sr, 2, max(W, 2 * depth))
    est = torch.tensor([max(win[1] / max(W, 2 * depth), 1e-3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)      # every rank must run the same number of frames: one estimate for all
    est_ms = float(est.item())
    R = int(min(200, max(5, math.ceil(args.min_seconds * 1e3 / (K * est_ms)))))    # windows of exactly K steps, >= min_seconds in total
    host_issue[0] = 0.0; host_issue[1] = 0
    sr.host_s, sr.host_n = [0.0, 0.0, 0.0, 0.0], 0
    sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(dev), "uuid", None)); sampler.start()
    launches0 = rt.kernel_launches()
    win, rays, timed_s = run_windows(sr, R, K)
    launches = rt.kernel_launches() - launches0
    host_parts = [x / max(sr.host_n, 1) * 1e6 for x in sr.host_s]
    clocks = sampler.stop()
    wt = torch.tensor(win, dtype=torch.float64, device=dev)
    tot = torch.tensor([float(rays), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(wt, op=dist.ReduceOp.MAX)          # every window: the slowest rank's device time
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    win = [float(x) for x in wt.tolist()]
    rays_per_frame = float(tot[0].item()) / (R * K)
    # The first window also fills the pipeline; the others are averaged: frames complete in bursts of `depth` (the GPU works
    # through the frames in flight breadth-first), so windows of K = 20 frames alternate between two values (C2, one GPU: 0.305
    # and 0.451 ms per frame, profiles/r2_bench_windows.txt) and their MEDIAN is arbitrary (0.315 in one run, 0.395 in the
    # next); the mean = device time of all steady windows / their frames is what repeats (0.3787, 0.3783).
    # (the last window, which sees the pipeline drain with nothing new competing, is left out like the first)
    steady = sorted(win[1:-1] if len(win) >= 4 else win[1:])
    window_ms = sum(steady) / len(steady)
    ms_per_step = window_ms / K
    value = rays_per_frame / (ms_per_step * 1e-3) / 1e6

    # the frame the other ranks helped to render must be the frame one GPU renders (rank 0, untimed)
    frame_equal = None
    if world > 1:
        barrier()
        sr.begin(0, params); sr.end(0)
        torch.cuda.current_stream().synchronize()
        barrier()
        if rank == 0:
            h_sharded = frame_hash(sr.frames[0])
            rt.set_shard(0, 1)
            solo = torch.zeros((rt.height, rt.width, 4), dtype=torch.float32, device=dev)
            torch.cuda.synchronize()
            rt.render_device(solo.data_ptr(), params)
            torch.cuda.synchronize()
            frame_equal = bool(h_sharded == frame_hash(solo))
            rt.set_shard(rank, world)
        barrier()

    # roofline of the dominant kernel: the same frames one at a time, the frame kernel bracketed by CUDA events on its stream
    trace_ms = 0.0; trace_launches = 0; prof_rays = 0; prof_frame_ms = 0.0
    n_prof = 10
    for k in range(n_prof):
        sr.begin(k, params, profile=1, before=l2_flush(k))
        st = sr.end(k)
        trace_ms += st["trace_ms"]; trace_launches += st["trace_launches"]; prof_rays += st["total"]; prof_frame_ms += st["frame_ms"]
    sr.begin(0, params, profile=3); st_count = sr.end(0)      # instrumented traversal: nodes fetched / triangles tested per ray
    barrier()
    l2 = rt.l2_bandwidth() if rank == 0 else None

    # e2e: the host-buffer C-ABI call every step (camera re-sent, frame in pinned host memory when the step ends), `depth` in flight
    c = sc.camera
    e2e_extra = {}
    if world == 1:
        def e2e_run(rgba8):
            if rgba8:
                hosts = [torch.empty((rt.height, rt.width, 4), dtype=torch.uint8).pin_memory() for _ in range(depth)]
            else:
                hosts = [torch.empty((rt.height, rt.width, 4), dtype=torch.float32).pin_memory() for _ in range(depth)]

            def frames(n):
                e_rays = 0
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for k in range(n):
                    if k >= depth:
                        e_rays += rt.render_end((k - depth) % depth)["total"]
                    rt.flush_l2(k % depth, FLUSH_BYTES, k & 0xFF)
                    rt.set_camera(c.width, c.height, c.fov_y, c.view_from, c.view_at)
                    rt.render_begin(k % depth, params, host_ptr=hosts[k % depth].data_ptr(), rgba8=rgba8)
                for k in range(max(0, n - depth), n):
                    e_rays += rt.render_end(k % depth)["total"]
                torch.cuda.synchronize()
                return time.perf_counter() - t0, e_rays

            frames(2 * depth)
            t, r = frames(R * K)
            return {"value": r / t / 1e6, "unit": UNIT, "h2d_bytes_per_step": 4 * (2 + 1 + 3 + 3) + 64,
                    "d2h_bytes_per_step": rt.width * rt.height * (4 if rgba8 else 16), "ms_per_step": t / (R * K) * 1e3,
                    "frame": "R8G8B8A8_UNORM (what the reference presents, pg1/simpleguidx11.cpp:229)" if rgba8 else "float RGBA (the reference's tex_data_, pg1/simpleguidx11.cpp:108-114)",
                    "frame_hash": frame_hash(hosts[(R * K - 1) % depth])}

        rt.set_shard(0, 1)
        e2e = e2e_run(False)
        e2e_extra["e2e_rgba8"] = e2e_run(True)
    else:
        # The frames live in host memory shared by all ranks (memfd registered with every device): each rank's frame kernel
        # stores its tiles through its own PCIe link.  Fallback: rank 0 copies the NVLink-gathered frame out.
        def e2e_run(rgba8):
            try:
                sr_e = ShardedRenderer(rt, rank, world, dev, depth=depth, mode="host", rgba8=rgba8)
            except (RuntimeError, ValueError):
                if rgba8:
                    return None
                sr_e = sr
            host = torch.empty((rt.height, rt.width, 4), dtype=torch.float32).pin_memory() if rank == 0 and sr_e is sr else None

            def frames(n):
                barrier()
                t0 = time.perf_counter(); rays = 0
                for k in range(n):
                    if k >= depth:
                        rays += sr_e.end(k - depth)["total"]
                    sr_e.begin(k, params, before=l2_flush(k))
                    if host is not None:
                        host.copy_(sr.frames[k % depth], non_blocking=True)   # on the consumer stream, behind that frame's completion
                for k in range(max(0, n - depth), n):
                    rays += sr_e.end(k)["total"]
                barrier()
                return time.perf_counter() - t0, rays

            frames(2 * depth)           # new destinations: let every slot re-capture its frame graph outside the timed region
            t, e_rays = frames(R * K)
            dt = torch.tensor([t], dtype=torch.float64, device=dev); er = torch.tensor([float(e_rays)], dtype=torch.float64, device=dev)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX); dist.all_reduce(er, op=dist.ReduceOp.SUM)
            out = {"value": float(er.item()) / float(dt.item()) / 1e6, "unit": UNIT, "h2d_bytes_per_step": 4 * (2 + 1 + 3 + 3) + 64,
                   "d2h_bytes_per_step": rt.width * rt.height * (4 if rgba8 else 16), "ms_per_step": float(dt.item()) / (R * K) * 1e3,
                   "frame": "R8G8B8A8_UNORM" if rgba8 else "float RGBA", "completion": sr_e.completion,
                   "path": "frames in shared host memory, every rank stores its tiles through its own PCIe link" if sr_e is not sr
                           else "NVLink gather to rank 0, device-to-host copy on rank 0"}
            if sr_e is not sr:
                if rank == 0:
                    out["frame_hash"] = frame_hash(sr_e.frames[(R * K - 1) % depth])   # the host frame is read, not just written
                sr_e.close()
            return out

        e2e = e2e_run(False)
        r8 = e2e_run(True)
        if r8 is not None:
            e2e_extra["e2e_rgba8"] = r8

    # the other configurations BASELINE.json names for scaling (C3, C5), one short measurement each at the same N
    extra = {}
    if args.extra and args.workload == "c2":
        for name, frames_n in (("c3", 6), ("c5", 6)):
            try:
                extra[name] = run_extra(name, frames_n, rank, world, local, dev, barrier)
            except Exception as e:      # never lose the headline line to an extra
                extra[name] = {"error": str(e)[:200]}

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        b_ray = algorithmic_bytes_per_ray(sc.ntris)
        achieved = prof_rays * b_ray / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else None
        frame_bytes = rt.width * rt.height * 16
        scene_bytes = sc.ntris * 48 + rt.build_stats["nodes"] * rt.build_stats["node_bytes"]
        traffic = ncu_traffic(args.workload)
        node_bytes = rt.build_stats["node_bytes"]
        req_bytes = st_count["nodes_visited"] * node_bytes + st_count["tris_tested"] * 48       # what the traversal asks L1/L2 for, per frame
        k_ms = trace_ms / max(trace_launches, 1)
        roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": (achieved / peaks["hbm_gbs"]) if achieved else None,
                "traffic": (traffic.get("k_frame_dram_bytes" if trace_launches == n_prof else "hybrid_trace_kernels_dram_bytes") if traffic else None),
                "kernel": ("k_frame (the whole of trace() for every sample: closest hit, shading, shadow and secondary rays, resolve)" if trace_launches == n_prof else
                           "k_trace + k_phong + k_frame (hybrid scheduler: closest-hit and shadow queries of level 0 as wavefront kernels, every ray of level >= 1 through the pool)"),
                "scheduler": "fused" if trace_launches == n_prof else "hybrid", "bytes_per_ray": b_ray,
                "launch_ms_avg": k_ms, "launches_per_frame": trace_launches / max(n_prof, 1), "peak_kind": peak_kind,
                "frame_ms_unpipelined": prof_frame_ms / max(n_prof, 1),
                "achieved_pipelined": value / world * 1e6 * b_ray / 1e9, "frac_pipelined": value / world * 1e6 * b_ray / 1e9 / peaks["hbm_gbs"],
                "algorithmic_min_frame_bytes": frame_bytes + scene_bytes,
                "whole_frame_dram_bytes": (traffic.get("frame_dram_bytes" if trace_launches == n_prof else "hybrid_frame_dram_bytes") if traffic else None),
                "traversal": {"nodes_per_ray": st_count["nodes_visited"] / max(st_count["total"], 1), "tris_per_ray": st_count["tris_tested"] / max(st_count["total"], 1),
                              "requested_bytes_per_frame": req_bytes, "requested_gbs": req_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None},
                "l2": {"peak_gbs": l2, "requested_frac_of_l2_peak": (req_bytes / (k_ms * 1e-3) / 1e9 / l2) if (l2 and k_ms > 0) else None,
                       "note": "peak = measured read bandwidth of a 32 MiB L2-resident buffer from all SMs (pgrt_debug_l2_bandwidth); requested = node and triangle "
                               "bytes the traversal loads per frame (instrumented run) over the frame kernel's time: the working set is L1/L2-resident, so this, "
                               "not HBM, is the memory roof that applies"},
                "fp32": {"flop_per_ray": flop_per_ray(sc.ntris), "achieved_tflops": value / world * 1e6 * flop_per_ray(sc.ntris) / 1e12,
                         "peak_tflops": 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12,
                         "note": "SURVEY 8(d) contract figure F_ray = 192*D + 180 on the per-GPU pipelined rate; peak = 148 SM x 128 lanes x 2 x max SM clock"},
                "note": "achieved = SURVEY 8(d) contract bytes (720 B/ray at this size) x rays of rank 0 over the traversal kernels' own time (fused scheduler: k_frame; hybrid: k_trace + k_phong + k_frame, summed), one frame at a time "
                        "(CUDA events on the launching stream, L2 flushed before each frame); *_pipelined = the same bytes over the per-GPU share of `value`"}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": {"workload": desc, "triangles": sc.ntris, "rays_per_frame": rays_per_frame,
                                                "parallelism": f"tiles32x8/rr x{world}", "frames_in_flight": depth, "gather": sr.mode, "completion": sr.completion,
                                                "timing": f"mean of {len(steady)} consecutive windows of {K} steps each in one continuous pipeline (the first window, which fills it, and the last, which drains it, are left out; frames complete in bursts of frames_in_flight, so single windows alternate between two values and the median is printed only for reference), "
                                                          f"CUDA events on the consumer stream, max over ranks per window; {timed_s:.2f} s between the barriers",
                                                "windows_ms_per_step": {"first": win[0] / K, "min": steady[0] / K, "median": steady[len(steady) // 2] / K, "mean": ms_per_step, "max": steady[-1] / K,
                                                                        "in_order": [round(x / K, 4) for x in win]},
                                                "host_issue_us_per_step": host_issue[0] / max(host_issue[1], 1) * 1e6,
                                                "host_issue_parts_us": dict(zip(("flush", "render_begin", "completion", "event"), host_parts)),
                                                "l2": f"flushed before every timed step on that step's stream ({FLUSH_BYTES >> 20} MiB fill = 1.125 x L2)",
                                                "bvh": rt.build_stats},
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(tot[1].item()), "roofline": roof}
        out.update(e2e_extra)
        if frame_equal is not None:
            out["frame_equal_to_1gpu"] = frame_equal
        if extra:
            out["config"]["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(sc, p)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_extra(name, n_frames, rank, world, local, dev, barrier):
    """A short measurement of another BASELINE.json configuration at the same N (own context; one continuous pipeline of
    n_frames frames after as many warm-up frames).  Returns rank 0's summary."""
    import torch
    import torch.distributed as dist

    from pgi_raytracing_b200 import raytracer_for, default_params
    from pgi_raytracing_b200.dist import ShardedRenderer

    sc, p, desc = workload(name)
    rt = raytracer_for(sc, device=local)
    params = default_params(**p)
    depth = 2
    sr = ShardedRenderer(rt, rank, world, dev, depth=depth)

    def frames(n):
        comm = torch.cuda.current_stream()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(comm)
        for s in sr.slot_streams:
            s.wait_event(e0)
        rays = 0
        for k in range(n):
            if k >= depth:
                rays += sr.end(k - depth)["total"]
            sr.begin(k, params)
        for k in range(max(0, n - depth), n):
            rays += sr.end(k)["total"]
        e1.record(comm)
        barrier()
        return e0.elapsed_time(e1), rays

    frames(depth + 1)
    ms, rays = frames(n_frames)
    t = torch.tensor([ms, float(rays)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, rays = float(mx[0].item()), float(t[1].item())
    out = {"workload": desc, "value": rays / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_frame": ms / n_frames, "frames": n_frames, "frames_in_flight": depth,
           "triangles": sc.ntris, "build_ms": rt.build_stats["build_ms"], "node_bytes": rt.build_stats["node_bytes"], "completion": sr.completion,
           "l2": "not flushed between frames (inputs per frame exceed L2 for c5; c3 writes 530 M samples per frame)"}
    sr.close()
    rt.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", dest="extra", action="store_false", help="skip the short C3 / C5 measurements appended to config.extra")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="the timed windows of --steps frames are repeated until they cover this much device time")
    ap.add_argument("--inflight", type=int, default=0, help="frames in flight per GPU (1 = one frame at a time; 0 = 16 for a frame or shard of 0.4 M primary samples or more, 32 below)")
    ap.add_argument("--watchdog", type=float, default=900.0, help="seconds after which a run that has not finished dumps every thread's stack and exits (a hang must not sit on a GPU box)")
    args = ap.parse_args()
    import faulthandler
    faulthandler.enable()
    if args.watchdog > 0:
        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
