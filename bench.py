#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 render loop (BASELINE.json metric: Mrays/s, avenger Whitted 1080p).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one frame = one pass of the hot path (pg1/simpleguidx11.cpp:95-118) over every pixel of the frame.
Workload (N=1 default) = BASELINE.json configs[1]: avenger (stand-in mesh, the reference's OBJ is absent) 1920x1080,
Whitted, depth cut-off 10, 1 spp un-jittered.  For N>1 the same frame is cut into 32x8 tiles dealt round-robin to
the ranks (strong scaling: total work fixed); every rank's resolve kernel stores its tiles straight into rank 0's frame over
NVLink (fallback: NCCL gather), a 4-byte NCCL all-reduce per step is the completion barrier.

value  : rays/s of K whole frames, scene resident in HBM, device time between two CUDA events around all K steps; the
         Producer loop renders frames forever (pg1/simpleguidx11.cpp:95-125), so --inflight frames (default 8; 16 for
         frames or shards below 1 M primary samples) are in flight per GPU (own stream each, pgrt_render*_begin /
         pgrt_render_end); L2 is flushed before every step on that step's stream (a fill of 1.125 x L2).
e2e    : the same metric through the host-buffer C-ABI call every step (pgrt_set_camera + pgrt_render_begin into pinned
         host memory + pgrt_render_end), wall clock.  N>1: the frames live in host memory shared by all ranks and every
         rank's tiles cross its own PCIe link (dist.ShardedRenderer mode "host").
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s (prim+secondary), avenger Whitted 1080p"
# dram__bytes_read.sum + dram__bytes_write.sum of one C2 frame's traversal launches, from the committed ncu --set full capture
# (profiles/r1_ncu_bvh8_f32w_k_trace_k_secondary.txt: k_trace 7.82 MB + 0.43 MB, k_secondary 10.87 MB + 0.02 MB; k_phong is negligible).
# The algorithmic figure for the same launches is 2.84 M rays x 720 B = 2.05 GB: the working set lives in L1/L2, not in HBM.
NCU_TRAFFIC_BYTES = int(7.819264e6 + 428544 + 10.873856e6 + 17920)
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


def run_reference(args):
    """CPU arm: the oracle restatement of the reference loop on all host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle.oracle import Oracle, make_params

    sc, p, desc = workload(args.workload)
    orc = Oracle(sc)
    cores = orc.max_threads()
    W, H = sc.camera.width, sc.camera.height
    # bounded sample: a centred horizontal band of the frame sized to keep each step around a second
    S = p["sampling_width"] ** 2
    rows = H if W * H * S <= 4_000_000 else max(8, int(4_000_000 / (W * S)) // 8 * 8)
    pr = make_params(**p)
    # ... and the whole run around a minute and a half whatever the host: one untimed calibration step, then the band
    # shrinks if (warmup + steps) of it would take longer
    y0 = (H - rows) // 2
    t0 = time.perf_counter()
    orc.render(pr, want_ids=False, threads=0, region=(0, y0, W, y0 + rows))
    t_step = time.perf_counter() - t0
    budget = 90.0
    if t_step * (args.steps + args.warmup) > budget:
        rows = max(8, int(rows * budget / (t_step * (args.steps + args.warmup))) // 8 * 8)
    y0 = (H - rows) // 2
    region = (0, y0, W, y0 + rows)
    for _ in range(args.warmup):
        orc.render(pr, want_ids=False, threads=0, region=region)
    t0 = time.perf_counter(); rays = 0
    for _ in range(args.steps):
        _, _, _, st = orc.render(pr, want_ids=False, threads=0, region=region)
        rays += st["total"]
    dt = time.perf_counter() - t0
    v = rays / dt / 1e6
    sample = f"rows {y0}..{y0 + rows} of {H} ({rows * W} px x {S} spp) per step, all host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "note": "CPU restatement of the reference loop (oracle/): the Windows/D3D11/Embree reference cannot be built here"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(sc, p, budget_s=12.0):
    """Oracle timed on the host cores on a bounded sample of the same workload."""
    from oracle.oracle import Oracle, make_params

    orc = Oracle(sc)
    cores = orc.max_threads()
    W, H = sc.camera.width, sc.camera.height
    S = p["sampling_width"] ** 2
    rows = H if W * H * S <= 4_000_000 else max(8, int(4_000_000 / (W * S)) // 8 * 8)
    y0 = (H - rows) // 2
    region = (0, y0, W, y0 + rows)
    pr = make_params(**p)
    orc.render(pr, want_ids=False, threads=0, region=region)
    t0 = time.perf_counter(); rays = 0; n = 0
    while True:
        _, _, _, st = orc.render(pr, want_ids=False, threads=0, region=region)
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
            "sample": f"{n} x rows {y0}..{y0 + rows} of {H} at {W} px x {S} spp, all {cores} host threads; single-thread {st1['total'] / dt1 / 1e6:.3f} Mrays/s on 8 rows"}


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
    # frames in flight: a frame (or shard) of < 1 M primary samples is latency-bound and wants a deeper pipeline
    shard_samples = sc.camera.width * sc.camera.height * p.get("sampling_width", 1) ** 2 / world
    auto_depth = 8 if shard_samples >= 1e6 else 16
    depth = max(1, min(args.inflight if args.inflight > 0 else auto_depth, 16))
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

    def run_frames(n, timed):
        """n frames, `depth` in flight; every frame is preceded by an L2 flush on its own stream.  Returns (device ms, rays)."""
        comm = torch.cuda.current_stream()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        barrier()
        t0.record(comm)
        for s in sr.slot_streams:
            s.wait_event(t0)
        rays = 0
        for k in range(n):
            if k >= depth:
                rays += sr.end(k - depth)["total"]
            h0 = time.perf_counter()
            sr.begin(k, params, before=l2_flush(k))
            host_issue[0] += time.perf_counter() - h0; host_issue[1] += 1
        for k in range(max(0, n - depth), n):
            rays += sr.end(k)["total"]
        for s in sr.slot_streams:
            comm.wait_stream(s)
        t1.record(comm)
        barrier()
        return t0.elapsed_time(t1), rays

    run_frames(max(W, 2 * depth), False)   # every slot allocates its queues on first use: keep that out of the timed region
    host_issue[0] = 0.0; host_issue[1] = 0
    sr.host_s, sr.host_n = [0.0, 0.0, 0.0, 0.0], 0
    sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(dev), "uuid", None)); sampler.start()
    launches0 = rt.kernel_launches()
    total_ms, rays = run_frames(K, True)
    launches = rt.kernel_launches() - launches0
    host_parts = [x / max(sr.host_n, 1) * 1e6 for x in sr.host_s]
    clocks = sampler.stop()
    step_ms = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(rays), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms = float(step_ms.item()); total_rays = float(tot[0].item())
    value = total_rays / (total_ms * 1e-3) / 1e6

    # roofline of the dominant kernels: the same frames one at a time with every launch bracketed by CUDA events on its stream
    trace_ms = 0.0; trace_launches = 0; prof_rays = 0; lv = None; prof_frame_ms = 0.0
    n_prof = min(K, 10)
    for k in range(n_prof):
        sr.begin(k, params, profile=1, before=l2_flush(k))
        st = sr.end(k)
        trace_ms += st["trace_ms"]; trace_launches += st["trace_launches"]; prof_rays += st["total"]; prof_frame_ms += st["frame_ms"]
    lv = rt.level_stats()
    barrier()

    # e2e: the host-buffer C-ABI call every step (camera re-sent, frame copied back to pinned host memory), `depth` in flight
    c = sc.camera
    if world == 1:
        hosts = [torch.empty((rt.height, rt.width, 4), dtype=torch.float32).pin_memory() for _ in range(depth)]

        def e2e_frames(n):
            e_rays = 0
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(n):
                if k >= depth:
                    e_rays += rt.render_end((k - depth) % depth)["total"]
                rt.flush_l2(k % depth, FLUSH_BYTES, k & 0xFF)
                rt.set_camera(c.width, c.height, c.fov_y, c.view_from, c.view_at)
                rt.render_begin(k % depth, params, host_ptr=hosts[k % depth].data_ptr())
            for k in range(max(0, n - depth), n):
                e_rays += rt.render_end(k % depth)["total"]
            torch.cuda.synchronize()
            return time.perf_counter() - t0, e_rays

        rt.set_shard(0, 1)
        e2e_frames(2 * depth)
        t_e2e, e_rays = e2e_frames(K)
        e2e = {"value": e_rays / t_e2e / 1e6, "unit": UNIT, "h2d_bytes_per_step": 4 * (2 + 1 + 3 + 3) + 64,
               "d2h_bytes_per_step": rt.width * rt.height * 16, "ms_per_step": t_e2e / K * 1e3}
    else:
        # multi-GPU e2e: rank 0 additionally copies the gathered frame to pinned host memory every step
        # The frames live in host memory shared by all ranks (memfd registered with every device): each rank's resolve
        # kernel stores its tiles through its own PCIe link.  Fallback: rank 0 copies the NVLink-gathered frame out.
        try:
            sr_e = ShardedRenderer(rt, rank, world, dev, depth=depth, mode="host")
        except RuntimeError:
            sr_e = sr
        host = torch.empty((rt.height, rt.width, 4), dtype=torch.float32).pin_memory() if rank == 0 and sr_e is sr else None

        def e2e_frames(n):
            barrier()
            t0 = time.perf_counter(); rays = 0
            for k in range(n):
                if k >= depth:
                    rays += sr_e.end(k - depth)["total"]
                sr_e.begin(k, params, before=l2_flush(k))
                if host is not None:
                    host.copy_(sr.frames[k % depth], non_blocking=True)   # on the communication stream, after that frame's barrier
            for k in range(max(0, n - depth), n):
                rays += sr_e.end(k)["total"]
            barrier()
            return t0, rays

        if sr_e is not sr:
            e2e_frames(2 * depth)       # new destinations: let every slot re-capture its frame graph outside the timed region
        t0, e_rays = e2e_frames(K)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        er = torch.tensor([float(e_rays)], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX); dist.all_reduce(er, op=dist.ReduceOp.SUM)
        e2e = {"value": float(er.item()) / float(dt.item()) / 1e6, "unit": UNIT, "h2d_bytes_per_step": 4 * (2 + 1 + 3 + 3) + 64,
               "d2h_bytes_per_step": rt.width * rt.height * 16, "ms_per_step": float(dt.item()) / K * 1e3,
               "path": "frames in shared host memory, every rank stores its tiles through its own PCIe link" if sr_e is not sr
                       else "NVLink gather to rank 0, device-to-host copy on rank 0"}
        if sr_e is not sr:
            if rank == 0:
                e2e["checksum"] = float(sr_e.frames[(K - 1) % depth].double().sum().item())   # the host frame is read, not just written
            sr_e.close()

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        b_ray = algorithmic_bytes_per_ray(sc.ntris)
        achieved = prof_rays * b_ray / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else None
        roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": (achieved / peaks["hbm_gbs"]) if achieved else None,
                "traffic": NCU_TRAFFIC_BYTES if args.workload == "c2" else None,
                "kernel": "k_trace + k_phong + k_secondary (every closest-hit query of the frame)", "bytes_per_ray": b_ray,
                "launch_ms_avg": trace_ms / max(trace_launches, 1), "launches_per_frame": trace_launches / max(n_prof, 1), "peak_kind": peak_kind,
                "frame_ms_unpipelined": prof_frame_ms / max(n_prof, 1),
                "achieved_pipelined": value * 1e6 * b_ray / 1e9, "frac_pipelined": value * 1e6 * b_ray / 1e9 / peaks["hbm_gbs"],
                "fp32": {"flop_per_ray": flop_per_ray(sc.ntris), "achieved_tflops": value * 1e6 * flop_per_ray(sc.ntris) / 1e12,
                         "peak_tflops": 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12,
                         "note": "SURVEY 8(d) contract figure F_ray = 192*D + 180 on the pipelined whole-frame rate; peak = 148 SM x 128 lanes x 2 x max SM clock"},
                "level0_trace_ms": lv[0]["trace_ms"] if lv else None, "secondary_ms": lv[1]["trace_ms"] if lv and len(lv) > 1 else None,
                "note": "algorithmic bytes = SURVEY 8(d) contract figure (720 B/ray at this size) x rays of rank 0, over the summed device time of the "
                        "traversal launches measured one frame at a time (CUDA events on the launching stream, L2 flushed before each frame); the "
                        "working set is L2-resident and the kernels are issue-/latency-bound, see DESIGN.md 3.3.  achieved_pipelined = the same "
                        "bytes over the timed, pipelined whole-frame rate (`value`): frames overlap, so it exceeds the one-at-a-time figure"}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": {"workload": desc, "triangles": sc.ntris, "rays_per_frame": total_rays / K,
                                                "parallelism": f"tiles32x8/rr x{world}", "frames_in_flight": depth, "gather": sr.mode,
                                                "host_issue_us_per_step": host_issue[0] / max(host_issue[1], 1) * 1e6,
                                                "host_issue_parts_us": dict(zip(("flush", "render_begin", "barrier", "event"), host_parts)),
                                                "l2": f"flushed before every timed step on that step's stream ({FLUSH_BYTES >> 20} MiB fill = 1.125 x L2)",
                                                "bvh": rt.build_stats},
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(tot[1].item()), "roofline": roof}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(sc, p)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=0, help="frames in flight per GPU (1 = one frame at a time; 0 = 8, or 16 when a frame or shard has fewer than 1 M primary samples)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
