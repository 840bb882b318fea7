#!/usr/bin/env python
"""bench.py -- throughput of the DiPs frame-difference hot path on B200 (one process per GPU).

    python bench.py --gpus N --steps K --warmup W            our arm (N>1 under torch.distributed.run)
    python bench.py --impl reference --gpus N ...            CPU arm: the oracle port on the host cores (rank 0 only)

A step = one pass of the hot path over one synthetic clip (default workload: BASELINE.json configs[1], 1920x1080 RGB8,
1800 frames per GPU, overall difference + threshold).  `value` = frames/s over all GPUs with the clip resident in HBM;
`e2e` = the same through dipsb_run_clip_host from pinned HOST memory (H2D inside the timed region, results read back).
Weak scaling: every rank owns an 1800-frame shard of an N*1800-frame clip; frame 0 is broadcast (overall mode) or the
one-frame halo is sent to the next rank (per-frame mode) and the accumulators are all-reduced once per step (NCCL).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x44695073
WORKLOADS = {
    # name: (description, width, height, fmt, mode, tau, frames per GPU)
    "c1": ("C1: synthetic 640x480 RGB8 300-frame clip, overall-difference vs first frame", 640, 480, 0, 0, 32, 300),
    "c2": ("C2: synthetic 1920x1080 RGB8 1800-frame clip, overall-difference + threshold", 1920, 1080, 0, 0, 32, 1800),
    "c3": ("C3: synthetic 1920x1080 RGB8 1800-frame clip, per-frame difference + per-frame scalars", 1920, 1080, 0, 1, 32, 1800),
    "c4": ("C4-shard: synthetic 3840x2160 RGBx8 450-frame shard (3600-frame clip over 8 GPUs), overall-difference", 3840, 2160, 1, 0, 32, 450),
    "c5": ("C5-shard: synthetic 7680x4320 RGB8 150-frame shard (1200-frame clip over 8 GPUs)", 7680, 4320, 0, 0, 32, 150),
}
MODE_NAMES = {0: "overall", 1: "per-frame"}
FMT_NAMES = {0: "RGB8", 1: "RGBx8", 2: "BGR8", 3: "BGRx8"}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--workload", choices=sorted(WORKLOADS), default="c2")
    p.add_argument("--frames", type=int, default=0, help="override frames per GPU")
    p.add_argument("--mode", choices=["overall", "per-frame"], default=None)
    p.add_argument("--e2e-steps", type=int, default=3)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--exchange", choices=["replicated", "collective"], default="replicated",
                   help="what a shard needs before its first frame (the clip's frame 0 / the one-frame halo): replicated = "
                        "each shard carries its own copy, loaded with it; collective = broadcast / send-recv every clip")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--stages", type=int, default=0)
    p.add_argument("--tile-px", type=int, default=0)
    p.add_argument("--segments", type=int, default=0)
    p.add_argument("--regs", type=int, default=0)
    p.add_argument("--kernel", type=int, default=-1, help="0 clip_kernel, 1 clip_kernel_ws, -1 automatic (library default)")
    return p.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload: str):
    """dram bytes per launch of the clip kernel from the committed ncu --set full capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": (max(self.power) if self.power else None)}


# --------------------------------------------------------------------------------------------------------------------
def host_cores() -> int:
    """threads the CPU arm uses: every core this process may run on (torchrun exports OMP_NUM_THREADS=1; the thread count
    is therefore passed to the oracle explicitly)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_port_throughput(wl, sample_frames: int, clip_host=None, budget_s: float = 20.0):
    """The oracle (C port of the reference's shader arithmetic) on the host cores: frames/s on a bounded sample."""
    import numpy as np
    from oracle import oracle as O
    desc, w, h, fmt, mode, tau, _ = wl
    cores = host_cores()
    if clip_host is None:
        clip_host = O.synth_clip(sample_frames, w, h, fmt, seed=SEED, profile=O.SYNTH_SCENE)
    clip_host = np.ascontiguousarray(clip_host[:sample_frames])
    O.run_clip(clip_host[: min(8, sample_frames)], fmt, mode, tau, nthreads=cores)        # warm the threads and caches
    t = time.perf_counter()
    reps = 0
    while True:
        O.run_clip(clip_host, fmt, mode, tau, nthreads=cores)
        reps += 1
        dt = time.perf_counter() - t
        if dt > budget_s / 2 or reps >= 20:
            break
    return reps * sample_frames / dt, cores, f"{sample_frames}-frame prefix of the workload x{reps} passes, {cores} OpenMP threads, oracle/dips_oracle.c"


def reference_arm(args, wl, rank, out):
    """--impl reference: the reference has no CPU path and cannot be built here (Rust + WGSL/wgpu); per the task contract
    this arm times the oracle port on the host cores, all threads, on a bounded sample of the same workload."""
    if rank != 0:
        return 0
    import numpy as np
    from oracle import oracle as O
    desc, w, h, fmt, mode, tau, frames = wl
    fb = w * h * O.bpp(fmt)
    sample = max(8, min(frames, int(1.5e9 // fb)))          # ~1.5 GB of frames per step
    cores = host_cores()
    clip = O.synth_clip(sample, w, h, fmt, seed=SEED, profile=O.SYNTH_SCENE, nthreads=cores)
    for _ in range(max(1, min(args.warmup, 2))):
        O.run_clip(clip[: max(8, sample // 8)], fmt, mode, tau, nthreads=cores)
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        O.run_clip(clip, fmt, mode, tau, nthreads=cores)
        done += 1
        if time.perf_counter() - t0 > 120 and done >= 3:      # keep the whole run within a few minutes
            break
    dt = time.perf_counter() - t0
    fps = done * sample / dt
    line = {
        "impl": "reference", "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_of(wl, args, 1, sample_note=f"each step = {sample}-frame prefix of the workload"),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{sample}-frame prefix x{done} steps, {cores} OpenMP threads (oracle/dips_oracle.c; the reference has no CPU implementation)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)
    return 0


def config_of(wl, args, world, sample_note=None):
    desc, w, h, fmt, mode, tau, frames = wl
    cfg = {"workload": desc, "width": w, "height": h, "format": FMT_NAMES[fmt], "mode": MODE_NAMES[mode],
           "threshold_i2": tau, "frames_per_gpu": frames, "frames_total": frames * world, "seed": hex(SEED),
           "profile": "scene", "parallelism": f"frame-shard x{world}",
           "exchange": ("none (single GPU)" if world == 1 else
                        ("reference / halo frame replicated with each shard at load, " if args.exchange == "replicated" else
                         "reference plane broadcast / halo send-recv per clip, ") + "one packed accumulator all-reduce per clip"),
           "l2": "clip per GPU >> 126 MB L2 (inputs larger than L2)"}
    if sample_note:
        cfg["sample"] = sample_note
    return cfg


def _claim_stdout():
    """Libraries (NCCL prints its version banner) may write to fd 1; the contract is ONE JSON line on stdout.  Keep a private
    duplicate of the real stdout for that line and point fd 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    args = parse_args()
    out = _claim_stdout()
    wl = list(WORKLOADS[args.workload])
    if args.frames:
        wl[6] = args.frames
    if args.mode:
        wl[4] = 0 if args.mode == "overall" else 1
    wl = tuple(wl)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, wl, rank, out)

    import numpy as np
    import torch
    import torch.distributed as dist

    import dips_b200
    from dips_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; dips_b200 has no CPU fallback")
    desc, w, h, fmt, mode, tau, frames = wl
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        # a rank that never arrives at a collective should end the run after minutes, not hold the box for the default 10+
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(minutes=4))
    bpp = dips_b200.bytes_per_pixel(fmt)
    npx, fb = w * h, w * h * bpp
    stream = torch.cuda.Stream(device=dev)       # one explicit stream for our kernels, torch copies and NCCL ordering
    torch.cuda.set_stream(stream)

    # ---- synthetic shard, generated on the device (outside every timed region) ------------------------------------
    t0_frame = rank * frames
    clip = torch.empty((frames, fb), dtype=torch.uint8, device=dev)
    dips_b200.synth_fill_device(local_rank, clip.data_ptr(), t0_frame, frames, w, h, fmt, SEED, dips_b200.SYNTH_SCENE,
                                stream.cuda_stream)
    torch.cuda.synchronize()

    ctx = dips_b200.Context(w, h, fmt, mode, tau, device=local_rank)
    if args.kernel >= 0:
        ctx.set_kernel(args.kernel)
    if args.stages or args.tile_px or args.segments or args.regs:
        ctx.set_tuning(args.stages, args.tile_px, args.segments, args.regs)
    ctx.set_stream(stream.cuda_stream)
    reference = None
    replicated = world > 1 and args.exchange == "replicated"          # the same on every rank
    if replicated and rank > 0:
        # the shard's copy of the frame it needs before its first one: the clip's frame 0 (overall) or its predecessor
        # t0-1 (per-frame); part of the resident input, like the shard itself
        ref_frame = torch.empty(fb, dtype=torch.uint8, device=dev)
        dips_b200.synth_fill_device(local_rank, ref_frame.data_ptr(), 0 if mode == sharding.MODE_OVERALL else t0_frame - 1,
                                    1, w, h, fmt, SEED, dips_b200.SYNTH_SCENE, stream.cuda_stream)
        torch.cuda.synchronize()
        reference = {mode: ref_frame}
    engine = sharding.GpuShardEngine(ctx, clip, torch, total_frames=world * frames, replicated=replicated, reference=reference)

    phase_events = []

    def step(record=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        if ev: ev[0].record(stream)
        ctx.reset()
        sharding.exchange_reference(engine, mode, rank, world, dist if world > 1 else None)
        if ev: ev[1].record(stream)
        engine.run(t0_frame)
        if ev: ev[2].record(stream)
        if world > 1:
            dist.all_reduce(engine.acc_tensor(), op=dist.ReduceOp.SUM)
            engine.after_reduce()
        if ev:
            ev[3].record(stream)
            phase_events.append(ev)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    fence()
    ctx.enable_timing(True)
    ctx.clip_kernel_time()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = dips_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    e0.record(stream)
    for _ in range(args.steps):
        step(record=True)
    e1.record(stream)
    fence()
    clocks = sampler.stop()
    launches = dips_b200.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    kern_ms, kern_n = ctx.clip_kernel_time()
    ctx.enable_timing(False)
    plan = ctx.last_plan()
    phases = {"reset+exchange_ms": sum(e[0].elapsed_time(e[1]) for e in phase_events) / len(phase_events),
              "prime+clip+finalize_ms": sum(e[1].elapsed_time(e[2]) for e in phase_events) / len(phase_events),
              "allreduce_ms": sum(e[2].elapsed_time(e[3]) for e in phase_events) / len(phase_events)}
    t_all = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms_total = float(t_all.item())
    value = world * frames * args.steps / (ms_total / 1e3)
    per_rank = [kern_ms / max(kern_n, 1)]
    if world > 1:                                   # the slowest GPU sets the step time: report every rank's kernel time
        g_all = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(g_all, torch.tensor([per_rank[0]], dtype=torch.float64, device=dev))
        per_rank = [float(x.item()) for x in g_all]

    # sanity of the last step's results (cheap integer identity; parity proper lives in tests/)
    sad, cnt = ctx.get_scalars(t0_frame, frames)
    acc_sum, acc_cnt = ctx.get_accumulators()
    tot = torch.tensor([int(sad.sum()), int(cnt.sum())], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)          # per-frame scalars are per shard; the all-reduced maps must add up to all of them
    assert int(acc_sum.astype(np.uint64).sum()) == int(tot[0]) and int(acc_cnt.astype(np.uint64).sum()) == int(tot[1]), \
        "checksum of checksums failed"
    if world > 1 and mode == sharding.MODE_OVERALL:   # every rank differenced against the same reference plane
        plane = ctx.get_state_plane().astype(np.int64)
        sig = torch.tensor([int(plane.sum()), int((plane * (np.arange(plane.size) % 65521 + 1)).sum())], dtype=torch.int64, device=dev)
        sigs = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        assert all(bool((x == sigs[0]).all()) for x in sigs), "reference planes differ between ranks"

    # ---- roofline of the dominant kernel --------------------------------------------------------------------------
    alg_bytes = frames * fb + npx * 2 + npx * 8      # every input byte once + reference plane + accumulators once
    kern_avg_ms = kern_ms / max(kern_n, 1)
    achieved = alg_bytes / (kern_avg_ms / 1e3) / 1e9
    peak, peak_src = measured_peak()
    probe_ms = ctx.stream_probe(clip.data_ptr(), frames, clip.stride(0), 5)      # compute-free TMA stream, same tiles
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.workload),
                "kernel": "clip_kernel_ws" if plan.get("kernel") == 1 else "clip_kernel", "kernel_ms": kern_avg_ms,
                "kernel_share_of_step": kern_avg_ms / (ms_total / args.steps), "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms_per_rank": per_rank, "peak_source": peak_src, "frac_of_8TBps_nominal": achieved / 8000.0,
                "stream_probe_GBps": frames * fb / (probe_ms / 1e3) / 1e9,
                "frac_of_stream_probe": (frames * fb / (kern_avg_ms / 1e3) / 1e9) / (frames * fb / (probe_ms / 1e3) / 1e9)}

    # ---- end to end: pinned host clip -> dipsb_run_clip_host -> results back on the host ---------------------------
    e2e = None
    if not args.no_e2e:
        import psutil
        e2e_frames = frames
        avail = psutil.virtual_memory().available
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        while e2e_frames > 32 and e2e_frames * fb * local_world * 1.5 > avail:
            e2e_frames //= 2
        host = torch.empty((e2e_frames, fb), dtype=torch.uint8, pin_memory=True)
        host.copy_(clip[:e2e_frames])
        torch.cuda.synchronize()
        h_sum = torch.empty(npx, dtype=torch.int32, pin_memory=True)
        h_cnt = torch.empty(npx, dtype=torch.int32, pin_memory=True)

        class HostEngine(sharding.GpuShardEngine):
            def first_frame(self):
                return host[0].to(dev, non_blocking=True)

            def last_frame(self):
                return host[e2e_frames - 1].to(dev, non_blocking=True)

            def run(self, first_frame_index):
                ctx.run_clip_host(host.data_ptr(), e2e_frames, fb, first_frame_index)

        hengine = HostEngine(ctx, clip, torch, total_frames=world * e2e_frames)

        def e2e_step():
            ctx.reset()
            sharding.run_sharded(hengine, mode, t0_frame, rank, world, dist if world > 1 else None)
            ctx.get_accumulators_into(h_sum.data_ptr(), h_cnt.data_ptr())                           # D2H of the maps
            return ctx.get_scalars(t0_frame, e2e_frames)                                            # D2H of the scalars

        e2e_step()
        fence()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        fence()
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        dt = float(t_e.item())
        e2e = {"value": world * e2e_frames * args.e2e_steps / dt, "unit": "frames/s",
               "h2d_bytes_per_step": e2e_frames * fb, "d2h_bytes_per_step": npx * 8 + e2e_frames * 16,
               "frames_per_gpu": e2e_frames, "steps": args.e2e_steps,
               "api": "dipsb_run_clip_host (pinned host clip, chunked H2D overlapped with the clip kernel) + dipsb_get_accumulators + dipsb_get_scalars",
               "h2d_GBps_per_gpu": e2e_frames * fb * args.e2e_steps / dt / 1e9}
        if world == 1:
            # the same call from ORDINARY host memory (what a caller without page-locked buffers has): every chunk is first
            # staged into a page-locked bounce buffer by the library's threaded host copy
            pn = min(e2e_frames, max(32, int(2.5e9 // fb)))
            pageable = np.empty((pn, fb), np.uint8)
            pageable[:] = host[:pn].numpy()
            ctx.reset(); ctx.run_clip_host(pageable.ctypes.data, pn, fb, 0); ctx.synchronize()
            t0 = time.perf_counter()
            for _ in range(2):
                ctx.reset()
                ctx.run_clip_host(pageable.ctypes.data, pn, fb, 0)
                ctx.get_accumulators_into(h_sum.data_ptr(), h_cnt.data_ptr())
                ctx.get_scalars(0, pn)
            e2e["pageable_host_fps"] = 2 * pn / (time.perf_counter() - t0)
            e2e["pageable_host_frames"] = pn
            del pageable
        del host

    # ---- streaming boundary (the reference's per-frame callback shape): RGBA8 frame in -> RGBA8 difference frame out ----
    stream_info = None
    if not args.no_e2e and world == 1:
        sw, sh, sn = 1920, 1080, 120
        frames_rgba = np.empty((8, sw * sh * 4), np.uint8)
        tmp = torch.empty(8 * sw * sh * 4, dtype=torch.uint8, device=dev)
        dips_b200.synth_fill_device(local_rank, tmp.data_ptr(), 0, 8, sw, sh, dips_b200.FMT_RGBX8, SEED,
                                    dips_b200.SYNTH_SCENE, stream.cuda_stream)
        torch.cuda.synchronize()
        frames_rgba[:] = tmp.cpu().numpy().reshape(8, -1)
        del tmp
        stream_info = {"geometry": "1920x1080 RGBx8 in, RGBA8 difference frame out, per-frame call", "frames": sn,
                       "buffers": "pageable = ordinary host memory (staged by the library's threaded host copy each way); pinned = "
                                  "dipsb_host_alloc buffers (copy engine reads/writes them directly)"}
        pin_in = dips_b200.PinnedBuffer(8 * sw * sh * 4, device=local_rank)
        pin_out = dips_b200.PinnedBuffer(2 * sw * sh * 4, device=local_rank)
        pin_in.array[:] = frames_rgba.reshape(-1)
        for kind in ("pageable", "pinned"):
            src = frames_rgba if kind == "pageable" else pin_in.array.reshape(8, -1)
            dst = (np.empty((2, sw * sh * 4), np.uint8) if kind == "pageable" else pin_out.array.reshape(2, -1))
            dst[:] = 0
            for name in ("dipsb_push_frame", "dipsb_push_frame_pipelined"):
                with dips_b200.Context(sw, sh, dips_b200.FMT_RGBX8, 0, tau, device=local_rank) as sctx:
                    fn = sctx.push_frame if name == "dipsb_push_frame" else sctx.push_frame_pipelined
                    for k in range(4):
                        fn(src[k % 8], out=dst[k & 1])
                    t0 = time.perf_counter()
                    for k in range(sn):
                        fn(src[k % 8], out=dst[k & 1])
                    if name != "dipsb_push_frame":
                        sctx.flush_frame(out=dst[sn & 1])
                    key = name + ("_fps" if kind == "pageable" else "_pinned_fps")
                    stream_info[key] = sn / (time.perf_counter() - t0)
        pin_in.close()
        pin_out.close()

    # ---- CPU baseline (rank 0, N == 1 only) -----------------------------------------------------------------------
    cpu = None
    if not args.no_cpu and world == 1:
        sample = max(8, min(frames, int(1.0e9 // fb)))
        host_sample = clip[:sample].cpu().numpy()
        v, cores, what = cpu_port_throughput(wl, sample, host_sample)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": what}

    if rank == 0:
        cfg = config_of(wl, args, world)
        cfg["plan"] = plan
        line = {
            "metric": "frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
            "hbm_GBps_per_gpu_whole_step": frames * fb * args.steps / (ms_total / 1e3) / 1e9,
            "phases": phases, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "stream": stream_info,
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
        }
        print(json.dumps(line), file=out, flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
