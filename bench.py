#!/usr/bin/env python
"""bench.py -- throughput of the DiPs frame-difference hot path on B200 (one process per GPU).

    python bench.py --gpus N --steps K --warmup W            our arm (N>1 under torch.distributed.run)
    python bench.py --impl reference --gpus N ...            CPU arm: the oracle port on the host cores (rank 0 only)

Headline workload (BASELINE.json configs[3], the configuration the target sentence is quoted on): ONE synthetic
3840x2160 RGBx8 clip of 3600 frames, overall difference + threshold, frame-sharded over the N GPUs -- strong scaling: rank r
owns frames dipsb_shard_range(3600, N, r); at N=1 the whole 119.4 GB clip is resident on the one GPU.  A step = one pass
over the clip: dipsb_reset + dipsb_run_clip_sharded_device, i.e. inside the library and inside the timed region: rank 0
builds the reference plane from frame 0 and ncclBroadcast's it, every rank runs its shard through the clip kernel, and the
per-pixel accumulators are combined by the library's reduce-scatter over peer memory (NVLink).  `value` = frames/s of the
whole job with the clip resident in HBM; `e2e` = the same through dipsb_run_clip_sharded_host from pinned HOST memory (H2D
inside the timed region, maps and scalars read back).  `configs` carries one row per other BASELINE configuration (C2, C3,
C5 overall, C5 per-frame with the halo exchange), each with its own roofline, phases and a sustained (>= 1 s) figure.
torch is used for device memory, streams and (gloo) for handing the NCCL unique id around and taking the max over ranks;
every collective of the data path is issued by libdips_b200.so.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x44695073
WORKLOADS = {
    # name: (description, width, height, fmt, mode, tau, frames of the whole clip)
    "c1": ("C1: synthetic 640x480 RGB8 300-frame clip, overall-difference vs first frame", 640, 480, 0, 0, 32, 300),
    "c2": ("C2: synthetic 1920x1080 RGB8 1800-frame clip, overall-difference + threshold", 1920, 1080, 0, 0, 32, 1800),
    "c3": ("C3: synthetic 1920x1080 RGB8 1800-frame clip, per-frame difference + per-frame scalars", 1920, 1080, 0, 1, 32, 1800),
    "c4": ("C4: synthetic 3840x2160 RGBx8 3600-frame clip, overall-difference, frame-sharded with accumulator reduction", 3840, 2160, 1, 0, 32, 3600),
    "c4p": ("C4 geometry, per-frame mode: synthetic 3840x2160 RGBx8 3600-frame clip, per-frame difference + scalars", 3840, 2160, 1, 1, 32, 3600),
    "c5o": ("C5: synthetic 7680x4320 RGB8 1200-frame clip, overall-difference, frame-sharded", 7680, 4320, 0, 0, 32, 1200),
    "c5p": ("C5: synthetic 7680x4320 RGB8 1200-frame clip, per-frame difference, halo exchange at shard boundaries", 7680, 4320, 0, 1, 32, 1200),
}
ROWS = ["c2", "c3", "c5o", "c5p"]
MODE_NAMES = {0: "overall", 1: "per-frame"}
FMT_NAMES = {0: "RGB8", 1: "RGBx8", 2: "BGR8", 3: "BGRx8"}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--workload", choices=sorted(WORKLOADS), default="c4")
    p.add_argument("--frames", type=int, default=0, help="override the frames of the whole clip")
    p.add_argument("--rows", default=",".join(ROWS), help="comma-separated extra workloads reported under `configs` ('' = none)")
    p.add_argument("--sustained-s", type=float, default=1.0, help="length of the back-to-back sustained measurement per workload")
    p.add_argument("--e2e-steps", type=int, default=3)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-stream", action="store_true")
    p.add_argument("--no-preflight", action="store_true")
    p.add_argument("--no-numa-bind", action="store_true", help="N>1: leave the ranks' threads and host buffers wherever the scheduler puts them")
    p.add_argument("--only-e2e", action="store_true", help="skip the device-resident measurement; print the e2e object alone (for host-side experiments)")
    p.add_argument("--no-comm-probes", dest="comm_probes", action="store_false")
    p.add_argument("--reduce", choices=["auto", "p2p", "nccl"], default="auto",
                   help="accumulator exchange: the library's peer-memory kernels or pack + ncclAllReduce + unpack")
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


def shard_of(total: int, world: int, rank: int):
    """first frame and frame count of `rank` (the same arithmetic as dipsb_shard_range, without loading the library)"""
    t0 = (rank * total) // world
    return t0, ((rank + 1) * total) // world - t0


def cpu_sample_frames(wl) -> int:
    """frames of the bounded sample the CPU port is timed on (C1/C2/C3 in full, 4K/8K on a >= 64-frame
    prefix)"""
    _, w, h, fmt, _, _, total = wl
    fb = w * h * (3 if fmt in (0, 2) else 4)
    return total if total * fb <= 12e9 else min(total, max(64, int(2.2e9 // fb)))


def cpu_port_throughput(wl, clip_host, budget_s: float = 12.0):
    """The oracle (C port of the reference's shader arithmetic) on the host cores: frames/s on a bounded sample."""
    from oracle import oracle as O
    _, w, h, fmt, mode, tau, _ = wl
    cores = host_cores()
    n = clip_host.shape[0]
    O.run_clip(clip_host[: min(8, n)], fmt, mode, tau, nthreads=cores)        # warm the threads and caches
    t = time.perf_counter()
    reps = 0
    while True:
        O.run_clip(clip_host, fmt, mode, tau, nthreads=cores)
        reps += 1
        dt = time.perf_counter() - t
        if dt > budget_s or reps >= 20:
            break
    return reps * n / dt, cores, f"{n}-frame prefix of the workload x{reps} passes, {cores} OpenMP threads, oracle/dips_oracle.c"


def config_of(wl, world):
    """the `config` object of the JSON line -- identical for both arms"""
    desc, w, h, fmt, mode, tau, total = wl
    shards = [shard_of(total, world, r)[1] for r in range(world)]
    return {"workload": desc, "width": w, "height": h, "format": FMT_NAMES[fmt], "mode": MODE_NAMES[mode],
            "threshold_i2": tau, "frames_total": total, "frames_per_gpu": shards[0] if len(set(shards)) == 1 else shards,
            "seed": hex(SEED), "profile": "scene", "parallelism": f"frame-shard x{world} (strong scaling: one clip, {world} contiguous frame ranges)",
            "exchange": ("none (single GPU)" if world == 1 else
                         "per clip, issued by libdips_b200.so: " +
                         ("reference plane of frame 0 scattered by rank 0's prime kernel and all-gathered over peer memory" if mode == 0 else
                          "one-frame halo pushed over NVLink by the copy engine during the pass") +
                         " + accumulator reduce-scatter over peer memory (totals stay sharded by pixel range; gathered on read-out)"),
            "l2": "clip shard per GPU >> 126 MB L2 (inputs larger than L2)"}


def reference_arm(args, wl, rank, out):
    """--impl reference: the reference has no CPU path and cannot be built here (Rust + WGSL/wgpu); per the task contract
    this arm times the oracle port on the host cores, all threads, on a bounded sample of the same workload."""
    if rank != 0:
        return 0
    from oracle import oracle as O
    desc, w, h, fmt, mode, tau, total = wl
    sample = cpu_sample_frames(wl)
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
        if time.perf_counter() - t0 > 150 and done >= 3:      # keep the whole run within a few minutes
            break
    dt = time.perf_counter() - t0
    fps = done * sample / dt
    what = (f"each step = the first {sample} of the clip's {total} frames" if sample < total else f"each step = all {total} frames of the clip")
    line = {
        "impl": "reference", "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_of(wl, args.gpus),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{what}, x{done} steps, {cores} OpenMP threads (oracle/dips_oracle.c; the reference has no CPU implementation)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)
    return 0


def _claim_stdout():
    """Libraries (NCCL prints its version banner) may write to fd 1; the contract is ONE JSON line on stdout.  Keep a private
    duplicate of the real stdout for that line and point fd 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def log(rank, *a):
    if rank == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------------------------------
def bind_near_gpu(torch, local_rank: int):
    """Multi-GPU runs only: keep this rank's threads (and so the page-locked host buffers they allocate and the library's
    host-copy helpers they spawn) on the CPU socket the GPU hangs off.  Without it the ranks' host buffers land on whatever
    node the scheduler picked and half of the 8 uploads cross the socket link.  Returns what was done, for the JSON line."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"node": None, "why": "the platform reports no NUMA node for the GPU"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"node": node, "why": "none of the node's CPUs are available to this process"}
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus), "gpu": bdf}
    except (OSError, ValueError, AttributeError) as e:
        return {"node": None, "why": f"{type(e).__name__}: {e}"}


class Bench:
    """One process = one GPU = one rank.  Holds the device buffer the synthetic shards are generated into."""

    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist

        import dips_b200
        self.torch, self.dist, self.lib = torch, dist, dips_b200
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.host_numa = bind_near_gpu(torch, local_rank) if world > 1 and not args.no_numa_bind else None
        if world > 1:
            import datetime
            # gloo only carries the NCCL unique id, barriers and max-over-ranks of the timings; every collective of the data
            # path is issued by libdips_b200.so on its own NCCL communicator / peer-memory kernels
            dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(minutes=4))
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.buf = None
        self.reduce_path = {"auto": dips_b200.REDUCE_AUTO, "p2p": dips_b200.REDUCE_P2P, "nccl": dips_b200.REDUCE_NCCL}[args.reduce]

    # ---- plumbing -------------------------------------------------------------------------------------------------
    def fence(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, xs):
        if self.world == 1:
            return list(xs)
        t = self.torch.tensor(list(xs), dtype=self.torch.int64)
        self.dist.all_reduce(t)
        return [int(v) for v in t]

    def gather_floats(self, x: float):
        if self.world == 1:
            return [x]
        outs = [None] * self.world
        self.dist.all_gather_object(outs, x)
        return [float(v) for v in outs]

    def gather_objects(self, x):
        if self.world == 1:
            return [x]
        outs = [None] * self.world
        self.dist.all_gather_object(outs, x)
        return outs

    def unique_id(self) -> bytes:
        box = [self.lib.comm_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def ensure_buf(self, nbytes: int):
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.torch.cuda.empty_cache()
            self.buf = self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.dev)
        return self.buf

    def make_context(self, wl, tune=True):
        _, w, h, fmt, mode, tau, _ = wl
        a = self.args
        ctx = self.lib.Context(w, h, fmt, mode, tau, device=self.local_rank)
        if tune and a.kernel >= 0:
            ctx.set_kernel(a.kernel)
        if tune and (a.stages or a.tile_px or a.segments or a.regs):
            ctx.set_tuning(a.stages, a.tile_px, a.segments, a.regs)
        ctx.set_stream(self.stream.cuda_stream)
        if self.world > 1:
            ctx.comm_init_rank(self.world, self.rank, self.unique_id())
            if self.reduce_path != self.lib.REDUCE_AUTO:
                ctx.comm_set_reduce(self.reduce_path)
        return ctx

    # ---- one workload: device-resident shard, K timed steps, sustained run, roofline --------------------------------
    def measure(self, name, wl, steps, warmup, with_probe=False, keep=False):
        torch, lib = self.torch, self.lib
        desc, w, h, fmt, mode, tau, total = wl
        bpp = lib.bytes_per_pixel(fmt)
        npx, fb = w * h, w * h * bpp
        t0_frame, n_local = shard_of(total, self.world, self.rank)
        clip = self.ensure_buf(n_local * fb)[: n_local * fb].view(n_local, fb)
        lib.synth_fill_device(self.local_rank, clip.data_ptr(), t0_frame, n_local, w, h, fmt, SEED, lib.SYNTH_SCENE,
                              self.stream.cuda_stream)
        torch.cuda.synchronize()
        ctx = self.make_context(wl)
        sharded = self.world > 1

        def step():
            ctx.reset()
            if sharded:
                ctx.run_clip_sharded_device(clip.data_ptr(), n_local, t0_frame, total, fb)
            else:
                ctx.run_clip_device(clip.data_ptr(), n_local, fb, 0)

        for _ in range(max(warmup, 3)):
            step()
        self.fence()
        ctx.enable_timing(True)
        ctx.clip_kernel_time()
        if sharded:
            ctx.comm_phase_times()
        sampler = ClockSampler(self.local_rank)
        sampler.start()
        launches0 = lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.fence()
        e0.record(self.stream)
        for _ in range(steps):
            step()
        e1.record(self.stream)
        self.fence()
        clocks = sampler.stop()
        launches = lib.launch_count() - launches0
        ms_total = self.max_over_ranks(e0.elapsed_time(e1))
        kern_ms, kern_n = ctx.clip_kernel_time()
        kern_avg_ms = kern_ms / max(kern_n, 1)
        phases = {"step_ms": ms_total / steps, "clip_kernel_ms": kern_avg_ms}
        if sharded:
            (ex_ms, run_ms, red_ms), n_pass = ctx.comm_phase_times()
            n_pass = max(n_pass, 1)
            mine = {"exchange_before_pass_ms": ex_ms / n_pass, "prime+clip+finalize_ms": run_ms / n_pass,
                    "accumulator_exchange_ms": red_ms / n_pass}
            phases.update(mine)
            phases["note"] = "event pairs on this rank's stream (rank 0); a phase that waits for a slower rank contains the wait"
            allp = [None] * self.world
            self.dist.all_gather_object(allp, mine)
            phases["per_rank"] = {k: [round(p[k], 4) for p in allp] for k in mine}
            ctx.comm_check()
        ctx.enable_timing(False)
        per_rank = self.gather_floats(kern_avg_ms)
        value = total * steps / (ms_total / 1e3)

        # ---- sanity of the last step's results (cheap integer identity; parity proper lives in tests/ and the preflight) ----
        import numpy as np
        sad, cnt = ctx.get_scalars(t0_frame, n_local)
        if sharded:
            t_g = time.perf_counter()
            ctx.gather_accumulators()
            ctx.synchronize()
            gather_ms = 1e3 * (time.perf_counter() - t_g)
            ctx.comm_check()
        acc_sum, acc_cnt = ctx.get_accumulators()
        tot = self.sum_over_ranks([int(sad.sum()), int(cnt.sum())])
        got = (int(acc_sum.astype(np.uint64).sum()), int(acc_cnt.astype(np.uint64).sum()))
        # reported, not asserted: a failed identity must not take the measurements of the other rows down with it
        checks = {"sum_of_maps_equals_sum_of_scalars": got[0] == tot[0] and got[1] == tot[1]}
        if not checks["sum_of_maps_equals_sum_of_scalars"]:
            checks["detail"] = f"maps {got} vs scalars {tuple(tot)}"
            log(self.rank, f"{name}: CHECKSUM MISMATCH maps {got} vs scalars {tuple(tot)} (rank {self.rank}: scalars {int(sad.sum())}, {int(cnt.sum())})")
        if sharded:   # every rank must hold the same gathered maps
            sig = [int(acc_sum[::97].astype(np.uint64).sum()), int(acc_cnt[::89].astype(np.uint64).sum())]
            allsig = self.sum_over_ranks(sig)
            checks["gathered_maps_equal_on_all_ranks"] = allsig[0] == sig[0] * self.world and allsig[1] == sig[1] * self.world
            phases["gather_on_readout_ms"] = gather_ms

        # ---- sustained: back-to-back steps for >= sustained_s (the power-capped regime; short runs measure burst clocks) ----
        sustained = None
        if self.args.sustained_s > 0:
            n_sus = max(steps, int(math.ceil(self.args.sustained_s * 1e3 / (ms_total / steps))))
            ctx.enable_timing(True)
            ctx.clip_kernel_time()
            s2 = ClockSampler(self.local_rank)
            s2.start()
            self.fence()
            e0.record(self.stream)
            for _ in range(n_sus):
                step()
            e1.record(self.stream)
            self.fence()
            c2 = s2.stop()
            ms_sus = self.max_over_ranks(e0.elapsed_time(e1))
            k_ms, k_n = ctx.clip_kernel_time()
            ctx.enable_timing(False)
            if sharded:
                ctx.comm_phase_times()
                ctx.comm_check()
            k_avg = k_ms / max(k_n, 1)
            sustained = {"steps": n_sus, "seconds": ms_sus / 1e3, "value": total * n_sus / (ms_sus / 1e3), "unit": "frames/s",
                         "ms_per_step": ms_sus / n_sus, "clip_kernel_ms": k_avg,
                         "kernel_GBps": (n_local * fb + npx * 10) / (k_avg / 1e3) / 1e9, "clocks": c2}

        # ---- roofline of the dominant kernel ------------------------------------------------------------------------
        alg_bytes = n_local * fb + npx * 2 + npx * 8      # every input byte once + reference plane + accumulators once
        achieved = alg_bytes / (kern_avg_ms / 1e3) / 1e9
        peak, peak_src = measured_peak()
        plan = ctx.last_plan()
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(name),
                    "kernel": "clip_kernel_ws" if plan.get("kernel") == 1 else "clip_kernel", "kernel_ms": kern_avg_ms,
                    "kernel_share_of_step": kern_avg_ms / (ms_total / steps), "algorithmic_bytes_per_launch": alg_bytes,
                    "kernel_ms_per_rank": per_rank, "peak_source": peak_src, "frac_of_8TBps_nominal": achieved / 8000.0}
        if sustained:
            roofline["sustained"] = {"achieved": sustained["kernel_GBps"], "frac": sustained["kernel_GBps"] / peak,
                                     "seconds": sustained["seconds"], "sm_mhz": sustained["clocks"]["sm_mhz"],
                                     "reasons": sustained["clocks"]["reasons"]}
        if with_probe:
            probe_ms = ctx.stream_probe(clip.data_ptr(), n_local, fb, 3)      # compute-free TMA stream, same tiles
            roofline["stream_probe_GBps"] = n_local * fb / (probe_ms / 1e3) / 1e9
            roofline["frac_of_stream_probe"] = probe_ms / kern_avg_ms
        if sharded and self.args.comm_probes:
            # the exchanges alone, back to back after a barrier (no pass in between, so no waiting for a slower rank's kernel)
            info = ctx.comm_info()
            probes = {}
            for key, what, need in (("reduce_scatter_p2p_ms", 0, info["peer_memory"]), ("reference_broadcast_nccl_ms", 1, info["nccl"]),
                                    ("allgather_p2p_ms", 2, info["peer_memory"]), ("pack_allreduce_unpack_nccl_ms", 3, info["nccl"]),
                                    ("reference_broadcast_p2p_ms", 4, info["peer_memory"] and mode == 0)):
                if need:
                    self.fence()
                    probes[key] = self.max_over_ranks(ctx.comm_probe(what, total, 10))
            ctx.reset()
            ctx.comm_check()
            phases["exchange_probes"] = probes
        res = {"name": name, "value": value, "unit": "frames/s", "ms_per_step": ms_total / steps, "steps": steps,
               "hbm_GBps_per_gpu_whole_step": n_local * fb * steps / (ms_total / 1e3) / 1e9,
               "phases": phases, "roofline": roofline, "sustained": sustained, "clocks": clocks, "plan": plan,
               "gpu_launches": int(launches) * self.world, "comm": ctx.comm_info() if sharded else None, "checks": checks}
        if keep:
            return res, ctx, clip
        ctx.close()
        return res, None, None

    # ---- multi-GPU parity against the oracle (outside every timed region) -------------------------------------------
    def preflight(self):
        """A small clip through the real sharded path on the real GPUs -- both modes, both accumulator paths -- compared bit
        for bit with the CPU oracle run over the whole clip."""
        import numpy as np
        from oracle import oracle as O      # the checker; nothing it computes is measured or shipped
        lib, torch = self.lib, self.torch
        w, h, fmt, tau = 640, 360, 1, 24
        total = 6 * self.world + 3
        fb = w * h * 4
        cases = []
        for mode in (0, 1):
            for path, pname in ((lib.REDUCE_P2P, "p2p"), (lib.REDUCE_NCCL, "nccl")):
                t0, n = shard_of(total, self.world, self.rank)
                clip = self.ensure_buf(n * fb)[: n * fb].view(n, fb)
                lib.synth_fill_device(self.local_rank, clip.data_ptr(), t0, n, w, h, fmt, SEED, lib.SYNTH_SCENE, self.stream.cuda_stream)
                torch.cuda.synchronize()
                ctx = lib.Context(w, h, fmt, mode, tau, device=self.local_rank)
                ctx.set_stream(self.stream.cuda_stream)
                ctx.comm_init_rank(self.world, self.rank, self.unique_id())
                info = ctx.comm_info()
                if path == lib.REDUCE_P2P and not info["peer_memory"]:
                    ctx.close()
                    cases.append({"mode": MODE_NAMES[mode], "reduce": pname, "skipped": "peer memory not mapped"})
                    continue
                ctx.comm_set_reduce(path)
                for _ in range(2):                      # twice: the second pass runs on the other window parity
                    ctx.reset()
                    ctx.run_clip_sharded_device(clip.data_ptr(), n, t0, total, fb)
                ctx.gather_accumulators()
                ctx.synchronize()
                ctx.comm_check()
                acc_sum, acc_cnt = ctx.get_accumulators()
                sad, cnt = ctx.get_scalars(t0, n)
                ctx.close()
                whole = O.synth_clip(total, w, h, fmt, seed=SEED, profile=O.SYNTH_SCENE)
                want = O.run_clip(whole, fmt, mode, tau)
                ok = (np.array_equal(acc_sum, want.acc_sum) and np.array_equal(acc_cnt, want.acc_cnt) and
                      np.array_equal(sad, want.sad[t0:t0 + n]) and np.array_equal(cnt, want.cnt[t0:t0 + n]))
                n_ok = self.sum_over_ranks([1 if ok else 0])[0]
                cases.append({"mode": MODE_NAMES[mode], "reduce": pname, "ranks_bit_exact": n_ok, "ranks": self.world})
                assert n_ok == self.world, f"multi-GPU parity failed: {cases[-1]}"
        return {"checked": True, "clip": f"{w}x{h} RGBx8 x{total} frames over {self.world} ranks, 2 passes", "against": "oracle/dips_oracle.c over the whole clip",
                "cases": cases, "bit_exact": True}

    # ---- end to end: pinned host shard -> dipsb_run_clip_sharded_host -> maps and scalars back on the host -----------
    def e2e(self, wl, ctx):
        import psutil
        torch, lib = self.torch, self.lib
        desc, w, h, fmt, mode, tau, total = wl
        fb, npx = w * h * lib.bytes_per_pixel(fmt), w * h
        t0_frame, n_local = shard_of(total, self.world, self.rank)
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(self.world)))
        avail = psutil.virtual_memory().available
        frames = min(n_local, int(16e9 // fb))                       # at most 16 GB of page-locked memory per process
        while frames > 16 and frames * fb * local_world * 1.5 > avail:
            frames //= 2
        frames = int(self.sum_over_ranks([frames])[0] // self.world) if self.world > 1 else frames   # same on every rank
        frames = min(frames, n_local)
        e_total, e_first = frames * self.world, frames * self.rank
        host = torch.empty((frames, fb), dtype=torch.uint8, pin_memory=True)
        lib.synth_fill_device(self.local_rank, self.buf.data_ptr(), e_first, frames, w, h, fmt, SEED, lib.SYNTH_SCENE, self.stream.cuda_stream)
        host.copy_(self.buf[: frames * fb].view(frames, fb))
        torch.cuda.synchronize()
        h_sum = torch.empty(npx, dtype=torch.int32, pin_memory=True)
        h_cnt = torch.empty(npx, dtype=torch.int32, pin_memory=True)

        def step():
            ctx.reset()
            if self.world > 1:
                ctx.run_clip_sharded_host(host.data_ptr(), frames, e_first, e_total, fb)
                ctx.gather_accumulators()
            else:
                ctx.run_clip_host(host.data_ptr(), frames, fb, 0)
            ctx.get_accumulators_into(h_sum.data_ptr(), h_cnt.data_ptr())            # D2H of the maps
            return ctx.get_scalars(e_first, frames)                                  # D2H of the scalars

        step()
        self.fence()
        t0 = time.perf_counter()
        for _ in range(self.args.e2e_steps):
            step()
        self.fence()
        dt = self.max_over_ranks(time.perf_counter() - t0)
        if self.world > 1:
            ctx.comm_check()
        res = {"value": e_total * self.args.e2e_steps / dt, "unit": "frames/s",
               "h2d_bytes_per_step": frames * fb * self.world, "d2h_bytes_per_step": (npx * 8 + frames * 16) * self.world,
               "frames_per_gpu": frames, "frames_total": e_total, "steps": self.args.e2e_steps,
               "clip": f"a {e_total}-frame clip of the workload's geometry ({frames} frames per GPU: what fits page-locked host memory), same call path",
               "api": ("dipsb_run_clip_sharded_host + dipsb_gather_accumulators" if self.world > 1 else "dipsb_run_clip_host") +
                      " (pinned host clip, chunked H2D overlapped with the clip kernel) + dipsb_get_accumulators + dipsb_get_scalars",
               "h2d_GBps_per_gpu": frames * fb * self.args.e2e_steps / dt / 1e9,
               "h2d_GBps_aggregate": frames * fb * self.world * self.args.e2e_steps / dt / 1e9,
               "limit": "host side: PCIe Gen5 x16 per GPU (54 GB/s measured alone, tools/native/pcie_probe.cu; 52 GB/s per GPU at N=4); "
                        "the aggregate stops at 185-210 GB/s on this pool's hosts (single-node 32-vCPU VMs: 210 GB/s at N=4, 184 GB/s at N=8, "
                        "no NUMA placement to choose), which is the host's memory / root complexes, not the GPUs"}
        if self.host_numa is not None:
            res["host_numa"] = self.gather_objects(self.host_numa)
        if self.world == 1:
            import numpy as np
            # the same call from ORDINARY host memory (what a caller without page-locked buffers has): every chunk is first
            # staged into a page-locked bounce buffer by the library's threaded host copy
            pn = min(frames, max(16, int(2.5e9 // fb)))
            pageable = np.empty((pn, fb), np.uint8)
            pageable[:] = host[:pn].numpy()
            ctx.reset(); ctx.run_clip_host(pageable.ctypes.data, pn, fb, 0); ctx.synchronize()
            t0 = time.perf_counter()
            for _ in range(2):
                ctx.reset()
                ctx.run_clip_host(pageable.ctypes.data, pn, fb, 0)
                ctx.get_accumulators_into(h_sum.data_ptr(), h_cnt.data_ptr())
                ctx.get_scalars(0, pn)
            res["pageable_host_fps"] = 2 * pn / (time.perf_counter() - t0)
            res["pageable_host_frames"] = pn
            del pageable
        del host
        return res

    # ---- streaming boundary (the reference's per-frame callback shape): RGBA8 frame in -> RGBA8 difference frame out ----
    def stream_boundary(self, tau=32):
        import numpy as np
        torch, lib = self.torch, self.lib
        sw, sh, sn = 1920, 1080, 120
        frames_rgba = np.empty((8, sw * sh * 4), np.uint8)
        tmp = torch.empty(8 * sw * sh * 4, dtype=torch.uint8, device=self.dev)
        lib.synth_fill_device(self.local_rank, tmp.data_ptr(), 0, 8, sw, sh, lib.FMT_RGBX8, SEED, lib.SYNTH_SCENE, self.stream.cuda_stream)
        torch.cuda.synchronize()
        frames_rgba[:] = tmp.cpu().numpy().reshape(8, -1)
        del tmp
        info = {"geometry": "1920x1080 RGBx8 in, RGBA8 difference frame out, per-frame call", "frames": sn,
                "buffers": "pageable = ordinary host memory (staged by the library's threaded host copy each way); pinned = "
                           "dipsb_host_alloc buffers (copy engine reads/writes them directly)"}
        pin_in = lib.PinnedBuffer(8 * sw * sh * 4, device=self.local_rank)
        pin_out = lib.PinnedBuffer(2 * sw * sh * 4, device=self.local_rank)
        pin_in.array[:] = frames_rgba.reshape(-1)
        for kind in ("pageable", "pinned"):
            src = frames_rgba if kind == "pageable" else pin_in.array.reshape(8, -1)
            dst = (np.empty((2, sw * sh * 4), np.uint8) if kind == "pageable" else pin_out.array.reshape(2, -1))
            dst[:] = 0
            for name in ("dipsb_push_frame", "dipsb_push_frame_pipelined", "dipsb_stage_frame+dipsb_dispatch_staged"):
                with lib.Context(sw, sh, lib.FMT_RGBX8, 0, tau, device=self.local_rank) as sctx:
                    if name.startswith("dipsb_stage"):
                        # the call pair the C++ / Rust mirrors of ComputeState::{add_texture, dispatch} are built on
                        def fn(frame, out, _c=sctx):
                            _c.stage_frame(frame)
                            return _c.dispatch_staged(out=out)
                    else:
                        fn = sctx.push_frame if name == "dipsb_push_frame" else sctx.push_frame_pipelined
                    for k in range(4):
                        fn(src[k % 8], out=dst[k & 1])
                    t0 = time.perf_counter()
                    for k in range(sn):
                        fn(src[k % 8], out=dst[k & 1])
                    if name == "dipsb_push_frame_pipelined":
                        sctx.flush_frame(out=dst[sn & 1])
                    info[name + ("_fps" if kind == "pageable" else "_pinned_fps")] = sn / (time.perf_counter() - t0)
        pin_in.close()
        pin_out.close()
        # the same boundary without the Python binding: the C++ mirror of frame_callback / ComputeState and the raw C calls
        exe = os.path.join(ROOT, "build", "stream_rate")
        if os.path.exists(exe):
            try:
                import subprocess
                r = subprocess.run([exe, "200"], capture_output=True, text=True, timeout=120)
                if r.returncode == 0:
                    info["native"] = json.loads(r.stdout.strip().splitlines()[-1])
                    info["native"]["what"] = ("tools/native/stream_rate.cpp: mirror_frame_callback = dips::frame_callback over dips::ComputeState "
                                              "(dips_b200/host/dips_host.hpp, add_texture + dispatch, fresh vector out per frame); the rest are the C entry points")
            except Exception as e:      # a measurement aid must not take the bench line down
                info["native"] = {"error": str(e)}
        return info


def ring_batch(B, tau=32, n=600):
    """SURVEY row N1 in batch: the reference's ring-of-4 (`dips`) and ring-of-2 (`dips_alt`) semantics over a device-resident
    1080p RGBx clip through dipsb_run_clip_device -- warm-up frames on the per-frame kernel, the steady state in one launch of
    ring_clip_kernel.  Device-timed, the clip (5 GB) larger than L2."""
    torch, lib = B.torch, B.lib
    w, h, fmt = 1920, 1080, lib.FMT_RGBX8
    fb = w * h * 4
    clip = B.ensure_buf(n * fb)[: n * fb]
    lib.synth_fill_device(B.local_rank, clip.data_ptr(), 0, n, w, h, fmt, SEED, lib.SYNTH_SCENE, B.stream.cuda_stream)
    torch.cuda.synchronize()
    res = {"geometry": f"1920x1080 RGBx8 x {n} frames, device-resident, accumulators + per-frame scalars", "unit": "frames/s"}
    for name, flavor in (("dips_ring4", lib.FLAVOR_DIPS_RING4), ("dips_alt_ring2", lib.FLAVOR_ALT_RING2)):
        with lib.Context(w, h, fmt, 0, tau, device=B.local_rank, flavor=flavor) as ctx:
            ctx.set_stream(B.stream.cuda_stream)
            ctx.reset(); ctx.run_clip_device(clip.data_ptr(), n, fb, 0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            reps = 5
            e0.record(B.stream)
            for _ in range(reps):
                ctx.reset(); ctx.run_clip_device(clip.data_ptr(), n, fb, 0)
            e1.record(B.stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            res[name] = {"value": n / ms * 1e3, "ms_per_clip": ms, "frame_GBps": n * fb / ms / 1e6, "ring_clip_kernel": bool(ctx.last_plan()["ring_clip"])}
    return res


def main():
    args = parse_args()
    out = _claim_stdout()
    wl = list(WORKLOADS[args.workload])
    if args.frames:
        wl[6] = args.frames
    wl = tuple(wl)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, wl, rank, out)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; dips_b200 has no CPU fallback")
    B = Bench(args, rank, world, local_rank)
    t_start = time.perf_counter()

    parity = None
    if world > 1 and not args.no_preflight:
        parity = B.preflight()
        log(rank, "preflight parity ok", json.dumps(parity["cases"]))

    # ---- headline workload ----------------------------------------------------------------------------------------
    if args.only_e2e:
        head, ctx, clip = B.measure(args.workload, wl, 3, 3, keep=True)
        e2e = B.e2e(wl, ctx)
        ctx.close()
        if rank == 0:
            out.write(json.dumps({"only_e2e": True, "n_gpus": world, "workload": args.workload, "e2e": e2e}) + "\n")
        return 0
    head, ctx, clip = B.measure(args.workload, wl, args.steps, args.warmup, with_probe=True, keep=True)
    log(rank, f"{args.workload}: {head['value']:.0f} frames/s, {head['ms_per_step']:.3f} ms/step, kernel {head['roofline']['achieved']:.0f} GB/s "
              f"({time.perf_counter() - t_start:.0f} s)")
    e2e = None
    if not args.no_e2e:
        e2e = B.e2e(wl, ctx)
        log(rank, f"e2e: {e2e['value']:.0f} frames/s ({time.perf_counter() - t_start:.0f} s)")
    cpu = None
    if not args.no_cpu and world == 1:
        desc, w, h, fmt, mode, tau, total = wl
        n_s = min(cpu_sample_frames(wl), max(16, int(1.1e9 // clip.shape[1])))
        host_sample = clip[:n_s].cpu().numpy()
        v, cores, what = cpu_port_throughput(wl, host_sample)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": what}
        del host_sample
    ctx.close()
    del clip

    # ---- the other BASELINE configurations, one row each ------------------------------------------------------------
    rows = []
    for name in [r for r in args.rows.split(",") if r and r != args.workload]:
        if name in ("c2", "c3") and world > 1:
            continue                    # C2 / C3 are BASELINE's one-GPU configurations; their rows belong to the N=1 line
        rwl = WORKLOADS[name]
        rsteps = max(5, min(args.steps, 40))
        r, _, _ = B.measure(name, rwl, rsteps, args.warmup)
        r["config"] = config_of(rwl, world)
        rows.append(r)
        log(rank, f"{name}: {r['value']:.0f} frames/s, {r['ms_per_step']:.3f} ms/step, kernel {r['roofline']['achieved']:.0f} GB/s "
                  f"({time.perf_counter() - t_start:.0f} s)")

    stream_info = None
    if not args.no_stream and not args.no_e2e and world == 1:
        stream_info = B.stream_boundary()
        stream_info["ring_batch"] = ring_batch(B)

    if rank == 0:
        line = {
            "metric": "frames/sec", "value": head["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config_of(wl, world),
            "plan": head["plan"], "hbm_GBps_per_gpu_whole_step": head["hbm_GBps_per_gpu_whole_step"],
            "phases": head["phases"], "roofline": head["roofline"], "sustained": head["sustained"], "cpu_baseline": cpu, "e2e": e2e,
            "configs": rows, "multi_gpu_parity": parity, "comm": head["comm"], "stream": stream_info, "checks": head["checks"],
            "gpu_launches": head["gpu_launches"], "clocks": head["clocks"],
        }
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        B.dist.barrier()
        B.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
