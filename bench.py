#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""bench.py -- world-steps/sec and ray-segment tests/sec of the batched
World.step hot path (BASELINE.json metric) on N B200s, beside the reference's
CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config config2] [--worlds B]
    python bench.py --impl reference ...      # the pure-Python CPU path, all host cores

A "step" is one World.step (move + all sensors + camera picture) of every world
of the batch: ONE fused kernel launch per GPU.  N > 1 is weak scaling: every
rank steps its own `--worlds` worlds, no collective on the data path.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_TEST = 13
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # SMs x lanes x 2 x max clock


def make_scene(rb, config, n_worlds, first_world):
    g = rb.scenes
    if config == "config1":
        return g.config1(n_worlds=n_worlds, first_world=first_world)
    if config == "config2":
        return g.config2(n_worlds=n_worlds, first_world=first_world)
    if config == "config3":
        return g.config3(n_worlds=n_worlds, first_world=first_world)
    if config == "config4":
        return g.config4(n_worlds=n_worlds, first_world=first_world)
    if config.startswith("config5"):
        _, S, R = config.split("-")  # config5-<segments>-<rays>
        return g.config5(n_worlds=n_worlds, n_segments=int(S), n_rays=int(R), first_world=first_world)
    raise SystemExit("unknown config %r" % config)


DEFAULT_WORLDS = {"config1": 4096, "config2": 4096, "config3": 65536, "config4": 131072}


def describe(config, scene):
    import aitk_robots_b200 as rb

    cams = scene.devices(rb.CameraSpec)
    return {
        "workload": "%s: %d robots/world, %d segments/world, %d RangeSensors, %s, %d LightSensors x %d bulbs"
        % (config, scene.n_robots, int(scene.seg_count[0]), len(scene.devices(rb.RangeSpec)),
           ("%d Camera %dx%d" % (len(cams), cams[0][1].width, cams[0][1].height)) if cams else "no camera",
           len(scene.devices(rb.LightSpec)), 0 if scene.bulbs is None else scene.bulbs.shape[1]),
        "worlds_per_gpu": scene.n_worlds,
    }


# ------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed regions (B200_PROFILING.md), read
    in-process through NVML (nvidia_ml_py) -- spawning nvidia-smi every few milliseconds
    perturbs a 0.1 ms kernel -- with nvidia-smi as the fallback."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES renumbers devices: resolve the CUDA ordinal through its UUID
            import torch

            uuid = str(torch.cuda.get_device_properties(index).uuid)
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [getattr(n, "nvmlClocksEventReasonHwSlowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                getattr(n, "nvmlClocksEventReasonSwPowerCap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4))]
        return [str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for b in bits]

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                         timeout=5).stdout
                    parts = [p.strip() for p in out.strip().split(",")]
                    if len(parts) >= 6:
                        self.rows.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.002 if self.nvml is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k] == "Active" for r in self.rows)]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------ CPU legs
def _cpu_worker(args):
    """Steps `count` worlds of the scene slice once through the pure-Python oracle
    (World.step + every Camera.take_picture); returns (world_steps, tests, seconds)."""
    config, first, count, n_steps = args
    import aitk_robots_b200 as rb
    from oracle import scene_adapter as sa

    scene = make_scene(rb, config, count, first)
    t0 = time.perf_counter()
    tests = 0
    for i in range(count):
        world, devs = sa.build_world(scene, i)
        world.counted_tests = 0
        for _ in range(n_steps):
            world.step()
            for cam in devs["camera"]:
                cam.take_picture()
        tests += world.counted_tests
    return count * n_steps, tests, time.perf_counter() - t0


def _real_aitk():
    """The real reference engine, if the driver installed one (baseline/_ref)."""
    try:
        import aitk.robots  # noqa: F401

        return True
    except Exception:
        pass
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref):
        sys.path.insert(0, ref)
        try:
            import aitk.robots  # noqa: F401

            return True
        except Exception:
            sys.path.remove(ref)
    return False


_POOL = None


def _pool(cores):
    global _POOL
    if _POOL is None:
        import multiprocessing as mp

        import atexit

        _POOL = mp.get_context("spawn").Pool(cores)
        atexit.register(lambda: (_POOL.close(), _POOL.join()))
        _POOL.map(_cpu_worker, [("config1", 0, 1, 1)] * cores)  # imports done before anything is timed
    return _POOL


def cpu_throughput(config, target_seconds, cores=None, first_world=10_000_000):
    """Oracle ("port") throughput over all host cores on a bounded sample of the workload."""
    cores = cores or os.cpu_count() or 1
    # calibrate on one world-step
    ws, tests, dt = _cpu_worker((config, first_world, 1, 1))
    per_world = max(dt, 1e-4)
    n_per_core = max(1, int(target_seconds / per_world))
    jobs = [(config, first_world + 1 + c * n_per_core, n_per_core, 1) for c in range(cores)]
    pool = _pool(cores)
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    # workers import + build scenes inside their own timers: use the slowest worker's compute time
    busy = max(r[2] for r in res)
    wsteps, tests = sum(r[0] for r in res), sum(r[1] for r in res)
    return {"world_steps_per_sec": wsteps / busy, "tests_per_sec": tests / busy, "cores": cores,
            "sample": "%d worlds x 1 World.step (+take_picture) of %s, %d per core, pure-Python float64 port"
                      % (wsteps, config, n_per_core), "wall_s": wall, "busy_s": busy}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The real
    aitk.robots cannot be installed here: the in-repo port stands in."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = "port"
    if _real_aitk():
        kind = "port"  # the adapter to the real package is tests/test_real_aitk.py; timing still uses the port
    per_step_seconds = max(1.0, min(8.0, 150.0 / max(args.steps + args.warmup, 1)))
    vals, tests = [], []
    sample = ""
    for k in range(args.warmup + args.steps):
        r = cpu_throughput(args.config, per_step_seconds if k >= args.warmup else 1.0)
        if k >= args.warmup:
            vals.append(r["world_steps_per_sec"])
            tests.append(r["tests_per_sec"])
            sample, cores = r["sample"], r["cores"]
    v = sum(vals) / len(vals)
    import aitk_robots_b200 as rb

    cfg = describe(args.config, make_scene(rb, args.config, 2, 0))
    cfg["worlds_per_gpu"] = args.worlds
    line = {
        "impl": "reference", "metric": "world-steps/sec", "value": v, "unit": "world-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        # a "step" of this arm is a bounded sample; ms_per_step extrapolates it to the full batch
        "ms_per_step": 1e3 * args.worlds * max(args.gpus, 1) / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "ray_segment_tests_per_sec": sum(tests) / len(tests),
        "cpu_baseline": {"value": v, "unit": "world-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="config2")
    ap.add_argument("--worlds", type=int, default=0, help="worlds per GPU (default: the BASELINE.json size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-large-batch", action="store_true",
                    help="skip the supplementary run at 4x the worlds (shows the launch prologue/tail share)")
    ap.add_argument("--gather", action="store_true", help="also time the optional NCCL observation all-gather")
    args = ap.parse_args()
    if not args.worlds:
        args.worlds = DEFAULT_WORLDS.get(args.config, 262144 // 8)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        ge.build()
    if world_size > 1:
        dist.barrier()
    import aitk_robots_b200 as rb

    B = args.worlds
    scene = make_scene(rb, args.config, B, rank * B)
    eng = rb.BatchEngine(scene, device=local_rank)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    flags = rb.STEP_ALL
    tests_per_step = eng.count_tests(flags, sh)
    assert tests_per_step == scene.tests_per_step(), (tests_per_step, scene.tests_per_step())

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps, one kernel launch each -------------------
    for _ in range(max(args.warmup, 3)):
        eng.step(flags=flags, stream=sh)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get("BRS_BENCH_NO_CLOCKS") else None
    if sampler:
        sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record(stream)
    for k in range(args.steps):
        eng.step(flags=flags, stream=sh)
        evs[k + 1].record(stream)
    barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    launch_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]

    # ---- end to end: host buffers in, host buffers out, through the C ABI ----------
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)  # noqa: E731
    h_tvel = pin((B, eng.n_robots, 3), torch.float64)
    h_tvel.copy_(torch.from_numpy(scene.target_vel))
    h_pose = pin((B, eng.n_robots, 3), torch.float64)
    h_stalled = pin((B, eng.n_robots), torch.uint8)
    h_range = pin((B, max(eng.n_range, 1)), torch.float32)
    h_cam = pin((B, max(eng.n_camera, 1), max(eng.cam_h, 1), max(eng.cam_w, 1), 4), torch.uint8)
    h_light = pin((B, max(eng.n_light, 1)), torch.float32)
    h_smell = pin((B, max(eng.n_smell, 1)), torch.float32)
    h2d = h_tvel.numel() * 8
    d2h = (h_pose.numel() * 8 + h_stalled.numel() + eng.n_range * B * 4 + eng.n_camera * B * eng.cam_h * eng.cam_w * 4
           + eng.n_light * B * 4 + eng.n_smell * B * 4)

    def e2e_step():
        eng.step_host(flags=flags, target_vel=h_tvel.data_ptr(), pose=h_pose.data_ptr(), stalled=h_stalled.data_ptr(),
                      range_=h_range.data_ptr() if eng.n_range else 0, camera=h_cam.data_ptr() if eng.n_camera else 0,
                      light=h_light.data_ptr() if eng.n_light else 0, smell=h_smell.data_ptr() if eng.n_smell else 0,
                      stream=sh)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(3):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None  # sampled over the device-timed and the end-to-end regions

    # ---- optional observation all-gather (the only collective of the path) ---------
    gather = None
    if args.gather and world_size > 1:
        # the kernels write this rank's slice of the gather buffer directly (brs_bind_output),
        # the all-gather is in place: no staging copy precedes the collective
        name = "camera" if eng.n_camera else "range"
        local = eng.tensor(name)
        full = torch.empty((B * world_size,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        mine = full[rank * B:(rank + 1) * B]
        eng.bind_output(name, mine)
        for _ in range(3):
            eng.step(flags=flags, stream=sh)
            dist.all_gather_into_tensor(full, mine)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(args.steps):
            eng.step(flags=flags, stream=sh)
            dist.all_gather_into_tensor(full, mine)
        g1.record(stream)
        barrier()
        eng.bind_output(name, None)
        gather = {"observation": name, "bytes_per_rank_per_step": int(mine.numel() * mine.element_size()),
                  "ms_per_step_with_gather": g0.elapsed_time(g1) / args.steps, "in_place": True}

    # ---- max over ranks ---------------------------------------------------------------
    t = torch.tensor([total_ms, e2e_ms, float(tests_per_step)], dtype=torch.float64, device="cuda")
    if world_size > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms, all_tests = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        all_tests = float(tests_per_step)

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world_size * B * args.steps / (total_ms * 1e-3)
        tests_per_sec = all_tests * args.steps / (total_ms * 1e-3)
        e2e_value = world_size * B * e2e_steps / (e2e_ms * 1e-3)
        # roofline of the (only) kernel: algorithmic HBM bytes per launch, this rank
        R = eng.n_robots
        bytes_read = B * (5 * eng.seg_cap * 4 + 4 + 8 + 3 * R * 24) + B * eng.n_light * eng.n_bulbs * 16
        bytes_write = B * (2 * R * 24 + R + eng.n_range * 4 + eng.n_camera * eng.cam_h * eng.cam_w * 4
                           + eng.n_light * 4 + eng.n_smell * 4)
        alg_bytes = bytes_read + bytes_write
        avg_launch_ms = sum(launch_ms) / len(launch_ms)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            hbm_peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
        else:
            hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = alg_bytes / (avg_launch_ms * 1e-3) / 1e9
        try:
            ffma_tflops, _ = rb.measure_fp32_peak(local_rank, 4096)
        except Exception:
            ffma_tflops = None
        fp32_tflops = tests_per_step * FLOPS_PER_TEST / (avg_launch_ms * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.config)
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        hbm = {"achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
               "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes}
        try:
            # context: what a plain streaming-store kernel reaches on this GPU (the path is write-only)
            wc = rb.measure_write_bw(local_rank, 512 << 20, 0, 5)
            hbm["write_only_ceiling_gbs"] = wc
            hbm["frac_of_write_only_ceiling"] = achieved / wc
        except Exception:
            pass
        fp32 = {"achieved": fp32_tflops, "peak": ffma_tflops or NOMINAL_FP32_TFLOPS, "unit": "TFLOP/s",
                "frac": fp32_tflops / (ffma_tflops or NOMINAL_FP32_TFLOPS),
                "peak_source": "measured on this GPU (brs_measure_fp32_peak, dependent FFMA chains)" if ffma_tflops
                else "nominal 148 SMs x 128 lanes x 2 x 1.965 GHz",
                "peak_nominal": NOMINAL_FP32_TFLOPS, "flops_per_test": FLOPS_PER_TEST,
                "algorithmic_tests_per_launch": tests_per_step}
        # the bound of this workload is the roofline it sits closer to: pictures are HBM-write
        # bound (config2), ray-heavy worlds with small pictures FP32-issue bound (config4/5)
        dom = hbm if hbm["frac"] >= fp32["frac"] else fp32
        roofline = {"bound": "hbm" if dom is hbm else "fp32", "achieved": dom["achieved"], "peak": dom["peak"],
                    "unit": dom["unit"], "frac": dom["frac"], "traffic": traffic, "peak_source": dom["peak_source"],
                    "kernel": "brs_step_kernel", "avg_launch_ms": avg_launch_ms, "hbm": hbm, "fp32": fp32}
        line = {
            "metric": "world-steps/sec", "value": value, "unit": "world-steps/s", "n_gpus": world_size,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(describe(args.config, scene),
                           l2="no explicit flush: every step streams %.0f MB of fresh observations through a 126 MB L2"
                              % (bytes_write / 1e6),
                           precision="FP32 ray search + FP64 winner refinement, FP64 motion/collision",
                           launch=eng.launch_info()),
            "ray_segment_tests_per_sec": tests_per_sec,
            "ray_segment_tests_per_step": all_tests,
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": "world-steps/s", "h2d_bytes_per_step": int(h2d) * world_size,
                    "d2h_bytes_per_step": int(d2h) * world_size, "ms_per_step": e2e_ms / e2e_steps,
                    "api": "brs_step_host (C ABI, pinned host buffers)"},
            "gpu_launches": args.steps,
            "clocks": clocks,
        }
        if gather:
            line["gather"] = gather
        if world_size == 1 and not args.no_large_batch and 4 * alg_bytes < 16e9:
            # supplementary: the same workload at 4x the worlds.  A launch of the named size lasts
            # ~0.1 ms, of which the pipeline fill (first tile of every CTA) and the tail are a fixed
            # ~20 us; the larger batch shows the kernel's steady-state rate.  Not the headline.
            eng.close()
            eng = None
            big = rb.BatchEngine(make_scene(rb, args.config, 4 * B, 50_000_000), device=local_rank)
            for _ in range(3):
                big.step(flags=flags, stream=sh)
            torch.cuda.synchronize()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record(stream)
            for _ in range(args.steps):
                big.step(flags=flags, stream=sh)
            b1.record(stream)
            torch.cuda.synchronize()
            bms = b0.elapsed_time(b1) / args.steps
            line["large_batch"] = {"worlds_per_gpu": 4 * B, "ms_per_step": bms, "value": 4 * B / (bms * 1e-3),
                                   "hbm_frac": 4 * alg_bytes / (bms * 1e-3) / 1e9 / hbm_peak,
                                   "fp32_frac": 4 * tests_per_step * FLOPS_PER_TEST / (bms * 1e-3) / 1e12
                                   / (ffma_tflops or NOMINAL_FP32_TFLOPS)}
            big.close()
        if not args.no_cpu_baseline and world_size == 1:
            r = cpu_throughput(args.config, args.cpu_seconds)
            line["cpu_baseline"] = {"value": r["world_steps_per_sec"], "unit": "world-steps/s", "cores": r["cores"],
                                    "kind": "port", "sample": r["sample"],
                                    "ray_segment_tests_per_sec": r["tests_per_sec"]}
        print(json.dumps(line))
    if eng is not None:
        eng.close()
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
