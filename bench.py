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
    if config == "config2-boxes":
        return g.config2_boxes(n_worlds=n_worlds, first_world=first_world)
    if config == "config3":
        return g.config3(n_worlds=n_worlds, first_world=first_world)
    if config == "config4":
        return g.config4(n_worlds=n_worlds, first_world=first_world)
    if config == "config4-random":  # the random-segment variant of the 100-segment worlds
        return g.config4(n_worlds=n_worlds, first_world=first_world, variant="random")
    if config.startswith("config5"):
        _, S, R = config.split("-")  # config5-<segments>-<rays>
        return g.config5(n_worlds=n_worlds, n_segments=int(S), n_rays=int(R), first_world=first_world)
    raise SystemExit("unknown config %r" % config)


DEFAULT_WORLDS = {"config1": 4096, "config2": 4096, "config2-boxes": 4096, "config3": 65536, "config4": 131072,
                  "config4-random": 131072}


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
def _ref_module(use_real):
    """The module that supplies World / Robot / devices on the CPU legs: the real
    ``aitk.robots`` when it is importable (also from baseline/_ref) and asked for, else the
    in-repo float64 restatement (oracle/)."""
    from oracle import aitk_oracle as orc
    from oracle import scene_adapter as sa

    if use_real:
        mod = sa.find_real_aitk(ROOT)
        if mod is not None:
            return mod, "aitk"
    return orc, "port"


def _cpu_worker(args):
    """Steps `count` worlds of the scene slice `n_steps` times on the CPU (World.step + every
    Camera.take_picture); returns (world_steps, tests, seconds)."""
    config, first, count, n_steps, use_real = args
    import aitk_robots_b200 as rb
    from oracle import scene_adapter as sa

    mod, kind = _ref_module(use_real)
    scene = make_scene(rb, config, count, first)
    t0 = time.perf_counter()
    tests = 0
    for i in range(count):
        world, devs = sa.build_world(scene, i, mod)
        if kind == "port":
            world.counted_tests = 0
        for _ in range(n_steps):
            world.step()
            for cam in devs["camera"]:
                cam.take_picture()
        tests += world.counted_tests if kind == "port" else scene.slice(i, 1).tests_per_step() * n_steps
    return count * n_steps, tests, time.perf_counter() - t0


_POOL = None


def _pool(cores, use_real):
    global _POOL
    if _POOL is None:
        import multiprocessing as mp

        import atexit

        _POOL = mp.get_context("spawn").Pool(cores)
        atexit.register(lambda: (_POOL.close(), _POOL.join()))
        _POOL.map(_cpu_worker, [("config1", 0, 1, 1, use_real)] * cores)  # imports done before anything is timed
    return _POOL


def cpu_throughput(config, target_seconds, cores=None, first_world=10_000_000, use_real=True):
    """Reference-CPU throughput over all host cores on a bounded sample of the workload."""
    cores = cores or os.cpu_count() or 1
    _, kind = _ref_module(use_real)
    # calibrate on one world-step
    ws, tests, dt = _cpu_worker((config, first_world, 1, 1, use_real))
    per_world = max(dt, 1e-4)
    n_per_core = max(1, int(target_seconds / per_world))
    jobs = [(config, first_world + 1 + c * n_per_core, n_per_core, 1, use_real) for c in range(cores)]
    pool = _pool(cores, use_real)
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    # workers import + build scenes inside their own timers: use the slowest worker's compute time
    busy = max(r[2] for r in res)
    wsteps, tests = sum(r[0] for r in res), sum(r[1] for r in res)
    what = "the real aitk.robots package" if kind == "aitk" else "pure-Python float64 port (oracle/)"
    return {"world_steps_per_sec": wsteps / busy, "tests_per_sec": tests / busy, "cores": cores, "kind": kind,
            "sample": "%d worlds x 1 World.step (+take_picture) of %s, %d per core, %s"
                      % (wsteps, config, n_per_core, what), "wall_s": wall, "busy_s": busy}


def config1_as_written(use_real=True, n_steps=1000):
    """BASELINE.json configs[0] exactly as written: ONE 200x150 world, 1 robot with 1
    RangeSensor (max 100), 4 walls, 1000 World.steps, a single CPU process, noise off."""
    ws, tests, dt = _cpu_worker(("config1", 0, 1, n_steps, use_real))
    _, kind = _ref_module(use_real)
    return {"world_steps_per_sec": ws / dt, "seconds_for_1000_steps": dt * 1000.0 / n_steps, "steps": n_steps,
            "tests_per_sec": tests / dt, "cores": 1, "kind": kind}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host
    cores -- the real aitk.robots package when it is importable (kind "aitk"), else the
    in-repo float64 port (kind "port"; the mounted reference tree ships no engine)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step_seconds = max(1.0, min(8.0, 150.0 / max(args.steps + args.warmup, 1)))
    vals, tests = [], []
    sample, cores, kind = "", 1, "port"
    for k in range(args.warmup + args.steps):
        r = cpu_throughput(args.config, per_step_seconds if k >= args.warmup else 1.0)
        if k >= args.warmup:
            vals.append(r["world_steps_per_sec"])
            tests.append(r["tests_per_sec"])
            sample, cores, kind = r["sample"], r["cores"], r["kind"]
    v = sum(vals) / len(vals)
    import aitk_robots_b200 as rb

    cfg = describe(args.config, make_scene(rb, args.config, 2, 0))
    cfg["worlds_per_gpu"] = args.worlds
    line = {
        "impl": "reference", "metric": "world-steps/sec", "value": v, "unit": "world-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        # a "step" of this arm is a bounded sample of the batch: ms_per_step is that sample's rate
        # EXTRAPOLATED to the whole N-GPU batch, not a measured duration
        "ms_per_step": 1e3 * args.worlds * max(args.gpus, 1) / v, "ms_per_step_extrapolated": True,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "ray_segment_tests_per_sec": sum(tests) / len(tests),
        "cpu_baseline": {"value": v, "unit": "world-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.config == "config2" and not args.no_extra:
        line["config1_as_written"] = config1_as_written()
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm
def bind_to_gpu_numa_node(torch, local_rank):
    """Run this rank on the CPUs next to its GPU, BEFORE any pinned allocation, so that the
    host buffers of the end-to-end path are first-touched on the GPU's NUMA node (8 ranks
    draining pictures into one node's memory is what capped the round-1 aggregate at ~92 GB/s)."""
    info = {"bound": False}
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        node = int(open(base + "/numa_node").read().strip())
        cpulist = open(base + "/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        info.update({"pci": bus, "numa_node": node, "local_cpus": len(cpus)})
        if use and node >= 0:
            os.sched_setaffinity(0, use)
            info["bound"] = True
    except Exception as e:  # no sysfs / no permission: stay where we are
        info["error"] = str(e)[:80]
    return info


class Bench:
    """One rank of the GPU arm: shared plumbing of the per-workload measurements."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- this engine has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.numa = bind_to_gpu_numa_node(torch, self.local_rank)
        if self.world_size > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        import __graft_entry__ as ge

        if self.rank == 0:
            ge.build()
        if self.world_size > 1:
            dist.barrier()
        import aitk_robots_b200 as rb

        self.rb = rb
        self.stream = torch.cuda.current_stream()
        self.sh = self.stream.cuda_stream
        self.launches = 0
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            self.hbm_peak, self.peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
        else:
            self.hbm_peak, self.peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            self.ffma_tflops, _ = rb.measure_fp32_peak(self.local_rank, 4096)
        except Exception:
            self.ffma_tflops = None

    def barrier(self):
        if self.world_size > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op="max"):
        """max / sum over ranks of a list of floats."""
        if self.world_size == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t]

    def time_steps(self, eng, flags, steps, warmup, per_launch=False):
        """K device-timed steps (CUDA events on the launching stream, barrier + synchronize on
        both sides); returns (total ms of this rank, per-step ms list or None)."""
        torch = self.torch
        for _ in range(max(warmup, 3)):
            eng.step(flags=flags, stream=self.sh)
        self.barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range((steps + 1) if per_launch else 2)]
        self.barrier()
        evs[0].record(self.stream)
        for k in range(steps):
            eng.step(flags=flags, stream=self.sh)
            if per_launch:
                evs[k + 1].record(self.stream)
        if not per_launch:
            evs[1].record(self.stream)
        self.barrier()
        self.launches += steps * eng.launch_info()["launches_per_step"]
        total = evs[0].elapsed_time(evs[-1])
        return total, ([evs[k].elapsed_time(evs[k + 1]) for k in range(steps)] if per_launch else None)

    def time_steps_call(self, eng, flags, steps, warmup):
        """The same K steps as ONE call of brs_steps (upstream World.steps(K): K consecutive
        steps, every one writing all its observations; one launch where the plan allows, the
        pipeline running on across the step boundaries).  Returns (total ms, launches)."""
        torch = self.torch
        for _ in range(max(warmup, 3)):
            eng.step(flags=flags, stream=self.sh)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        eng.steps(steps, flags=flags, stream=self.sh)
        e1.record(self.stream)
        self.barrier()
        n = eng.steps_launches(steps, flags)
        self.launches += n
        return e0.elapsed_time(e1), n

    def rooflines(self, eng, ms, tests_per_step, executed):
        """Both rooflines of one launch of `eng` that took `ms`."""
        B, R = eng.n_worlds, eng.n_robots
        bytes_read = B * (5 * eng.seg_cap * 4 + 4 + 8 + 3 * R * 24) + B * eng.n_light * eng.n_bulbs * 16
        bytes_write = B * (2 * R * 24 + R + eng.n_range * 4 + eng.n_camera * eng.cam_h * eng.cam_w * 4
                           + eng.n_light * 4 + eng.n_smell * 4)
        alg = bytes_read + bytes_write
        achieved = alg / (ms * 1e-3) / 1e9
        hbm = {"achieved": achieved, "peak": self.hbm_peak, "unit": "GB/s", "frac": achieved / self.hbm_peak,
               "peak_source": self.peak_src, "algorithmic_bytes_per_launch": alg, "bytes_written_per_launch": bytes_write}
        peak = self.ffma_tflops or NOMINAL_FP32_TFLOPS
        tf = tests_per_step * FLOPS_PER_TEST / (ms * 1e-3) / 1e12
        fp32 = {"achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                "peak_source": "measured on this GPU (brs_measure_fp32_peak, dependent FFMA chains)" if self.ffma_tflops
                else "nominal 148 SMs x 128 lanes x 2 x 1.965 GHz",
                "peak_nominal": NOMINAL_FP32_TFLOPS, "flops_per_test": FLOPS_PER_TEST,
                "algorithmic_tests_per_launch": tests_per_step,
                "note": "frac counts every (ray, segment) pair of the workload at 13 flops (SURVEY.md 8(d): early exits "
                        "still count); the executed_* keys count what the search loops really ran after the culls"}
        if executed is not None:
            ray_t, edge_t = executed
            fp32["executed_ray_tests_per_launch"] = ray_t
            fp32["executed_edge_tests_per_launch"] = edge_t
            fp32["executed_tests_per_sec"] = (ray_t + edge_t) / (ms * 1e-3)
            fp32["executed_frac"] = (ray_t + edge_t) * FLOPS_PER_TEST / (ms * 1e-3) / 1e12 / peak
            fp32["culled_share"] = 1.0 - (ray_t + edge_t) / max(tests_per_step, 1)
        return hbm, fp32

    def workload(self, config, B, steps, warmup, e2e=False, per_launch=False, keep=False):
        """Device-timed steps of one configuration on every rank (weak scaling: B worlds each).
        Returns (summary dict on rank 0 else None, engine if keep)."""
        rb = self.rb
        scene = make_scene(rb, config, B, self.rank * B)
        eng = rb.BatchEngine(scene, device=self.local_rank)
        flags = {"all": rb.STEP_ALL, "move": rb.STEP_MOVE, "camera": rb.STEP_CAMERA, "sense": rb.STEP_SENSE,
                 "move+camera": rb.STEP_MOVE | rb.STEP_CAMERA}[self.args.flags]
        tests_per_step = eng.count_tests(flags, self.sh)
        if flags == rb.STEP_ALL:
            assert tests_per_step == scene.tests_per_step(), (tests_per_step, scene.tests_per_step())
        total_ms, launch_ms = self.time_steps(eng, flags, steps, warmup, per_launch)
        executed = None
        try:
            executed = eng.step_counted(flags=flags, stream=self.sh)
        except Exception:
            pass
        # the same K steps as one brs_steps call, on a fresh engine stepped through the same
        # states (a step's cost drifts as the robots move, so both timings see the same steps)
        eng2 = rb.BatchEngine(scene, device=self.local_rank)
        call_ms, call_launches = self.time_steps_call(eng2, flags, steps, warmup)
        eng2.close()
        total_ms, call_ms = self.reduce([total_ms, call_ms])
        all_tests, = self.reduce([float(tests_per_step)], "sum")
        ms = total_ms / steps
        out = None
        if self.rank == 0:
            avg_launch = sum(launch_ms) / len(launch_ms) if launch_ms else ms
            hbm, fp32 = self.rooflines(eng, avg_launch, tests_per_step, executed)
            cms = call_ms / steps
            chbm, cfp32 = self.rooflines(eng, cms, tests_per_step, executed)
            # one launch = `steps` world-steps per world: per-launch figures scale, rates do not
            chbm["steps_per_launch"] = cfp32["steps_per_launch"] = steps if call_launches == 1 else 1
            if call_launches == 1:
                for d, keys in ((chbm, ("algorithmic_bytes_per_launch", "bytes_written_per_launch")),
                                (cfp32, ("algorithmic_tests_per_launch", "executed_ray_tests_per_launch",
                                         "executed_edge_tests_per_launch"))):
                    for k in keys:
                        if k in d:
                            d[k] *= steps
            out = {"config": describe(config, scene), "ms_per_step": ms, "avg_launch_ms": avg_launch,
                   "value": self.world_size * B / (ms * 1e-3), "unit": "world-steps/s",
                   "ray_segment_tests_per_sec": all_tests / (ms * 1e-3), "ray_segment_tests_per_step": all_tests,
                   "hbm": hbm, "fp32": fp32, "launch": eng.launch_info(), "scene": scene,
                   "steps_call": {"api": "brs_steps(K): World.steps(K)", "launches": call_launches, "steps": steps,
                                  "shape": eng.steps_info(),
                                  "launch_ms": call_ms, "ms_per_step": cms,
                                  "value": self.world_size * B / (cms * 1e-3),
                                  "ray_segment_tests_per_sec": all_tests / (cms * 1e-3), "hbm": chbm, "fp32": cfp32}}
        if keep:
            return out, eng, scene
        eng.close()
        return out, None, None


def e2e_measure(bench, eng, scene, steps):
    """The same step through brs_step_host: target velocities host->device, every observation
    device->host, pinned buffers, all inside the timed region."""
    torch, rb = bench.torch, bench.rb
    import numpy as np

    B = eng.n_worlds
    # pinned host buffers from the C ABI: first-touched on this GPU's NUMA node (brs_host_alloc)
    hb = lambda shape, dt: torch.from_numpy(eng.host_buffer(shape, dt))  # noqa: E731
    h_tvel = hb((B, eng.n_robots, 3), np.float64)
    h_tvel.copy_(torch.from_numpy(scene.target_vel))
    h_pose = hb((B, eng.n_robots, 3), np.float64)
    h_stalled = hb((B, eng.n_robots), np.uint8)
    h_range = hb((B, max(eng.n_range, 1)), np.float32)
    h_cam = hb((B, max(eng.n_camera, 1), max(eng.cam_h, 1), max(eng.cam_w, 1), 4), np.uint8)
    h_light = hb((B, max(eng.n_light, 1)), np.float32)
    h_smell = hb((B, max(eng.n_smell, 1)), np.float32)
    h2d = h_tvel.numel() * 8
    d2h = (h_pose.numel() * 8 + h_stalled.numel() + eng.n_range * B * 4 + eng.n_camera * B * eng.cam_h * eng.cam_w * 4
           + eng.n_light * B * 4 + eng.n_smell * B * 4)
    flags = rb.STEP_ALL

    def e2e_step():
        eng.step_host(flags=flags, target_vel=h_tvel.data_ptr(), pose=h_pose.data_ptr(), stalled=h_stalled.data_ptr(),
                      range_=h_range.data_ptr() if eng.n_range else 0, camera=h_cam.data_ptr() if eng.n_camera else 0,
                      light=h_light.data_ptr() if eng.n_light else 0, smell=h_smell.data_ptr() if eng.n_smell else 0,
                      stream=bench.sh)

    n = max(3, min(steps, 10))
    for _ in range(3):
        e2e_step()
    bench.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(bench.stream)
    for _ in range(n):
        e2e_step()
    e1.record(bench.stream)
    bench.barrier()
    bench.launches += n * eng.launch_info()["launches_per_step"]
    e2e_ms = e0.elapsed_time(e1)
    # the host ceiling: the same bytes as ONE plain device->host copy per rank, all ranks at
    # once, no kernel -- what the PCIe links and the host memory system deliver on this box
    big, big_name = (h_cam, "camera") if eng.n_camera else (h_pose, "pose")

    def plain_copy():
        from aitk_robots_b200 import _cabi as cabi
        from aitk_robots_b200.engine import _BUF_META

        cabi.check(eng._lib.brs_download(eng._h, _BUF_META[big_name][0], big.data_ptr(), 0, B, bench.sh))

    for _ in range(2):
        plain_copy()
    bench.barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(bench.stream)
    for _ in range(n):
        plain_copy()
    c1.record(bench.stream)
    bench.barrier()
    copy_ms = c0.elapsed_time(c1)
    # context: the same loop for a consumer that lives on the GPU (an RL policy reading the zero-copy
    # picture tensor): actions host->device, poses / stall flags / range readings device->host, the
    # pictures painted every step but left in HBM
    small_ms, small_d2h = None, 0
    if eng.n_camera:
        def small_step():
            eng.step_host(flags=flags, target_vel=h_tvel.data_ptr(), pose=h_pose.data_ptr(), stalled=h_stalled.data_ptr(),
                          range_=h_range.data_ptr() if eng.n_range else 0, camera=0,
                          light=h_light.data_ptr() if eng.n_light else 0, smell=h_smell.data_ptr() if eng.n_smell else 0,
                          stream=bench.sh)

        for _ in range(3):
            small_step()
        bench.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(bench.stream)
        for _ in range(n):
            small_step()
        s1.record(bench.stream)
        bench.barrier()
        bench.launches += n * eng.launch_info()["launches_per_step"]
        small_ms = s0.elapsed_time(s1)
        small_d2h = d2h - eng.n_camera * B * eng.cam_h * eng.cam_w * 4
    e2e_ms, copy_ms, small_ms = bench.reduce([e2e_ms, copy_ms, small_ms if small_ms is not None else 0.0])
    ws = bench.world_size
    nbytes = big.numel() * big.element_size()
    extra = {}
    if eng.n_camera:
        extra["pictures_left_on_device"] = {
            "value": ws * B * n / (small_ms * 1e-3), "unit": "world-steps/s", "ms_per_step": small_ms / n,
            "h2d_bytes_per_step": int(h2d) * ws, "d2h_bytes_per_step": int(small_d2h) * ws,
            "note": "same brs_step_host loop, every picture painted each step but read by a consumer on the GPU "
                    "(zero-copy tensor) instead of being shipped to the host: what the PCIe drain costs"}
    return {**extra, "value": ws * B * n / (e2e_ms * 1e-3), "unit": "world-steps/s", "h2d_bytes_per_step": int(h2d) * ws,
            "d2h_bytes_per_step": int(d2h) * ws, "ms_per_step": e2e_ms / n,
            "api": "brs_step_host (C ABI, pinned host buffers; kernel of chunk c+1 overlaps the copies of chunk c)",
            "achieved_d2h_gbs": d2h * ws * n / (e2e_ms * 1e-3) / 1e9,
            "host_ceiling_gbs": nbytes * ws * n / (copy_ms * 1e-3) / 1e9,
            "host_ceiling_note": "aggregate GB/s of one plain cudaMemcpyAsync D2H of the pictures per rank, all ranks "
                                 "at once, same pinned buffers",
            "numa": bench.numa}


def gather_measure(bench, eng, steps, name, overlap):
    """World.step + the observation all-gather (the path's only collective), in place: the
    kernels write this rank's slice of the gather buffer (brs_bind_output).  `overlap`: two
    gather buffers, the all-gather of step t runs on NCCL's stream while step t+1 computes."""
    torch, dist, rb = bench.torch, bench.dist, bench.rb
    B, ws, rank = eng.n_worlds, bench.world_size, bench.rank
    local = eng.tensor(name)
    nbuf = 2 if overlap else 1
    full = [torch.empty((B * ws,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) for _ in range(nbuf)]
    mine = [f[rank * B:(rank + 1) * B] for f in full]
    bindable = name in ("camera", "range", "light", "smell")
    works = [None] * nbuf

    def one(t):
        b = t % nbuf
        if works[b] is not None:
            works[b].wait()  # the buffer's previous gather is done before the kernel overwrites its slice
            works[b] = None
        if bindable:
            eng.bind_output(name, mine[b])
        eng.step(flags=rb.STEP_ALL, stream=bench.sh)
        if not bindable:
            mine[b].copy_(local)
        if overlap:
            works[b] = dist.all_gather_into_tensor(full[b], mine[b], async_op=True)
        else:
            dist.all_gather_into_tensor(full[b], mine[b])

    for t in range(4):
        one(t)
    for w in works:
        if w is not None:
            w.wait()
    works = [None] * nbuf
    bench.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(bench.stream)
    for t in range(steps):
        one(t)
    for w in works:
        if w is not None:
            w.wait()
    g1.record(bench.stream)
    bench.barrier()
    bench.launches += steps * eng.launch_info()["launches_per_step"]
    if bindable:
        eng.bind_output(name, None)
    ms, = bench.reduce([g0.elapsed_time(g1) / steps])
    per_rank = int(mine[0].numel() * mine[0].element_size())
    inbound = per_rank * (ws - 1)
    del full, mine
    torch.cuda.empty_cache()
    return {"observation": name, "bytes_per_rank_per_step": per_rank, "ms_per_step_with_gather": ms, "in_place": bindable,
            "overlapped_with_next_step": overlap,
            "nvlink_floor_ms": inbound / 900e9 * 1e3,
            "nvlink_floor_note": "(N-1) x bytes_per_rank inbound per GPU / 900 GB/s per direction (NVLink 5)"}


def fused_gather_measure(bench, eng, steps, name="camera"):
    """World.step with the picture all-gather INSIDE the step kernel: every rank maps every other
    rank's gather buffer (CUDA IPC over NVLink, brs_peer_alloc / brs_peer_open) and PAINT stores
    each picture quad locally and into all peers (brs_bind_peer_outputs).  What remains of the
    collective is a one-element all-reduce as the cross-rank stream barrier."""
    import numpy as np

    torch, dist, rb = bench.torch, bench.dist, bench.rb
    B, ws, rank = eng.n_worlds, bench.world_size, bench.rank
    local = eng.tensor(name)
    shape = (B * ws,) + tuple(local.shape[1:])
    full, handle = eng.peer_alloc(shape, np.uint8)
    mine_h = torch.tensor(list(handle), dtype=torch.uint8, device=local.device)
    all_h = torch.empty((ws, 64), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(all_h, mine_h)
    all_h = all_h.cpu().numpy()
    stride = int(np.prod(shape[1:]))
    ptrs = [eng.peer_open(all_h[r].tobytes()) + rank * B * stride for r in range(ws) if r != rank]
    mine = full[rank * B:(rank + 1) * B]
    eng.bind_output(name, mine)
    eng.bind_peer_outputs(name, ptrs)
    token = torch.zeros(1, device=local.device)

    def one():
        eng.step(flags=rb.STEP_ALL, stream=bench.sh)
        dist.all_reduce(token)  # every rank's step (and with it its stores into this GPU) is complete

    for _ in range(4):
        one()
    bench.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(bench.stream)
    for _ in range(steps):
        one()
    g1.record(bench.stream)
    bench.barrier()
    bench.launches += steps * eng.launch_info()["launches_per_step"]
    # every rank must now hold every rank's pictures: compare a few remote slices with their owners'
    probe = torch.stack([full[r * B + (B // 2)].to(torch.int64).sum() for r in range(ws)])
    owners = [torch.zeros_like(probe) for _ in range(ws)]
    dist.all_gather(owners, probe)
    complete = all(int(owners[r][r]) == int(probe[r]) and int(probe[r]) > 0 for r in range(ws))
    eng.bind_peer_outputs(name, [])
    eng.bind_output(name, None)
    ms, = bench.reduce([g0.elapsed_time(g1) / steps])
    per_rank = int(mine.numel() * mine.element_size())
    return {"observation": name, "bytes_per_rank_per_step": per_rank, "ms_per_step_with_gather": ms, "in_place": True,
            "fused_into_step_kernel": True, "every_rank_holds_every_picture": bool(complete),
            "nvlink_floor_ms": per_rank * (ws - 1) / 900e9 * 1e3,
            "note": "PAINT stores each quad locally and into the (N-1) peer gather buffers over NVLink; the collective "
                    "that remains is a one-element all-reduce as the cross-rank barrier"}


def summarize(w):
    """Sub-object of the JSON line for one extra workload."""
    return {"workload": w["config"]["workload"], "worlds_per_gpu": w["config"]["worlds_per_gpu"],
            "ms_per_step": w["ms_per_step"], "world_steps_per_sec": w["value"],
            "ray_segment_tests_per_sec": w["ray_segment_tests_per_sec"],
            "hbm_frac": w["hbm"]["frac"], "fp32_frac": w["fp32"]["frac"],
            "fp32_executed_frac": w["fp32"].get("executed_frac"), "culled_share": w["fp32"].get("culled_share"),
            "executed_tests_per_sec": w["fp32"].get("executed_tests_per_sec"),
            "steps_call": {"launches": w["steps_call"]["launches"], "ms_per_step": w["steps_call"]["ms_per_step"],
                           "world_steps_per_sec": w["steps_call"]["value"],
                           "ray_segment_tests_per_sec": w["steps_call"]["ray_segment_tests_per_sec"],
                           "hbm_frac": w["steps_call"]["hbm"]["frac"], "fp32_frac": w["steps_call"]["fp32"]["frac"],
                           "fp32_executed_frac": w["steps_call"]["fp32"].get("executed_frac")},
            "launch": {k: w["launch"][k] for k in ("path", "launches_per_step", "threads", "ctas", "smem_bytes")}}


def pipe_utilisation(config):
    """ncu pipe counters of this workload's step kernel from the committed profile summary
    (profiles/r2/pipes.json, written by tools/ncu_pipes.py from the same commit's capture):
    a bench run cannot read hardware counters itself."""
    pth = os.path.join(ROOT, "profiles", "r2", "pipes.json")
    if os.path.exists(pth):
        try:
            return json.load(open(pth)).get(config)
        except Exception:
            return None
    return None


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
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the other BASELINE.json configurations (config1/3/4 sub-objects of the default line)")
    ap.add_argument("--flags", default="all", choices=["all", "move", "camera", "sense", "move+camera"],
                    help="tuning: which parts of World.step the device-timed steps run (the default line always uses all)")
    ap.add_argument("--gather", action="store_true", help="also time the optional NCCL observation all-gather")
    args = ap.parse_args()
    if not args.worlds:
        args.worlds = DEFAULT_WORLDS.get(args.config, 262144 // 8)
    if args.impl == "reference":
        return run_reference(args)

    bench = Bench(args)
    rb, torch = bench.rb, bench.torch
    rank, world_size, local_rank = bench.rank, bench.world_size, bench.local_rank
    B = args.worlds
    extras = args.config == "config2" and not args.no_extra

    # ---- the headline workload: device-resident timing, one launch per step -----------
    sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get("BRS_BENCH_NO_CLOCKS") else None
    if sampler:
        sampler.start()
    head, eng, scene = bench.workload(args.config, B, args.steps, args.warmup, per_launch=True, keep=True)
    # ---- end to end: host buffers in, host buffers out, through the C ABI ----------
    e2e = e2e_measure(bench, eng, scene, args.steps)
    clocks = sampler.stop() if sampler else None  # sampled over the device-timed and the end-to-end regions
    gather = None
    if args.gather and world_size > 1:
        gather = gather_measure(bench, eng, args.steps, "camera" if eng.n_camera else "range", overlap=False)
    eng.close()
    eng = None

    line = None
    if rank == 0:
        # headline: the K timed steps are ONE call of brs_steps (World.steps(K)); the same K steps
        # as K calls of brs_step are reported beside it (`per_step_launches`)
        sc = head["steps_call"]
        hbm, fp32 = sc["hbm"], sc["fp32"]
        try:
            # context: what a plain streaming-store kernel reaches on this GPU (the path is write-only)
            modes = {"grid-stride st.cs": 0, "cudaMemsetAsync": 2, "CTA-contiguous 128 KB chunks of st.cs": 3}
            got = {k: rb.measure_write_bw(local_rank, 2048 << 20, m, 5) for k, m in modes.items()}
            wc = max(got.values())
            hbm["write_only_ceiling_gbs"] = wc
            hbm["write_only_ceiling_by_pattern"] = got
            hbm["frac_of_write_only_ceiling"] = hbm["achieved"] / wc
            hbm["note"] = ("peak is MEASURED_PEAKS.json's COPY bandwidth (read + write bytes of b.copy_(a)); this path only "
                           "writes, and a write-only stream can exceed a copy's byte rate -- write_only_ceiling_gbs is the "
                           "best plain store kernel / memset on this GPU, measured in this run")
        except Exception:
            pass
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r2", "traffic_%s.json" % args.config)
        if os.path.exists(tp):
            tj = json.load(open(tp))
            # (per launch, like `achieved`: the capture's per-step DRAM bytes x the steps of this launch)
            per_step = tj.get("dram_bytes_per_step", tj.get("dram_bytes_per_launch"))
            traffic = per_step * (args.steps // sc["launches"]) if per_step is not None else None
            traffic_src = tj.get("source")
        pipes = pipe_utilisation(args.config)
        if pipes:
            fp32["ncu_pipes"] = pipes
        # the bound of this workload is the roofline it sits closer to by EXECUTED work: pictures
        # are HBM-write bound (config2), ray-heavy worlds with small pictures issue bound (config4/5)
        fp_exec = fp32.get("executed_frac", fp32["frac"])
        dom = hbm if hbm["frac"] >= fp_exec else fp32
        roofline = {"bound": "hbm" if dom is hbm else "fp32", "achieved": dom["achieved"], "peak": dom["peak"],
                    "unit": dom["unit"], "frac": dom["frac"], "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": dom["peak_source"], "kernel": "brs_step_kernel" if head["launch"]["path"].startswith("fused")
                    else ("brs_mini_kernel" if head["launch"]["path"].startswith("mini") else
                          "brs_scan_kernel" if head["launch"]["path"].startswith("solo range") else
                          "brs_solo_kernel" if head["launch"]["path"].startswith("solo") else "brs_wide_kernel"),
                    "avg_launch_ms": sc["launch_ms"] / sc["launches"], "world_steps_per_launch": B * args.steps // sc["launches"],
                    "hbm": hbm, "fp32": fp32}
        line = {
            "metric": "world-steps/sec", "value": sc["value"], "unit": "world-steps/s", "n_gpus": world_size,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": sc["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": head["config"],
            "l2": "no explicit flush: every step streams %.0f MB of fresh observations through a 126 MB L2"
                  % (head["hbm"]["bytes_written_per_launch"] / 1e6),
            "launch_mode": "the %d timed steps are one brs_steps call (upstream World.steps(n)) = %d kernel launch(es): every "
                           "step writes all of its observations, a tile's steps are ordered by per-tile epochs, the "
                           "kernel's pipeline runs on across step boundaries" % (args.steps, sc["launches"]),
            "steps_launch_shape": sc["shape"],
            "per_step_launches": {"api": "brs_step x %d" % args.steps, "ms_per_step": head["ms_per_step"],
                                  "value": head["value"], "avg_launch_ms": head["avg_launch_ms"],
                                  "hbm_frac": head["hbm"]["frac"], "fp32_frac": head["fp32"]["frac"],
                                  "ray_segment_tests_per_sec": head["ray_segment_tests_per_sec"]},
            "precision": "FP32 ray search + FP64 winner refinement, FP64 motion/collision",
            "launch": head["launch"],
            "ray_segment_tests_per_sec": sc["ray_segment_tests_per_sec"],
            "ray_segment_tests_per_step": head["ray_segment_tests_per_step"],
            "roofline": roofline,
            "e2e": e2e,
            "clocks": clocks,
        }
        if gather:
            line["gather"] = gather

    # ---- the other BASELINE.json configurations, same run (sub-objects of the line) ----
    if world_size == 1 and not args.no_large_batch and rank == 0 and 4 * head["hbm"]["algorithmic_bytes_per_launch"] < 16e9:
        # supplementary: the same workload at 4x the worlds.  A launch of the named size lasts
        # ~0.1 ms, of which the pipeline fill (first tile of every CTA) and the tail are a fixed
        # ~20 us; the larger batch shows the kernel's steady-state rate.  Not the headline.
        big, _, _ = bench.workload(args.config, 4 * B, args.steps, 3)
        line["large_batch"] = {"worlds_per_gpu": 4 * B, "ms_per_step": big["steps_call"]["ms_per_step"],
                               "value": big["steps_call"]["value"],
                               "hbm_frac": big["steps_call"]["hbm"]["frac"], "fp32_frac": big["steps_call"]["fp32"]["frac"],
                               "per_step_launches": {"ms_per_step": big["ms_per_step"], "value": big["value"],
                                                     "hbm_frac": big["hbm"]["frac"]},
                               "note": "hbm_frac is against the measured COPY peak (reads + writes); a write-only stream "
                                       "can exceed it"}
    if extras:
        ex = {}
        steps_x = max(5, min(args.steps, 20))
        if world_size == 1:
            # config 1 on the GPU: ONE world, 1000 sequential World.steps (launch-latency bound)
            scene1 = make_scene(rb, "config1", 1, 0)
            e1 = rb.BatchEngine(scene1, device=local_rank)
            for _ in range(20):
                e1.step(stream=bench.sh)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a0.record(bench.stream)
            for _ in range(1000):
                e1.step(stream=bench.sh)
            a1.record(bench.stream)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            bench.launches += 1000
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            b0.record(bench.stream)
            e1.steps(1000, stream=bench.sh)  # World.steps(1000): config 1 exactly as BASELINE.json writes it
            b1.record(bench.stream)
            torch.cuda.synchronize()
            wall_call = time.perf_counter() - t0
            bench.launches += e1.steps_launches(1000)
            ex["config1"] = {"workload": describe("config1", scene1)["workload"], "worlds": 1, "steps": 1000,
                             "gpu_seconds_for_1000_steps": b0.elapsed_time(b1) * 1e-3, "host_wall_seconds": wall_call,
                             "world_steps_per_sec": 1000.0 / (b0.elapsed_time(b1) * 1e-3),
                             "api": "World.steps(1000) = one brs_steps call, %d launch" % e1.steps_launches(1000),
                             "per_step_launches": {"gpu_seconds_for_1000_steps": a0.elapsed_time(a1) * 1e-3,
                                                   "host_wall_seconds": wall,
                                                   "world_steps_per_sec": 1000.0 / (a0.elapsed_time(a1) * 1e-3)},
                             "note": "one world cannot fill a GPU: this is latency; the batched path is config1 x 4096"}
            e1.close()
            w, _, _ = bench.workload("config1", 4096, steps_x, 3)
            ex["config1_x4096"] = summarize(w)
            w, _, _ = bench.workload("config2-boxes", 4096, steps_x, 3)
            ex["config2_boxes"] = summarize(w)  # the "20 boxes = 80 segments" reading of configs[1]
            w, _, _ = bench.workload("config3", DEFAULT_WORLDS["config3"], steps_x, 3)
            ex["config3"] = summarize(w)
            w, _, _ = bench.workload("config4-random", DEFAULT_WORLDS["config4-random"], steps_x, 3)
            ex["config4_random_segments"] = summarize(w)  # config 4 with 100 random segments instead of a grid maze
        # config 4: 1,048,576 worlds split over the ranks (N = 1: one eighth, the 8-GPU shard)
        per_rank = (1 << 20) // max(world_size, 8 if world_size == 1 else world_size)
        w, e4, s4 = bench.workload("config4", per_rank, steps_x, 3, keep=True)
        if rank == 0:
            ex["config4"] = summarize(w)
            ex["config4"]["worlds_total"] = per_rank * world_size
            ex["config4"]["sharding"] = ("1,048,576 worlds / %d GPUs" % world_size) if world_size > 1 \
                else "one GPU's shard of the 8-GPU split (131,072 worlds)"
            pipes = pipe_utilisation("config4")
            if pipes:
                ex["config4"]["ncu_pipes"] = pipes
        if world_size > 1:
            g_serial = gather_measure(bench, e4, steps_x, "camera", overlap=False)
            g_over = gather_measure(bench, e4, steps_x, "camera", overlap=True)
            g_small = gather_measure(bench, e4, steps_x, "pose", overlap=True)
            try:
                g_fused = fused_gather_measure(bench, e4, steps_x)
            except Exception as exc:  # no peer access on this box: the NCCL numbers stand
                g_fused = {"unavailable": str(exc)[:200]}
            if rank == 0:
                ex["config4"]["gather_pictures"] = g_serial
                ex["config4"]["gather_pictures_overlapped"] = g_over
                ex["config4"]["gather_pictures_fused"] = g_fused
                ex["config4"]["gather_poses_only"] = g_small
        e4.close()
        if rank == 0:
            line["configs"] = ex
    if rank == 0:
        line["gpu_launches"] = bench.launches
        if not args.no_cpu_baseline and world_size == 1:
            r = cpu_throughput(args.config, args.cpu_seconds)
            line["cpu_baseline"] = {"value": r["world_steps_per_sec"], "unit": "world-steps/s", "cores": r["cores"],
                                    "kind": r["kind"], "sample": r["sample"],
                                    "ray_segment_tests_per_sec": r["tests_per_sec"]}
            if extras:
                line["configs"]["config1_cpu_as_written"] = config1_as_written()
        for v in line.get("configs", {}).values():
            v.pop("scene", None)
        print(json.dumps(line))
    if world_size > 1:
        bench.dist.barrier()
        bench.dist.destroy_process_group()


if __name__ == "__main__":
    main()
