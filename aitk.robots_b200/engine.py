# -*- coding: utf-8 -*-
"""BatchEngine: Python host driver of libbrs.so for one batch (shard) of worlds.

Python only configures, uploads and launches; every result comes from the
sm_100a kernels behind the C ABI (include/brs.h).  PyTorch appears solely in
``tensor()`` for the zero-copy observation hand-off and for the stream handle.
"""

import ctypes as C
import math

import numpy as np

from . import _cabi as cabi
from .scenes import CameraSpec, LightSpec, RangeSpec, SceneBatch, SmellSpec

PI_OVER_180 = math.pi / 180.0

_BUF_META = {
    # name: (id, numpy dtype, shape-after-world as a function of the engine)
    "seg": (cabi.BUF_SEG, np.uint32, lambda e: (5, e.seg_cap)),
    "seg_count": (cabi.BUF_SEG_COUNT, np.int32, lambda e: ()),
    "world_size": (cabi.BUF_WORLD_SIZE, np.float32, lambda e: (2,)),
    "bulbs": (cabi.BUF_BULBS, np.float32, lambda e: (e.n_bulbs, 4)),
    "grid": (cabi.BUF_GRID, np.float32, lambda e: (e.grid_h, e.grid_w)),
    "pose": (cabi.BUF_POSE, np.float64, lambda e: (e.n_robots, 3)),
    "vel": (cabi.BUF_VEL, np.float64, lambda e: (e.n_robots, 3)),
    "target_vel": (cabi.BUF_TARGET_VEL, np.float64, lambda e: (e.n_robots, 3)),
    "stalled": (cabi.BUF_STALLED, np.uint8, lambda e: (e.n_robots,)),
    "range": (cabi.BUF_RANGE, np.float32, lambda e: (e.n_range,)),
    "camera": (cabi.BUF_CAMERA, np.uint8, lambda e: (e.n_camera, e.cam_h, e.cam_w, 4)),
    "light": (cabi.BUF_LIGHT, np.float32, lambda e: (e.n_light,)),
    "smell": (cabi.BUF_SMELL, np.float32, lambda e: (e.n_smell,)),
}


def _mount(position):
    """Mount point polar form exactly as upstream devices compute it."""
    x, y = position
    return math.sqrt((0 - x) ** 2 + (0 - y) ** 2), math.atan2(-x, y)


class _CudaBuffer:
    """Minimal __cuda_array_interface__ holder so torch can view engine memory."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {
            "shape": tuple(int(s) for s in shape),
            "typestr": typestr,
            "data": (int(ptr), False),
            "version": 2,
            "strides": None,
        }
        self._owner = owner  # keeps the engine alive while a tensor views it


class BatchEngine:
    """One engine = one CUDA device = one contiguous shard of worlds."""

    def __init__(self, scene: SceneBatch, device=0, sky=(0, 0, 128), ground=(0, 128, 0), sky_grad=(0, 0, 0),
                 ground_grad=(0, 0, 0), light_multiplier=1000.0):
        self._lib = cabi.load()
        self._h = None
        self._host_allocs = []
        self._peer_allocs, self._peer_maps = [], []
        self.scene = scene
        self.device = int(device)
        self.n_worlds = scene.n_worlds
        self.n_robots = scene.n_robots
        self.seg_cap = scene.seg_cap
        self.time_step = scene.time_step
        self.ranges = scene.devices(RangeSpec)
        self.cameras = scene.devices(CameraSpec)
        self.lights = scene.devices(LightSpec)
        self.smells = scene.devices(SmellSpec)
        self.n_range, self.n_camera = len(self.ranges), len(self.cameras)
        self.n_light, self.n_smell = len(self.lights), len(self.smells)
        self.n_bulbs = 0 if scene.bulbs is None else int(scene.bulbs.shape[1])
        self.grid_h, self.grid_w = (0, 0) if scene.grid is None else (int(scene.grid.shape[1]), int(scene.grid.shape[2]))
        self.cam_w = self.cameras[0][1].width if self.cameras else 0
        self.cam_h = self.cameras[0][1].height if self.cameras else 0

        cfg = cabi.Config()
        cfg.abi_version = cabi.BRS_ABI_VERSION
        cfg.device = self.device
        cfg.n_worlds, cfg.n_robots, cfg.seg_cap = self.n_worlds, self.n_robots, self.seg_cap
        cfg.n_bulbs, cfg.grid_w, cfg.grid_h = self.n_bulbs, self.grid_w, self.grid_h
        cfg.n_range, cfg.n_camera, cfg.n_light, cfg.n_smell = self.n_range, self.n_camera, self.n_light, self.n_smell
        cfg.smell_cell_size = float(scene.smell_cell_size)
        cfg.light_multiplier = float(light_multiplier)
        for k in range(3):
            cfg.sky_rgb[k], cfg.sky_grad[k] = float(sky[k]), float(sky_grad[k])
            cfg.ground_rgb[k], cfg.ground_grad[k] = float(ground[k]), float(ground_grad[k])
        robots = (cabi.RobotDesc * self.n_robots)()
        for r, spec in enumerate(scene.robots):
            for k in range(4):
                robots[r].bbox[k] = float(spec.bbox[k])
                robots[r].rgba[k] = int(spec.rgba[k])
            robots[r].height = float(spec.height)
            robots[r].vx_ramp, robots[r].vy_ramp, robots[r].va_ramp = spec.vx_ramp, spec.vy_ramp, spec.va_ramp
        ranges = (cabi.RangeDesc * max(self.n_range, 1))()
        for i, (r, d) in enumerate(self.ranges):
            ranges[i].robot = r
            ranges[i].dist_from_center, ranges[i].dir_from_center = _mount(d.position)
            ranges[i].a = d.a * PI_OVER_180
            ranges[i].max = float(d.max)
            ranges[i].width = d.width * PI_OVER_180
            ranges[i].n_rays = 0
        cameras = (cabi.CameraDesc * max(self.n_camera, 1))()
        for i, (r, d) in enumerate(self.cameras):
            cameras[i].robot, cameras[i].width, cameras[i].height = r, d.width, d.height
            cameras[i].dist_from_center, cameras[i].dir_from_center = _mount(d.position)
            cameras[i].fov = d.a * PI_OVER_180
            cameras[i].max_range = float(d.max_range)
            cameras[i].color_fade = float(d.colorsFadeWithDistance)
            cameras[i].size_fade = float(d.sizeFadeWithDistance)
        lights = (cabi.PointDesc * max(self.n_light, 1))()
        for i, (r, d) in enumerate(self.lights):
            lights[i].robot = r
            lights[i].dist_from_center, lights[i].dir_from_center = _mount(d.position)
        smells = (cabi.PointDesc * max(self.n_smell, 1))()
        for i, (r, d) in enumerate(self.smells):
            smells[i].robot = r
            smells[i].dist_from_center, smells[i].dir_from_center = _mount(d.position)
        cfg.robots = robots
        cfg.ranges = ranges if self.n_range else None
        cfg.cameras = cameras if self.n_camera else None
        cfg.lights = lights if self.n_light else None
        cfg.smells = smells if self.n_smell else None
        h = C.c_void_p()
        cabi.check(self._lib.brs_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._upload_scene(scene)

    # -- lifetime --------------------------------------------------------------
    def close(self):
        if self._h is not None:
            for ptr in self._host_allocs:
                self._lib.brs_host_free(self._h, ptr)
            self._host_allocs = []
            self._lib.brs_synchronize(self._h, None)
            for ptr in self._peer_maps:
                self._lib.brs_peer_close(self._h, ptr)
            for ptr in self._peer_allocs:
                self._lib.brs_peer_free(self._h, ptr)
            self._peer_allocs, self._peer_maps = [], []
            self._lib.brs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- buffers ---------------------------------------------------------------
    def _shape(self, name):
        _, dtype, tail = _BUF_META[name]
        return (self.n_worlds,) + tuple(tail(self)), dtype

    def upload(self, name, array, first_world=0, stream=0):
        bid, dtype, tail = _BUF_META[name]
        a = np.ascontiguousarray(array, dtype=dtype)
        n = a.shape[0]
        assert a.shape[1:] == tuple(tail(self)), (name, a.shape, tail(self))
        cabi.check(self._lib.brs_upload(self._h, bid, a.ctypes.data, first_world, n, stream))
        self.synchronize(stream)  # pageable source: keep `a` alive until the copy is done

    def download(self, name, first_world=0, n_worlds=None, stream=0):
        bid, dtype, tail = _BUF_META[name]
        n = self.n_worlds - first_world if n_worlds is None else n_worlds
        out = np.empty((n,) + tuple(tail(self)), dtype=dtype)
        cabi.check(self._lib.brs_download(self._h, bid, out.ctypes.data, first_world, n, stream))
        self.synchronize(stream)
        return out

    def device_ptr(self, name):
        p, nbytes = C.c_void_p(), C.c_size_t()
        cabi.check(self._lib.brs_device_ptr(self._h, _BUF_META[name][0], C.byref(p), C.byref(nbytes)))
        return p.value, nbytes.value

    def tensor(self, name):
        """Zero-copy torch view of an engine buffer (observation hand-off)."""
        import torch

        shape, dtype = self._shape(name)
        ptr, nbytes = self.device_ptr(name)
        if nbytes == 0 or 0 in shape:
            return torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), device="cuda:%d" % self.device)
        if dtype == np.uint32:  # torch has no general uint32 support: view the wall store as int32
            typestr = "<i4"
        else:
            typestr = np.dtype(dtype).str
        return torch.as_tensor(_CudaBuffer(ptr, shape, typestr, self), device="cuda:%d" % self.device)

    def bind_output(self, name, tensor_or_ptr):
        """Make the kernels write observation `name` straight into caller memory."""
        ptr = tensor_or_ptr if isinstance(tensor_or_ptr, int) or tensor_or_ptr is None else tensor_or_ptr.data_ptr()
        cabi.check(self._lib.brs_bind_output(self._h, _BUF_META[name][0], ptr))

    def peer_alloc(self, shape, dtype):
        """(zero-copy torch tensor, 64-byte CUDA IPC handle) of a fresh device buffer other
        processes can map (brs_peer_alloc); freed with the engine."""
        import torch

        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
        cabi.check(self._lib.brs_peer_alloc(self._h, n, C.byref(ptr), handle))
        self._peer_allocs.append(ptr.value)
        t = torch.as_tensor(_CudaBuffer(ptr.value, tuple(shape), np.dtype(dtype).str, self), device="cuda:%d" % self.device)
        return t, bytes(handle)

    def peer_open(self, handle):
        """Device address, in this process, of a peer's brs_peer_alloc buffer (NVLink peer access)."""
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        ptr = C.c_void_p()
        cabi.check(self._lib.brs_peer_open(self._h, buf, C.byref(ptr)))
        self._peer_maps.append(ptr.value)
        return ptr.value

    def render_world(self, first_world=0, n_worlds=None, width=256, height=None, robot=-1, view=None, line_px=0.0,
                     bulb_px=0.0, stream=0):
        """Top-down RGBA pictures [n, height, width, 4] (torch uint8, on the device) of the
        worlds' current state (brs_render_world).  ``robot`` >= 0 with ``view=(w, h)`` in world
        units: the window under / around that robot, row 0 towards its front (GroundCamera)."""
        import torch

        n = self.n_worlds - first_world if n_worlds is None else n_worlds
        if height is None:
            ws = self.scene.world_size[first_world] if view is None else view
            height = max(1, int(round(width * float(ws[1]) / float(ws[0]))))
        vw, vh = (0.0, 0.0) if view is None else (float(view[0]), float(view[1]))
        out = torch.empty((n, height, width, 4), dtype=torch.uint8, device="cuda:%d" % self.device)
        cabi.check(self._lib.brs_render_world(self._h, int(first_world), int(n), int(width), int(height), int(robot),
                                              vw, vh, float(line_px), float(bulb_px), out.data_ptr(), stream))
        return out

    def bind_peer_outputs(self, name, ptrs):
        """Fused gather: observation `name` is also stored at each address of `ptrs` (peer-GPU
        memory mapped into this process); an empty list unbinds.  See brs_bind_peer_outputs."""
        arr = (C.c_void_p * max(len(ptrs), 1))(*[int(p) for p in ptrs])
        cabi.check(self._lib.brs_bind_peer_outputs(self._h, _BUF_META[name][0], len(ptrs), arr))

    def _upload_scene(self, scene):
        B, cap = self.n_worlds, self.seg_cap
        store = np.empty((B, 5, cap), dtype=np.uint32)
        store[:, :4, :] = np.ascontiguousarray(scene.seg, dtype=np.float32).view(np.uint32)
        store[:, 4, :] = np.ascontiguousarray(scene.seg_rgba, dtype=np.uint8).reshape(B, cap, 4).view(np.uint32)[:, :, 0]
        self.upload("seg", store)
        self.upload("seg_count", scene.seg_count)
        self.upload("world_size", scene.world_size)
        if self.n_bulbs:
            self.upload("bulbs", scene.bulbs)
        if scene.grid is not None:
            self.upload("grid", scene.grid)
        self.upload("pose", scene.pose)
        self.upload("target_vel", scene.target_vel)
        self.upload("vel", np.zeros_like(scene.pose))

    # -- stepping --------------------------------------------------------------
    def step(self, time_step=None, flags=cabi.STEP_ALL, stream=0):
        """World.step over the whole batch: one fused kernel launch."""
        dt = self.time_step if time_step is None else time_step
        cabi.check(self._lib.brs_step(self._h, float(dt), int(flags), stream))

    def steps(self, n, time_step=None, flags=cabi.STEP_ALL, stream=0):
        """World.steps(n) over the whole batch: n consecutive steps with the current target
        velocities, in ONE launch where the plan allows (brs_steps); bit-identical to n step()s."""
        dt = self.time_step if time_step is None else time_step
        cabi.check(self._lib.brs_steps(self._h, float(dt), int(flags), int(n), stream))

    def steps_launches(self, n, flags=cabi.STEP_ALL):
        """Kernel launches steps(n, flags) issues on this engine as it is configured now."""
        k = C.c_int()
        cabi.check(self._lib.brs_steps_launches(self._h, int(flags), int(n), C.byref(k)))
        return int(k.value)

    def steps_info(self):
        """Launch shape of a one-launch steps(): worlds per tile, tiles, CTAs, shared memory."""
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        cabi.check(self._lib.brs_steps_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"worlds_per_tile": a.value, "tiles": b.value, "ctas": c.value, "smem_bytes": d.value}

    def step_host(self, time_step=None, flags=cabi.STEP_ALL, target_vel=0, pose=0, stalled=0, range_=0, camera=0,
                  light=0, smell=0, stream=0):
        """The same step with HOST (pinned) buffers given as raw addresses (0 = skip)."""
        dt = self.time_step if time_step is None else time_step
        cabi.check(
            self._lib.brs_step_host(self._h, float(dt), int(flags), target_vel, pose, stalled, range_, camera, light,
                                    smell, stream)
        )

    def host_buffer(self, shape, dtype):
        """Pinned host array for step_host(), allocated by the C ABI on the NUMA node of this
        engine's GPU (brs_host_alloc); freed with the engine."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        cabi.check(self._lib.brs_host_alloc(self._h, n, C.byref(ptr)))
        self._host_allocs.append(ptr.value)
        buf = (C.c_char * max(n, 1)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def set_noise(self, seed, range_std=0.0, light_rel_std=0.0, first_world=0):
        """Seeded sensor noise (off by default): Philox4x32-10 keyed by `seed`, one sample per
        (first_world + world, sensor, step); see include/brs.h brs_set_noise."""
        cabi.check(self._lib.brs_set_noise(self._h, int(seed), int(first_world), float(range_std), float(light_rel_std)))

    def step_counted(self, time_step=None, flags=cabi.STEP_ALL, stream=0):
        """One real step with the executed work counted on the device:
        (ray-segment tests the search loops ran, box-edge tests the collision classifier ran)."""
        dt = self.time_step if time_step is None else time_step
        a, b = C.c_ulonglong(), C.c_ulonglong()
        cabi.check(self._lib.brs_step_counted(self._h, float(dt), int(flags), stream, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def count_tests(self, flags=cabi.STEP_ALL, stream=0):
        n = C.c_ulonglong()
        cabi.check(self._lib.brs_count_tests(self._h, int(flags), stream, C.byref(n)))
        return int(n.value)

    def launch_info(self):
        t, c, s, m = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        cabi.check(self._lib.brs_launch_info(self._h, C.byref(t), C.byref(c), C.byref(s), C.byref(m)))
        tw, nt, nw, nr, ng = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        cabi.check(self._lib.brs_tile_info(self._h, C.byref(tw), C.byref(nt), C.byref(nw), C.byref(nr), C.byref(ng)))
        wide, nl = C.c_int(), C.c_int()
        cabi.check(self._lib.brs_path_info(self._h, C.byref(wide), C.byref(nl)))
        return {"path": ("fused (one kernel)", "wide (prep + step kernel, PDL)", "fused fed by the prep kernel (PDL)",
                         "solo (one warp per world, one kernel)",
                         "solo range (one warp per world, walls streamed, one kernel)",
                         "solo multi (one warp per world, several robots, one kernel)",
                         "mini range (a team of 2..16 lanes per world, one kernel)")[wide.value],
                "launches_per_step": nl.value, "threads": t.value, "ctas": c.value, "smem_bytes": s.value,
                "sm_count": m.value, "worlds_per_tile": tw.value, "tiles": nt.value, "workers": nw.value,
                "cols_per_thread": nr.value, "ray_groups": ng.value}

    def synchronize(self, stream=0):
        cabi.check(self._lib.brs_synchronize(self._h, stream))


def measure_fp32_peak(device=0, iters=4096):
    lib = cabi.load()
    tf, mhz = C.c_double(), C.c_double()
    cabi.check(lib.brs_measure_fp32_peak(int(device), int(iters), C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value


def measure_write_bw(device=0, nbytes=512 << 20, mode=0, reps=5):
    """GB/s of plain streaming 16-byte stores (the PAINT phase without its work)."""
    lib = cabi.load()
    g = C.c_double()
    cabi.check(lib.brs_measure_write_bw(int(device), int(nbytes), int(mode), int(reps), C.byref(g)))
    return g.value
