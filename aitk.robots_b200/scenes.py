# -*- coding: utf-8 -*-
"""Scene description of a batch of worlds + seeded synthetic generators.

``SceneBatch`` is the host-side, numpy form of the HBM layout (DESIGN.md
"Data layout"): wall segments as SoA float32 rows, robot poses as float64,
device lists shared by every world of the batch.  The generators produce the
five BASELINE.json configurations; all coordinates are
float32-representable so the float64 reference sees bit-identical inputs.
"""

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

PI_OVER_180 = math.pi / 180.0

PALETTE = np.array(
    [
        (255, 0, 0, 255),
        (0, 255, 0, 255),
        (0, 0, 255, 255),
        (255, 255, 0, 255),
        (0, 255, 255, 255),
        (255, 0, 255, 255),
        (255, 165, 0, 255),
        (128, 0, 128, 255),
    ],
    dtype=np.uint8,
)


@dataclass
class RangeSpec:
    """RangeSensor(position, a [deg], max, width [deg]) -- upstream devices/rangesensors.py."""

    position: Tuple[float, float] = (8.0, 0.0)
    a: float = 0.0
    max: float = 20.0
    width: float = 1.0
    name: str = "sensor"


@dataclass
class CameraSpec:
    """Camera(width, height, a [fov deg], ...) -- upstream devices/cameras.py."""

    width: int = 64
    height: int = 32
    a: float = 60.0
    colorsFadeWithDistance: float = 0.9
    sizeFadeWithDistance: float = 1.0
    max_range: float = 1000.0
    position: Tuple[float, float] = (0.0, 0.0)
    name: str = "camera"


@dataclass
class LightSpec:
    position: Tuple[float, float] = (0.0, 0.0)
    name: str = "light"


@dataclass
class SmellSpec:
    position: Tuple[float, float] = (0.0, 0.0)
    name: str = "smell"


@dataclass
class RobotSpec:
    """Robot slot shared by every world of a batch -- upstream robot.py Robot.__init__."""

    bbox: Tuple[float, float, float, float] = (-7.5, 7.5, -6.67, 6.67)
    height: float = 0.25
    rgba: Tuple[int, int, int, int] = (255, 0, 0, 255)
    max_trans_velocity: float = 2.0
    max_rotate_velocity: float = math.pi / 4
    vx_ramp: float = 1.0
    vy_ramp: float = 1.0
    va_ramp: float = 1.0
    devices: list = field(default_factory=list)

    def radius(self):
        return max(math.hypot(x, y) for x in self.bbox[:2] for y in self.bbox[2:])


@dataclass
class SceneBatch:
    world_size: np.ndarray  # float32 [B, 2] width, height
    seg: np.ndarray  # float32 [B, 4, cap] rows x1, y1, x2, y2
    seg_rgba: np.ndarray  # uint8 [B, cap, 4]
    seg_count: np.ndarray  # int32 [B]
    pose: np.ndarray  # float64 [B, R, 3] x, y, a (radians)
    target_vel: np.ndarray  # float64 [B, R, 3] tvx, tvy, tva
    robots: List[RobotSpec]
    bulbs: Optional[np.ndarray] = None  # float32 [B, NB, 4] x, y, z, brightness
    grid: Optional[np.ndarray] = None  # float32 [B, GH, GW] smell field
    smell_cell_size: float = 0.0
    food: Optional[np.ndarray] = None  # float32 [B, NF, 3] x, y, std (grid source, for the oracle)
    time_step: float = 0.1
    name: str = "scene"

    @property
    def n_worlds(self):
        return int(self.world_size.shape[0])

    @property
    def n_robots(self):
        return len(self.robots)

    @property
    def seg_cap(self):
        return int(self.seg.shape[2])

    def devices(self, kind):
        """[(robot_slot, spec)] of one device class over all robot slots, slot-major."""
        out = []
        for r, rob in enumerate(self.robots):
            for d in rob.devices:
                if isinstance(d, kind):
                    out.append((r, d))
        return out

    def slice(self, first, count):
        """Contiguous world range as its own batch (multi-GPU sharding is a slice)."""
        s = slice(first, first + count)
        return SceneBatch(
            world_size=self.world_size[s],
            seg=self.seg[s],
            seg_rgba=self.seg_rgba[s],
            seg_count=self.seg_count[s],
            pose=self.pose[s],
            target_vel=self.target_vel[s],
            robots=self.robots,
            bulbs=None if self.bulbs is None else self.bulbs[s],
            grid=None if self.grid is None else self.grid[s],
            smell_cell_size=self.smell_cell_size,
            food=None if self.food is None else self.food[s],
            time_step=self.time_step,
            name=self.name,
        )

    def tests_per_step(self, camera=True, sense=True, move=True):
        """Algorithmic ray-segment tests of one World.step over the batch."""
        R = self.n_robots
        rays = 0
        if move:
            rays += 4 * R
        if sense:
            for _, d in self.devices(RangeSpec):
                rays += 3 if d.width != 0 else 1
            nb = 0 if self.bulbs is None else self.bulbs.shape[1]
            rays += len(self.devices(LightSpec)) * nb
        if camera:
            for _, d in self.devices(CameraSpec):
                rays += d.width
        return int(rays * (self.seg_count.astype(np.int64) + 4 * (R - 1)).sum())


def smell_grid(world_size, food, cell):
    """Smell field sampled at cell centres: sum of gaussians around the food (float64 -> float32)."""
    B = world_size.shape[0]
    gw = int(math.ceil(float(world_size[:, 0].max()) / cell))
    gh = int(math.ceil(float(world_size[:, 1].max()) / cell))
    cx = (np.arange(gw, dtype=np.float64) + 0.5) * cell
    cy = (np.arange(gh, dtype=np.float64) + 0.5) * cell
    grid = np.zeros((B, gh, gw), dtype=np.float64)
    for f in range(food.shape[1]):
        fx = food[:, f, 0].astype(np.float64)[:, None, None]
        fy = food[:, f, 1].astype(np.float64)[:, None, None]
        sd = food[:, f, 2].astype(np.float64)[:, None, None]
        d2 = (cx[None, None, :] - fx) ** 2 + (cy[None, :, None] - fy) ** 2
        grid += np.exp(-d2 / (2.0 * sd * sd))
    return grid.astype(np.float32)


# ----------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------


def _boundary(w, h):
    return [(0.0, 0.0, w, 0.0), (0.0, 0.0, 0.0, h), (0.0, h, w, h), (w, 0.0, w, h)]


def _rand_segments(rng, B, n, w, h, min_len=10.0, max_len=None):
    """n random segments per world with both endpoints inside the world."""
    x1 = rng.uniform(0, w, (B, n)).astype(np.float32)
    y1 = rng.uniform(0, h, (B, n)).astype(np.float32)
    ang = rng.uniform(0, 2 * math.pi, (B, n))
    hi = max_len if max_len else 0.5 * min(w, h)
    ln = rng.uniform(min_len, hi, (B, n))
    x2 = np.clip(x1 + ln * np.cos(ang), 0, w).astype(np.float32)
    y2 = np.clip(y1 + ln * np.sin(ang), 0, h).astype(np.float32)
    return x1, y1, x2, y2


def _point_seg_dist(px, py, x1, y1, x2, y2):
    """Distance of points [B, P] to segments [B, S] -> [B, P, S] (float64)."""
    px, py = px[:, :, None].astype(np.float64), py[:, :, None].astype(np.float64)
    x1, y1, x2, y2 = (a[:, None, :].astype(np.float64) for a in (x1, y1, x2, y2))
    dx, dy = x2 - x1, y2 - y1
    l2 = dx * dx + dy * dy
    t = np.where(l2 > 0, ((px - x1) * dx + (py - y1) * dy) / np.where(l2 > 0, l2, 1.0), 0.0)
    t = np.clip(t, 0.0, 1.0)
    return np.hypot(px - (x1 + t * dx), py - (y1 + t * dy))


def _place_robots(rng, seg, seg_count, w, h, R, radius, tries=200, clear_extra=2.0):
    """Rejection-sample R collision-free robot centres per world (float32-representable)."""
    B = seg.shape[0]
    clear = radius + clear_extra
    px = np.zeros((B, R), dtype=np.float32)
    py = np.zeros((B, R), dtype=np.float32)
    valid_seg = np.arange(seg.shape[2])[None, :] < seg_count[:, None]
    for r in range(R):
        todo = np.ones(B, dtype=bool)
        for _ in range(tries):
            n = int(todo.sum())
            if n == 0:
                break
            cx = rng.uniform(clear, w - clear, n).astype(np.float32)
            cy = rng.uniform(clear, h - clear, n).astype(np.float32)
            s = seg[todo]
            d = _point_seg_dist(cx[:, None], cy[:, None], s[:, 0], s[:, 1], s[:, 2], s[:, 3])[:, 0, :]
            d = np.where(valid_seg[todo], d, np.inf)
            ok = d.min(axis=1) > clear
            if r > 0:
                dr = np.hypot(px[todo, :r].astype(np.float64) - cx[:, None], py[todo, :r].astype(np.float64) - cy[:, None])
                ok &= dr.min(axis=1) > 2 * clear
            good = np.flatnonzero(todo)[ok]
            px[good, r] = cx[ok]
            py[good, r] = cy[ok]
            todo[good] = False
        if todo.any():
            raise RuntimeError("could not place robot %d in %d worlds" % (r, int(todo.sum())))
    return px, py


def _assemble(name, w, h, segs, colors, robots, rng, B, bulbs=None, food=None, smell_cell=0.0, pose=None, vel=None,
              tries=200, clear_extra=2.0):
    x1, y1, x2, y2 = segs
    n = x1.shape[1]
    cap = (n + 3) // 4 * 4
    seg = np.zeros((B, 4, cap), dtype=np.float32)
    seg[:, 0, :n], seg[:, 1, :n], seg[:, 2, :n], seg[:, 3, :n] = x1, y1, x2, y2
    rgba = np.zeros((B, cap, 4), dtype=np.uint8)
    rgba[:, :n] = colors
    count = np.full(B, n, dtype=np.int32)
    R = len(robots)
    if pose is None:
        radius = max(r.radius() for r in robots)
        px, py = _place_robots(rng, seg, count, w, h, R, radius, tries=tries, clear_extra=clear_extra)
        pa = rng.uniform(0, 2 * math.pi, (B, R))
        pose = np.stack([px.astype(np.float64), py.astype(np.float64), pa], axis=2)
    if vel is None:
        tr = rng.uniform(-1, 1, (B, R))
        ro = rng.uniform(-1, 1, (B, R))
        vel = np.zeros((B, R, 3), dtype=np.float64)
        for r, rob in enumerate(robots):
            vel[:, r, 0] = tr[:, r] * rob.max_trans_velocity
            vel[:, r, 2] = ro[:, r] * rob.max_rotate_velocity
    ws = np.tile(np.array([[w, h]], dtype=np.float32), (B, 1))
    grid = None
    if food is not None:
        grid = smell_grid(ws, food, smell_cell)
    return SceneBatch(
        world_size=ws,
        seg=seg,
        seg_rgba=rgba,
        seg_count=count,
        pose=np.ascontiguousarray(pose, dtype=np.float64),
        target_vel=np.ascontiguousarray(vel, dtype=np.float64),
        robots=robots,
        bulbs=bulbs,
        grid=grid,
        smell_cell_size=smell_cell,
        food=food,
        name=name,
    )


def _rng(seed, first_world):
    return np.random.Generator(np.random.PCG64([seed, first_world]))


# ----------------------------------------------------------------------------
# BASELINE.json configurations
# ----------------------------------------------------------------------------


def config1(n_worlds=1, seed=1, first_world=0):
    """Single 200x150 world, 1 robot with 1 RangeSensor (max 100), 4 walls + boundary."""
    rng = _rng(seed, first_world)
    w, h, B = 200.0, 150.0, n_worlds
    bx = np.array(_boundary(w, h), dtype=np.float32)
    r = _rand_segments(rng, B, 4, w, h, min_len=10.0)
    segs = [np.concatenate([np.tile(bx[None, :, k], (B, 1)), r[k]], axis=1) for k in range(4)]
    colors = np.concatenate(
        [np.tile(np.array([[(128, 0, 128, 255)] * 4], dtype=np.uint8), (B, 1, 1)), PALETTE[rng.integers(0, 8, (B, 4))]],
        axis=1,
    )
    robot = RobotSpec(devices=[RangeSpec(position=(8.0, 0.0), a=0.0, max=100.0, width=0.0)])
    vel = np.zeros((B, 1, 3))
    vel[:, 0, 0] = 0.5 * robot.max_trans_velocity
    vel[:, 0, 2] = 0.2 * robot.max_rotate_velocity
    scene = _assemble("config1", w, h, segs, colors, [robot], rng, B, vel=vel)
    return scene


def config2(n_worlds=4096, seed=2000, first_world=0, cam_w=256, cam_h=128):
    """1 robot with 2 RangeSensors + 256x128 Camera in a 20-wall room."""
    rng = _rng(seed, first_world)
    w, h, B = 300.0, 200.0, n_worlds
    bx = np.array(_boundary(w, h), dtype=np.float32)
    r = _rand_segments(rng, B, 16, w, h, min_len=15.0, max_len=80.0)
    segs = [np.concatenate([np.tile(bx[None, :, k], (B, 1)), r[k]], axis=1) for k in range(4)]
    colors = np.concatenate(
        [np.tile(np.array([[(128, 0, 128, 255)] * 4], dtype=np.uint8), (B, 1, 1)), PALETTE[rng.integers(0, 8, (B, 16))]],
        axis=1,
    )
    robot = RobotSpec(
        devices=[
            RangeSpec(position=(7.0, -3.0), a=30.0, max=100.0, width=0.0, name="left-ir"),
            RangeSpec(position=(7.0, 3.0), a=-30.0, max=100.0, width=0.0, name="right-ir"),
            CameraSpec(width=cam_w, height=cam_h, a=60.0),
        ]
    )
    return _assemble("config2", w, h, segs, colors, [robot], rng, B)


def config2_boxes(n_worlds=4096, seed=2100, first_world=0, cam_w=256, cam_h=128, n_boxes=20):
    """The other reading of BASELINE.json configs[1] ("a 20-wall room" = 20 box-shaped walls, as
    upstream ``World.add_wall(..., box=True)`` makes them): 80 wall segments + the boundary."""
    rng = _rng(seed, first_world)
    w, h, B = 300.0, 200.0, n_worlds
    bx = np.array(_boundary(w, h), dtype=np.float32)
    cx = rng.uniform(15, w - 15, (B, n_boxes))
    cy = rng.uniform(15, h - 15, (B, n_boxes))
    hw = rng.uniform(2.0, 12.0, (B, n_boxes))
    hh = rng.uniform(2.0, 12.0, (B, n_boxes))
    x0, x1_, y0, y1_ = cx - hw, cx + hw, cy - hh, cy + hh
    # four edges per box, in add_wall(box=True) order
    ex1 = np.stack([x0, x1_, x1_, x0], axis=2).reshape(B, -1)
    ey1 = np.stack([y0, y0, y1_, y1_], axis=2).reshape(B, -1)
    ex2 = np.stack([x1_, x1_, x0, x0], axis=2).reshape(B, -1)
    ey2 = np.stack([y0, y1_, y1_, y0], axis=2).reshape(B, -1)
    r = [a.astype(np.float32) for a in (ex1, ey1, ex2, ey2)]
    segs = [np.concatenate([np.tile(bx[None, :, k], (B, 1)), r[k]], axis=1) for k in range(4)]
    box_col = PALETTE[rng.integers(0, 8, (B, n_boxes))]
    colors = np.concatenate(
        [np.tile(np.array([[(128, 0, 128, 255)] * 4], dtype=np.uint8), (B, 1, 1)), np.repeat(box_col, 4, axis=1)], axis=1)
    robot = RobotSpec(
        devices=[
            RangeSpec(position=(7.0, -3.0), a=30.0, max=100.0, width=0.0, name="left-ir"),
            RangeSpec(position=(7.0, 3.0), a=-30.0, max=100.0, width=0.0, name="right-ir"),
            CameraSpec(width=cam_w, height=cam_h, a=60.0),
        ]
    )
    return _assemble("config2-boxes", w, h, segs, colors, [robot], rng, B, tries=2000)


def config3(n_worlds=65536, seed=3000, first_world=0):
    """4 robots each with a 16-ray range fan + LightSensor, 2 bulbs, 12 wall segments."""
    rng = _rng(seed, first_world)
    w, h, B = 300.0, 300.0, n_worlds
    bx = np.array(_boundary(w, h), dtype=np.float32)
    r = _rand_segments(rng, B, 8, w, h, min_len=15.0, max_len=70.0)
    segs = [np.concatenate([np.tile(bx[None, :, k], (B, 1)), r[k]], axis=1) for k in range(4)]
    colors = np.concatenate(
        [np.tile(np.array([[(128, 0, 128, 255)] * 4], dtype=np.uint8), (B, 1, 1)), PALETTE[rng.integers(0, 8, (B, 8))]],
        axis=1,
    )
    robots = []
    for k in range(4):
        devs = []
        for i in range(16):
            ang = -90.0 + 180.0 * i / 15.0
            px, py = 7.5 * math.cos(ang * PI_OVER_180), 6.0 * math.sin(ang * PI_OVER_180)
            devs.append(RangeSpec(position=(float(np.float32(px)), float(np.float32(py))), a=-ang, max=60.0, width=0.0, name="ir%d" % i))
        devs.append(LightSpec(position=(6.0, 0.0)))
        robots.append(RobotSpec(rgba=tuple(int(c) for c in PALETTE[k]), devices=devs))
    bulbs = np.zeros((B, 2, 4), dtype=np.float32)
    bulbs[:, :, 0] = rng.uniform(10, w - 10, (B, 2))
    bulbs[:, :, 1] = rng.uniform(10, h - 10, (B, 2))
    bulbs[:, :, 3] = 1.0
    return _assemble("config3", w, h, segs, colors, robots, rng, B, bulbs=bulbs)


def grid_maze_segments(rng, B, cells=11, cell=24.0, n_interior=96):
    """Axis-aligned maze pieces: n_interior of the interior cell edges, chosen per world."""
    edges = []
    for i in range(1, cells):  # vertical edges at x = i*cell
        for j in range(cells):
            edges.append((i * cell, j * cell, i * cell, (j + 1) * cell))
    for j in range(1, cells):  # horizontal edges at y = j*cell
        for i in range(cells):
            edges.append((i * cell, j * cell, (i + 1) * cell, j * cell))
    edges = np.array(edges, dtype=np.float32)  # [E, 4]
    keys = rng.random((B, edges.shape[0]))
    pick = np.argsort(keys, axis=1)[:, :n_interior]
    return edges[pick]  # [B, n_interior, 4]


def config4(n_worlds=1 << 20, seed=4000, first_world=0, cam_w=64, cam_h=32, variant="maze"):
    """1 robot with a 64x32 Camera in a 100-segment maze."""
    rng = _rng(seed, first_world)
    cells, cell = 11, 24.0
    w = h = cells * cell
    B = n_worlds
    bx = np.array(_boundary(w, h), dtype=np.float32)
    if variant == "maze":
        m = grid_maze_segments(rng, B, cells, cell, 96)
        r = [m[:, :, k] for k in range(4)]
    else:
        r = _rand_segments(rng, B, 96, w, h, min_len=10.0, max_len=40.0)
    segs = [np.concatenate([np.tile(bx[None, :, k], (B, 1)), r[k]], axis=1) for k in range(4)]
    colors = np.concatenate(
        [np.tile(np.array([[(128, 0, 128, 255)] * 4], dtype=np.uint8), (B, 1, 1)), PALETTE[rng.integers(0, 8, (B, 96))]],
        axis=1,
    )
    robot = RobotSpec(devices=[CameraSpec(width=cam_w, height=cam_h, a=60.0)])
    pose = None
    if variant == "maze":
        # cell centres are always free: the robot (radius ~10) fits a 24-unit cell
        ci = rng.integers(0, cells, (B, 1))
        cj = rng.integers(0, cells, (B, 1))
        pose = np.stack(
            [((ci + 0.5) * cell).astype(np.float64), ((cj + 0.5) * cell).astype(np.float64), rng.uniform(0, 2 * math.pi, (B, 1))],
            axis=2,
        )
    return _assemble("config4-" + variant, w, h, segs, colors, [robot], rng, B, pose=pose)


def config5(n_worlds=262144, n_segments=100, n_rays=16, seed=5000, first_world=0):
    """Worst-case sweep cell: n_segments random walls, n_rays single-ray range sensors."""
    rng = _rng(seed + 7919 * n_segments + n_rays, first_world)
    w = h = 400.0
    B = n_worlds
    bx = np.array(_boundary(w, h), dtype=np.float32)
    n_in = max(n_segments - 4, 0)
    r = _rand_segments(rng, B, n_in, w, h, min_len=5.0, max_len=30.0)
    segs = [np.concatenate([np.tile(bx[None, :, k], (B, 1)), r[k]], axis=1) for k in range(4)]
    colors = np.concatenate(
        [np.tile(np.array([[(128, 0, 128, 255)] * 4], dtype=np.uint8), (B, 1, 1)), PALETTE[rng.integers(0, 8, (B, n_in))]],
        axis=1,
    )
    devs = [
        RangeSpec(position=(0.0, 0.0), a=360.0 * i / n_rays, max=200.0, width=0.0, name="ray%d" % i) for i in range(n_rays)
    ]
    robot = RobotSpec(bbox=(-2.0, 2.0, -2.0, 2.0), devices=devs)
    # dense cells (2000 walls) leave little free space: small clearance, many tries
    return _assemble("config5-S%d-R%d" % (n_segments, n_rays), w, h, segs, colors, [robot], rng, B, tries=4000,
                     clear_extra=0.25)


def demo_scene(n_worlds=8, seed=7, first_world=0):
    """Small everything-on scene for tests: 3 robots, every device kind, bulbs, food."""
    rng = _rng(seed, first_world)
    w, h, B = 240.0, 180.0, n_worlds
    bx = np.array(_boundary(w, h), dtype=np.float32)
    r = _rand_segments(rng, B, 9, w, h, min_len=10.0, max_len=60.0)
    segs = [np.concatenate([np.tile(bx[None, :, k], (B, 1)), r[k]], axis=1) for k in range(4)]
    colors = np.concatenate(
        [np.tile(np.array([[(128, 0, 128, 255)] * 4], dtype=np.uint8), (B, 1, 1)), PALETTE[rng.integers(0, 8, (B, 9))]],
        axis=1,
    )
    robots = [
        RobotSpec(
            rgba=(255, 0, 0, 255),
            devices=[
                RangeSpec(position=(8.0, 0.0), a=0.0, max=50.0, width=10.0),
                RangeSpec(position=(6.0, -4.0), a=45.0, max=40.0, width=0.0),
                CameraSpec(width=32, height=16, a=60.0),
                LightSpec(position=(6.0, 0.0)),
                SmellSpec(position=(6.0, 0.0)),
            ],
        ),
        RobotSpec(
            rgba=(0, 0, 255, 255),
            height=0.5,
            devices=[RangeSpec(position=(8.0, 0.0), a=0.0, max=80.0, width=0.0), CameraSpec(width=32, height=16, a=90.0)],
        ),
        RobotSpec(rgba=(255, 255, 0, 255), devices=[LightSpec(position=(0.0, 5.0)), SmellSpec(position=(-5.0, 0.0))]),
    ]
    bulbs = np.zeros((B, 2, 4), dtype=np.float32)
    bulbs[:, :, 0] = rng.uniform(10, w - 10, (B, 2))
    bulbs[:, :, 1] = rng.uniform(10, h - 10, (B, 2))
    bulbs[:, :, 3] = rng.uniform(0.5, 2.0, (B, 2))
    food = np.zeros((B, 2, 3), dtype=np.float32)
    food[:, :, 0] = rng.uniform(10, w - 10, (B, 2))
    food[:, :, 1] = rng.uniform(10, h - 10, (B, 2))
    food[:, :, 2] = rng.uniform(10, 40, (B, 2))
    return _assemble("demo", w, h, segs, colors, robots, rng, B, bulbs=bulbs, food=food, smell_cell=12.0)
