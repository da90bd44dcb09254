# -*- coding: utf-8 -*-
"""Drop-in facade: World / Robot / add_device / get_distance / take_picture.

Mirrors the public names of upstream aitk/robots (world.py, robot.py,
devices/*.py; BASELINE.json north_star lists them) for ONE world; every number
it returns was computed by the CUDA engine on a batch of one (or, through
``World.batch``, on a batch of identical copies).  No arithmetic of the hot
path lives here, and nothing falls back to the CPU: stepping without a CUDA
device raises.
"""

import math

import numpy as np

from . import _cabi as cabi
from .scenes import CameraSpec, LightSpec, RangeSpec, RobotSpec, SceneBatch, SmellSpec, smell_grid

PI_OVER_180 = math.pi / 180.0


class FacadeDefaults:
    """Constants of the public API that never reach the kernel and are recollected, NOT
    verified against upstream aitk.  They are named here --
    and mirrored by ``oracle.aitk_oracle.Switches`` under the same names -- so that the day
    the real package is seen they are flipped in one place."""

    # Robot(max_rotate_velocity=...): radians per second at rotate == 1            [unverified]
    ROBOT_MAX_ROTATE_VELOCITY = math.pi / 4
    # Robot(max_trans_velocity=...): world units per step at translate == 1       [unverified]
    ROBOT_MAX_TRANS_VELOCITY = 2.0
    # Robot.move(): upstream may round the target velocities (e.g. to 1 decimal); None = no rounding  [unverified]
    MOVE_ROUND_DIGITS = None
    # Camera(a=...): default field of view in degrees                               [unverified]
    CAMERA_DEFAULT_FOV = 60
    # Camera(width, height) defaults                                                [unverified]
    CAMERA_DEFAULT_SIZE = (64, 32)


def _round_target(v):
    d = FacadeDefaults.MOVE_ROUND_DIGITS
    return v if d is None else round(v, d)


COLORS = {
    "black": (0, 0, 0), "white": (255, 255, 255), "red": (255, 0, 0), "green": (0, 128, 0),
    "lime": (0, 255, 0), "blue": (0, 0, 255), "yellow": (255, 255, 0), "cyan": (0, 255, 255),
    "magenta": (255, 0, 255), "orange": (255, 165, 0), "purple": (128, 0, 128), "pink": (255, 192, 203),
    "grey": (128, 128, 128), "gray": (128, 128, 128), "brown": (165, 42, 42),
}


def to_rgba(color):
    if isinstance(color, str):
        r, g, b = COLORS[color.lower()]
        return (r, g, b, 255)
    c = tuple(int(v) for v in color)
    return c if len(c) == 4 else c + (255,)


class _Device:
    _spec_cls = None

    def __init__(self):
        self.robot = None
        self._index = None  # index among devices of the same kind in the world

    def _world(self):
        if self.robot is None or self.robot.world is None:
            raise RuntimeError("device is not attached to a robot in a world")
        return self.robot.world


class RangeSensor(_Device):
    def __init__(self, position=(8, 0), a=0, max=20, width=1.0, name="sensor"):
        super().__init__()
        self.spec = RangeSpec(position=(float(position[0]), float(position[1])), a=float(a), max=float(max),
                              width=float(width), name=name)
        self.name = name

    def get_max(self):
        return self.spec.max

    def get_distance(self):
        w = self._world()
        w._ensure_sensed()
        return float(w._obs["range"][0, self._index])

    def get_reading(self):
        return self.get_distance() / self.spec.max


class Camera(_Device):
    def __init__(self, width=None, height=None, a=None, colorsFadeWithDistance=0.9, sizeFadeWithDistance=1.0,
                 max_range=1000, position=(0, 0), name="camera"):
        super().__init__()
        width = FacadeDefaults.CAMERA_DEFAULT_SIZE[0] if width is None else width
        height = FacadeDefaults.CAMERA_DEFAULT_SIZE[1] if height is None else height
        a = FacadeDefaults.CAMERA_DEFAULT_FOV if a is None else a
        self.spec = CameraSpec(width=int(width), height=int(height), a=float(a),
                               colorsFadeWithDistance=float(colorsFadeWithDistance),
                               sizeFadeWithDistance=float(sizeFadeWithDistance), max_range=float(max_range),
                               position=(float(position[0]), float(position[1])), name=name)
        self.name = name

    def take_picture(self, as_array=False):
        """RGBA picture of what the camera sees now (lazy, like upstream)."""
        w = self._world()
        img = w._render()[0, self._index]
        if as_array:
            return img.copy()
        from PIL import Image

        return Image.fromarray(img, "RGBA")

    get_image = take_picture


class GroundCamera(_Device):
    """Downward-looking camera (upstream devices/cameras.py GroundCamera [RECALL]): a
    ``width`` x ``height`` pixel picture of the ``view`` (world units) window centred on the
    robot, row 0 towards its front; the robot itself is not in it.  Painted by the engine's
    top-down renderer (brs_render_world) from the current state -- no per-step cost."""

    def __init__(self, width=15, height=15, view=None, name="ground-camera"):
        super().__init__()
        self.width, self.height = int(width), int(height)
        self.view = (float(width), float(height)) if view is None else (float(view[0]), float(view[1]))
        self.name = name
        self.spec = None  # nothing for the step kernels to do

    def take_picture(self, as_array=False):
        w = self._world()
        w._ensure_engine()
        img = w._engine.render_world(0, 1, self.width, self.height, robot=self.robot._slot, view=self.view)[0].cpu().numpy()
        if as_array:
            return img
        from PIL import Image

        return Image.fromarray(img, "RGBA")

    get_image = take_picture


class LightSensor(_Device):
    def __init__(self, position=(0, 0), name="light"):
        super().__init__()
        self.spec = LightSpec(position=(float(position[0]), float(position[1])), name=name)
        self.name = name

    def get_brightness(self):
        w = self._world()
        w._ensure_sensed()
        return float(w._obs["light"][0, self._index])

    get_reading = get_brightness


class SmellSensor(_Device):
    def __init__(self, position=(0, 0), name="smell"):
        super().__init__()
        self.spec = SmellSpec(position=(float(position[0]), float(position[1])), name=name)
        self.name = name

    def get_reading(self):
        w = self._world()
        w._ensure_sensed()
        return float(w._obs["smell"][0, self._index])


class Bulb:
    def __init__(self, color, x, y, z=0.0, brightness=1.0, name="bulb"):
        self.color, self.x, self.y, self.z, self.brightness, self.name = to_rgba(color), x, y, z, brightness, name


class Robot:
    """Robot(x, y, a) with ``a`` in degrees, as upstream."""

    def __init__(self, x=0, y=0, a=0, color="red", name="Robbie", height=0.25, max_trans_velocity=None,
                 max_rotate_velocity=None, bbox=None):
        if max_trans_velocity is None:
            max_trans_velocity = FacadeDefaults.ROBOT_MAX_TRANS_VELOCITY
        if max_rotate_velocity is None:
            max_rotate_velocity = FacadeDefaults.ROBOT_MAX_ROTATE_VELOCITY
        self.world = None
        self.name = name
        self._slot = None
        self._pose = [float(x), float(y), float(a) * PI_OVER_180]
        self._pose0 = list(self._pose)  # what World.reset() returns to
        self._tvel = [0.0, 0.0, 0.0]
        self.color = to_rgba(color)
        self.height = float(height)
        self.max_trans_velocity = float(max_trans_velocity)
        self.max_rotate_velocity = float(max_rotate_velocity)
        self.bbox = tuple(float(v) for v in (bbox if bbox is not None else RobotSpec().bbox))
        self._devices = []
        self._pending = True  # constructor / set_pose values not yet uploaded

    # state (read back from the engine after every step)
    def _state(self, key, k):
        w = self.world
        if w is not None and w._engine is not None and not self._pending:
            return float(w._state[key][0, self._slot, k])
        return self._pose[k]

    @property
    def x(self):
        return self._state("pose", 0)

    @property
    def y(self):
        return self._state("pose", 1)

    @property
    def a(self):
        return self._state("pose", 2)

    @property
    def stalled(self):
        w = self.world
        if w is not None and w._engine is not None:
            return bool(w._state["stalled"][0, self._slot])
        return False

    def get_pose(self):
        return (self.x, self.y, self.a)

    def set_pose(self, x=None, y=None, a=None):
        """a in degrees."""
        cur = list(self.get_pose())
        if x is not None:
            cur[0] = float(x)
        if y is not None:
            cur[1] = float(y)
        if a is not None:
            cur[2] = float(a) * PI_OVER_180
        self._pose = cur
        self._pending = True
        if self.world is None or self.world._engine is None:
            self._pose0 = list(cur)  # placed before the first step: that is the initial pose
        if self.world is not None:
            # a pose change is an upload of one row into the live engine, not a rebuild:
            # velocities, stall flags and every other robot keep their state
            self.world._pose_dirty = True
            self.world._sensed = False

    def add_device(self, device):
        device.robot = self
        self._devices.append(device)
        if self.world is not None:
            self.world._dirty = True
        return device

    def __getitem__(self, item):
        if isinstance(item, int):
            return self._devices[item]
        for d in self._devices:
            if d.name == item:
                return d
        raise KeyError(item)

    def move(self, translate, rotate):
        self._tvel[0] = _round_target(translate * self.max_trans_velocity)
        self._tvel[2] = _round_target(rotate * self.max_rotate_velocity)

    def forward(self, translate):
        self._tvel[0] = _round_target(translate * self.max_trans_velocity)

    def backward(self, translate):
        self._tvel[0] = _round_target(-translate * self.max_trans_velocity)

    def turn(self, rotate):
        self._tvel[2] = _round_target(rotate * self.max_rotate_velocity)

    def stop(self):
        self._tvel = [0.0, 0.0, 0.0]

    def _spec(self):
        return RobotSpec(bbox=self.bbox, height=self.height, rgba=self.color,
                         max_trans_velocity=self.max_trans_velocity, max_rotate_velocity=self.max_rotate_velocity,
                         devices=[d.spec for d in self._devices if d.spec is not None])


class World:
    def __init__(self, width=500, height=250, seed=0, boundary_wall=True, boundary_wall_color="purple",
                 ground_color="green", smell_cell_size=None, time_step=0.1, device=0):
        self.width, self.height = float(width), float(height)
        self.seed = seed
        self.time = 0.0
        self.time_step = time_step
        self.ground_color = to_rgba(ground_color)
        self.smell_cell_size = smell_cell_size if smell_cell_size else max(width, height) / 20.0
        self.device = device
        self._segments = []  # (x1, y1, x2, y2, rgba)
        self._bulbs = []
        self._food = []
        self._robots = []
        self._engine = None
        self._dirty = True
        self._pose_dirty = False
        self._sensed = False
        self._state = {}
        self._obs = {}
        self._noise = None  # (range_std, light_rel_std): seeded sensor noise, see set_sensor_noise
        self._boundary_wall = bool(boundary_wall)
        self._boundary_color = to_rgba(boundary_wall_color)
        if boundary_wall:
            c = to_rgba(boundary_wall_color)
            w, h = self.width, self.height
            for s in ((0, 0, w, 0), (0, 0, 0, h), (0, h, w, h), (w, 0, w, h)):
                self._segments.append(s + (c,))

    # -- construction ------------------------------------------------------------
    def add_wall(self, color, x1, y1, x2, y2, box=True):
        c = to_rgba(color)
        if box:
            pts = [(x1, y1), (x2, y1), (x2, y2), (x1, y2)]
            for k in range(4):
                a, b = pts[k], pts[(k + 1) % 4]
                self._segments.append((a[0], a[1], b[0], b[1], c))
        else:
            self._segments.append((x1, y1, x2, y2, c))
        self._dirty = True

    def add_bulb(self, color, x, y, z=0.0, brightness=1.0, name="bulb"):
        b = Bulb(color, x, y, z, brightness, name)
        self._bulbs.append(b)
        self._dirty = True
        return b

    def add_food(self, x, y, standard_deviation):
        self._food.append((x, y, standard_deviation))
        self._dirty = True

    def add_robot(self, robot):
        if len(self._robots) >= cabi.BRS_MAX_ROBOTS:
            raise ValueError("at most %d robots per world" % cabi.BRS_MAX_ROBOTS)
        robot.world = self
        robot._slot = len(self._robots)
        self._robots.append(robot)
        self._dirty = True
        return robot

    # -- batch form --------------------------------------------------------------
    def to_scene(self, copies=1):
        """This world as a SceneBatch of `copies` identical worlds."""
        if not self._robots:
            raise RuntimeError("world has no robots")
        n = len(self._segments)
        cap = max((n + 3) // 4 * 4, 4)
        seg = np.zeros((1, 4, cap), dtype=np.float32)
        rgba = np.zeros((1, cap, 4), dtype=np.uint8)
        for j, (x1, y1, x2, y2, c) in enumerate(self._segments):
            seg[0, :, j] = (x1, y1, x2, y2)
            rgba[0, j] = c
        pose = np.array([[r.get_pose() for r in self._robots]], dtype=np.float64)
        tvel = np.array([[r._tvel for r in self._robots]], dtype=np.float64)
        bulbs = None
        if self._bulbs:
            bulbs = np.array([[(b.x, b.y, b.z, b.brightness) for b in self._bulbs]], dtype=np.float32)
        ws = np.array([[self.width, self.height]], dtype=np.float32)
        food, grid = None, None
        if self._food:
            food = np.array([self._food], dtype=np.float32)
            grid = smell_grid(ws, food, self.smell_cell_size)
        rep = lambda a: None if a is None else np.repeat(a, copies, axis=0)  # noqa: E731
        return SceneBatch(world_size=rep(ws), seg=rep(seg), seg_rgba=rep(rgba),
                          seg_count=np.full(copies, n, dtype=np.int32), pose=rep(pose), target_vel=rep(tvel),
                          robots=[r._spec() for r in self._robots], bulbs=rep(bulbs), grid=rep(grid),
                          smell_cell_size=self.smell_cell_size if self._food else 0.0, food=rep(food),
                          time_step=self.time_step, name="world")

    def batch(self, copies, device=None):
        """BatchEngine stepping `copies` replicas of this world on the GPU."""
        from .engine import BatchEngine

        return BatchEngine(self.to_scene(copies), device=self.device if device is None else device)

    # -- robots as a sequence (upstream: ``world[0]``, ``for robot in world``) ------------
    def __iter__(self):
        return iter(self._robots)

    def __len__(self):
        return len(self._robots)

    def __getitem__(self, item):
        if isinstance(item, str):
            for r in self._robots:
                if r.name == item:
                    return r
            raise KeyError(item)
        return self._robots[item]

    def set_sensor_noise(self, range_std=0.0, light_rel_std=0.0):
        """Seeded sensor noise, keyed by the world's ``seed`` (upstream: ``World(seed=...)`` makes
        noisy experiments reproducible).  Gaussian: range readings += range_std * N(0,1) clamped
        to [0, max]; light values *= 1 + light_rel_std * N(0,1).  (0, 0) switches it off."""
        self._noise = (float(range_std), float(light_rel_std)) if (range_std or light_rel_std) else None
        if self._engine is not None:
            self._engine.set_noise(int(self.seed), *(self._noise or (0.0, 0.0)))

    # -- engine plumbing -----------------------------------------------------------
    def _build(self):
        """(Re)create the engine after a structural change (walls, bulbs, devices, robots).
        The dynamic state of the robots that already existed -- actual velocities (the
        ramp) and stall flags -- is carried over; only constructor / set_pose poses are new."""
        from .engine import BatchEngine

        carry = None
        if self._engine is not None:
            carry = (self._engine.download("vel"), self._engine.download("stalled"))
            self._engine.close()
        idx = {RangeSensor: 0, Camera: 0, LightSensor: 0, SmellSensor: 0, GroundCamera: 0}
        for r in self._robots:
            for d in r._devices:
                d._index = idx[type(d)]
                idx[type(d)] += 1
        self._engine = BatchEngine(self.to_scene(1), device=self.device)
        if self._noise is not None:
            self._engine.set_noise(int(self.seed), *self._noise)
        if carry is not None:
            n = min(carry[0].shape[1], len(self._robots))
            vel = np.zeros((1, len(self._robots), 3), dtype=np.float64)
            stalled = np.zeros((1, len(self._robots)), dtype=np.uint8)
            vel[:, :n], stalled[:, :n] = carry[0][:, :n], carry[1][:, :n]
            self._engine.upload("vel", vel)
            self._engine.upload("stalled", stalled)
        for r in self._robots:
            r._pending = False
        self._dirty = False
        self._pose_dirty = False
        self._sensed = False
        self._pull_state()

    def _push_poses(self):
        """Robot.set_pose on a live engine: upload the pose row, keep everything else."""
        pose = np.array([[r.get_pose() for r in self._robots]], dtype=np.float64)
        self._engine.upload("pose", pose)
        for r in self._robots:
            r._pending = False
        self._pose_dirty = False
        self._sensed = False
        self._state["pose"] = pose

    def _pull_state(self):
        e = self._engine
        self._state = {"pose": e.download("pose"), "stalled": e.download("stalled")}

    def _pull_obs(self):
        e = self._engine
        self._obs = {k: e.download(k) for k in ("range", "light", "smell")}
        self._sensed = True

    def _ensure_engine(self):
        if self._dirty or self._engine is None:
            self._build()
        elif self._pose_dirty:
            self._push_poses()

    def _ensure_sensed(self):
        self._ensure_engine()
        if not self._sensed:
            self.update()

    def _render(self):
        self._ensure_engine()
        self._engine.step(self.time_step, cabi.STEP_CAMERA)
        return self._engine.download("camera")

    def take_picture(self, width=None, height=None, as_array=False):
        """The world from above as it is now (upstream World.take_picture / display [RECALL]):
        ground, bulbs, walls, robots -- painted by the engine (brs_render_world).  Default
        size: one pixel per world unit, at most 1024 wide."""
        self._ensure_engine()
        if width is None:
            width = int(min(1024, max(1, round(self.width))))
        img = self._engine.render_world(0, 1, int(width), None if height is None else int(height))[0].cpu().numpy()
        if as_array:
            return img
        from PIL import Image

        return Image.fromarray(img, "RGBA")

    get_image = take_picture

    # -- stepping --------------------------------------------------------------------
    def step(self, time_step=None):
        """One tick: move every robot, refresh every device, advance the clock."""
        self._ensure_engine()
        dt = self.time_step if time_step is None else time_step
        self._engine.upload("target_vel", np.array([[r._tvel for r in self._robots]], dtype=np.float64))
        self._engine.step(dt, cabi.STEP_MOVE | cabi.STEP_SENSE)
        self.time = round(self.time + dt, 10)
        self._pull_state()
        self._pull_obs()

    def update(self):
        """Refresh the devices without moving (upstream World.update)."""
        self._ensure_engine()
        self._engine.step(self.time_step, cabi.STEP_SENSE)
        self._pull_obs()

    def steps(self, n, time_step=None):
        """n ticks with the robots' current move() commands (upstream World.steps): one engine
        call (brs_steps), state and observations read back once at the end."""
        n = int(n)
        if n <= 0:
            return
        self._ensure_engine()
        dt = self.time_step if time_step is None else time_step
        self._engine.upload("target_vel", np.array([[r._tvel for r in self._robots]], dtype=np.float64))
        self._engine.steps(n, dt, cabi.STEP_MOVE | cabi.STEP_SENSE)
        for _ in range(n):
            self.time = round(self.time + dt, 10)
        self._pull_state()
        self._pull_obs()

    # -- scene files ---------------------------------
    def to_json(self):
        """The scene as a plain dict (json.dump-able).  Key names follow the upstream
        ``World.to_json`` layout as recollected ([RECALL]: width / height /
        boundary_wall / walls[{color, p1, p2}] / bulbs / food / robots[{x, y, a, devices[
        {class, ...}]}]); it could not be checked against a real aitk scene file here."""
        nb = 4 if self._boundary_wall else 0
        walls = [{"color": list(c), "p1": {"x": x1, "y": y1}, "p2": {"x": x2, "y": y2}, "wtype": "line"}
                 for (x1, y1, x2, y2, c) in self._segments[nb:]]
        robots = []
        for r in self._robots:
            x, y, a = r.get_pose()
            devs = []
            for d in r._devices:
                if isinstance(d, GroundCamera):
                    devs.append({"class": "GroundCamera", "width": d.width, "height": d.height,
                                 "view": list(d.view), "name": d.name})
                    continue
                spec = dict(vars(d.spec))
                spec["position"] = list(spec["position"])
                spec["class"] = type(d).__name__
                devs.append(spec)
            robots.append({"class": "Robot", "name": r.name, "x": x, "y": y, "a": a / PI_OVER_180,
                           "color": list(r.color), "height": r.height, "bbox": list(r.bbox),
                           "max_trans_velocity": r.max_trans_velocity,
                           "max_rotate_velocity": r.max_rotate_velocity, "devices": devs})
        return {"width": self.width, "height": self.height, "seed": self.seed,
                "boundary_wall": self._boundary_wall, "boundary_wall_color": list(self._boundary_color),
                "ground_color": list(self.ground_color), "smell_cell_size": self.smell_cell_size,
                "time_step": self.time_step, "walls": walls,
                "bulbs": [{"color": list(b.color), "x": b.x, "y": b.y, "z": b.z, "brightness": b.brightness,
                           "name": b.name} for b in self._bulbs],
                "food": [list(f) for f in self._food], "robots": robots}

    @classmethod
    def from_json(cls, config, device=0):
        """Rebuild a World from ``to_json`` output (or a path to a JSON file of it)."""
        if isinstance(config, str):
            import json

            with open(config) as f:
                config = json.load(f)
        world = cls(config["width"], config["height"], seed=config.get("seed", 0),
                    boundary_wall=config.get("boundary_wall", True),
                    boundary_wall_color=config.get("boundary_wall_color", "purple"),
                    ground_color=config.get("ground_color", "green"),
                    smell_cell_size=config.get("smell_cell_size"), time_step=config.get("time_step", 0.1),
                    device=device)
        for w in config.get("walls", []):
            world.add_wall(w["color"], w["p1"]["x"], w["p1"]["y"], w["p2"]["x"], w["p2"]["y"],
                           box=w.get("wtype", "line") == "box")
        for b in config.get("bulbs", []):
            world.add_bulb(b["color"], b["x"], b["y"], b.get("z", 0.0), b.get("brightness", 1.0), b.get("name", "bulb"))
        for f in config.get("food", []):
            world.add_food(*f)
        kinds = {"RangeSensor": RangeSensor, "Camera": Camera, "LightSensor": LightSensor, "SmellSensor": SmellSensor,
                 "GroundCamera": GroundCamera}
        for rc in config.get("robots", []):
            robot = Robot(rc["x"], rc["y"], rc["a"], color=rc.get("color", "red"), name=rc.get("name", "Robbie"),
                          height=rc.get("height", 0.25), max_trans_velocity=rc.get("max_trans_velocity", 2.0),
                          max_rotate_velocity=rc.get("max_rotate_velocity", math.pi / 4), bbox=rc.get("bbox"))
            for dc in rc.get("devices", []):
                dc = dict(dc)
                robot.add_device(kinds[dc.pop("class")](**dc))
            world.add_robot(robot)
        return world

    def seconds(self, seconds, time_step=None):
        dt = self.time_step if time_step is None else time_step
        self.steps(round(seconds / dt), time_step)

    def run(self, function=None, time_step=None, max_steps=None):
        """Step until ``function(world)`` returns a true value (upstream World.run: the
        controller loop [RECALL]; the real-time / display options of upstream have no meaning
        here).  The controller is called before every step; without one, runs ``max_steps``."""
        n = 0
        while max_steps is None or n < max_steps:
            if function is not None and function(self):
                break
            if function is None and max_steps is None:
                raise ValueError("World.run needs a controller function or max_steps")
            self.step(time_step)
            n += 1
        return n

    def reset(self):
        """Back to the state the world was built with: robots at their constructor poses (or
        the last set_pose before the first step), zero velocities, clock at 0 (upstream
        World.reset [RECALL])."""
        for r in self._robots:
            r._pose = list(r._pose0)
            r._tvel = [0.0, 0.0, 0.0]
            r._pending = True
        self.time = 0.0
        self._sensed = False
        if self._engine is not None:
            self._engine.upload("pose", np.array([[r._pose for r in self._robots]], dtype=np.float64))
            zeros = np.zeros((1, len(self._robots), 3), dtype=np.float64)
            self._engine.upload("vel", zeros)
            self._engine.upload("target_vel", zeros)
            self._engine.upload("stalled", np.zeros((1, len(self._robots)), dtype=np.uint8))
            for r in self._robots:
                r._pending = False
            self._pose_dirty = False
            self._pull_state()

    def get_robot(self, item):
        """Robot by index or name (upstream: ``world[item]``)."""
        return self[item]
