# -*- coding: utf-8 -*-
"""
CPU restatement oracle for the aitk.robots ``World.step`` hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``aitk.robots_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or as the timed CPU baseline.

PARITY UNPINNED.  The mounted reference (/root/reference) is a two-file
packaging stub (setup.py:29 ``install_requires=["aitk"]``, no code).  The engine
lives in the un-vendored, un-pinned third-party distribution ``aitk``
(subpackage ``aitk.robots``) which is not importable here and has no wheel on
disk.  The reference ships no tests, fixtures or golden vectors for this path
(SURVEY.md section 4).  This file therefore restates the *published behaviour* of
``aitk.robots`` (world.py / robot.py / utils.py / hit.py /
devices/{rangesensors,cameras,lightsensors,smellsensors}.py -- names from
BASELINE.json north_star, internals from recollection)
in plain float64 Python.  Every recollection-dependent choice is a named
switch in :class:`Switches` so that it can be flipped the moment the real
package is seen (tests/test_real_aitk.py runs the same scenes through
``aitk.robots`` whenever ``import aitk.robots`` succeeds).

The arithmetic is deliberately written the way the reference computes it:
one Python object per wall / line / hit, one float64 operation at a time, a
list of ``Hit`` objects per ray sorted by distance.  That makes this file both
the parity checker and a faithful stand-in for the reference's pure-Python CPU
cost (``cpu_baseline.kind == "port"``).
"""

import math

PI_OVER_2 = math.pi / 2.0
PI_OVER_180 = math.pi / 180.0


class Switches:
    """Named, flippable restatement choices.

    The CUDA engine implements exactly the default setting of every switch;
    tests assert that the defaults are in force before comparing.
    """

    # U2: ``tvx = vx*sin(-pa+o) + vy*cos(-pa+o) * dt`` -- operator precedence
    # applies ``dt`` to the sideways (vy) term only; forward speed is per step.
    MOTION_DT_ON_VY_ONLY = True
    # U3: velocities ramp towards their targets, at most ``ramp * dt`` per step.
    VELOCITY_RAMP = True
    # U4: a stalled robot keeps its pose and has its actual velocities zeroed.
    STALL_ZEROES_VELOCITY = True
    # U7: World.step moves every robot (in list order, each seeing the already
    # moved boxes of the robots before it) and only then refreshes devices.
    MOVE_ALL_THEN_SENSE = True
    # U8: a RangeSensor with width != 0 casts rays at -w/2, 0, +w/2.
    RANGE_FAN_INCLUSIVE = True
    # U9: camera constants.
    CAMERA_SKY = (0, 0, 128)
    CAMERA_GROUND = (0, 128, 0)
    # optional vertical gradient on sky / ground: colour = base + grad * dist,
    # dist = |j - H/2| / (H/2); zero = flat colours.
    CAMERA_SKY_GRAD = (0, 0, 0)
    CAMERA_GROUND_GRAD = (0, 0, 0)
    # U10: light = sum over visible bulbs of brightness * multiplier / dist**2.
    LIGHT_MULTIPLIER = 1000.0
    # U12: smell grid samples sum of gaussians at cell centres; walls ignored.
    SMELL_WALLS_BLOCK = False
    # U3 / U9 / U11 -- facade constants that never reach the kernel (mirrored, under the same
    # names, by aitk.robots_b200.world.FacadeDefaults; all UNVERIFIED recollection):
    ROBOT_MAX_ROTATE_VELOCITY = math.pi / 4  # rad/s at rotate == 1
    ROBOT_MAX_TRANS_VELOCITY = 2.0           # units per step at translate == 1
    MOVE_ROUND_DIGITS = None                 # Robot.move rounds its targets to this many digits (None: no rounding)
    CAMERA_DEFAULT_FOV = 60                  # degrees
    CAMERA_DEFAULT_SIZE = (64, 32)
    # Seeded sensor noise.  None = off (BASELINE.json states parity with noise disabled).  Else a
    # dict(seed=int, range_std=float, light_rel_std=float): the ENGINE'S OWN model (upstream's is
    # not in the mounted tree), restated here so that the CUDA noise has a checker -- one
    # Philox4x32-10 block per (world, sensor, step), Box-Muller; see include/brs.h brs_set_noise.
    SENSOR_NOISE = None


SW = Switches


# ----------------------------------------------------------------------------
# Seeded noise: Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy
# as 1, 2, 3", SC'11; known-answer vectors of the Random123 distribution in tests/test_oracle.py)
# ----------------------------------------------------------------------------
NOISE_RANGE, NOISE_LIGHT = 1, 2


def philox4x32_10(counter, key):
    """counter: 4 x uint32, key: 2 x uint32 -> 4 x uint32."""
    c0, c1, c2, c3 = (int(v) & 0xFFFFFFFF for v in counter)
    k0, k1 = (int(v) & 0xFFFFFFFF for v in key)
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        hi0, lo0 = p0 >> 32, p0 & 0xFFFFFFFF
        hi1, lo1 = p1 >> 32, p1 & 0xFFFFFFFF
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return [c0, c1, c2, c3]


def noise_normal(seed, world, kind, index, step):
    """N(0,1) of one (world, sensor kind / index, step): the engine's counter layout."""
    x = philox4x32_10([world, (kind << 24) | index, step & 0xFFFFFFFF, step >> 32], [seed & 0xFFFFFFFF, seed >> 32])
    u1 = ((x[0] >> 8) + 0.5) / 16777216.0
    u2 = ((x[1] >> 8) + 0.5) / 16777216.0
    return math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)


# ----------------------------------------------------------------------------
# L0 geometry primitives (upstream aitk/robots/utils.py, hit.py)
# ----------------------------------------------------------------------------


def distance(x1, y1, x2, y2):
    """Euclidean distance (upstream utils.distance)."""
    return math.sqrt((x1 - x2) ** 2 + (y1 - y2) ** 2)


def rotate_around(x1, y1, length, angle):
    """Swing a line of ``length`` around (x1, y1) (upstream utils.rotate_around)."""
    return [x1 + length * math.cos(-angle), y1 - length * math.sin(-angle)]


def intersect_hit(x1, y1, x2, y2, x3, y3, x4, y4):
    """Segment (x1,y1)-(x2,y2) against segment (x3,y3)-(x4,y4).

    Returns the intersection point or None.  Parametric line-line form (the
    13-flop test of SURVEY.md 8(d)): denominator, two numerators, two divides,
    both parameters must lie in [0, 1] inclusive.
    """
    if (x1 == x2 and y1 == y2) or (x3 == x4 and y3 == y4):
        return None
    denominator = (y4 - y3) * (x2 - x1) - (x4 - x3) * (y2 - y1)
    if denominator == 0:
        return None
    ua = ((x4 - x3) * (y1 - y3) - (y4 - y3) * (x1 - x3)) / denominator
    ub = ((x2 - x1) * (y1 - y3) - (y2 - y1) * (x1 - x3)) / denominator
    if ua < 0 or ua > 1 or ub < 0 or ub > 1:
        return None
    return (x1 + ua * (x2 - x1), y1 + ua * (y2 - y1))


def intersect(x1, y1, x2, y2, x3, y3, x4, y4):
    """Boolean-use variant for collision checks (same arithmetic)."""
    return intersect_hit(x1, y1, x2, y2, x3, y3, x4, y4)


def arange(start, stop, step):
    """Inclusive float range used by the RangeSensor fan (U8)."""
    current = start
    while current <= stop:
        yield current
        current += step


_COLOR_TABLE = {
    "black": (0, 0, 0),
    "white": (255, 255, 255),
    "red": (255, 0, 0),
    "green": (0, 128, 0),
    "lime": (0, 255, 0),
    "blue": (0, 0, 255),
    "yellow": (255, 255, 0),
    "cyan": (0, 255, 255),
    "magenta": (255, 0, 255),
    "orange": (255, 165, 0),
    "purple": (128, 0, 128),
    "pink": (255, 192, 203),
    "grey": (128, 128, 128),
    "gray": (128, 128, 128),
    "brown": (165, 42, 42),
}


class Color:
    def __init__(self, red, green=None, blue=None, alpha=255):
        if isinstance(red, Color):
            red, green, blue, alpha = red.red, red.green, red.blue, red.alpha
        elif isinstance(red, str):
            red, green, blue = _COLOR_TABLE[red.lower()]
        elif green is None:
            red, green, blue = red[0], red[1], red[2]
        self.red, self.green, self.blue, self.alpha = red, green, blue, alpha

    def to_tuple(self):
        return (int(self.red), int(self.green), int(self.blue), int(self.alpha))


class Point:
    __slots__ = ("x", "y")

    def __init__(self, x, y):
        self.x = x
        self.y = y


class Line:
    __slots__ = ("p1", "p2")

    def __init__(self, p1, p2):
        self.p1 = p1
        self.p2 = p2


class Wall:
    """A colour plus 1 (plain / boundary) or 4 (box, robot) line segments."""

    def __init__(self, color, robot, *lines):
        self.color = color
        self.robot = robot
        self.lines = list(lines)


class Hit:
    """One ray-segment hit (upstream hit.py)."""

    __slots__ = ("robot", "height", "x", "y", "distance", "color", "start_x", "start_y", "boundary")

    def __init__(self, robot, height, x, y, distance, color, start_x, start_y, boundary):
        self.robot = robot
        self.height = height
        self.x = x
        self.y = y
        self.distance = distance
        self.color = color
        self.start_x = start_x
        self.start_y = start_y
        self.boundary = boundary


class Bulb:
    """A static light source in the world (upstream devices/bulbs.py)."""

    def __init__(self, color, x, y, z=0.0, brightness=1.0, name="bulb"):
        self.color = Color(color)
        self.x, self.y, self.z = x, y, z
        self.brightness = brightness
        self.name = name
        self.robot = None


# ----------------------------------------------------------------------------
# L1 devices
# ----------------------------------------------------------------------------


class _Device:
    def _place(self, position):
        # upstream: distance / direction of the mount point from the robot centre
        self.position = [position[0], position[1]]
        self.dist_from_center = distance(0, 0, position[0], position[1])
        self.dir_from_center = math.atan2(-position[0], position[1])

    def _mount_point(self):
        return rotate_around(
            self.robot.x,
            self.robot.y,
            self.dist_from_center,
            self.robot.a + self.dir_from_center + PI_OVER_2,
        )


class RangeSensor(_Device):
    """IR / sonar style range finder (upstream devices/rangesensors.py).

    position in robot frame (cm), ``a`` heading offset in DEGREES, ``max`` range
    in cm, ``width`` beam width in DEGREES (0 = single ray).
    """

    def __init__(self, position=(8, 0), a=0, max=20, width=1.0, name="sensor"):
        self.robot = None
        self.name = name
        self._place(position)
        self.a = a * PI_OVER_180
        self.max = max
        self.width = width * PI_OVER_180
        self.reading = 1.0
        self.distance = self.reading * self.max
        self.counted_tests = 0

    def ray_count(self):
        if self.width != 0:
            return sum(1 for _ in arange(-self.width / 2, self.width / 2, self.width / 2))
        return 1

    def set_reading(self, reading):
        self.reading = reading
        self.distance = reading * self.max

    def get_reading(self):
        return self.reading

    def get_distance(self):
        return self.distance

    def get_max(self):
        return self.max

    def update(self):
        self.set_reading(1.0)
        p = self._mount_point()
        if self.width != 0:
            incrs = list(arange(-self.width / 2, self.width / 2, self.width / 2))
        else:
            incrs = [0.0]
        for incr in incrs:
            if self.width != 0:
                angle = -self.robot.a - self.a + PI_OVER_2 + incr
            else:
                angle = -self.robot.a - self.a + PI_OVER_2
            hits = self.robot.cast_ray(p[0], p[1], angle, self.max)
            if hits:
                # closest hit is last (list is sorted furthest first)
                if hits[-1].distance < self.get_reading() * self.max:
                    self.set_reading(hits[-1].distance / self.max)
        if SW.SENSOR_NOISE and SW.SENSOR_NOISE.get("range_std"):
            world = self.robot.world
            z = noise_normal(SW.SENSOR_NOISE["seed"], world.noise_world_index, NOISE_RANGE, self.noise_index, world.noise_step)
            d = min(max(self.distance + SW.SENSOR_NOISE["range_std"] * z, 0.0), self.max)
            self.reading, self.distance = d / self.max, d


class Camera(_Device):
    """Forward facing pinhole-column camera (upstream devices/cameras.py).

    One ray per image column; ``take_picture`` paints sky / wall / ground per
    column and overdraws other robots back to front.
    """

    def __init__(
        self,
        width=None,
        height=None,
        a=None,
        colorsFadeWithDistance=0.9,
        sizeFadeWithDistance=1.0,
        max_range=1000,
        position=(0, 0),
        name="camera",
    ):
        self.robot = None
        self.name = name
        self._place(position)
        width = SW.CAMERA_DEFAULT_SIZE[0] if width is None else width
        height = SW.CAMERA_DEFAULT_SIZE[1] if height is None else height
        a = SW.CAMERA_DEFAULT_FOV if a is None else a
        self.cameraShape = [width, height]
        self.angle = a * PI_OVER_180  # field of view
        self.colorsFadeWithDistance = colorsFadeWithDistance
        self.sizeFadeWithDistance = sizeFadeWithDistance
        self.max_range = max_range
        self.hits = [[] for _ in range(width)]

    def update(self):
        # lazy upstream: the rays are cast when the picture is asked for
        pass

    def _update(self):
        p = self._mount_point()
        for i in range(self.cameraShape[0]):
            angle = i / self.cameraShape[0] * self.angle - self.angle / 2
            self.hits[i] = self.robot.cast_ray(
                p[0], p[1], PI_OVER_2 - self.robot.a - angle, self.max_range
            )

    def take_picture(self):
        """Returns H rows of W (r, g, b, a) tuples (a PIL RGBA image upstream)."""
        self._update()
        W, H = self.cameraShape
        pixels = [[(0, 0, 0, 0)] * W for _ in range(H)]
        size = max(self.robot.world.width, self.robot.world.height)
        horizon = H / 2
        # walls first
        for i in range(W):
            hits = [hit for hit in self.hits[i] if hit.height == 1.0]
            if len(hits) == 0:
                continue
            hit = hits[-1]  # closest
            s = max(min(1.0 - hit.distance / size * self.sizeFadeWithDistance, 1.0), 0.0)
            sc = max(min(1.0 - hit.distance / size * self.colorsFadeWithDistance, 1.0), 0.0)
            hcolor = (
                int(hit.color.red * sc),
                int(hit.color.green * sc),
                int(hit.color.blue * sc),
                255,
            )
            high = (1.0 - s) * H
            for j in range(H):
                dist = max(min(abs(j - horizon) / horizon, 1.0), 0.0)
                if j < high / 2:
                    pixels[j][i] = (
                        int(SW.CAMERA_SKY[0] + SW.CAMERA_SKY_GRAD[0] * dist),
                        int(SW.CAMERA_SKY[1] + SW.CAMERA_SKY_GRAD[1] * dist),
                        int(SW.CAMERA_SKY[2] + SW.CAMERA_SKY_GRAD[2] * dist),
                        255,
                    )
                elif j < H - high / 2:
                    pixels[j][i] = hcolor
                else:
                    pixels[j][i] = (
                        int(SW.CAMERA_GROUND[0] + SW.CAMERA_GROUND_GRAD[0] * dist),
                        int(SW.CAMERA_GROUND[1] + SW.CAMERA_GROUND_GRAD[1] * dist),
                        int(SW.CAMERA_GROUND[2] + SW.CAMERA_GROUND_GRAD[2] * dist),
                        255,
                    )
        # other robots on top, back to front (hits are sorted furthest first)
        for i in range(W):
            hits = [hit for hit in self.hits[i] if hit.height < 1.0]
            for hit in hits:
                s = max(min(1.0 - hit.distance / size * self.sizeFadeWithDistance, 1.0), 0.0)
                sc = max(min(1.0 - hit.distance / size * self.colorsFadeWithDistance, 1.0), 0.0)
                distance_to = H / 2 * (1.0 - sc)
                height = round(hit.height * H / 2.0 * s)
                hcolor = (
                    int(hit.color.red * sc),
                    int(hit.color.green * sc),
                    int(hit.color.blue * sc),
                    255,
                )
                for j in range(height):
                    row = H - j - 1 - round(distance_to)
                    if 0 <= row < H:
                        pixels[row][i] = hcolor
        return pixels


class LightSensor(_Device):
    """Sums brightness / distance**2 over bulbs with a clear line of sight."""

    def __init__(self, position=(0, 0), name="light"):
        self.robot = None
        self.name = name
        self._place(position)
        self.value = 0.0

    def get_brightness(self):
        return self.value

    def get_reading(self):
        return self.value

    def update(self):
        self.value = 0.0
        p = self._mount_point()
        for bulb in self.robot.world._bulbs:
            x, y = bulb.x, bulb.y
            angle = math.atan2(x - p[0], y - p[1])
            dist = distance(x, y, p[0], p[1])
            hits = self.robot.cast_ray(p[0], p[1], angle, dist)
            if len(hits) == 0 and dist > 0:
                self.value += bulb.brightness * SW.LIGHT_MULTIPLIER / (dist ** 2)
        if SW.SENSOR_NOISE and SW.SENSOR_NOISE.get("light_rel_std"):
            world = self.robot.world
            z = noise_normal(SW.SENSOR_NOISE["seed"], world.noise_world_index, NOISE_LIGHT, self.noise_index, world.noise_step)
            self.value = max(self.value * (1.0 + SW.SENSOR_NOISE["light_rel_std"] * z), 0.0)


class SmellSensor(_Device):
    """Reads the world's smell grid at the mount point."""

    def __init__(self, position=(0, 0), name="smell"):
        self.robot = None
        self.name = name
        self._place(position)
        self.value = 0.0

    def get_reading(self):
        return self.value

    def update(self):
        p = self._mount_point()
        self.value = self.robot.world._grid_get(p[0], p[1])


# ----------------------------------------------------------------------------
# L2 robot
# ----------------------------------------------------------------------------


class Robot:
    """Wheeled robot with a 4-edge bounding box (upstream robot.py)."""

    # default body extents in the robot frame (cm): length 15, width ~13.3
    DEFAULT_BBOX = (-7.5, 7.5, -6.67, 6.67)

    def __init__(
        self,
        x=0,
        y=0,
        a=0,
        color="red",
        name="Robbie",
        height=0.25,
        max_trans_velocity=None,
        max_rotate_velocity=None,
        bbox=None,
    ):
        if max_trans_velocity is None:
            max_trans_velocity = SW.ROBOT_MAX_TRANS_VELOCITY
        if max_rotate_velocity is None:
            max_rotate_velocity = SW.ROBOT_MAX_ROTATE_VELOCITY
        self.world = None
        self.name = name
        self.x, self.y = x, y
        self.a = a * PI_OVER_180  # constructor takes degrees
        self.color = Color(color)
        self.height = height
        self.max_trans_velocity = max_trans_velocity
        self.max_rotate_velocity = max_rotate_velocity
        self.vx = self.vy = self.va = 0.0
        self.tvx = self.tvy = self.tva = 0.0
        self.vx_ramp = self.vy_ramp = self.va_ramp = 1.0
        self.stalled = False
        self.bbox = tuple(bbox) if bbox is not None else Robot.DEFAULT_BBOX
        self._devices = []
        self.bounding_lines = [Line(Point(0, 0), Point(0, 0)) for _ in range(4)]
        self.update_boundingbox(*self.compute_boundingbox(self.x, self.y, self.a))
        self.counted_tests = 0

    # -- API -----------------------------------------------------------------
    def add_device(self, device):
        device.robot = self
        self._devices.append(device)
        return device

    def __getitem__(self, item):
        if isinstance(item, int):
            return self._devices[item]
        for device in self._devices:
            if device.name == item:
                return device
        raise KeyError(item)

    @staticmethod
    def _target(v):
        return v if SW.MOVE_ROUND_DIGITS is None else round(v, SW.MOVE_ROUND_DIGITS)

    def move(self, translate, rotate):
        self.tvx = self._target(translate * self.max_trans_velocity)
        self.tva = self._target(rotate * self.max_rotate_velocity)

    def forward(self, translate):
        self.tvx = self._target(translate * self.max_trans_velocity)

    def backward(self, translate):
        self.tvx = self._target(-translate * self.max_trans_velocity)

    def turn(self, rotate):
        self.tva = self._target(rotate * self.max_rotate_velocity)

    def stop(self):
        self.tvx = self.tvy = self.tva = 0.0

    def get_pose(self):
        return (self.x, self.y, self.a)

    # -- geometry ------------------------------------------------------------
    def compute_boundingbox(self, px, py, pa):
        min_x, max_x, min_y, max_y = self.bbox
        ps = []
        for x, y in ((min_x, max_y), (min_x, min_y), (max_x, min_y), (max_x, max_y)):
            dist = distance(0, 0, x, y)
            angle = math.atan2(-x, y)
            ps.append(rotate_around(px, py, dist, pa + angle + PI_OVER_2))
        return ps

    def update_boundingbox(self, p1, p2, p3, p4):
        pts = (p1, p2, p3, p4)
        for k in range(4):
            a, b = pts[k], pts[(k + 1) % 4]
            line = self.bounding_lines[k]
            line.p1.x, line.p1.y = a[0], a[1]
            line.p2.x, line.p2.y = b[0], b[1]

    def cast_ray(self, x1, y1, a, maxRange, x2=None, y2=None):
        """One ray against every segment in the world; hits furthest first."""
        hits = []
        if x2 is None:
            x2 = math.sin(a) * maxRange + x1
        if y2 is None:
            y2 = math.cos(a) * maxRange + y1
        for wall in self.world._walls:
            if wall.robot is self:
                continue
            for line in wall.lines:
                self.world.counted_tests += 1
                pos = intersect_hit(x1, y1, x2, y2, line.p1.x, line.p1.y, line.p2.x, line.p2.y)
                if pos is not None:
                    dist = distance(pos[0], pos[1], x1, y1)
                    height = 1.0 if wall.robot is None else wall.robot.height
                    color = wall.robot.color if wall.robot else wall.color
                    boundary = len(wall.lines) == 1
                    hits.append(Hit(wall.robot, height, pos[0], pos[1], dist, color, x1, y1, boundary))
        hits.sort(key=lambda h: h.distance, reverse=True)
        return hits

    # -- motion --------------------------------------------------------------
    @staticmethod
    def _deltav(tv, v, maxdv, time_step):
        if not SW.VELOCITY_RAMP:
            return tv
        lim = maxdv * time_step
        return v + min(max(tv - v, -lim), lim)

    def step(self, time_step):
        """Integrate, test the proposed bounding box, commit or stall."""
        va = self._deltav(self.tva, self.va, self.va_ramp, time_step)
        vx = self._deltav(self.tvx, self.vx, self.vx_ramp, time_step)
        vy = self._deltav(self.tvy, self.vy, self.vy_ramp, time_step)
        offset = PI_OVER_2
        pa = self.a - va * time_step
        if SW.MOTION_DT_ON_VY_ONLY:
            tvx = vx * math.sin(-pa + offset) + vy * math.cos(-pa + offset) * time_step
            tvy = vx * math.cos(-pa + offset) - vy * math.sin(-pa + offset) * time_step
        else:
            tvx = (vx * math.sin(-pa + offset) + vy * math.cos(-pa + offset)) * time_step
            tvy = (vx * math.cos(-pa + offset) - vy * math.sin(-pa + offset)) * time_step
        px = self.x + tvx
        py = self.y + tvy
        p1, p2, p3, p4 = self.compute_boundingbox(px, py, pa)
        self.stalled = False
        for wall in self.world._walls:
            if wall.robot is self:
                continue
            for line in wall.lines:
                w1, w2 = line.p1, line.p2
                self.world.counted_tests += 4
                if (
                    intersect(p1[0], p1[1], p2[0], p2[1], w1.x, w1.y, w2.x, w2.y)
                    or intersect(p2[0], p2[1], p3[0], p3[1], w1.x, w1.y, w2.x, w2.y)
                    or intersect(p3[0], p3[1], p4[0], p4[1], w1.x, w1.y, w2.x, w2.y)
                    or intersect(p4[0], p4[1], p1[0], p1[1], w1.x, w1.y, w2.x, w2.y)
                ):
                    self.stalled = True
        if not self.stalled:
            self.va, self.vx, self.vy = va, vx, vy
            self.x, self.y, self.a = px, py, pa
            self.update_boundingbox(p1, p2, p3, p4)
        elif SW.STALL_ZEROES_VELOCITY:
            self.va = self.vx = self.vy = 0.0
        else:
            self.va, self.vx, self.vy = va, vx, vy

    def update(self):
        for device in self._devices:
            device.update()


# ----------------------------------------------------------------------------
# L3 world
# ----------------------------------------------------------------------------


class World:
    """Rectangular 2-D world: walls, bulbs, food (smell), robots, clock."""

    def __init__(
        self,
        width=500,
        height=250,
        seed=0,
        boundary_wall=True,
        boundary_wall_color="purple",
        ground_color="green",
        smell_cell_size=None,
        time_step=0.1,
    ):
        self.width = width
        self.height = height
        self.seed = seed
        self.time = 0.0
        self.time_step = time_step
        self.ground_color = Color(ground_color)
        self._walls = []
        self._bulbs = []
        self._food = []
        self._robots = []
        self._grid = None
        self.smell_cell_size = smell_cell_size if smell_cell_size else max(width, height) / 20.0
        self.counted_tests = 0
        if boundary_wall:
            c = Color(boundary_wall_color)
            w, h = width, height
            self._walls.append(Wall(c, None, Line(Point(0, 0), Point(w, 0))))
            self._walls.append(Wall(c, None, Line(Point(0, 0), Point(0, h))))
            self._walls.append(Wall(c, None, Line(Point(0, h), Point(w, h))))
            self._walls.append(Wall(c, None, Line(Point(w, 0), Point(w, h))))

    # -- construction --------------------------------------------------------
    def add_wall(self, color, x1, y1, x2, y2, box=True):
        c = Color(color)
        if box:
            p1, p2, p3, p4 = Point(x1, y1), Point(x2, y1), Point(x2, y2), Point(x1, y2)
            wall = Wall(c, None, Line(p1, p2), Line(p2, p3), Line(p3, p4), Line(p4, p1))
        else:
            wall = Wall(c, None, Line(Point(x1, y1), Point(x2, y2)))
        # keep robot boxes at the end of the list: static walls first
        n_static = sum(1 for w in self._walls if w.robot is None)
        self._walls.insert(n_static, wall)
        self._grid = None
        return wall

    def add_bulb(self, color, x, y, z=0.0, brightness=1.0, name="bulb"):
        bulb = Bulb(color, x, y, z, brightness, name)
        self._bulbs.append(bulb)
        return bulb

    def add_food(self, x, y, standard_deviation):
        self._food.append((x, y, standard_deviation))
        self._grid = None

    def add_robot(self, robot):
        robot.world = self
        self._robots.append(robot)
        self._walls.append(Wall(robot.color, robot, *robot.bounding_lines))
        return robot

    # -- smell grid ----------------------------------------------------------
    def _build_grid(self):
        cs = self.smell_cell_size
        gw = int(math.ceil(self.width / cs))
        gh = int(math.ceil(self.height / cs))
        grid = [[0.0] * gw for _ in range(gh)]
        for gy in range(gh):
            for gx in range(gw):
                cx, cy = (gx + 0.5) * cs, (gy + 0.5) * cs
                v = 0.0
                for (fx, fy, sd) in self._food:
                    d = distance(cx, cy, fx, fy)
                    v += math.exp(-(d * d) / (2.0 * sd * sd))
                grid[gy][gx] = v
        self._grid = grid

    def _grid_get(self, x, y):
        if self._grid is None:
            self._build_grid()
        cs = self.smell_cell_size
        gx, gy = int(x / cs), int(y / cs)
        if gy < 0 or gy >= len(self._grid) or gx < 0 or gx >= len(self._grid[0]):
            return 0.0
        return self._grid[gy][gx]

    # -- stepping ------------------------------------------------------------
    def step(self, time_step=None):
        """One tick: move robots, refresh devices, advance the clock."""
        time_step = self.time_step if time_step is None else time_step
        if SW.MOVE_ALL_THEN_SENSE:
            for robot in self._robots:
                robot.step(time_step)
            self.time = round(self.time + time_step, 10)
            for robot in self._robots:
                robot.update()
        else:
            for robot in self._robots:
                robot.step(time_step)
                robot.update()
            self.time = round(self.time + time_step, 10)
        self.noise_step = getattr(self, "noise_step", 0) + 1  # (one noise block per sensor and sensing step)

    def update(self):
        for robot in self._robots:
            robot.update()

    def steps(self, n, time_step=None):
        for _ in range(n):
            self.step(time_step)

    # -- accounting (SURVEY.md 8(d) unit definitions) ------------------------
    def segment_count_for(self, robot):
        return sum(len(w.lines) for w in self._walls if w.robot is not robot)


def assert_default_switches():
    """The CUDA engine implements the defaults; refuse to compare otherwise."""
    assert SW.MOTION_DT_ON_VY_ONLY and SW.VELOCITY_RAMP and SW.STALL_ZEROES_VELOCITY
    assert SW.MOVE_ALL_THEN_SENSE and SW.RANGE_FAN_INCLUSIVE
    assert SW.LIGHT_MULTIPLIER == 1000.0 and not SW.SMELL_WALLS_BLOCK
    assert SW.SENSOR_NOISE is None, "parity is stated with sensor noise disabled"
    # facade constants: the product's FacadeDefaults must carry the same values
    from aitk_robots_b200.world import FacadeDefaults as FD

    for k in ("ROBOT_MAX_ROTATE_VELOCITY", "ROBOT_MAX_TRANS_VELOCITY", "MOVE_ROUND_DIGITS", "CAMERA_DEFAULT_FOV",
              "CAMERA_DEFAULT_SIZE"):
        assert getattr(SW, k) == getattr(FD, k), k
