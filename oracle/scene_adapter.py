# -*- coding: utf-8 -*-
"""SceneBatch world i -> oracle World objects.  TEST INFRASTRUCTURE ONLY (see
oracle/aitk_oracle.py); imported by tests/, smoke() and bench.py's CPU legs."""

import math

from . import aitk_oracle as orc

_KINDS = {"RangeSpec": "range", "CameraSpec": "camera", "LightSpec": "light", "SmellSpec": "smell"}


def find_real_aitk(root=None):
    """The real ``aitk.robots`` module if it can be imported (also tried from
    ``baseline/_ref`` under the repo root), else None.  In the build container it cannot
    (SURVEY.md section 0); the first environment that has it pins the oracle."""
    import importlib
    import os
    import sys

    try:
        return importlib.import_module("aitk.robots")
    except Exception:
        pass
    root = root or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = os.path.join(root, "baseline", "_ref")
    if os.path.isdir(ref):
        sys.path.insert(0, ref)
        try:
            return importlib.import_module("aitk.robots")
        except Exception:
            sys.path.remove(ref)
    return None


def _is_oracle(mod):
    return mod is orc


def build_world(scene, i, mod=None):
    """World equal to world `i` of the batch, built with module `mod`: the restatement
    oracle (default) or the REAL ``aitk.robots`` (tests/test_real_aitk.py, bench.py --impl
    reference when it is importable).  Returns (world, devices-by-kind).

    With the real package only its public constructors and methods are used where they are
    known; the few places that need more (exact radian heading, per-robot ramps, a smell
    grid given as numbers, box extents) are guarded with hasattr and documented -- they are
    recollection and the first run against the real package is what validates them."""
    mod = orc if mod is None else mod
    real = not _is_oracle(mod)
    w, h = float(scene.world_size[i, 0]), float(scene.world_size[i, 1])
    if real:
        try:
            world = mod.World(w, h, boundary_wall=False)
        except TypeError:
            world = mod.World(w, h)
        if hasattr(world, "time_step"):
            world.time_step = scene.time_step
    else:
        world = mod.World(w, h, boundary_wall=False, time_step=scene.time_step)
    if not real:
        world.noise_world_index, world.noise_step = i, 0
    Color = getattr(mod, "Color", None)
    if Color is None and real:
        try:
            from aitk.robots.colors import Color  # noqa: N806  [RECALL placement]
        except Exception:
            Color = None
    mkcolor = (lambda r, g, b: Color(int(r), int(g), int(b))) if Color is not None else (lambda r, g, b: (int(r), int(g), int(b)))
    n = int(scene.seg_count[i])
    for j in range(n):
        x1, y1, x2, y2 = (float(v) for v in scene.seg[i, :, j])
        c = scene.seg_rgba[i, j]
        world.add_wall(mkcolor(c[0], c[1], c[2]), x1, y1, x2, y2, box=False)
    if scene.bulbs is not None:
        for b in scene.bulbs[i]:
            world.add_bulb("white", float(b[0]), float(b[1]), float(b[2]), float(b[3]))
    if scene.grid is not None:
        if real:
            # the real package diffuses smell from add_food(); the batch carries the food list
            for (fx, fy, sd) in ([] if scene.food is None else scene.food[i]):
                world.add_food(float(fx), float(fy), float(sd))
        else:
            world.smell_cell_size = scene.smell_cell_size
            world._grid = [[float(v) for v in row] for row in scene.grid[i]]
    devs = {"range": [], "camera": [], "light": [], "smell": []}
    for r, spec in enumerate(scene.robots):
        x, y, a = (float(v) for v in scene.pose[i, r])
        if real:
            robot = mod.Robot(x, y, math.degrees(a))
            for k, v in (("height", spec.height), ("max_trans_velocity", spec.max_trans_velocity),
                         ("max_rotate_velocity", spec.max_rotate_velocity), ("vx_ramp", spec.vx_ramp),
                         ("vy_ramp", spec.vy_ramp), ("va_ramp", spec.va_ramp)):
                if hasattr(robot, k):
                    setattr(robot, k, v)
            if hasattr(robot, "a"):
                robot.a = a  # exact radians (the constructor converts from degrees)
        else:
            robot = mod.Robot(x, y, 0.0, color=mod.Color(*spec.rgba[:3]), height=spec.height,
                              max_trans_velocity=spec.max_trans_velocity, max_rotate_velocity=spec.max_rotate_velocity,
                              bbox=spec.bbox)
            robot.a = a
            robot.update_boundingbox(*robot.compute_boundingbox(robot.x, robot.y, robot.a))
            robot.vx_ramp, robot.vy_ramp, robot.va_ramp = spec.vx_ramp, spec.vy_ramp, spec.va_ramp
        for d in spec.devices:
            kind = _KINDS[type(d).__name__]
            if kind == "range":
                dev = mod.RangeSensor(position=d.position, a=d.a, max=d.max, width=d.width, name=d.name)
            elif kind == "camera":
                dev = mod.Camera(width=d.width, height=d.height, a=d.a, colorsFadeWithDistance=d.colorsFadeWithDistance,
                                 sizeFadeWithDistance=d.sizeFadeWithDistance, max_range=d.max_range,
                                 position=d.position, name=d.name)
            elif kind == "light":
                dev = mod.LightSensor(position=d.position, name=d.name)
            else:
                dev = mod.SmellSensor(position=d.position, name=d.name)
            robot.add_device(dev)
            if not real:
                dev.noise_index = len(devs[kind])  # index among the world's sensors of this kind, as in the engine
            devs[kind].append(dev)
        world.add_robot(robot)
        tv = [float(v) for v in scene.target_vel[i, r]]
        if real:
            # Robot.move takes fractions of the maximum velocities
            robot.move(tv[0] / spec.max_trans_velocity if spec.max_trans_velocity else 0.0,
                       tv[2] / spec.max_rotate_velocity if spec.max_rotate_velocity else 0.0)
        else:
            robot.tvx, robot.tvy, robot.tva = tv
    return world, devs


def step_and_observe(scene, i, n_steps=1, camera=True, mod=None):
    """Run n_steps of World.step on world i and return its final observations."""
    world, devs = build_world(scene, i, mod)
    for _ in range(n_steps):
        world.step()
    return observe(world, devs, camera)


def _picture_rows(pic):
    """Camera.take_picture() result as H rows of W (r, g, b, a): the oracle returns that
    already, the real package returns a PIL image."""
    if isinstance(pic, list):
        return pic
    import numpy as np

    return np.asarray(pic.convert("RGBA") if hasattr(pic, "convert") else pic).tolist()


def observe(world, devs, camera=True):
    robots = getattr(world, "_robots", None)
    if robots is None:
        robots = list(world)  # the real package iterates robots
    out = {
        "pose": [[r.x, r.y, r.a] for r in robots],
        "vel": [[r.vx, r.vy, r.va] for r in robots],
        "stalled": [int(bool(r.stalled)) for r in robots],
        "range": [d.get_distance() for d in devs["range"]],
        "light": [d.get_brightness() for d in devs["light"]],
        "smell": [d.get_reading() for d in devs["smell"]],
    }
    if camera:
        out["camera"] = [_picture_rows(d.take_picture()) for d in devs["camera"]]
    return out


def endpoint_margin(world, robot, x1, y1, x2, y2):
    """Smallest distance, over the segments the ray can see, between the ray
    and a segment endpoint or between a segment and the ray's far end -- the
    geometric quantity the hit / no-hit tolerance of BASELINE.json is stated in."""
    best = math.inf
    dx, dy = x2 - x1, y2 - y1
    l2 = dx * dx + dy * dy
    for wall in world._walls:
        if wall.robot is robot:
            continue
        for line in wall.lines:
            for (px, py) in ((line.p1.x, line.p1.y), (line.p2.x, line.p2.y)):
                t = 0.0 if l2 == 0 else max(0.0, min(1.0, ((px - x1) * dx + (py - y1) * dy) / l2))
                best = min(best, math.hypot(px - (x1 + t * dx), py - (y1 + t * dy)))
            ex, ey = line.p2.x - line.p1.x, line.p2.y - line.p1.y
            e2 = ex * ex + ey * ey
            for (px, py) in ((x1, y1), (x2, y2)):
                t = 0.0 if e2 == 0 else max(0.0, min(1.0, ((px - line.p1.x) * ex + (py - line.p1.y) * ey) / e2))
                best = min(best, math.hypot(px - (line.p1.x + t * ex), py - (line.p1.y + t * ey)))
    return best
