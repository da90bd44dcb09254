# -*- coding: utf-8 -*-
"""SceneBatch world i -> oracle World objects.  TEST INFRASTRUCTURE ONLY (see
oracle/aitk_oracle.py); imported by tests/, smoke() and bench.py's CPU legs."""

import math

from . import aitk_oracle as orc

_KINDS = {"RangeSpec": "range", "CameraSpec": "camera", "LightSpec": "light", "SmellSpec": "smell"}


def build_world(scene, i):
    """Oracle world equal to world `i` of the batch; returns (world, devices-by-kind)."""
    w, h = float(scene.world_size[i, 0]), float(scene.world_size[i, 1])
    world = orc.World(w, h, boundary_wall=False, time_step=scene.time_step)
    n = int(scene.seg_count[i])
    for j in range(n):
        x1, y1, x2, y2 = (float(v) for v in scene.seg[i, :, j])
        c = scene.seg_rgba[i, j]
        world.add_wall(orc.Color(int(c[0]), int(c[1]), int(c[2])), x1, y1, x2, y2, box=False)
    if scene.bulbs is not None:
        for b in scene.bulbs[i]:
            world.add_bulb("white", float(b[0]), float(b[1]), float(b[2]), float(b[3]))
    if scene.grid is not None:
        world.smell_cell_size = scene.smell_cell_size
        world._grid = [[float(v) for v in row] for row in scene.grid[i]]
    devs = {"range": [], "camera": [], "light": [], "smell": []}
    for r, spec in enumerate(scene.robots):
        x, y, a = (float(v) for v in scene.pose[i, r])
        robot = orc.Robot(x, y, 0.0, color=orc.Color(*spec.rgba[:3]), height=spec.height,
                          max_trans_velocity=spec.max_trans_velocity, max_rotate_velocity=spec.max_rotate_velocity,
                          bbox=spec.bbox)
        robot.a = a
        robot.update_boundingbox(*robot.compute_boundingbox(robot.x, robot.y, robot.a))
        robot.vx_ramp, robot.vy_ramp, robot.va_ramp = spec.vx_ramp, spec.vy_ramp, spec.va_ramp
        robot.tvx, robot.tvy, robot.tva = (float(v) for v in scene.target_vel[i, r])
        for d in spec.devices:
            kind = _KINDS[type(d).__name__]
            if kind == "range":
                dev = orc.RangeSensor(position=d.position, a=d.a, max=d.max, width=d.width, name=d.name)
            elif kind == "camera":
                dev = orc.Camera(width=d.width, height=d.height, a=d.a, colorsFadeWithDistance=d.colorsFadeWithDistance,
                                 sizeFadeWithDistance=d.sizeFadeWithDistance, max_range=d.max_range,
                                 position=d.position, name=d.name)
            elif kind == "light":
                dev = orc.LightSensor(position=d.position, name=d.name)
            else:
                dev = orc.SmellSensor(position=d.position, name=d.name)
            robot.add_device(dev)
            devs[kind].append(dev)
        world.add_robot(robot)
    return world, devs


def step_and_observe(scene, i, n_steps=1, camera=True):
    """Run n_steps of World.step on world i and return its final observations."""
    world, devs = build_world(scene, i)
    for _ in range(n_steps):
        world.step()
    return observe(world, devs, camera)


def observe(world, devs, camera=True):
    out = {
        "pose": [[r.x, r.y, r.a] for r in world._robots],
        "vel": [[r.vx, r.vy, r.va] for r in world._robots],
        "stalled": [int(r.stalled) for r in world._robots],
        "range": [d.get_distance() for d in devs["range"]],
        "light": [d.get_brightness() for d in devs["light"]],
        "smell": [d.get_reading() for d in devs["smell"]],
    }
    if camera:
        out["camera"] = [d.take_picture() for d in devs["camera"]]
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
