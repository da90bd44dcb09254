# -*- coding: utf-8 -*-
"""Shared parity checker: CUDA engine (through the C ABI) vs the float64 oracle."""
import math

import numpy as np

import aitk_robots_b200 as rb
from oracle import aitk_oracle as orc
from oracle import scene_adapter as sa

# BASELINE.json tolerances
REL_DIST = 1e-4  # distances within 1e-4 relative
# hit / no-hit may differ only where a ray passes within 1e-6 (relative to the
# world extent) of a segment endpoint
ENDPOINT_REL = 1e-6


def run_engine(scene, n_steps, flags=rb.STEP_ALL, device=0):
    eng = rb.BatchEngine(scene, device=device)
    for _ in range(n_steps):
        eng.step(flags=flags)
    eng.synchronize()
    out = {k: eng.download(k) for k in ("pose", "vel", "stalled", "range", "light", "smell")}
    if eng.n_camera:
        out["camera"] = eng.download("camera")
    eng.close()
    return out


def range_rays(world, dev):
    """(x1, y1, x2, y2) of every ray of an oracle RangeSensor at the current pose."""
    p = dev._mount_point()
    incrs = list(orc.arange(-dev.width / 2, dev.width / 2, dev.width / 2)) if dev.width != 0 else [0.0]
    rays = []
    for incr in incrs:
        a = -dev.robot.a - dev.a + orc.PI_OVER_2 + incr
        rays.append((p[0], p[1], math.sin(a) * dev.max + p[0], math.cos(a) * dev.max + p[1]))
    return rays


def compare(scene, n_steps=1, worlds=None, camera=True, ref_mod=None):
    """Step both engines n_steps and compare everything.  Returns a stats dict;
    raises AssertionError on a parity violation.  `ref_mod`: the module that supplies the
    reference observations -- the restatement oracle (default) or the real ``aitk.robots``
    (tests/test_real_aitk.py); the geometric margins that excuse an end-point flip are
    always measured on an oracle world stepped in lockstep."""
    orc.assert_default_switches()
    worlds = list(range(scene.n_worlds)) if worlds is None else list(worlds)
    gpu = run_engine(scene, n_steps)
    return compare_outputs(scene, gpu, n_steps, worlds, camera, ref_mod)


def compare_outputs(scene, gpu, n_steps, worlds, camera=True, ref_mod=None):
    """`gpu`: dict of engine outputs after n_steps (run_engine layout), checked world by world."""
    stats = {"worlds": len(worlds), "ranges": 0, "range_hit_flips": 0, "max_rel_dist": 0.0, "pixels": 0,
             "pixel_mismatch": 0, "max_pose_err": 0.0, "lights": 0, "light_flips": 0}
    for i in worlds:
        world, devs = sa.build_world(scene, i)
        for _ in range(n_steps):
            world.step()
        if ref_mod is None or ref_mod is orc:
            ref = sa.observe(world, devs, camera=camera)
        else:
            rworld, rdevs = sa.build_world(scene, i, ref_mod)
            for _ in range(n_steps):
                rworld.step()
            ref = sa.observe(rworld, rdevs, camera=camera)
        size = max(world.width, world.height)
        # collision flags: exact
        assert list(gpu["stalled"][i]) == ref["stalled"], ("stalled", i, gpu["stalled"][i], ref["stalled"])
        # pose / velocity: float64 on both sides
        perr = np.abs(gpu["pose"][i] - np.array(ref["pose"])).max()
        stats["max_pose_err"] = max(stats["max_pose_err"], float(perr))
        assert perr < 1e-9, ("pose", i, perr)
        assert np.abs(gpu["vel"][i] - np.array(ref["vel"])).max() < 1e-12, ("vel", i)
        # range sensors
        for k, dev in enumerate(devs["range"]):
            g, r = float(gpu["range"][i, k]), ref["range"][k]
            stats["ranges"] += 1
            ghit, rhit = g < dev.max * (1 - 1e-7), r < dev.max
            rel = abs(g - r) / max(abs(r), 1e-12)
            if ghit != rhit or rel > REL_DIST:
                margin = min(sa.endpoint_margin(world, dev.robot, *ray) for ray in range_rays(world, dev))
                assert margin <= ENDPOINT_REL * size * 4, ("range", i, k, g, r, margin)
                stats["range_hit_flips"] += 1
            else:
                stats["max_rel_dist"] = max(stats["max_rel_dist"], rel)
        # light sensors: a flip of one bulb's visibility changes the sum
        for k, dev in enumerate(devs["light"]):
            g, r = float(gpu["light"][i, k]), ref["light"][k]
            stats["lights"] += 1
            if abs(g - r) > REL_DIST * max(abs(r), 1e-9) + 1e-12:
                p = dev._mount_point()
                margin = min(sa.endpoint_margin(world, dev.robot, p[0], p[1], b.x, b.y) for b in world._bulbs)
                assert margin <= ENDPOINT_REL * size * 4, ("light", i, k, g, r, margin)
                stats["light_flips"] += 1
        # smell: float32 field lookup
        for k, dev in enumerate(devs["smell"]):
            g, r = float(gpu["smell"][i, k]), ref["smell"][k]
            assert abs(g - r) <= 1e-6 * max(abs(r), 1.0), ("smell", i, k, g, r)
        # camera pixels: exact away from wall edges
        if camera and devs["camera"]:
            for k, dev in enumerate(devs["camera"]):
                img = gpu["camera"][i, k]
                refimg = np.array(ref["camera"][k], dtype=np.uint8)
                assert img.shape == refimg.shape, (img.shape, refimg.shape)
                diff = np.any(img != refimg, axis=2)
                stats["pixels"] += diff.size
                bad_cols = np.flatnonzero(diff.any(axis=0))
                stats["pixel_mismatch"] += int(diff.sum())
                for col in bad_cols:
                    # a differing column must be one whose ray grazes a wall / box edge
                    W = dev.cameraShape[0]
                    angle = col / W * dev.angle - dev.angle / 2
                    p = dev._mount_point()
                    a = orc.PI_OVER_2 - dev.robot.a - angle
                    x2, y2 = math.sin(a) * dev.max_range + p[0], math.cos(a) * dev.max_range + p[1]
                    margin = sa.endpoint_margin(world, dev.robot, p[0], p[1], x2, y2)
                    assert margin <= ENDPOINT_REL * size * 4, ("camera", i, k, int(col), margin)
    return stats
