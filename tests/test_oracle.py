# -*- coding: utf-8 -*-
"""Known-answer tests of the CPU restatement oracle (hand-derivable results,
SURVEY.md section 4 item 1) and its golden fixtures.  CPU only."""
import json
import math
import os

import numpy as np
import pytest

from oracle import aitk_oracle as orc
from oracle import scene_adapter as sa
import aitk_robots_b200 as rb

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_intersect_axis_aligned():
    # ray along +x from (0,5) of length 100 against the vertical wall x=40
    assert orc.intersect_hit(0, 5, 100, 5, 40, 0, 40, 10) == (40.0, 5.0)
    # parallel: no hit
    assert orc.intersect_hit(0, 5, 100, 5, 0, 7, 100, 7) is None
    # too short
    assert orc.intersect_hit(0, 5, 30, 5, 40, 0, 40, 10) is None
    # through the shared endpoint of two walls: both report the corner (inclusive ends)
    assert orc.intersect_hit(0, 0, 10, 10, 5, 5, 5, 20) == (5.0, 5.0)
    assert orc.intersect_hit(0, 0, 10, 10, 5, 5, 20, 5) == (5.0, 5.0)
    # degenerate segment
    assert orc.intersect_hit(0, 0, 10, 0, 5, 5, 5, 5) is None


def test_rotate_around_and_distance():
    x, y = orc.rotate_around(10, 20, 5, 0.0)
    assert (x, y) == (15.0, 20.0)
    x, y = orc.rotate_around(0, 0, 2, math.pi / 2)
    assert abs(x) < 1e-15 and abs(y - 2) < 1e-15
    assert orc.distance(0, 0, 3, 4) == 5.0


def _simple_world():
    world = orc.World(200, 150)
    robot = orc.Robot(100, 75, 0)
    robot.add_device(orc.RangeSensor(position=(8, 0), a=0, max=100, width=0))
    world.add_robot(robot)
    return world, robot


def test_range_sensor_known_answer():
    world, robot = _simple_world()
    world.update()
    # heading 0 looks along +x: mount at x=108, boundary wall at x=200
    assert robot[0].get_distance() == pytest.approx(92.0, abs=1e-12)
    world.add_wall("blue", 150, 0, 150, 150, box=False)
    world.update()
    assert robot[0].get_distance() == pytest.approx(42.0, abs=1e-12)
    # out of range -> max
    r2 = orc.RangeSensor(position=(8, 0), a=0, max=20, width=0)
    robot.add_device(r2)
    world.update()
    assert r2.get_distance() == 20 and r2.get_reading() == 1.0


def test_range_fan_has_three_rays():
    s = orc.RangeSensor(width=10.0)
    assert s.ray_count() == 3
    assert orc.RangeSensor(width=0).ray_count() == 1


def test_drive_into_wall_stalls():
    world, robot = _simple_world()
    robot.move(1.0, 0.0)
    stalled_at = None
    for step in range(200):
        world.step()
        if robot.stalled:
            stalled_at = step
            break
    assert stalled_at is not None
    # front of the box (7.5 ahead of the centre) stops short of the wall at x=200
    assert robot.x + 7.5 <= 200.0 and robot.x + 7.5 > 200.0 - 2.0 - 1e-9
    assert robot.vx == 0.0
    # the stall zeroed the velocity: the robot creeps up again (ramp) but never crosses
    stalls = 0
    for _ in range(100):
        world.step()
        stalls += robot.stalled
        assert robot.x + 7.5 <= 200.0
    assert stalls > 0


def test_motion_ramp_and_heading():
    world, robot = _simple_world()
    robot.move(1.0, 1.0)
    world.step()
    # one step: velocities ramp by ramp*dt = 0.1
    assert robot.vx == pytest.approx(0.1) and robot.va == pytest.approx(0.1)
    assert robot.a == pytest.approx(-0.01)
    assert robot.x == pytest.approx(100 + 0.1 * math.cos(-0.01))
    assert robot.y == pytest.approx(75 + 0.1 * math.sin(-0.01))


def test_light_and_occlusion():
    world, robot = _simple_world()
    light = robot.add_device(orc.LightSensor(position=(0, 0)))
    world.add_bulb("white", 150, 75, 0, 2.0)
    world.update()
    assert light.get_brightness() == pytest.approx(2.0 * 1000.0 / 50.0 ** 2)
    world.add_wall("blue", 120, 0, 120, 150, box=False)
    world.update()
    assert light.get_brightness() == 0.0


def test_smell_grid_lookup():
    world, robot = _simple_world()
    smell = robot.add_device(orc.SmellSensor(position=(0, 0)))
    world.add_food(100, 75, 20.0)
    world.update()
    cs = world.smell_cell_size
    cx, cy = (int(100 / cs) + 0.5) * cs, (int(75 / cs) + 0.5) * cs
    d2 = (cx - 100) ** 2 + (cy - 75) ** 2
    assert smell.get_reading() == pytest.approx(math.exp(-d2 / (2 * 400.0)))


def test_camera_columns_known_answer():
    world = orc.World(200, 200)
    robot = orc.Robot(100, 100, 0)
    cam = robot.add_device(orc.Camera(width=8, height=16, a=60))
    world.add_robot(robot)
    pic = cam.take_picture()
    assert len(pic) == 16 and len(pic[0]) == 8
    # centre column (i = 4, angle 0) looks along +x and hits the purple boundary at distance 100
    s = 1.0 - 100.0 / 200.0 * 1.0
    sc = 1.0 - 100.0 / 200.0 * 0.9
    high = (1 - s) * 16
    col = [pic[j][4] for j in range(16)]
    for j in range(16):
        if j < high / 2:
            assert col[j] == (0, 0, 128, 255)
        elif j < 16 - high / 2:
            assert col[j] == (int(128 * sc), 0, int(128 * sc), 255)
        else:
            assert col[j] == (0, 128, 0, 255)


def test_camera_sees_other_robot_strip():
    world = orc.World(200, 200)
    robot = orc.Robot(50, 100, 0)
    cam = robot.add_device(orc.Camera(width=8, height=16, a=60))
    world.add_robot(robot)
    other = orc.Robot(120, 100, 0, color="blue", height=0.5)
    world.add_robot(other)
    pic = cam.take_picture()
    blues = [pic[j][4] for j in range(16) if pic[j][4][2] > 0 and pic[j][4][0] == 0 and pic[j][4][1] == 0 and pic[j][4][2] != 128]
    assert len(blues) > 0


def test_counted_tests_match_formula():
    scene = rb.scenes.demo_scene(2)
    world, devs = sa.build_world(scene, 0)
    world.counted_tests = 0
    world.step()
    for cam in devs["camera"]:
        cam.take_picture()
    assert world.counted_tests == scene.slice(0, 1).tests_per_step()


@pytest.mark.parametrize("name", ["config1_1000steps.json", "demo_3steps.json"])
def test_golden_fixtures(name):
    """The oracle still reproduces the committed golden vectors (tests/golden/make_golden.py)."""
    with open(os.path.join(GOLDEN, name)) as f:
        gold = json.load(f)
    scene = getattr(rb.scenes, gold["generator"])(**gold["kwargs"])
    for entry in gold["worlds"]:
        obs = sa.step_and_observe(scene, entry["world"], gold["steps"], camera=False)
        for key in ("pose", "range", "light", "smell", "stalled"):
            np.testing.assert_allclose(np.array(obs[key], dtype=np.float64), np.array(entry[key], dtype=np.float64),
                                       rtol=1e-12, atol=1e-12, err_msg=key)


def test_philox4x32_10_known_answers():
    """Random123 known-answer vectors (kat_vectors: philox4x32 10 rounds) pin the generator the
    seeded sensor noise is built on -- the CUDA side is then checked against this one."""
    from oracle import aitk_oracle as orc

    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        assert orc.philox4x32_10(ctr, key) == want, (ctr, key)
    # the normal variates built on it: mean 0, deviation 1
    zs = [orc.noise_normal(7, w, orc.NOISE_RANGE, 3, 5) for w in range(4000)]
    m = sum(zs) / len(zs)
    sd = (sum((z - m) ** 2 for z in zs) / len(zs)) ** 0.5
    assert abs(m) < 0.06 and abs(sd - 1.0) < 0.05, (m, sd)


def test_topdown_restatement_known_pixels():
    """oracle/topdown.py: hand-checkable picture (one wall, one bulb, one robot)."""
    import numpy as np

    from oracle import topdown

    seg = np.zeros((4, 4), dtype=np.float32)
    seg[:, 0] = (10, 10, 90, 10)
    rgba = np.zeros((4, 4), dtype=np.uint8)
    rgba[0] = (0, 0, 255, 7)
    poses = np.array([[50.0, 30.0, 0.0]])
    img, fragile = topdown.render((100, 50), seg, rgba, 1, poses, [(-7.5, 7.5, -6.67, 6.67)], [(255, 0, 0, 255)],
                                  bulbs=[(20, 40, 0, 1)], img_w=100, img_h=50)
    assert tuple(img[10, 50]) == (0, 0, 255, 255) and tuple(img[9, 50]) == (0, 0, 255, 255)  # line 1.5 px thick
    assert tuple(img[12, 50]) == (0, 128, 0, 255)
    assert tuple(img[30, 50]) == (255, 0, 0, 255) and tuple(img[30, 58]) == (0, 128, 0, 255)
    assert tuple(img[40, 20]) == (255, 255, 0, 255)
    assert not fragile.all()
    # the window under the robot: robot not drawn, wall 20 units to its left (heading 0 = +x, wall is at -y)
    img, _ = topdown.render((100, 50), seg, rgba, 1, poses, [(-7.5, 7.5, -6.67, 6.67)], [(255, 0, 0, 255)],
                            img_w=50, img_h=50, robot=0, view=(50.0, 50.0))
    assert tuple(img[25, 25]) == (0, 128, 0, 255)
    assert tuple(img[25, 5]) == (0, 0, 255, 255) and tuple(img[25, 45]) == (0, 128, 0, 255)
