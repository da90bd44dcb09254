# -*- coding: utf-8 -*-
"""Parity proper: CUDA engine through the C ABI vs the float64 oracle on the
same seeded scenes (BASELINE.json tolerances are in tests/parity.py)."""
import json
import os

import numpy as np
import pytest

import aitk_robots_b200 as rb
import parity

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_demo_scene_every_device_kind():
    stats = parity.compare(rb.scenes.demo_scene(32), n_steps=3)
    assert stats["max_rel_dist"] < 1e-5
    assert stats["pixel_mismatch"] <= 0.001 * stats["pixels"]


def test_config1_trajectory_1000_steps():
    """Config 1 as written: 1000 World.steps, pose/collision trajectory identical."""
    scene = rb.scenes.config1(n_worlds=4, seed=1)
    stats = parity.compare(scene, n_steps=1000)
    assert stats["max_pose_err"] < 1e-9


def test_config1_golden_fixture():
    with open(os.path.join(GOLDEN, "config1_1000steps.json")) as f:
        gold = json.load(f)
    scene = rb.scenes.config1(**gold["kwargs"])
    out = parity.run_engine(scene, gold["steps"])
    for entry in gold["worlds"]:
        i = entry["world"]
        np.testing.assert_allclose(out["pose"][i], np.array(entry["pose"]), atol=1e-9)
        np.testing.assert_array_equal(out["stalled"][i], np.array(entry["stalled"]))
        np.testing.assert_allclose(out["range"][i], np.array(entry["range"]), rtol=1e-4)


def test_demo_golden_fixture():
    with open(os.path.join(GOLDEN, "demo_3steps.json")) as f:
        gold = json.load(f)
    scene = rb.scenes.demo_scene(**gold["kwargs"])
    out = parity.run_engine(scene, gold["steps"])
    for entry in gold["worlds"]:
        i = entry["world"]
        np.testing.assert_allclose(out["pose"][i], np.array(entry["pose"]), atol=1e-9)
        np.testing.assert_array_equal(out["stalled"][i], np.array(entry["stalled"]))
        np.testing.assert_allclose(out["range"][i], np.array(entry["range"]), rtol=1e-4)
        np.testing.assert_allclose(out["light"][i], np.array(entry["light"]), rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(out["smell"][i], np.array(entry["smell"]), rtol=1e-6, atol=1e-9)


def test_config2_camera_and_range():
    scene = rb.scenes.config2(n_worlds=64, cam_w=256, cam_h=128)
    stats = parity.compare(scene, n_steps=2, worlds=range(0, 64, 8))
    assert stats["max_rel_dist"] < 1e-5
    assert stats["pixel_mismatch"] <= 0.001 * stats["pixels"]


def test_config2_boxes_reading():
    """The "20 boxes = 80 segments" reading of BASELINE.json configs[1]."""
    scene = rb.scenes.config2_boxes(n_worlds=32, cam_w=64, cam_h=32)
    stats = parity.compare(scene, n_steps=3, worlds=range(0, 32, 4))
    assert stats["max_rel_dist"] < 1e-5
    assert stats["pixel_mismatch"] <= 0.002 * stats["pixels"]


def test_config3_multi_robot_collisions_and_light():
    scene = rb.scenes.config3(n_worlds=64)
    stats = parity.compare(scene, n_steps=20, worlds=range(0, 64, 4), camera=False)
    assert stats["ranges"] == 16 * 64 and stats["max_rel_dist"] < 1e-5


def test_config4_maze_camera():
    for variant in ("maze", "random"):
        scene = rb.scenes.config4(n_worlds=32, variant=variant)
        stats = parity.compare(scene, n_steps=3, worlds=range(0, 32, 4))
        assert stats["pixel_mismatch"] <= 0.002 * stats["pixels"]


@pytest.mark.parametrize("S,rays", [(12, 1), (300, 4), (2000, 1), (1000, 64)])
def test_config5_sweep_cells(S, rays):
    scene = rb.scenes.config5(n_worlds=8, n_segments=S, n_rays=rays)
    stats = parity.compare(scene, n_steps=1, worlds=range(0, 8, 4), camera=False)
    assert stats["max_rel_dist"] < 1e-5


def test_edge_cases_empty_and_ragged():
    """Worlds with zero walls, ragged segment counts, a robot boxed in."""
    scene = rb.scenes.config2(n_worlds=8, cam_w=32, cam_h=16)
    scene.seg_count[:] = [20, 0, 1, 4, 7, 19, 20, 3]
    stats = parity.compare(scene, n_steps=2)
    assert stats["worlds"] == 8


def test_full_size_properties_config2():
    """BASELINE size (4096 worlds): size-independent properties instead of the oracle."""
    scene = rb.scenes.config2(n_worlds=4096)
    eng = rb.BatchEngine(scene)
    eng.step()
    eng.synchronize()
    rng1 = eng.download("range")
    cam1 = eng.download("camera")
    pose1 = eng.download("pose")
    # idempotence: sensing again without moving reproduces every observation bit for bit
    eng.step(flags=rb.STEP_SENSE | rb.STEP_CAMERA)
    eng.synchronize()
    assert np.array_equal(rng1, eng.download("range"))
    assert np.array_equal(cam1, eng.download("camera"))
    assert np.array_equal(pose1, eng.download("pose"))
    # ranges are bounded by max and positive; every pixel is opaque (walls enclose the room)
    assert (rng1 > 0).all() and (rng1 <= 100.0).all()
    assert (cam1[..., 3] == 255).all()
    # batch independence: world i of the big batch == the same world stepped alone
    for i in (0, 1777, 4095):
        alone = parity.run_engine(scene.slice(i, 1), 1)
        assert np.array_equal(alone["range"][0], rng1[i])
        assert np.array_equal(alone["camera"][0], cam1[i])
    # the first image column band is sky, the last ground
    assert (cam1[:, 0, 0, :, 2] == 128).all() and (cam1[:, 0, -1, :, 1] == 128).all()
    eng.close()


def test_facade_dropin_api():
    world = rb.World(200, 150)
    world.add_wall("blue", 150, 0, 150, 150, box=False)
    robot = rb.Robot(100, 75, 0)
    robot.add_device(rb.RangeSensor(position=(8, 0), a=0, max=100, width=0))
    cam = robot.add_device(rb.Camera(width=64, height=32))
    world.add_robot(robot)
    world.step()
    assert robot[0].get_distance() == pytest.approx(42.0, rel=1e-6)
    pic = cam.take_picture()
    assert pic.size == (64, 32) and pic.mode == "RGBA"
    robot.move(1.0, 0.0)
    for _ in range(5):
        world.step()
    assert robot.x > 100 and not robot.stalled


def test_facade_set_pose_keeps_velocities_and_state():
    """Robot.set_pose on a live world uploads one pose row: the velocity ramp of every robot
    continues (upstream keeps vx / vy / va), and a structural change (add_wall) carries the
    dynamic state into the rebuilt engine."""
    from oracle import aitk_oracle as orc

    def build(mod):
        w = mod.World(300, 200)
        r1, r2 = mod.Robot(60, 100, 0), mod.Robot(220, 100, 180)
        r1.add_device(mod.RangeSensor(position=(8, 0), a=0, max=100, width=0))
        w.add_robot(r1)
        w.add_robot(r2)
        r1.move(1.0, 0.3)
        r2.move(0.5, -0.2)
        return w, r1, r2

    (gw, g1, g2), (ow, o1, o2) = build(rb), build(orc)
    for _ in range(5):
        gw.step()
        ow.step()
    # teleport robot 1 mid-ramp
    g1.set_pose(100, 60, 90)
    o1.x, o1.y, o1.a = 100.0, 60.0, 90 * orc.PI_OVER_180
    o1.update_boundingbox(*o1.compute_boundingbox(o1.x, o1.y, o1.a))
    assert g1.x == 100.0 and g1.y == 60.0
    for _ in range(5):
        gw.step()
        ow.step()
    for g, o in ((g1, o1), (g2, o2)):
        assert abs(g.x - o.x) < 1e-9 and abs(g.y - o.y) < 1e-9 and abs(g.a - o.a) < 1e-9, (g.get_pose(), o.get_pose())
    # structural change: the engine is rebuilt, velocities survive
    gw.add_wall("blue", 10, 10, 20, 10, box=False)
    ow.add_wall("blue", 10, 10, 20, 10, box=False)
    for _ in range(5):
        gw.step()
        ow.step()
    for g, o in ((g1, o1), (g2, o2)):
        assert abs(g.x - o.x) < 1e-9 and abs(g.y - o.y) < 1e-9 and abs(g.a - o.a) < 1e-9, (g.get_pose(), o.get_pose())
    assert g1[0].get_distance() == pytest.approx(o1[0].get_distance(), rel=1e-6)


def test_zero_copy_tensor_and_bind_output():
    import torch

    scene = rb.scenes.config2(n_worlds=16, cam_w=32, cam_h=16)
    eng = rb.BatchEngine(scene)
    t = eng.tensor("range")
    eng.step()
    eng.synchronize()
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), eng.download("range"))
    ext = torch.zeros((16, 1, 16, 32, 4), dtype=torch.uint8, device="cuda")
    eng.bind_output("camera", ext)
    eng.step(flags=rb.STEP_CAMERA)
    eng.synchronize()
    assert int(ext[..., 3].min()) == 255
    eng.bind_output("camera", None)
    eng.close()


def _outputs(scene, steps, env):
    import os

    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return parity.run_engine(scene, steps)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("maker", [
    lambda: rb.scenes.config2(n_worlds=256),
    lambda: rb.scenes.config4(n_worlds=2048),
    lambda: rb.scenes.config4(n_worlds=128, variant="random"),
    lambda: rb.scenes.config5(n_worlds=64, n_segments=300, n_rays=16),
    lambda: rb.scenes.demo_scene(64),
    lambda: rb.scenes.config3(n_worlds=96),
])
def test_culling_and_tiling_never_change_results(maker):
    """Cone / range culling, table compaction, tile size and launch shape only remove or
    regroup work: every output is bit-identical with them switched off or changed."""
    scene = maker()
    # the base run is the fused kernel; BRS_WIDE=1 variants run the two-kernel wide path
    # (one-warp teams, wider teams, several worlds per tile, walls from L2, no PDL)
    base = _outputs(scene, 3, {"BRS_WIDE": "0", "BRS_SOLO": "0"})
    # (BRS_SOLO=1 only takes effect for the shapes the solo path accepts: config 4 here)
    for env in ({"BRS_SOLO": "1"}, {"BRS_SOLO": "1", "BRS_SOLO_CTAS": "8", "BRS_NO_CONE": "1"},
                {"BRS_SOLO": "1", "BRS_SOLO_CTAS": "4", "BRS_SOLO_MAX_CTAS": "1", "BRS_NO_RANGE_CULL": "1"},
                {"BRS_SOLO": "1", "BRS_SOLO_CTAS": "7"}, {"BRS_SOLO": "1", "BRS_SOLO_CTAS": "5"},
                {"BRS_WIDE": "0", "BRS_NO_CONE": "1", "BRS_NO_RANGE_CULL": "1"}, {"BRS_WIDE": "0", "BRS_TW": "1"},
                {"BRS_WIDE": "0", "BRS_TW": "5", "BRS_BIG": "1"},
                {"BRS_WIDE": "0", "BRS_NH": "2"}, {"BRS_WIDE": "0", "BRS_DIRECT_ON_HELPER": "1"},
                {"BRS_WIDE": "0", "BRS_RLIST": "1"}, {"BRS_WIDE": "1", "BRS_RLIST": "1", "BRS_WNT": "64"},
                {"BRS_WIDE": "0", "BRS_RMASK": "0"}, {"BRS_WIDE": "0", "BRS_RMASK": "0", "BRS_TW": "3"},
                {"BRS_WIDE": "0", "BRS_DIRECT_ON_HELPER": "0", "BRS_TW": "2"},
                {"BRS_WIDE": "0", "BRS_NR": "1", "BRS_GROUP_MIN": "1000"},
                {"BRS_WIDE": "0", "BRS_PREP": "1"}, {"BRS_WIDE": "0", "BRS_PREP": "1", "BRS_TW": "3", "BRS_BIG": "1"},
                {"BRS_WIDE": "1"}, {"BRS_WIDE": "1", "BRS_WNT": "32", "BRS_WTW": "1"},
                {"BRS_WIDE": "1", "BRS_WNT": "32", "BRS_WTW": "3"}, {"BRS_WIDE": "1", "BRS_WNT": "96", "BRS_WTW": "2"},
                {"BRS_WIDE": "1", "BRS_WNT": "256", "BRS_WTW": "5", "BRS_WBIG": "1"},
                {"BRS_WIDE": "1", "BRS_PDL": "0", "BRS_NO_CONE": "1"},
                {"BRS_WIDE": "1", "BRS_NR": "1", "BRS_GROUP_MIN": "1000"}):
        other = _outputs(scene, 3, dict({"BRS_SOLO": "0"}, **env))
        for k in base:
            if k in ("range", "light") and env.get("BRS_GROUP_MIN"):
                # group path vs direct path are two FP32 formulations of the same search
                np.testing.assert_allclose(other[k], base[k], rtol=1e-6, err_msg=str((k, env)))
            else:
                assert np.array_equal(other[k], base[k]), (k, env)


def _with_env(env, fn):
    import os

    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("cam_w,cam_h", [(64, 32), (32, 16), (128, 20)])
def test_solo_path_matches_the_oracle(cam_w, cam_h):
    """The warp-per-world path (one robot, one camera; csrc/brs_solo.cuh) against the oracle:
    maze and random walls, batches that are not a multiple of its PREP pass, ragged and empty
    worlds, columns that see nothing (short camera range), robots driven into walls."""
    for variant in ("maze", "random"):
        scene = rb.scenes.config4(n_worlds=41, variant=variant, cam_w=cam_w, cam_h=cam_h)
        scene.seg_count[3] = 0
        scene.seg_count[5] = 37
        scene.target_vel[:, :, 0] = np.linspace(-2.0, 2.0, 41)[:, None]

        def run():
            eng = rb.BatchEngine(scene)
            assert eng.launch_info()["path"].startswith("solo"), eng.launch_info()
            eng.close()
            return parity.compare(scene, n_steps=4, worlds=range(0, 41, 3))

        stats = _with_env({"BRS_SOLO": "1"}, run)
        assert stats["pixel_mismatch"] <= 0.002 * stats["pixels"]
    # a camera that cannot see the far walls leaves blank (0, 0, 0, 0) columns
    scene = rb.scenes.config4(n_worlds=20, variant="maze", cam_w=cam_w, cam_h=cam_h)
    for spec in scene.robots[0].devices:
        if isinstance(spec, rb.CameraSpec):
            spec.max_range = 30.0
    stats = _with_env({"BRS_SOLO": "1"}, lambda: parity.compare(scene, n_steps=2))
    assert stats["pixel_mismatch"] <= 0.002 * stats["pixels"]
    # 60 steps at full speed: robots stall against the maze walls, flags and poses exact
    scene = rb.scenes.config4(n_worlds=16, cam_w=cam_w, cam_h=cam_h)
    scene.target_vel[:, :, 0] = 2.0
    stats = _with_env({"BRS_SOLO": "1"}, lambda: parity.compare(scene, n_steps=60, camera=False))
    assert stats["max_pose_err"] < 1e-9


def test_solo_path_step_flags():
    """Move-only and camera-only steps of the solo path equal the fused kernel's, bit for bit."""
    scene = rb.scenes.config4(n_worlds=100)
    scene.target_vel[:, :, 0] = 2.0

    def run():
        eng = rb.BatchEngine(scene)
        outs = []
        for flags in (rb.STEP_CAMERA, rb.STEP_MOVE, rb.STEP_MOVE, rb.STEP_SENSE, rb.STEP_CAMERA, rb.STEP_ALL):
            eng.step(flags=flags)
            eng.synchronize()
            outs.append({k: eng.download(k) for k in ("pose", "vel", "stalled", "camera")})
        eng.close()
        return outs

    a = _with_env({"BRS_SOLO": "0", "BRS_WIDE": "0"}, run)
    b = _with_env({"BRS_SOLO": "1"}, run)
    for x, y in zip(a, b):
        for k in x:
            assert np.array_equal(x[k], y[k]), k


def test_solo_path_sky_gradient():
    """Non-flat sky / ground rows on the solo path (its generic paint loop)."""
    from oracle import aitk_oracle as orc
    from oracle import scene_adapter as sa

    scene = rb.scenes.config4(n_worlds=6)
    sky_grad, gnd_grad = (40, 20, 100), (30, 60, 10)
    old = (orc.SW.CAMERA_SKY_GRAD, orc.SW.CAMERA_GROUND_GRAD)
    orc.SW.CAMERA_SKY_GRAD, orc.SW.CAMERA_GROUND_GRAD = sky_grad, gnd_grad
    try:
        def run():
            eng = rb.BatchEngine(scene, sky_grad=sky_grad, ground_grad=gnd_grad)
            assert eng.launch_info()["path"].startswith("solo")
            eng.step()
            eng.synchronize()
            cam = eng.download("camera")
            eng.close()
            return cam

        cam = _with_env({"BRS_SOLO": "1"}, run)
        for i in range(6):
            ref = np.array(sa.step_and_observe(scene, i, 1)["camera"][0], dtype=np.uint8)
            diff = np.any(cam[i, 0] != ref, axis=2)
            assert diff.any(axis=0).sum() <= 1, i
    finally:
        orc.SW.CAMERA_SKY_GRAD, orc.SW.CAMERA_GROUND_GRAD = old


def _range_fan_scene(n_worlds, n_segments, n_sensors, width, seed=3):
    """config 5 with fans: every sensor casts 3 rays when width != 0."""
    scene = rb.scenes.config5(n_worlds=n_worlds, n_segments=n_segments, n_rays=n_sensors, seed=seed)
    scene.robots[0].devices = [rb.RangeSpec(position=(1.0, 0.5), a=360.0 * i / n_sensors, max=150.0 + i, width=width,
                                            name="ray%d" % i) for i in range(n_sensors)]
    return scene


@pytest.mark.parametrize("S,rays", [(12, 1), (12, 7), (100, 16), (100, 64), (300, 4), (300, 256), (1000, 33), (2000, 2)])
def test_scan_path_bit_identical_and_matches_the_oracle(S, rays):
    """The solo range path (one warp per world, walls streamed, table in blocks; brs_scan_kernel)
    against the fused kernel searching the same shared-origin table (BRS_GROUP_MIN=1), bit for
    bit, and against the oracle; ragged / empty worlds and robots driven into walls included."""
    B = 37 if S < 1000 else 9
    scene = rb.scenes.config5(n_worlds=B, n_segments=S, n_rays=rays)
    scene.seg_count[1] = 0
    scene.seg_count[2] = min(S, 5)
    scene.target_vel[:, :, 0] = np.linspace(-2.0, 2.0, B)[:, None]

    def solo():
        eng = rb.BatchEngine(scene)
        assert eng.launch_info()["path"].startswith("solo range"), eng.launch_info()
        eng.close()
        return parity.run_engine(scene, 6)

    a = _with_env({"BRS_SOLO": "0", "BRS_WIDE": "0", "BRS_GROUP_MIN": "1"}, lambda: parity.run_engine(scene, 6))
    for env in ({"BRS_SOLO": "1"}, {"BRS_SOLO": "1", "BRS_SOLO_INTERLEAVE": "0", "BRS_NO_CONE": "1"},
                {"BRS_SOLO": "1", "BRS_SCAN_SL": "1" if rays > 8 else "4", "BRS_SOLO_CTAS": "8"}):
        b = _with_env(env, solo)
        for k in a:
            assert np.array_equal(a[k], b[k]), (k, env)
    stats = parity.compare_outputs(scene, b, 6, range(0, B, 4), camera=False)
    assert stats["max_rel_dist"] < 1e-5 and stats["max_pose_err"] < 1e-9


def test_scan_path_fans_and_config1():
    """Fans of three rays per sensor (the reading is the minimum over the fan), an off-centre
    mount, and config 1 (one sensor, one ray) on the solo range path."""
    for scene in (_range_fan_scene(21, 40, 5, 30.0), _range_fan_scene(21, 200, 40, 10.0),
                  rb.scenes.config1(n_worlds=50)):
        a = _with_env({"BRS_SOLO": "0", "BRS_WIDE": "0", "BRS_GROUP_MIN": "1"}, lambda: parity.run_engine(scene, 5))

        def solo():
            eng = rb.BatchEngine(scene)
            assert eng.launch_info()["path"].startswith("solo range"), eng.launch_info()
            eng.close()
            return parity.run_engine(scene, 5)

        b = _with_env({"BRS_SOLO": "1"}, solo)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        stats = parity.compare_outputs(scene, b, 5, range(scene.n_worlds), camera=False)
        assert stats["max_rel_dist"] < 1e-5


def test_seeded_sensor_noise():
    """Philox sensor noise: every noisy reading equals the oracle's noise model sample for
    sample; the residuals have the requested deviation; a reading's noise depends only on (seed,
    global world index, sensor, step) -- not on the launch plan or on how the batch is sharded."""
    from oracle import aitk_oracle as orc
    from oracle import scene_adapter as sa

    seed, rstd, lrel = 0x1234_5678_9ABC, 1.5, 0.05
    scene = rb.scenes.config3(n_worlds=24)

    def run(sc, env=None, first_world=0, steps=3, noise=True):
        def go():
            eng = rb.BatchEngine(sc)
            if noise:
                eng.set_noise(seed, range_std=rstd, light_rel_std=lrel, first_world=first_world)
            for _ in range(steps):
                eng.step()
            eng.synchronize()
            out = {k: eng.download(k) for k in ("range", "light", "pose", "stalled")}
            eng.close()
            return out

        return _with_env(env or {}, go)

    noisy, clean = run(scene), run(scene, noise=False)
    assert np.array_equal(noisy["pose"], clean["pose"])  # noise touches observations only
    assert not np.array_equal(noisy["range"], clean["range"])
    # sample-for-sample against the oracle's restatement of the noise model
    orc.SW.SENSOR_NOISE = {"seed": seed, "range_std": rstd, "light_rel_std": lrel}
    try:
        for i in (0, 7, 23):
            ref = sa.step_and_observe(scene, i, 3, camera=False)
            np.testing.assert_allclose(noisy["range"][i], ref["range"], rtol=1e-4, atol=2e-4)
            np.testing.assert_allclose(noisy["light"][i], ref["light"], rtol=1e-4, atol=1e-6)
    finally:
        orc.SW.SENSOR_NOISE = None
    # reproducible, and independent of tiling / launch plan
    again = run(scene, {"BRS_TW": "3", "BRS_NH": "2"})
    assert np.array_equal(again["range"], noisy["range"]) and np.array_equal(again["light"], noisy["light"])
    # a shard that passes its world offset sees the noise of the same global worlds
    part = run(scene.slice(10, 9), first_world=10)
    assert np.array_equal(part["range"], noisy["range"][10:19]) and np.array_equal(part["light"], noisy["light"][10:19])
    # statistics on the solo range path: 64 rays x 512 worlds, residual deviation = range_std
    sc5 = rb.scenes.config5(n_worlds=512, n_segments=100, n_rays=64)
    n5 = run(sc5, {"BRS_SOLO": "1"}, steps=1)
    c5 = run(sc5, {"BRS_SOLO": "1"}, steps=1, noise=False)
    f5 = run(sc5, {"BRS_SOLO": "0", "BRS_GROUP_MIN": "1"}, steps=1)
    assert np.array_equal(n5["range"], f5["range"])  # fused kernel, same samples
    res = (n5["range"] - c5["range"])[(n5["range"] > 0) & (n5["range"] < 200.0) & (c5["range"] < 199.0)]
    assert res.size > 10000 and abs(res.mean()) < 0.05 and abs(res.std() - rstd) < 0.05, (res.size, res.mean(), res.std())


def test_world_json_round_trip_on_the_gpu():
    """World.to_json -> World.from_json rebuilds a world whose steps equal the original's."""
    w = rb.World(220, 180)
    w.add_wall("blue", 50, 40, 120, 45)
    w.add_wall("red", 160, 30, 170, 150)
    w.add_bulb("white", 100, 120, 0, 1.0)
    r = rb.Robot(60, 90, 30)
    r.add_device(rb.RangeSensor(position=(8, 0), a=0, max=80, width=20))
    r.add_device(rb.Camera(width=32, height=16))
    r.add_device(rb.LightSensor(position=(5, 0)))
    w.add_robot(r)
    r2 = rb.Robot(150, 60, 200)
    r2.add_device(rb.RangeSensor(position=(8, 0), a=0, max=60, width=0))
    w.add_robot(r2)
    r.move(0.7, 0.2)
    r2.move(0.4, -0.3)
    w2 = rb.World.from_json(w.to_json())
    list(w2)[0].move(0.7, 0.2)
    list(w2)[1].move(0.4, -0.3)
    for _ in range(12):
        w.step()
        w2.step()
    for a, b in zip(w, w2):
        assert (a.x, a.y, a.a, a.stalled) == (b.x, b.y, b.a, b.stalled)
        assert a[0].get_distance() == b[0].get_distance()
    assert np.array_equal(np.asarray(list(w)[0][1].take_picture()), np.asarray(list(w2)[0][1].take_picture()))


@pytest.mark.parametrize("maker", [
    lambda: rb.scenes.config3(n_worlds=70),
    lambda: _crowd_scene(13, 2, camera=False, extras=True),
    lambda: _crowd_scene(13, 5, camera=False, extras=True),
    lambda: _crowd_scene(9, 8, camera=False, extras=True),
    lambda: _crowd_scene(11, 3, camera=False),
])
def test_multi_path_bit_identical_and_matches_the_oracle(maker):
    """The solo multi path (one warp per world, several robots; brs_multi_kernel) against the
    fused kernel, bit for bit, over 40 steps of crowded rooms (robot-robot and robot-wall
    collisions, list-order replay, fans, lights behind robots, smell), and against the oracle."""
    scene = maker()
    steps = 40

    def multi():
        eng = rb.BatchEngine(scene)
        assert eng.launch_info()["path"].startswith("solo multi"), eng.launch_info()
        eng.close()
        return parity.run_engine(scene, steps)

    a = _with_env({"BRS_SOLO": "0", "BRS_WIDE": "0"}, lambda: parity.run_engine(scene, steps))
    for env in ({"BRS_MULTI": "1"}, {"BRS_MULTI": "1", "BRS_SOLO_INTERLEAVE": "0", "BRS_SOLO_CTAS": "8"}):
        b = _with_env(env, multi)
        for k in a:
            assert np.array_equal(a[k], b[k]), (k, env)
    assert a["stalled"].any() or scene.n_robots < 3
    stats = parity.compare_outputs(scene, b, steps, range(0, scene.n_worlds, 3), camera=False)
    assert stats["max_pose_err"] < 1e-9 and stats["max_rel_dist"] < 1e-5


def test_multi_path_step_flags():
    """Move-only and sense-only steps of the solo multi path equal the fused kernel's."""
    scene = _crowd_scene(20, 4, camera=False, extras=True)

    def run():
        eng = rb.BatchEngine(scene)
        outs = []
        for flags in (rb.STEP_SENSE, rb.STEP_MOVE, rb.STEP_MOVE, rb.STEP_SENSE, rb.STEP_ALL, rb.STEP_ALL):
            eng.step(flags=flags)
            eng.synchronize()
            outs.append({k: eng.download(k) for k in ("pose", "vel", "stalled", "range", "light", "smell")})
        eng.close()
        return outs

    a = _with_env({"BRS_SOLO": "0", "BRS_WIDE": "0"}, run)
    b = _with_env({"BRS_MULTI": "1"}, run)
    for x, y in zip(a, b):
        for k in x:
            assert np.array_equal(x[k], y[k]), k


def _fused_gather_worker(rank, ws, port, n_worlds):
    import os

    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), BRS_SOLO="1")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    from aitk_robots_b200.sharding import ShardedWorlds

    sw = ShardedWorlds(lambda first, count: rb.scenes.config4(n_worlds=count, first_world=first), n_worlds, device=rank)
    full = sw.gather_buffer("camera", fused=True)
    for _ in range(3):
        sw.step()
        got = sw.all_gather("camera")
    torch.cuda.synchronize()
    # the reference: every rank's own pictures, gathered by NCCL from the engines' private buffers
    mine = got[sw.first:sw.first + sw.count].clone()
    parts = [torch.empty_like(mine) for _ in range(ws)]
    dist.all_gather(parts, mine)
    assert torch.equal(torch.cat(parts, dim=0), full), "a rank does not hold every rank's pictures"
    assert int(full[..., 3].min()) == 255
    dist.barrier()
    sw.engine.close()
    dist.destroy_process_group()


def test_fused_picture_gather_two_gpus():
    """brs_bind_peer_outputs: the solo kernel stores its pictures into the other rank's gather
    buffer over NVLink; after the cross-rank barrier both ranks hold the whole batch.  Needs two
    GPUs with peer access (skipped on a one-GPU box; bench.py exercises the same path at N > 1)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_fused_gather_worker, args=(2, 29631, 64), nprocs=2, join=True)


def test_wide_fov_camera_has_no_cone():
    """A 200-degree camera cannot use the cone cull; results still match the oracle."""
    scene = rb.scenes.config2(n_worlds=8, cam_w=64, cam_h=32)
    for spec in scene.robots[0].devices:
        if isinstance(spec, rb.CameraSpec):
            spec.a = 200.0
    stats = parity.compare(scene, n_steps=2)
    assert stats["pixel_mismatch"] <= 0.002 * stats["pixels"]


def test_full_size_properties_config4_shard():
    """One GPU's shard of config 4 at a reduced but tile-rich size: idempotence, batch
    independence and opacity of every picture."""
    scene = rb.scenes.config4(n_worlds=32768)
    eng = rb.BatchEngine(scene)
    eng.step()
    eng.synchronize()
    cam1, pose1, st1 = eng.download("camera"), eng.download("pose"), eng.download("stalled")
    eng.step(flags=rb.STEP_CAMERA)
    eng.synchronize()
    assert np.array_equal(cam1, eng.download("camera")) and np.array_equal(pose1, eng.download("pose"))
    assert (cam1[..., 3] == 255).all()  # the boundary wall encloses every robot
    for i in (0, 12345, 32767):
        alone = parity.run_engine(scene.slice(i, 1), 1)
        assert np.array_equal(alone["camera"][0], cam1[i]) and np.array_equal(alone["stalled"][0], st1[i])
    # oracle on a handful of worlds of the big batch
    from oracle import scene_adapter as sa

    for i in (7, 20000):
        ref = sa.step_and_observe(scene, i, 1, camera=True)
        assert list(st1[i]) == ref["stalled"]
        diff = np.any(cam1[i, 0] != np.array(ref["camera"][0], dtype=np.uint8), axis=2)
        assert diff.any(axis=0).sum() <= 2  # at most a couple of edge-grazing columns
    eng.close()


def test_many_steps_stalled_robots_config4():
    """Robots run into maze walls and stall; flags and poses follow the oracle exactly."""
    scene = rb.scenes.config4(n_worlds=16)
    scene.target_vel[:, :, 0] = 2.0  # full speed ahead
    stats = parity.compare(scene, n_steps=60, camera=False)
    assert stats["max_pose_err"] < 1e-9


def _crowd_scene(n_worlds, n_robots, seed=11, camera=True, extras=False):
    """Many robots in a small room: robot-robot collisions are frequent (exercises the
    list-order replay on the pair masks up to BRS_MAX_ROBOTS)."""
    import math

    g = rb.scenes
    rng = g._rng(seed, 0)
    w, h, B = 120.0, 100.0, n_worlds
    bx = np.array(g._boundary(w, h), dtype=np.float32)
    segs = [np.tile(bx[None, :, k], (B, 1)) for k in range(4)]
    colors = np.tile(np.array([[(128, 0, 128, 255)] * 4], dtype=np.uint8), (B, 1, 1))
    robots = []
    for k in range(n_robots):
        devs = [g.RangeSpec(position=(8.0, 0.0), a=0.0, max=60.0, width=20.0 if k % 2 else 0.0)]
        if k == 0 and camera:
            devs.append(g.CameraSpec(width=32, height=16, a=90.0))
        if extras:
            devs.append(g.LightSpec(position=(4.0, 1.0 * k)))
            if k % 3 == 0:
                devs.append(g.SmellSpec(position=(3.0, 0.0)))
                devs.append(g.RangeSpec(position=(-5.0, 2.0), a=170.0, max=45.0, width=0.0))
        robots.append(g.RobotSpec(rgba=tuple(int(c) for c in g.PALETTE[k % 8]), height=0.3 + 0.05 * k, devices=devs))
    bulbs = food = None
    if extras:
        bulbs = np.zeros((B, 3, 4), dtype=np.float32)
        bulbs[:, :, 0] = rng.uniform(10, w - 10, (B, 3))
        bulbs[:, :, 1] = rng.uniform(10, h - 10, (B, 3))
        bulbs[:, :, 3] = rng.uniform(0.5, 2.0, (B, 3))
        food = np.stack([rng.uniform(10, w - 10, (B, 1)), rng.uniform(10, h - 10, (B, 1)), np.full((B, 1), 20.0)], axis=2)
    scene = g._assemble("crowd", w, h, segs, colors, robots, rng, B, bulbs=bulbs, food=food,
                        smell_cell=10.0 if extras else 0.0)
    scene.target_vel[:, :, 0] = rng.uniform(1.0, 2.0, (B, n_robots))
    return scene


@pytest.mark.parametrize("n_robots", [2, 5, 8])
def test_crowded_rooms_up_to_max_robots(n_robots):
    scene = _crowd_scene(12, n_robots)
    stats = parity.compare(scene, n_steps=40, camera=True)
    assert stats["max_pose_err"] < 1e-9
    # the rooms are crowded enough that robots do collide with each other
    out = parity.run_engine(scene, 40)
    assert out["stalled"].any()


def test_vec_env_matches_oracle_and_auto_resets():
    """The gym-style wrapper: device-resident actions in, zero-copy observations out; the
    trajectory equals the oracle driven with the same Robot.move() calls."""
    import torch

    from oracle import scene_adapter as sa

    scene = rb.scenes.config2(n_worlds=16, cam_w=32, cam_h=16)
    env = rb.VecEnv(scene, max_steps=4)
    obs = env.reset()
    assert obs["camera"].is_cuda and obs["camera"].shape == (16, 1, 16, 32, 4)
    gen = torch.Generator().manual_seed(0)
    acts = [torch.rand((16, 1, 2), generator=gen) * 2 - 1 for _ in range(3)]
    worlds = {i: sa.build_world(scene, i) for i in (0, 9)}
    for a in acts:
        obs, reward, done, info = env.step(a.cuda())
        assert not bool(done.any()) and reward.shape == (16,)
        for i, (w, devs) in worlds.items():
            w._robots[0].move(float(a[i, 0, 0]), float(a[i, 0, 1]))
            w.step()
            ref = sa.observe(w, devs, camera=False)
            assert abs(float(env.pose[i, 0, 0]) - ref["pose"][0][0]) < 1e-9
            assert int(env.stalled[i, 0]) == ref["stalled"][0]
            np.testing.assert_allclose(obs["range"][i].cpu().numpy(), ref["range"], rtol=1e-4)
    # 4th step ends every episode: worlds restart from their initial poses with fresh observations
    obs, reward, done, info = env.step(acts[0].cuda())
    assert bool(done.all()) and int(env.steps.max()) == 0
    assert torch.allclose(env.pose.cpu(), torch.as_tensor(scene.pose))
    first = rb.VecEnv(scene, max_steps=4)
    assert torch.equal(first.reset()["range"].cpu(), obs["range"].cpu())
    first.close()
    env.close()


def _random_scene(seed):
    """Random worlds with random device sets (hypothesis drives the seed): wall counts 0..40,
    1..3 robots, range sensors with random mounts / widths / ranges, cameras with random
    field of view (also > 180 degrees) and size, lights, smell."""
    import math

    g = rb.scenes
    rng = np.random.Generator(np.random.PCG64(seed))
    B = int(rng.integers(1, 7))
    w, h = float(rng.integers(120, 400)), float(rng.integers(100, 300))
    n_in = int(rng.integers(0, 37))
    bx = np.array(g._boundary(w, h), dtype=np.float32)
    r = g._rand_segments(rng, B, n_in, w, h, min_len=5.0, max_len=0.4 * min(w, h))
    segs = [np.concatenate([np.tile(bx[None, :, k], (B, 1)), r[k]], axis=1) for k in range(4)]
    colors = np.concatenate([np.tile(np.array([[(128, 0, 128, 255)] * 4], dtype=np.uint8), (B, 1, 1)),
                             g.PALETTE[rng.integers(0, 8, (B, n_in))]], axis=1)
    R = int(rng.integers(1, 4))
    cam_w, cam_h = int(rng.choice([8, 32, 64, 100])), int(rng.choice([6, 16, 31]))
    robots = []
    for k in range(R):
        devs = []
        shared = (float(np.float32(rng.uniform(-5, 8))), float(np.float32(rng.uniform(-5, 5))))
        for i in range(int(rng.integers(0, 6))):
            pos = shared if rng.random() < 0.5 else (float(np.float32(rng.uniform(-8, 8))), float(np.float32(rng.uniform(-6, 6))))
            devs.append(g.RangeSpec(position=pos, a=float(rng.uniform(-180, 180)), max=float(rng.uniform(5, 250)),
                                    width=float(rng.choice([0.0, 0.0, 10.0, 57.3])), name="r%d" % i))
        if rng.random() < 0.6:
            devs.append(g.CameraSpec(width=cam_w, height=cam_h, a=float(rng.choice([30.0, 60.0, 120.0, 200.0, 350.0])),
                                     position=shared if rng.random() < 0.5 else (0.0, 0.0),
                                     max_range=float(rng.choice([1000.0, 150.0]))))
        if rng.random() < 0.5:
            devs.append(g.LightSpec(position=(float(np.float32(rng.uniform(-6, 6))), 0.0)))
        if rng.random() < 0.4:
            devs.append(g.SmellSpec(position=(float(np.float32(rng.uniform(-6, 6))), 0.0)))
        robots.append(g.RobotSpec(rgba=tuple(int(c) for c in g.PALETTE[k]), height=float(rng.uniform(0.1, 0.9)),
                                  devices=devs))
    if not any(r_.devices for r_ in robots):
        robots[0].devices.append(g.RangeSpec(position=(8.0, 0.0), a=0.0, max=100.0, width=0.0))
    bulbs = np.zeros((B, 2, 4), dtype=np.float32)
    bulbs[:, :, 0] = rng.uniform(5, w - 5, (B, 2))
    bulbs[:, :, 1] = rng.uniform(5, h - 5, (B, 2))
    bulbs[:, :, 3] = rng.uniform(0.5, 2.0, (B, 2))
    food = np.zeros((B, 1, 3), dtype=np.float32)
    food[:, :, 0], food[:, :, 1], food[:, :, 2] = rng.uniform(5, w - 5, (B, 1)), rng.uniform(5, h - 5, (B, 1)), 30.0
    scene = g._assemble("random", w, h, segs, colors, robots, rng, B, bulbs=bulbs, food=food, smell_cell=10.0,
                        tries=2000, clear_extra=1.0)
    # ragged worlds
    scene.seg_count[:] = rng.integers(0, 4 + n_in + 1, B)
    return scene


def test_random_scenes_match_the_oracle():
    """Property test: whatever the device mix, wall count and field of view, every output of
    the engine matches the oracle under the contract's tolerances."""
    from hypothesis import given, settings, HealthCheck
    from hypothesis import strategies as st

    counts = {"ran": 0, "unplaceable": 0}

    @settings(max_examples=100, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(st.integers(0, 2 ** 31 - 1))
    def run(seed):
        try:
            scene = _random_scene(seed)
        except RuntimeError:  # could not place the robots in a cluttered draw
            counts["unplaceable"] += 1
            return
        counts["ran"] += 1
        parity.compare(scene, n_steps=3)

    run()
    # the draws that could not be placed are counted, not silently passed
    assert counts["ran"] >= 90 and counts["unplaceable"] <= 10, counts


def test_soak_1000_random_scenes():
    """The round-1 soak as a test: 1000 seeded random scenes (1-6 worlds each, every device
    mix) x 4 steps against the oracle; the number of unplaceable draws is asserted and hit /
    no-hit flips (only ever excused by the end-point margin) are counted."""
    ok, skipped = 0, 0
    tot = {"worlds": 0, "ranges": 0, "range_hit_flips": 0, "pixels": 0, "pixel_mismatch": 0, "lights": 0, "light_flips": 0}
    for seed in range(20_000, 21_000):
        try:
            scene = _random_scene(seed)
        except RuntimeError:
            skipped += 1
            continue
        st = parity.compare(scene, n_steps=4)
        for k in tot:
            tot[k] += st[k]
        ok += 1
    assert ok + skipped == 1000 and skipped <= 50, (ok, skipped)
    assert tot["worlds"] >= 2000 and tot["ranges"] > 5000 and tot["pixels"] > 1_000_000, tot
    assert tot["range_hit_flips"] <= 1e-3 * tot["ranges"] and tot["pixel_mismatch"] <= 1e-4 * tot["pixels"], tot
    print("soak:", ok, "scenes,", skipped, "unplaceable,", tot)


@pytest.mark.parametrize("config,n_worlds,steps", [("config2", 4096, 2), ("config3", 65536, 5), ("config4", 131072, 3)])
def test_full_size_batches_oracle_on_64_random_worlds(config, n_worlds, steps):
    """BASELINE.json sizes: the whole batch is stepped on the GPU, 64 randomly drawn worlds of
    it are stepped by the oracle and every output of those worlds is compared."""
    scene = getattr(rb.scenes, config)(n_worlds=n_worlds)
    gpu = parity.run_engine(scene, steps)
    rng = np.random.Generator(np.random.PCG64(99))
    worlds = sorted(int(i) for i in rng.choice(n_worlds, size=64, replace=False))
    stats = parity.compare_outputs(scene, gpu, steps, worlds, camera=True)
    assert stats["worlds"] == 64 and stats["max_pose_err"] < 1e-9
    assert stats["pixel_mismatch"] <= 0.002 * max(stats["pixels"], 1)
    assert stats["range_hit_flips"] <= 0.002 * max(stats["ranges"], 1)


@pytest.mark.parametrize("S", [10, 30, 100, 300, 1000, 2000])
def test_config5_every_sweep_cell(S):
    """All 30 cells of the config 5 sweep (6 wall counts x 5 ray counts): oracle on two worlds
    of every cell."""
    B = 8 if S < 1000 else 4
    base = rb.scenes.config5(n_worlds=B, n_segments=S, n_rays=256)
    for n in (1, 4, 16, 64, 256):
        base.robots[0].devices = [rb.RangeSpec(position=(0.0, 0.0), a=360.0 * i / n, max=200.0, width=0.0,
                                               name="ray%d" % i) for i in range(n)]
        stats = parity.compare(base, n_steps=2, worlds=(0, B - 1), camera=False)
        assert stats["ranges"] == 2 * n and stats["max_rel_dist"] < 1e-5, (S, n, stats)


def test_very_wide_camera_and_sky_gradient():
    """Pictures wider than 4 x the worker count (several column quads per thread) and the
    non-flat sky / ground rows (generic paint loop)."""
    from oracle import aitk_oracle as orc
    from oracle import scene_adapter as sa

    scene = rb.scenes.config2(n_worlds=6, cam_w=640, cam_h=9)
    stats = parity.compare(scene, n_steps=2)
    assert stats["pixel_mismatch"] <= 0.002 * stats["pixels"]
    # vertical gradients: colour = base + grad * |j - H/2| / (H/2)
    scene = rb.scenes.config2(n_worlds=4, cam_w=64, cam_h=32)
    sky_grad, gnd_grad = (40, 20, 100), (30, 60, 10)
    old = (orc.SW.CAMERA_SKY_GRAD, orc.SW.CAMERA_GROUND_GRAD)
    orc.SW.CAMERA_SKY_GRAD, orc.SW.CAMERA_GROUND_GRAD = sky_grad, gnd_grad
    try:
        eng = rb.BatchEngine(scene, sky_grad=sky_grad, ground_grad=gnd_grad)
        eng.step()
        eng.synchronize()
        cam = eng.download("camera")
        eng.close()
        for i in range(4):
            ref = np.array(sa.step_and_observe(scene, i, 1)["camera"][0], dtype=np.uint8)
            diff = np.any(cam[i, 0] != ref, axis=2)
            assert diff.any(axis=0).sum() <= 1, i
        assert len(np.unique(cam[0, 0, :3, :, 2])) > 1  # the sky really is a gradient
    finally:
        orc.SW.CAMERA_SKY_GRAD, orc.SW.CAMERA_GROUND_GRAD = old


@pytest.mark.parametrize("maker,size", [
    (lambda: rb.scenes.demo_scene(6), (240, 180)),
    (lambda: rb.scenes.config4(n_worlds=5), (200, 117)),
    (lambda: rb.scenes.config3(n_worlds=4), (97, 64)),
])
def test_topdown_renderer_matches_the_restatement(maker, size):
    """brs_render_world (whole-world view and the GroundCamera window under a robot) against
    oracle/topdown.py, pixel for pixel, after the robots have moved."""
    from oracle import topdown

    scene = maker()
    eng = rb.BatchEngine(scene)
    for _ in range(3):
        eng.step()
    poses = eng.download("pose")
    W, H = size
    img = eng.render_world(0, None, W, H).cpu().numpy()
    assert img.shape == (scene.n_worlds, H, W, 4)
    gv = eng.render_world(1, scene.n_worlds - 1, 40, 30, robot=0, view=(60.0, 45.0)).cpu().numpy()
    total = bad = fragile_n = 0
    for i in range(scene.n_worlds):
        ref, fragile = topdown.render_scene_world(scene, i, poses[i], img_w=W, img_h=H)
        diff = (img[i] != ref).any(axis=-1)
        assert not (diff & ~fragile).any(), (i, np.argwhere(diff & ~fragile)[:5])
        total, bad, fragile_n = total + diff.size, bad + int(diff.sum()), fragile_n + int(fragile.sum())
        assert len(np.unique(ref.reshape(-1, 4), axis=0)) >= 3  # ground, walls, a robot: not a blank picture
        if i >= 1:
            ref, fragile = topdown.render_scene_world(scene, i, poses[i], img_w=40, img_h=30, robot=0, view=(60.0, 45.0))
            diff = (gv[i - 1] != ref).any(axis=-1)
            assert not (diff & ~fragile).any(), (i, np.argwhere(diff & ~fragile)[:5])
            total, bad, fragile_n = total + diff.size, bad + int(diff.sum()), fragile_n + int(fragile.sum())
    assert bad <= 1e-3 * total and fragile_n <= 0.02 * total, (bad, fragile_n, total)
    eng.close()


def test_facade_world_picture_and_ground_camera():
    w = rb.World(200, 100)
    w.add_wall("blue", 50, 20, 60, 80)
    r = rb.Robot(100, 50, 0, color="red")
    g = r.add_device(rb.GroundCamera(width=20, height=20, view=(40, 40)))
    r.add_device(rb.RangeSensor(position=(8, 0), a=0, max=50, width=0))
    w.add_robot(r)
    w.step()
    pic = w.take_picture(as_array=True)
    assert pic.shape == (100, 200, 4)
    assert tuple(pic[50, 100]) == (255, 0, 0, 255)   # the robot
    assert tuple(pic[50, 55]) == (0, 128, 0, 255)    # inside the box wall: ground
    assert tuple(pic[50, 50]) == (0, 0, 255, 255)    # its left edge
    assert tuple(pic[5, 150]) == (0, 128, 0, 255)
    gp = g.take_picture(as_array=True)
    assert gp.shape == (20, 20, 4) and (gp.reshape(-1, 4) == (0, 128, 0, 255)).all()  # robot not in its own view
    r.set_pose(72, 50, 180)  # now facing the wall's right edge at x = 60: 12 units ahead
    gp = g.take_picture(as_array=True)
    assert tuple(gp[3, 10]) == (0, 0, 255, 255) and tuple(gp[15, 10]) == (0, 128, 0, 255)
    w2 = rb.World.from_json(w.to_json())
    assert isinstance(list(w2)[0][0], rb.GroundCamera) and list(w2)[0][0].view == (40.0, 40.0)


@pytest.mark.parametrize("maker,env,plan", [
    (lambda: rb.scenes.config1(n_worlds=1, seed=1), {"BRS_MINI": "0"}, "fused ("),
    (lambda: rb.scenes.config1(n_worlds=1, seed=1), {}, "mini range"),
    (lambda: rb.scenes.config5(n_worlds=5000, n_segments=30, n_rays=4), {}, "mini range"),
    (lambda: rb.scenes.config5(n_worlds=70, n_segments=200, n_rays=2), {"_N": "40"}, "mini range"),
    (lambda: rb.scenes.config1(n_worlds=700, seed=1), {"BRS_SOLO": "0"}, "fused ("),
    (lambda: rb.scenes.config2(n_worlds=1500, cam_w=64, cam_h=32), {}, "fused ("),
    (lambda: rb.scenes.config2(n_worlds=96), {"BRS_TW": "2"}, "fused ("),
    (lambda: rb.scenes.config3(n_worlds=9000), {}, "fused ("),
    (lambda: rb.scenes.demo_scene(50), {"BRS_BIG": "1"}, "fused ("),
    (lambda: rb.scenes.config4(n_worlds=9000), {"BRS_SOLO": "1"}, "solo ("),
    (lambda: rb.scenes.config4(n_worlds=40), {"BRS_SOLO": "1"}, "solo ("),
    (lambda: rb.scenes.config5(n_worlds=5000, n_segments=100, n_rays=16), {"BRS_SOLO": "1"}, "solo range"),
    (lambda: rb.scenes.config5(n_worlds=33, n_segments=300, n_rays=64), {"BRS_SOLO": "1"}, "solo range"),
    # many steps over few tiles: every step of a tile lands on another CTA, the epoch waits are exercised
    (lambda: rb.scenes.config2(n_worlds=600, cam_w=32, cam_h=16), {"_N": "150"}, "fused ("),
    (lambda: rb.scenes.config3(n_worlds=3000), {"_N": "60", "BRS_TW": "4"}, "fused ("),
    (lambda: rb.scenes.demo_scene(700), {"_N": "40"}, "fused ("),
])
def test_steps_in_one_launch_equal_single_steps(maker, env, plan):
    """brs_steps (World.steps(n): one launch, tiles / worlds statically owned, pipelines running
    on across the step boundaries) leaves every buffer bit-identical to n calls of brs_step --
    on every plan, for batches smaller and larger than the grid, for move-only steps too."""
    scene = maker()
    env = dict(env)
    n_steps = int(env.pop("_N", "7"))

    def run(how, flags=rb.STEP_ALL, n=n_steps):
        eng = rb.BatchEngine(scene)
        assert eng.launch_info()["path"].startswith(plan), eng.launch_info()
        eng.step()  # (a first step so that velocities have ramped and some robots have stalled)
        if how == "single":
            for _ in range(n):
                eng.step(flags=flags)
        else:
            eng.steps(n, flags=flags)
        eng.synchronize()
        out = {k: eng.download(k) for k in ("pose", "vel", "stalled", "range", "light", "smell")}
        if eng.n_camera:
            out["camera"] = eng.download("camera")
        eng.close()
        return out

    for flags in (rb.STEP_ALL, rb.STEP_MOVE):
        a = _with_env(env, lambda: run("single", flags))
        b = _with_env(env, lambda: run("steps", flags))
        c = _with_env(dict(env, BRS_STEPS_ONE_LAUNCH="0"), lambda: run("steps", flags))
        for k in a:
            assert np.array_equal(a[k], b[k]), (k, flags)
            assert np.array_equal(a[k], c[k]), (k, flags, "fallback")
    assert np.abs(a["pose"] - scene.pose).max() > 0.5  # they did move


def test_facade_steps_is_one_engine_call():
    """World.steps(n) == n x World.step() on the facade (config 1 as BASELINE.json writes it)."""
    def make():
        w = rb.World(200, 150)
        w.add_wall("blue", 40, 30, 80, 35)
        r = rb.Robot(100, 75, 0)
        r.add_device(rb.RangeSensor(position=(8, 0), a=0, max=100, width=0))
        w.add_robot(r)
        r.move(0.5, 0.2)
        return w, r

    w1, r1 = make()
    w2, r2 = make()
    for _ in range(100):
        w1.step()
    w2.steps(100)
    assert r1.get_pose() == r2.get_pose() and r1.stalled == r2.stalled
    assert r1[0].get_distance() == r2[0].get_distance()
    assert w1.time == w2.time == 10.0


def test_vec_env_frame_skip_equals_repeated_actions():
    """VecEnv(frame_skip=k): one env step == k env steps of a plain VecEnv with the same action."""
    import torch

    scene = rb.scenes.demo_scene(40)
    a = rb.VecEnv(scene, max_steps=10**6, frame_skip=4)
    b = rb.VecEnv(scene, max_steps=10**6)
    a.reset()
    b.reset()
    g = torch.Generator().manual_seed(5)
    for _ in range(6):
        act = torch.rand((scene.n_worlds, scene.n_robots, 2), generator=g) * 2 - 1
        oa, _, _, ia = a.step(act)
        for _ in range(4):
            ob, _, _, ib = b.step(act)
        assert torch.equal(a.pose, b.pose) and torch.equal(ia["stalled"], ib["stalled"])
        for k in oa:
            assert torch.equal(oa[k], ob[k]), k
    assert int(a.steps[0]) == 24
    a.close()
    b.close()


@pytest.mark.parametrize("S,rays", [(12, 1), (12, 7), (10, 16), (30, 4), (100, 16), (100, 3), (300, 1), (2000, 2)])
def test_mini_path_bit_identical_and_matches_the_oracle(S, rays):
    """The mini range path (a team of 2..16 lanes per world, everything in registers;
    brs_mini_kernel) against the fused kernel searching the same shared-origin table
    (BRS_GROUP_MIN=1) and the solo range path, bit for bit, with every team size, and against the
    oracle; ragged / empty worlds, robots driven into walls, batches that do not fill a warp."""
    B = 37 if S < 1000 else 9
    scene = rb.scenes.config5(n_worlds=B, n_segments=S, n_rays=rays)
    scene.seg_count[1] = 0
    scene.seg_count[2] = min(S, 5)
    scene.target_vel[:, :, 0] = np.linspace(-2.0, 2.0, B)[:, None]

    def mini():
        eng = rb.BatchEngine(scene)
        assert eng.launch_info()["path"].startswith("mini range"), eng.launch_info()
        eng.close()
        return parity.run_engine(scene, 6)

    a = _with_env({"BRS_SOLO": "0", "BRS_WIDE": "0", "BRS_GROUP_MIN": "1"}, lambda: parity.run_engine(scene, 6))
    s = _with_env({"BRS_SOLO": "1", "BRS_MINI": "0"}, lambda: parity.run_engine(scene, 6))
    for env in ({"BRS_MINI": "1"}, {"BRS_MINI": "1", "BRS_MINI_G": "2", "BRS_NO_CONE": "1"},
                {"BRS_MINI": "1", "BRS_MINI_G": "8"}, {"BRS_MINI": "1", "BRS_MINI_G": "16", "BRS_NO_RANGE_CULL": "1"},
                {"BRS_MINI": "1", "BRS_MINI_G": "4", "BRS_SOLO_MAX_CTAS": "1"}):
        b = _with_env(env, mini)
        for k in a:
            assert np.array_equal(a[k], b[k]), (k, env)
            assert np.array_equal(s[k], b[k]), (k, env, "vs solo range")
    stats = parity.compare_outputs(scene, b, 6, range(0, B, 4), camera=False)
    assert stats["max_rel_dist"] < 1e-5 and stats["max_pose_err"] < 1e-9


def test_mini_path_fans_config1_flags_noise_and_steps():
    """Fans of three rays per sensor, an off-centre mount and config 1 on the mini range path;
    sense-only / move-only steps; seeded noise; brs_steps in one launch."""
    for scene in (_range_fan_scene(21, 40, 5, 30.0), rb.scenes.config1(n_worlds=50), rb.scenes.config1(n_worlds=3000)):
        a = _with_env({"BRS_SOLO": "0", "BRS_WIDE": "0", "BRS_GROUP_MIN": "1"}, lambda: parity.run_engine(scene, 5))

        def mini(sc=scene):
            eng = rb.BatchEngine(sc)
            assert eng.launch_info()["path"].startswith("mini range"), eng.launch_info()
            eng.close()
            return parity.run_engine(sc, 5)

        b = _with_env({"BRS_MINI": "1"}, mini)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        stats = parity.compare_outputs(scene, b, 5, range(0, scene.n_worlds, max(1, scene.n_worlds // 40)), camera=False)
        assert stats["max_rel_dist"] < 1e-5
    scene = rb.scenes.config5(n_worlds=700, n_segments=30, n_rays=4)

    def run(env, how):
        def go():
            eng = rb.BatchEngine(scene)
            if how == "noise":
                eng.set_noise(99, 0.5, 0.0)
            eng.step()
            if how == "flags":
                eng.step(flags=rb.STEP_MOVE)
                eng.step(flags=rb.STEP_SENSE)
                eng.step(flags=rb.STEP_MOVE)
            elif how == "steps":
                eng.steps(9)
                assert eng.steps_launches(9) == 1
            elif how == "single":
                for _ in range(9):
                    eng.step()
            else:
                eng.step()
            out = {k: eng.download(k) for k in ("pose", "vel", "stalled", "range")}
            eng.close()
            return out
        return _with_env(env, go)

    fused = {"BRS_SOLO": "0", "BRS_WIDE": "0", "BRS_GROUP_MIN": "1"}
    for how in ("flags", "noise"):
        a, b = run(fused, how), run({"BRS_MINI": "1"}, how)
        for k in a:
            assert np.array_equal(a[k], b[k]), (how, k)
    a, b = run({"BRS_MINI": "1"}, "single"), run({"BRS_MINI": "1"}, "steps")
    for k in a:
        assert np.array_equal(a[k], b[k]), ("steps", k)


def test_steps_epoch_protocol_soak():
    """Stress of brs_steps on the fused kernel where the epoch words matter: batches of a few
    more tiles than CTAs (so that consecutive steps of a tile keep landing on other CTAs, some
    of them while the previous step's observations are still being written), random step
    counts and tile sizes, 40 cases each run four times -- always bit-identical to single steps."""
    rng = np.random.Generator(np.random.PCG64(2024))
    makers = [
        lambda n, s: rb.scenes.demo_scene(n, seed=s),
        lambda n, s: rb.scenes.config2(n_worlds=n, seed=s, cam_w=32, cam_h=16),
        lambda n, s: rb.scenes.config3(n_worlds=n, seed=s),
        lambda n, s: rb.scenes.config5(n_worlds=n, n_segments=40, n_rays=6, seed=s),
        lambda n, s: rb.scenes.config4(n_worlds=n, seed=s, cam_w=32, cam_h=16),
    ]
    cases = 0
    for it in range(40):
        mk = makers[it % len(makers)]
        tw = int(rng.choice([1, 1, 2, 3]))
        n = int(rng.integers(600, 1400)) * tw
        k = int(rng.integers(5, 40))
        scene = mk(n, int(rng.integers(1, 10 ** 6)))
        env = {"BRS_SOLO": "0", "BRS_WIDE": "0", "BRS_STEPS_TW": str(tw)}

        def run(how):
            eng = rb.BatchEngine(scene)
            assert eng.launch_info()["path"].startswith("fused ("), eng.launch_info()
            if how == "single":
                for _ in range(k):
                    eng.step()
            else:
                assert eng.steps_launches(k) == 1 and eng.steps_info()["tiles"] > eng.steps_info()["ctas"]
                eng.steps(k)
            out = {b: eng.download(b) for b in ("pose", "vel", "stalled", "range", "light", "smell")}
            if eng.n_camera:
                out["camera"] = eng.download("camera")
            eng.close()
            return out

        a = _with_env(env, lambda: run("single"))
        for rep in range(4):
            b = _with_env(env, lambda: run("steps"))
            for key in a:
                assert np.array_equal(a[key], b[key]), (it, rep, key, n, k, tw)
        cases += 1
    assert cases == 40


def test_facade_run_and_reset():
    w = rb.World(200, 150)
    r = rb.Robot(50, 75, 0)
    r.add_device(rb.RangeSensor(position=(8, 0), a=0, max=100, width=0))
    w.add_robot(r)
    r.move(1.0, 0.0)
    first = None

    def controller(world):
        return world[0][0].get_distance() < 30.0  # stop 30 units from the wall ahead

    n = w.run(controller)
    assert n > 10 and r[0].get_distance() < 30.0 and r.x > 100 and w.time == round(n * 0.1, 10)
    first = (r.x, r.y, r.a, n)
    w.reset()
    assert (r.x, r.y, r.a) == (50.0, 75.0, 0.0) and w.time == 0.0 and not r.stalled
    r.move(1.0, 0.0)
    assert w.run(controller) == n and (r.x, r.y, r.a, n) == first  # the same run again, bit for bit
    assert w.run(max_steps=5) == 5 and w.get_robot(0) is r
