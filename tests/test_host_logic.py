# -*- coding: utf-8 -*-
"""Host-side logic on CPU: scene generators, test accounting, sharding (gloo)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import aitk_robots_b200 as rb
from aitk_robots_b200.sharding import all_gather_worlds, partition


def test_generators_shapes_and_float32():
    for gen, kw in ((rb.scenes.config1, dict(n_worlds=3)), (rb.scenes.config2, dict(n_worlds=5)),
                    (rb.scenes.config3, dict(n_worlds=4)), (rb.scenes.config4, dict(n_worlds=6)),
                    (rb.scenes.config4, dict(n_worlds=3, variant="random")),
                    (rb.scenes.config5, dict(n_worlds=2, n_segments=300, n_rays=4)),
                    (rb.scenes.demo_scene, dict(n_worlds=2))):
        s = gen(**kw)
        B = kw["n_worlds"]
        assert s.seg.dtype == np.float32 and s.seg.shape[:2] == (B, 4) and s.seg_cap % 4 == 0
        assert s.pose.dtype == np.float64 and s.pose.shape == (B, s.n_robots, 3)
        assert (s.seg_count <= s.seg_cap).all()
        # poses are float32-representable in x, y (identical inputs for the float64 reference)
        assert (s.pose[:, :, :2] == s.pose[:, :, :2].astype(np.float32)).all()


def test_generators_deterministic_and_shardable():
    a = rb.scenes.config2(n_worlds=8, first_world=0)
    b = rb.scenes.config2(n_worlds=8, first_world=0)
    assert np.array_equal(a.seg, b.seg) and np.array_equal(a.pose, b.pose)
    c = rb.scenes.config2(n_worlds=8, first_world=8)
    assert not np.array_equal(a.seg, c.seg)
    sl = a.slice(2, 3)
    assert sl.n_worlds == 3 and np.array_equal(sl.seg, a.seg[2:5])


def test_config_shapes_named_in_baseline():
    s2 = rb.scenes.config2(n_worlds=2)
    assert int(s2.seg_count[0]) == 20 and len(s2.devices(rb.RangeSpec)) == 2
    cam = s2.devices(rb.CameraSpec)[0][1]
    assert (cam.width, cam.height) == (256, 128)
    s3 = rb.scenes.config3(n_worlds=2)
    assert s3.n_robots == 4 and len(s3.devices(rb.RangeSpec)) == 64 and s3.bulbs.shape[1] == 2
    assert int(s3.seg_count[0]) == 12
    s4 = rb.scenes.config4(n_worlds=2)
    assert int(s4.seg_count[0]) == 100 and s4.devices(rb.CameraSpec)[0][1].width == 64
    s1 = rb.scenes.config1()
    assert tuple(s1.world_size[0]) == (200.0, 150.0) and s1.devices(rb.RangeSpec)[0][1].max == 100.0


def test_tests_per_step_formula():
    s = rb.scenes.config3(n_worlds=2)
    R, S = 4, 12
    rays = 4 * R + 64 + 4 * 2
    assert s.tests_per_step() == 2 * rays * (S + 4 * (R - 1))


def test_partition_covers_everything():
    for n in (1, 7, 8, 4096, 1000003):
        for ws in (1, 2, 3, 8):
            spans = [partition(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1


def _gloo_worker(rank, ws, port, n_worlds):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        counts = [partition(n_worlds, ws, r)[1] for r in range(ws)]
        first, count = partition(n_worlds, ws, rank)
        scene = rb.scenes.config2(n_worlds=n_worlds, cam_w=8, cam_h=4).slice(first, count)
        local = torch.from_numpy(scene.pose.copy())
        full = all_gather_worlds(local, counts)
        ref = torch.from_numpy(rb.scenes.config2(n_worlds=n_worlds, cam_w=8, cam_h=4).pose)
        assert torch.equal(full, ref), "gathered shards differ from the global batch"
        # equal shards take the single-collective path
        if len(set(counts)) == 1:
            obs = torch.full((count, 3), float(rank))
            g = all_gather_worlds(obs)
            assert g.shape[0] == n_worlds and float(g[-1, 0]) == ws - 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_worlds", [8, 7])
def test_sharded_gather_gloo_world_size_2(n_worlds):
    port = 29500 + (os.getpid() % 2000) + n_worlds
    mp.spawn(_gloo_worker, args=(2, port, n_worlds), nprocs=2, join=True)


def test_world_json_round_trip(tmp_path):
    """World.to_json / from_json: the scene survives a file round trip."""
    import json

    w = rb.World(200, 150, boundary_wall_color="blue")
    w.add_wall("red", 50, 20, 50, 120, box=False)
    w.add_wall("cyan", 120, 40, 140, 60, box=True)
    w.add_bulb("white", 30, 30, 0, 2.0)
    w.add_food(160, 100, 25)
    r = rb.Robot(100, 75, 45, color="yellow", name="A")
    r.add_device(rb.RangeSensor(position=(8, 0), a=10, max=80, width=5, name="ir"))
    r.add_device(rb.Camera(width=32, height=16, a=70))
    r.add_device(rb.LightSensor(position=(6, 0)))
    r.add_device(rb.SmellSensor(position=(6, 0)))
    w.add_robot(r)
    w.add_robot(rb.Robot(40, 100, -90))
    path = tmp_path / "scene.json"
    path.write_text(json.dumps(w.to_json()))
    w2 = rb.World.from_json(str(path))
    a, b = w.to_scene(2), w2.to_scene(2)
    assert np.array_equal(a.seg, b.seg) and np.array_equal(a.seg_rgba, b.seg_rgba)
    assert np.array_equal(a.seg_count, b.seg_count) and int(a.seg_count[0]) == 4 + 1 + 4
    assert np.allclose(a.pose, b.pose) and np.array_equal(a.bulbs, b.bulbs) and np.array_equal(a.grid, b.grid)
    assert [type(d).__name__ for d in w2._robots[0]._devices] == ["RangeSensor", "Camera", "LightSensor", "SmellSensor"]
    assert w2._robots[0]["ir"].spec == r["ir"].spec
