# -*- coding: utf-8 -*-
"""The C-ABI library loads and exports every symbol include/brs.h declares;
host-side argument checking.  No compute calls (CPU container)."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as ge
import aitk_robots_b200 as rb
from aitk_robots_b200 import _cabi as cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    ge.build()
    return cabi.load()


def test_header_symbols_all_exported(lib):
    text = open(os.path.join(ROOT, "include", "brs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(brs_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations parsed"
    bound = {name for name, _, _ in cabi.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_struct_sizes(lib):
    assert lib.brs_abi_version() == cabi.BRS_ABI_VERSION
    # layouts the header promises (LP64)
    assert C.sizeof(cabi.RobotDesc) == 72
    assert C.sizeof(cabi.RangeDesc) == 48
    assert C.sizeof(cabi.CameraDesc) == 64
    assert C.sizeof(cabi.PointDesc) == 24


def test_create_rejects_bad_config(lib):
    h = C.c_void_p()
    cfg = cabi.Config()
    cfg.abi_version = 999
    assert lib.brs_create(C.byref(cfg), C.byref(h)) == -1
    assert b"ABI" in lib.brs_last_error()
    cfg.abi_version = cabi.BRS_ABI_VERSION
    cfg.n_worlds = 0
    assert lib.brs_create(C.byref(cfg), C.byref(h)) == -1
    cfg.n_worlds, cfg.n_robots, cfg.seg_cap = 1, 1, 6
    assert lib.brs_create(C.byref(cfg), C.byref(h)) == -1
    assert b"seg_cap" in lib.brs_last_error()
    assert lib.brs_create(None, C.byref(h)) == -1


def test_no_cpu_fallback_without_device(lib):
    """Without a CUDA device the engine refuses to exist (no silent CPU path)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    scene = rb.scenes.demo_scene(2)
    with pytest.raises(cabi.BrsError):
        rb.BatchEngine(scene)
    world = rb.World(100, 100)
    world.add_robot(rb.Robot(50, 50, 0))
    with pytest.raises(cabi.BrsError):
        world.step()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(cabi.BrsError):
        cabi.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "aitk.robots_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
