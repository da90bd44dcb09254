# -*- coding: utf-8 -*-
"""aitk.robots_b200 -- B200-native batched stepping engine for the aitk.robots
``World.step`` hot path (motion, collision, ray-cast sensors).

Drop-in names: World, Robot, RangeSensor, Camera, LightSensor, SmellSensor,
Bulb (single-world facade over a batch of one) plus BatchEngine / SceneBatch
for large batches.  Importable as ``aitk_robots_b200`` (the directory name
contains a dot; see ../aitk_robots_b200.py).
"""

from . import _cabi  # noqa: F401
from .engine import BatchEngine, measure_fp32_peak, measure_write_bw  # noqa: F401
from .scenes import (  # noqa: F401
    CameraSpec,
    LightSpec,
    RangeSpec,
    RobotSpec,
    SceneBatch,
    SmellSpec,
)
from . import scenes  # noqa: F401
from .world import Bulb, Camera, GroundCamera, LightSensor, RangeSensor, Robot, SmellSensor, World  # noqa: F401
from .sharding import ShardedWorlds  # noqa: F401
from .vec_env import VecEnv  # noqa: F401

STEP_MOVE, STEP_SENSE, STEP_CAMERA, STEP_ALL = _cabi.STEP_MOVE, _cabi.STEP_SENSE, _cabi.STEP_CAMERA, _cabi.STEP_ALL
__version__ = "0.1.0"
