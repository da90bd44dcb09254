# -*- coding: utf-8 -*-
"""CPU restatement of the top-down renderer (brs_render_world, include/brs.h).

TEST INFRASTRUCTURE ONLY: imported by tests/ (never by aitk.robots_b200/).

Upstream: the World's picture of itself (aitk/robots/world.py display / take_picture) and
the downward-looking GroundCamera (aitk/robots/devices/cameras.py) [RECALL names].  Neither
file is in the mounted reference tree, so the LOOK of the picture --
line thickness, bulb discs, robots as filled boxes -- is the engine's own definition; this
module restates that definition in numpy float32 (one IEEE operation per step, no
contraction), which is what the kernel computes with __fmul_rn / __fadd_rn / __fdiv_rn.
**Parity unpinned** against upstream; pinned against this restatement.

`render` also returns a mask of pixels whose inside/outside decision is within a few ulp of
flipping: only the robot boxes and the robot view's frame go through float64 sincos, where
the GPU's sincos and libm may differ in the last place, so only there can a pixel differ.
"""

import math

import numpy as np

F = np.float32


def _pack(c):
    return np.array([int(c[0]), int(c[1]), int(c[2]), 255], dtype=np.uint8)


def robot_corners(pose, cox, coy):
    """Four box corners (float64) of a robot at `pose`; corner = centre + R(a) (cox, coy)
    (Robot.compute_boundingbox of the restatement, oracle/aitk_oracle.py)."""
    x, y, a = (float(v) for v in pose)
    cs, sn = math.cos(a), math.sin(a)
    return [(x + (cs * ox - sn * oy), y + (sn * ox + cs * oy)) for ox, oy in zip(cox, coy)]


def corner_offsets(bbox):
    """bbox (min_x, max_x, min_y, max_y) -> corner offsets in the a = 0 frame, in the order
    and with the arithmetic of compute_boundingbox (dist * cos/sin(atan2(-x, y) + pi/2))."""
    min_x, max_x, min_y, max_y = bbox
    cox, coy = [], []
    for x, y in ((min_x, max_y), (min_x, min_y), (max_x, min_y), (max_x, max_y)):
        dist = math.sqrt(x * x + y * y)
        ang = math.atan2(-x, y)
        cox.append(dist * math.cos(ang + math.pi / 2))
        coy.append(dist * math.sin(ang + math.pi / 2))
    return cox, coy


def render(world_size, seg, seg_rgba, n_seg, poses, bboxes, robot_rgba, bulbs=None, ground=(0, 128, 0),
           img_w=256, img_h=128, robot=-1, view=None, line_px=1.5, bulb_px=2.5):
    """One world from above.  Returns (rgba uint8 [img_h, img_w, 4], fragile bool [img_h, img_w]).

    world_size (W, H); seg float32 [4, cap] rows x1 y1 x2 y2; seg_rgba uint8 [cap, 4];
    poses float64 [R, 3]; bboxes [R] of (min_x, max_x, min_y, max_y); robot_rgba [R] colours.
    """
    W, H = F(world_size[0]), F(world_size[1])
    cx = (np.arange(img_w, dtype=F) + F(0.5))[None, :]
    cy = (np.arange(img_h, dtype=F) + F(0.5))[:, None]
    if robot < 0:
        ox = oy = F(0)
        ux, uy, vx, vy = W / F(img_w), F(0), F(0), H / F(img_h)
        unit = max(ux, vy)
    else:
        X, Y, a = (float(v) for v in poses[robot])
        sn, cs = math.sin(a), math.cos(a)
        vw, vh = float(F(view[0])), float(F(view[1]))
        pu, pv = vw / img_w, vh / img_h
        fw, lt = (0.5 * img_h) * pv, (0.5 * img_w) * pu
        ox, oy = F(X + (cs * fw + sn * lt)), F(Y + (sn * fw - cs * lt))
        ux, uy, vx, vy = F(-(sn * pu)), F(cs * pu), F(-(cs * pv)), F(-(sn * pv))
        unit = max(F(view[0]) / F(img_w), F(view[1]) / F(img_h))
    x = ox + (cx * ux + cy * vx)
    y = oy + (cx * uy + cy * vy)
    assert x.dtype == F and y.dtype == F
    half = F(0.5 * line_px) * unit
    half2 = half * half
    br = F(bulb_px) * unit
    br2 = br * br
    img = np.empty((img_h, img_w, 4), dtype=np.uint8)
    img[:] = _pack(ground)
    fragile = np.zeros((img_h, img_w), dtype=bool)
    if robot >= 0:
        # the frame itself came through sincos: a pixel centre may move by an ulp or two
        eps = F(8) * np.spacing(np.abs(x) + np.abs(y)).astype(F) + F(1e-6)
    else:
        eps = F(0)
    if bulbs is not None:
        for b in np.asarray(bulbs, dtype=F):
            dx, dy = x - b[0], y - b[1]
            d2 = dx * dx + dy * dy
            img[d2 <= br2] = (255, 255, 0, 255)
            if robot >= 0:
                fragile |= np.abs(d2 - br2) <= F(4) * eps * (br + F(1))
    seg = np.asarray(seg, dtype=F)
    for j in range(int(n_seg)):
        x1, y1 = seg[0, j], seg[1, j]
        ex, ey = seg[2, j] - x1, seg[3, j] - y1
        dx, dy = x - x1, y - y1
        ee = ex * ex + ey * ey
        if ee > 0:
            h = (dx * ex + dy * ey) / ee
        else:
            h = np.zeros_like(dx)
        h = np.minimum(np.maximum(h, F(0)), F(1))
        wx, wy = dx - h * ex, dy - h * ey
        d2 = wx * wx + wy * wy
        img[d2 <= half2] = _pack(seg_rgba[j])
        if robot >= 0:
            fragile |= np.abs(d2 - half2) <= F(4) * eps * (half + F(1))
    for r in range(len(poses)):
        if r == robot:
            continue
        cox, coy = corner_offsets(bboxes[r])
        pts = [(F(px), F(py)) for px, py in robot_corners(poses[r], cox, coy)]
        pos = np.ones((img_h, img_w), dtype=bool)
        neg = np.ones((img_h, img_w), dtype=bool)
        for k in range(4):
            (ax, ay), (bx, by) = pts[k], pts[(k + 1) & 3]
            ex, ey = bx - ax, by - ay
            cr = ex * (y - ay) - ey * (x - ax)
            pos &= cr >= 0
            neg &= cr <= 0
            scale = (abs(ex) + abs(ey)) * (np.abs(x) + np.abs(y) + F(1))
            fragile |= np.abs(cr) <= F(1e-5) * scale
        img[pos | neg] = _pack(robot_rgba[r])
    return img, fragile


def render_scene_world(scene, i, poses=None, **kw):
    """`render` for world i of a SceneBatch (aitk.robots_b200.scenes.SceneBatch)."""
    poses = scene.pose[i] if poses is None else poses
    return render(scene.world_size[i], scene.seg[i], scene.seg_rgba[i], scene.seg_count[i], poses,
                  [r.bbox for r in scene.robots], [r.rgba for r in scene.robots],
                  None if scene.bulbs is None else scene.bulbs[i], **kw)
