# -*- coding: utf-8 -*-
"""Gym-style vectorised environment over a BatchEngine:
the caller the batched ``World.step`` throughput exists for.

Everything stays on the device: actions are written straight into the engine's
target-velocity buffer, observations are zero-copy torch views of the engine's
observation buffers, rewards / dones / resets are a handful of torch element-wise
ops on the pose and stalled views (plumbing, not the hot path -- the hot path is the
one ``brs_step`` launch per ``step``).

    env = VecEnv(rb.scenes.config2(n_worlds=4096))
    obs = env.reset()
    obs, reward, done, info = env.step(actions)      # actions: [B, R, 2] in [-1, 1]
"""

from . import _cabi as cabi
from .engine import BatchEngine


class VecEnv:
    """B independent worlds, R robots each, one CUDA device.

    actions[b, r] = (translate, rotate) in [-1, 1], scaled by the robot slot's
    max velocities exactly like upstream ``Robot.move(translate, rotate)``.

    reward_fn(env) -> [B] float tensor and done_fn(env) -> [B] bool tensor are
    optional hooks; the defaults reward forward progress (distance moved this
    step) minus ``stall_penalty`` per stalled robot, and end an episode after
    ``max_steps`` steps.  Finished worlds are reset in place (pose, velocities,
    step counter) before the next step, so ``step`` always returns live worlds.
    ``frame_skip=k`` repeats every action for k World.steps (k-1 move-only steps in one
    launch, then one full step); rewards and observations are those after the k-th.
    """

    def __init__(self, scene, device=0, max_steps=1000, stall_penalty=1.0, reward_fn=None, done_fn=None,
                 camera=True, frame_skip=1):
        import torch

        self.torch = torch
        self.engine = BatchEngine(scene, device=device)
        e = self.engine
        self.n_worlds, self.n_robots = e.n_worlds, e.n_robots
        self.max_steps, self.stall_penalty = int(max_steps), float(stall_penalty)
        self.reward_fn, self.done_fn = reward_fn, done_fn
        # action repeat: one env step = frame_skip World.steps with the same action; only the
        # last one senses (the others are move-only steps, all in one brs_steps launch)
        self.frame_skip = max(1, int(frame_skip))
        self.flags = cabi.STEP_MOVE | cabi.STEP_SENSE | (cabi.STEP_CAMERA if (camera and e.n_camera) else 0)
        dev = torch.device("cuda", e.device)
        # zero-copy views of engine memory
        self.pose, self.vel, self.target_vel = e.tensor("pose"), e.tensor("vel"), e.tensor("target_vel")
        self.stalled = e.tensor("stalled")
        self.obs = {k: e.tensor(k) for k in ("range", "light", "smell") if getattr(e, "n_" + k)}
        if camera and e.n_camera:
            self.obs["camera"] = e.tensor("camera")
        self._init_pose = torch.as_tensor(scene.pose, device=dev).clone()
        self._scale = torch.tensor([[r.max_trans_velocity, r.max_rotate_velocity] for r in scene.robots],
                                   dtype=torch.float64, device=dev)  # [R, 2]
        self.steps = torch.zeros(self.n_worlds, dtype=torch.int64, device=dev)
        self._prev_xy = self.pose[..., :2].clone()
        self._dev = dev

    # -- helpers -------------------------------------------------------------------
    def _stream(self):
        # the caller's CURRENT stream at the time of the call (torch.cuda.stream(...) contexts
        # change it), so that the launch is ordered with the torch ops around it
        return self.torch.cuda.current_stream(self._dev).cuda_stream

    def _sense(self):
        self.engine.step(flags=self.flags & ~cabi.STEP_MOVE, stream=self._stream())

    def observations(self):
        """Dict of zero-copy observation tensors (valid until the next step)."""
        return self.obs

    def reset(self, mask=None):
        """Reset every world (or the worlds selected by a [B] bool mask) to its initial pose."""
        t = self.torch
        if mask is None:
            mask = t.ones(self.n_worlds, dtype=t.bool, device=self.pose.device)
        m = mask.view(-1, 1, 1)
        self.pose.copy_(t.where(m, self._init_pose, self.pose))
        self.vel.copy_(t.where(m, t.zeros_like(self.vel), self.vel))
        self.target_vel.copy_(t.where(m, t.zeros_like(self.target_vel), self.target_vel))
        self.stalled.copy_(t.where(mask.view(-1, 1), t.zeros_like(self.stalled), self.stalled))
        self.steps.masked_fill_(mask, 0)
        self._prev_xy.copy_(self.pose[..., :2])
        self._sense()
        return self.obs

    def step(self, actions):
        """actions: [B, R, 2] (torch, any float dtype, CUDA or CPU) -> obs, reward, done, info."""
        t = self.torch
        a = t.as_tensor(actions, device=self.pose.device).to(t.float64).clamp(-1.0, 1.0)
        self.target_vel[..., 0] = a[..., 0] * self._scale[:, 0]
        self.target_vel[..., 2] = a[..., 1] * self._scale[:, 1]
        self._prev_xy.copy_(self.pose[..., :2])
        if self.frame_skip > 1:
            self.engine.steps(self.frame_skip - 1, flags=cabi.STEP_MOVE, stream=self._stream())
        self.engine.step(flags=self.flags, stream=self._stream())  # the step launch
        self.steps += self.frame_skip
        if self.reward_fn is not None:
            reward = self.reward_fn(self)
        else:
            moved = (self.pose[..., :2] - self._prev_xy).norm(dim=-1).sum(dim=1)
            reward = (moved - self.stall_penalty * self.stalled.to(t.float64).sum(dim=1)).to(t.float32)
        done = self.done_fn(self) if self.done_fn is not None else self.steps >= self.max_steps
        # (copies: the reset below clears the live views for the worlds that just finished)
        info = {"stalled": self.stalled.clone(), "steps": self.steps.clone()}
        if bool(done.any()):
            self.reset(done)  # finished worlds restart; their new observations are already sensed
        return self.obs, reward, done, info

    def close(self):
        self.engine.close()
