# -*- coding: utf-8 -*-
"""Multi-GPU data parallelism: worlds are independent, so a batch shards as
contiguous world ranges, one process per GPU, with no collective on the data
path.  The only exchange is the optional all-gather of observations into one
training batch (BASELINE.json north_star); the kernels write straight into
this rank's slice of the gather buffer (brs_bind_output), so no staging copy
precedes the collective (NCCL in-place all-gather).
"""

def _dist():
    import torch.distributed as dist  # lazy: the CPU baseline workers never touch torch

    return dist


def partition(n_worlds, world_size, rank):
    """Contiguous [first, first+count) range of `rank`; remainders go to the low ranks."""
    base, rem = divmod(int(n_worlds), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def all_gather_worlds(local, counts=None, group=None):
    """All-gather per-rank [n_r, ...] tensors along dim 0 (works on gloo and nccl)."""
    import torch

    dist = _dist()
    ws = dist.get_world_size(group)
    if counts is None or len(set(counts)) == 1:
        out = local.new_empty((local.shape[0] * ws,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # ragged shards: pad to the largest, gather once, trim
    m = max(counts)
    padded = local.new_zeros((m,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    out = local.new_empty((m * ws,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * m : r * m + c] for r, c in enumerate(counts)], dim=0)


class ShardedWorlds:
    """This rank's shard of a global batch + the observation all-gather.

    scene_fn(first_world, count) -> SceneBatch must be deterministic in
    (first_world, count) so that shards compose into the same global batch
    irrespective of the number of ranks (the generators in scenes.py are, for
    chunk-aligned shards).
    """

    def __init__(self, scene_fn, n_worlds, device=None, group=None):
        import torch

        from .engine import BatchEngine

        dist = _dist()
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_worlds = int(n_worlds)
        self.counts = [partition(n_worlds, self.world_size, r)[1] for r in range(self.world_size)]
        self.first, self.count = partition(n_worlds, self.world_size, self.rank)
        self.device = torch.cuda.current_device() if device is None else device
        self.engine = BatchEngine(scene_fn(self.first, self.count), device=self.device)
        self._gather = {}
        self._fused = set()
        self._token = None

    def step(self, time_step=None, flags=7, stream=0):
        self.engine.step(time_step, flags, stream)

    def gather_buffer(self, name, fused=False):
        """Allocate the global observation tensor and make the kernels write this
        rank's slice of it directly (equal shard sizes only).

        fused=True (pictures of the solo launch plan): the buffer is IPC-shareable, every rank
        maps every other rank's buffer over NVLink, and the step kernel itself stores this
        rank's pictures into all of them (brs_bind_peer_outputs): World.step leaves the whole
        batch on every GPU, all_gather() is then only a stream barrier across the ranks."""
        import numpy as np
        import torch

        if len(set(self.counts)) != 1:
            raise ValueError("in-place gather needs equal shards; use all_gather()")
        local = self.engine.tensor(name)
        shape = (self.n_worlds,) + tuple(local.shape[1:])
        if fused and self.world_size > 1:
            dist = _dist()
            full, handle = self.engine.peer_alloc(shape, np.dtype(str(local.dtype).replace("torch.", "")))
            mine_h = torch.tensor(list(handle), dtype=torch.uint8, device=local.device)
            all_h = torch.empty((self.world_size, 64), dtype=torch.uint8, device=local.device)
            dist.all_gather_into_tensor(all_h, mine_h, group=self.group)
            all_h = all_h.cpu().numpy()
            stride = int(np.prod(shape[1:])) * local.element_size()
            ptrs = [self.engine.peer_open(bytes(all_h[r].tobytes())) + self.first * stride
                    for r in range(self.world_size) if r != self.rank]
            mine = full[self.first:self.first + self.count]
            self.engine.bind_output(name, mine)
            self.engine.bind_peer_outputs(name, ptrs)
            self._gather[name] = (full, mine)
            self._fused.add(name)
            self._token = torch.zeros(1, device=local.device)
            return full
        full = torch.empty(shape, dtype=local.dtype, device=local.device)
        mine = full[self.first:self.first + self.count]
        self.engine.bind_output(name, mine)
        self._gather[name] = (full, mine)
        return full

    def all_gather(self, name):
        """Observation `name` of every world on every rank (NVTX range "brs_gather:<name>")."""
        import torch

        nvtx = torch.cuda.nvtx if torch.cuda.is_available() else None
        if nvtx is not None:
            nvtx.range_push("brs_gather:%s" % name)
        try:
            return self._all_gather(name)
        finally:
            if nvtx is not None:
                nvtx.range_pop()

    def _all_gather(self, name):
        if name in self._fused:
            # the step kernels already stored every rank's pictures everywhere: wait for all of them
            _dist().all_reduce(self._token, group=self.group)
            return self._gather[name][0]
        if name in self._gather:
            full, mine = self._gather[name]
            if self.world_size > 1:
                _dist().all_gather_into_tensor(full, mine, group=self.group)  # in place: mine is full's slice
            return full
        local = self.engine.tensor(name)
        if self.world_size == 1:
            return local
        return all_gather_worlds(local, self.counts, self.group)
