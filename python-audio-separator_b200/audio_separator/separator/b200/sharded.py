"""Time-sharding of ONE track across the GPUs of a box (BASELINE north_star: "chunks shard by time across the 8 GPUs
with overlap-region halo exchange via NCCL").

The reference has no multi-GPU path; this is new.  The chunk grid of MDXSeparator.demix
(architectures/mdx_separator.py:335-348) is cut into contiguous chunk ranges, one per rank.  Chunks are independent
given the padded mixture; only the windowed overlap-add couples neighbours: an output sample is covered by
ceil(chunk/step) consecutive chunks, so rank r needs the last k = ceil(chunk/step) - 1 chunk outputs of rank r-1 --
one `isend`/`irecv` pair of k*(2, chunk) floats (2 MB at the Inst_HQ_3 sizes) between time-neighbours.  Each rank then
overlap-adds and finalises ITS slice of the output; the slices go to rank 0 with point-to-point receives straight
into the full (N, 2) stem buffers (no staging copy).

`plan_shards` is pure host arithmetic (tested on CPU with gloo, world_size 2); ShardedMdxEngine needs CUDA + NCCL.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .engine import MdxEngine, _ptr, _stream, check, lib


@dataclass
class Shard:
    rank: int
    c0: int  # first own chunk
    c1: int  # one past the last own chunk
    halo: int  # chunks received from the left neighbour: [c0 - halo, c0)
    q0: int  # first output sample (unpadded coordinates) this rank finalises
    q1: int  # one past the last


def plan_shards(n_samples: int, total_len: int, step: int, chunk: int, trim: int, n_chunks: int, world: int) -> list[Shard]:
    """Contiguous, balanced chunk ranges; rank r finalises the padded positions [c0*step, c1*step) (last rank: to the end)."""
    k = -(-chunk // step) - 1  # chunks of the left neighbour that reach into a rank's range
    base, rem = divmod(n_chunks, world)
    shards, c0 = [], 0
    for r in range(world):
        c1 = c0 + base + (1 if r < rem else 0)
        p0 = c0 * step
        p1 = total_len if r == world - 1 else c1 * step
        q0 = min(max(p0 - trim, 0), n_samples)
        q1 = min(max(p1 - trim, 0), n_samples)
        halo = min(k, c0) if c1 > c0 else 0
        shards.append(Shard(r, c0, c1, halo, q0, q1))
        c0 = c1
    for s in shards[1:]:
        left = shards[s.rank - 1]
        if s.halo > left.c1 - left.c0 and s.c1 > s.c0:
            raise ValueError(f"track too short to shard over {world} ranks: rank {s.rank} needs {s.halo} halo chunks, its neighbour owns {left.c1 - left.c0}")
    return shards


class ShardedMdxEngine(MdxEngine):
    """MdxEngine whose demix runs on torch.distributed ranks (backend nccl, one process per GPU).  Every rank is given the
    same (2, N) mix; stems are returned on rank 0 (None elsewhere)."""

    def __init__(self, *args, group=None, **kw):
        super().__init__(*args, **kw)
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("ShardedMdxEngine needs torch.distributed to be initialised (backend nccl)")
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def demix_device(self, mix_dev, is_match_mix=False, out_scale=1.0, with_secondary=False, interleave=True):
        dist = self.dist
        assert interleave, "the sharded path produces (N, 2) stems"
        mix_dev = mix_dev.contiguous()
        N = mix_dev.shape[1]
        L, step, n_chunks, overlap = self.grid(N, is_match_mix)
        T = self.chunk_size
        sh = plan_shards(N, L, step, T, self.trim, n_chunks, self.world)[self.rank]
        n_own = sh.c1 - sh.c0
        mixture = torch.zeros((2, L), dtype=torch.float32, device=self.device)
        mixture[:, self.trim : self.trim + N] = mix_dev
        local = torch.empty((sh.halo + n_own, 2, T), dtype=torch.float32, device=self.device)  # [halo | own]
        net = None if is_match_mix else self.net.handle
        for b0 in range(0, n_own, self.batch):
            nb = min(self.batch, n_own - b0)
            g0 = sh.c0 + b0
            work = self._workspace(nb)
            check(
                lib.b200sep_mdx_run_model(self.plan.handle, net, mixture.data_ptr() + g0 * step * 4, step, L, L - g0 * step, nb, T, self.dim_f, int(self.enable_denoise), _ptr(local[sh.halo + b0 : sh.halo + b0 + nb]), _ptr(work), _stream()),
                "mdx_run_model",
            )
        # ---- halo: my last k chunk outputs -> right neighbour, left neighbour's -> my halo slots
        shards = plan_shards(N, L, step, T, self.trim, n_chunks, self.world)
        ops = []
        if self.rank + 1 < self.world and shards[self.rank + 1].halo > 0:
            k = shards[self.rank + 1].halo
            ops.append(dist.P2POp(dist.isend, local[sh.halo + n_own - k : sh.halo + n_own], self.rank + 1, self.group))
        if sh.halo > 0:
            ops.append(dist.P2POp(dist.irecv, local[: sh.halo], self.rank - 1, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        # ---- overlap-add + finalise my slice of the output (global indexing into full-size buffers)
        primary = torch.empty((N, 2), dtype=torch.float32, device=self.device)
        secondary = torch.empty((N, 2), dtype=torch.float32, device=self.device) if with_secondary else None
        if sh.q1 > sh.q0:
            check(
                lib.b200sep_demix_overlap_add_range(
                    _ptr(local), sh.c0 - sh.halo, sh.halo + n_own, n_chunks, T, step, L, self.trim, N, sh.q0, sh.q1, int(overlap != 0), float(out_scale),
                    _ptr(mix_dev) if with_secondary else None, self.compensate, 1, _ptr(primary), _ptr(secondary) if with_secondary else None, _stream(),
                ),
                "demix_overlap_add_range",
            )
        # ---- gather the slices on rank 0 (p2p straight into the destination rows)
        ops = []
        bufs = [primary] + ([secondary] if with_secondary else [])
        if self.rank == 0:
            for s in shards[1:]:
                if s.q1 > s.q0:
                    for buf in bufs:
                        ops.append(dist.P2POp(dist.irecv, buf[s.q0 : s.q1], s.rank, self.group))
        elif sh.q1 > sh.q0:
            for buf in bufs:
                ops.append(dist.P2POp(dist.isend, buf[sh.q0 : sh.q1], 0, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if self.rank != 0:
            return (None, None) if with_secondary else None
        return (primary, secondary) if with_secondary else primary
