"""CPU (gloo, world_size 2 and 3): the time-sharding protocol of audio_separator/separator/b200/sharded.py --
chunk-range partition, one-directional halo of k chunk outputs, per-rank overlap-add of its own output slice, gather on
rank 0 -- reproduces the single-process demix of the oracle sample-for-sample.  The per-chunk compute is the oracle's
run_model (this test is about the N>1 plumbing, not the kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mdx_oracle as O

SMALL = dict(n_fft=1536, hop_length=256, dim_f=768, dim_t=32, segment_size=32, g=8)


def ola_range(local, first_chunk, n_chunks, chunk, step, L, trim, N, q0, q1, use_window):
    """numpy restatement of b200sep_demix_overlap_add_range (gather form of mdx_separator.py:348-401)."""
    out = np.zeros((q1 - q0, 2), np.float32)
    for q in range(q0, q1):
        p = q + trim
        i_lo = 0 if p - chunk + 1 <= 0 else (p - chunk + step) // step
        i_hi = min(p // step, n_chunks - 1)
        res, div = np.zeros(2, np.float32), np.float32(0)
        for i in range(i_lo, i_hi + 1):
            s, e = i * step, min(i * step + chunk, L)
            if p >= e:
                continue
            w = np.float32(np.hanning(e - s)[p - s]) if use_window else np.float32(1)
            res += local[i - first_chunk][:, p - s] * w
            div += w
        out[q - q0] = res / div
    return out


def _worker(rank, world, port, n_samples, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audio_separator.separator.b200.sharded import plan_shards

    cfg = O.MDXConfig(**SMALL)
    mix = O.normalize(O.synth_music(n_samples, seed=5), 0.9, 0.0)
    L, step, starts = O.chunk_starts(n_samples, cfg)
    n_chunks, T = len(starts), cfg.chunk_size
    shards = plan_shards(n_samples, L, step, T, cfg.trim, n_chunks, world)
    sh = shards[rank]
    mixture = np.zeros((2, L), np.float32)
    mixture[:, cfg.trim : cfg.trim + n_samples] = mix
    local = np.zeros((sh.halo + sh.c1 - sh.c0, 2, T), np.float32)
    for c in range(sh.c0, sh.c1):
        part = np.zeros((1, 2, T), np.float32)
        e = min(c * step + T, L)
        part[0, :, : e - c * step] = mixture[:, c * step : e]
        local[sh.halo + c - sh.c0] = O.run_model(part, cfg, lambda s: s * 0.5)[0]
    reqs = []
    if rank + 1 < world and shards[rank + 1].halo:
        k = shards[rank + 1].halo
        reqs.append(dist.isend(torch.from_numpy(local[len(local) - k :].copy()), rank + 1))
    if sh.halo:
        buf = torch.empty((sh.halo, 2, T))
        dist.recv(buf, rank - 1)
        local[: sh.halo] = buf.numpy()
    for r in reqs:
        r.wait()
    mine = ola_range(local, sh.c0 - sh.halo, n_chunks, T, step, L, cfg.trim, n_samples, sh.q0, sh.q1, True)
    if rank == 0:
        full = np.zeros((n_samples, 2), np.float32)
        full[sh.q0 : sh.q1] = mine
        for s in shards[1:]:
            if s.q1 > s.q0:
                buf = torch.empty((s.q1 - s.q0, 2))
                dist.recv(buf, s.rank)
                full[s.q0 : s.q1] = buf.numpy()
        ref = O.demix(mix, cfg, lambda s: s * 0.5)
        q.put(float(np.abs(full.T - ref).max()))
    elif sh.q1 > sh.q0:
        dist.send(torch.from_numpy(mine.copy()), 0)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n_samples", [(2, 30000), (3, 52345)])
def test_time_sharded_demix_matches_single_process(world, n_samples):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_samples, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    for p_ in procs:
        p_.join(180)
        assert p_.exitcode == 0
    assert q.get(timeout=5) <= 2e-6


def test_plan_shards_properties():
    from audio_separator.separator.b200.sharded import plan_shards

    cfg = O.MDXConfig()
    N = 13_230_000
    L, step, starts = O.chunk_starts(N, cfg)
    for world in (1, 2, 4, 8):
        sh = plan_shards(N, L, step, cfg.chunk_size, cfg.trim, len(starts), world)
        assert sh[0].c0 == 0 and sh[-1].c1 == 68 and sh[0].q0 == 0 and sh[-1].q1 == N
        assert all(a.c1 == b.c0 and a.q1 == b.q0 for a, b in zip(sh, sh[1:]))  # contiguous, disjoint
        assert max(s.c1 - s.c0 for s in sh) - min(s.c1 - s.c0 for s in sh) <= 1  # balanced
        assert [s.halo for s in sh] == [0] + [1] * (world - 1)  # overlap 0.25 -> one chunk of halo
    with pytest.raises(ValueError):
        plan_shards(1000, 3000, 100, 950, 10, 3, 3)  # 9 halo chunks needed, neighbours own 1
