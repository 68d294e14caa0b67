"""world_size-2 gloo test of the N>1 host logic (CPU): stream sharding is a partition, the LUT
broadcast delivers rank 0's tables, and the MAX-over-ranks timing reduction works."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import __graft_entry__ as graft


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ob = graft.load_package()
    sh = ob.sharding
    mine = sh.streams_of_rank(n_streams, world, rank)
    rs = np.random.default_rng(0 if rank == 0 else 99)          # only rank 0 holds the real LUT
    d = rs.random((64, 3)).astype(np.float32)
    o = rs.random((64, 3)).astype(np.float32)
    td, to = sh.broadcast_lut(d, o, dist, src=0)
    t_max = sh.max_over_ranks(10.0 + rank, dist)
    # checksum of checksums over the shards: every stream is processed exactly once
    local = torch.tensor([sum(hash_stream(i) for i in mine)], dtype=torch.int64)
    dist.all_reduce(local)
    q.put((rank, mine, td.numpy().copy(), to.numpy().copy(), t_max, int(local.item())))
    dist.destroy_process_group()


def hash_stream(i):
    return (i * 2654435761) % 1000003


def test_two_rank_gloo_sharding_and_lut_broadcast():
    world, n_streams = 2, 64          # BASELINE configs[3]: 64 streams
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = res[0][1] + res[1][1]
    assert sorted(owned) == list(range(n_streams))
    assert not set(res[0][1]) & set(res[1][1])
    want = np.random.default_rng(0)
    d0 = want.random((64, 3)).astype(np.float32)
    o0 = want.random((64, 3)).astype(np.float32)
    for r in res:
        assert np.array_equal(r[2], d0) and np.array_equal(r[3], o0)
        assert r[4] == 11.0
        assert r[5] == sum(hash_stream(i) for i in range(n_streams))


def test_sharding_helpers():
    ob = graft.load_package()
    sh = ob.sharding
    for world in (1, 2, 4, 8):
        all_ids = sum((sh.streams_of_rank(64, world, r) for r in range(world)), [])
        assert sorted(all_ids) == list(range(64))
        assert all(sh.rank_of_stream(i, world) == r for r in range(world) for i in sh.streams_of_rank(64, world, r))
    with pytest.raises(ValueError):
        sh.streams_of_rank(8, 2, 2)
