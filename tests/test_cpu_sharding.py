"""CPU tests of the N>1 host logic with the gloo backend, world_size 2: scene
sharding, query-sharded kNN + all-gather (the local 'kernel' is the oracle here), gradient
all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pointops_oracle as O
        from pointcloudpdf_b200 import sharding as S
        g = torch.Generator().manual_seed(5)
        sizes = [700, 3, 450]
        xyz = torch.rand(sum(sizes), 3, generator=g)
        off_host = list(np.cumsum(sizes))
        offset = torch.tensor(off_host, dtype=torch.int32)
        full_idx, full_dist = O.knn_query(8, xyz, offset)
        idx, dst = S.sharded_knn_query(8, xyz, offset, off_host,
                                       knn_fn=lambda k, x, o, q, qo: O.knn_query(k, x, o, q, qo))
        ok_knn = torch.equal(idx, full_idx) and torch.equal(dst, full_dist)
        # gradient all-reduce = mean over ranks
        p = torch.nn.Parameter(torch.zeros(5))
        p.grad = torch.full((5,), float(rank + 1))
        S.allreduce_gradients([p])
        ok_grad = torch.allclose(p.grad, torch.full((5,), (1 + world) / 2))
        ret[rank] = (ok_knn, ok_grad, S.shard_scenes(5, rank, world))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_knn_and_allreduce():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert ret[0][0] and ret[1][0], "query-sharded kNN + all-gather differs from the single-process result"
        assert ret[0][1] and ret[1][1]
        assert ret[0][2] == [0, 1, 2] and ret[1][2] == [3, 4]


def test_shard_helpers():
    from pointcloudpdf_b200 import sharding as S
    assert S.shard_scenes(8, 3, 4) == [6, 7]
    assert S.shard_scenes(3, 3, 4) == []
    owned = sum((S.shard_scenes(7, r, 3) for r in range(3)), [])
    assert owned == list(range(7))
    spans = [S.query_slices([10, 13, 40], r, 4)[0] for r in range(4)]
    for s in range(3):  # slices of a scene tile it exactly
        cuts = [sp[s] for sp in spans]
        assert cuts[0][0] == [0, 10, 13][s] and cuts[-1][1] == [10, 13, 40][s]
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(3))
    coord = torch.arange(40.).view(40, 1).repeat(1, 3)
    c, f, off = S.slice_batch(coord, coord, [10, 13, 40], [0, 2])
    assert c.shape[0] == 37 and off == [10, 37] and c[10, 0] == 13
