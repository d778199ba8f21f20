"""Multi-GPU forms on real GPUs (skipped with fewer than 2 devices): the query-sharded large-scene kNN over NCCL
all-gather and over the fused P2P-store kernel must both equal the single-GPU result bit for bit; DDP training
steps on scene-sharded batches keep the replicas' parameters identical."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, k, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        from pointcloudpdf_b200 import synthetic as S, sharding
        from pointcloudpdf_b200.pointops import _common as C
        b = S.s3dis_batch([n], seed=2029)
        xyz, off = b["coord"].to(dev), b["offset"].to(dev)
        off_host = b["offset"].tolist()
        ref_idx, ref_dist, _ = C.get_grid(xyz, off).query(k, xyz, off, True, False)
        a_idx, a_dist = sharding.sharded_knn_query(k, xyz, off, off_host,
                                                   knn_fn=lambda ns, x, o, q, qo: C.get_grid(x, o).query(ns, q, qo, True, False)[:2])
        ok_gather = torch.equal(a_idx, ref_idx) and torch.equal(a_dist, ref_dist)
        ok_fused = None
        try:
            f_idx, f_dist = sharding.sharded_knn_query_fused(k, xyz, off, off_host)
            torch.cuda.synchronize()
            ok_fused = bool(torch.equal(f_idx, ref_idx) and torch.equal(f_dist, ref_dist))
            f_idx, f_dist = sharding.sharded_knn_query_fused(k, xyz, off, off_host)   # buffers are reused
            torch.cuda.synchronize()
            ok_fused = ok_fused and bool(torch.equal(f_idx, ref_idx))
        except Exception as e:  # noqa: BLE001 -- reported, not swallowed: the parent asserts on it
            ok_fused = f"{type(e).__name__}: {e}"
        out[rank] = (bool(ok_gather), ok_fused)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(300_000, 16), (120_000, 32)])
def test_sharded_knn_equals_single_gpu(cuda, n, k):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29500 + (os.getpid() % 2000), n, k, out), nprocs=world, join=True)
    for r in range(world):
        gather_ok, fused_ok = out[r]
        assert gather_ok, f"rank {r}: all-gather form differs from the single-GPU result"
        assert fused_ok is True, f"rank {r}: fused P2P form: {fused_ok}"
