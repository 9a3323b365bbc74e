"""CPU, world_size 2 over gloo: the multi-GPU host logic — contiguous root sharding and the single
all_gather of final root statistics (the only communication of a search)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hanabizero_b200.dist import AsyncStatsGather, gather_root_stats, shard_range


def test_shard_range_partitions_exactly():
    for total in (1, 7, 4096, 16384, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, A, ragged, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    visits_all = rng.integers(0, 50, (total, A)).astype(np.int32)
    values_all = rng.standard_normal(total).astype(np.float32)
    lo, hi = shard_range(total, rank, world)
    v, val = gather_root_stats(torch.from_numpy(visits_all[lo:hi]), torch.from_numpy(values_all[lo:hi]))
    ok = bool((v.numpy() == visits_all).all() and (val.numpy().view(np.uint32) == values_all.view(np.uint32)).all())
    if not ragged:   # the overlapped gather of the timed path (equal shards): two tickets in flight, results in order
        ag = AsyncStatsGather(hi - lo, A, torch.device("cpu"), depth=2)
        t0 = ag.submit(torch.from_numpy(visits_all[lo:hi]), torch.from_numpy(values_all[lo:hi]))
        t1 = ag.submit(torch.from_numpy(visits_all[lo:hi] + 1), torch.from_numpy(values_all[lo:hi] * 2))
        v0, val0 = ag.result(t0)
        v1, val1 = ag.result(t1)
        ok = ok and bool((v0.numpy() == visits_all).all() and (val0.numpy() == values_all).all()
                         and (v1.numpy() == visits_all + 1).all() and (val1.numpy() == values_all * 2).all())
        # the packing-free variant SearchPipeline uses: one buffer [n * A visits | n value bits] per rank
        n = hi - lo
        flat = torch.cat((torch.from_numpy(visits_all[lo:hi]).reshape(-1),
                          torch.from_numpy(values_all[lo:hi].view(np.int32).copy())))
        t2 = ag.submit_flat(flat)
        v2, val2 = ag.result(t2)
        ok = ok and bool((v2.numpy() == visits_all).all() and (val2.numpy().view(np.uint32) == values_all.view(np.uint32)).all())
    out_q.put((rank, ok, tuple(v.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total,ragged", [(64, False), (37, True)])
def test_gather_root_stats_world2_gloo(total, ragged):
    world, A = 2, 20
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (1 if ragged else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, A, ragged, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok and shape == (total, A), (rank, ok, shape)
