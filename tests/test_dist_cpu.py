"""world_size-2 gloo test of the sharding / weight-broadcast host logic (the N>1 path of bench.py)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import kbnet_b200
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                      # different random weights per rank before the broadcast
    m = kbnet_b200.SparseToDensePool(2, [3, 5], [7], 8, 2, 'xavier_normal', 'leaky_relu')
    nbytes = kbnet_b200.dist.broadcast_parameters(m.parameters(), src=0)
    flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    full = torch.arange(7 * 3, dtype=torch.float32).view(7, 3)   # 7 samples over 2 ranks: 4 + 3
    (mine,) = kbnet_b200.dist.shard_batch([full], rank, world)
    back = kbnet_b200.dist.gather_outputs(mine * 2, 7)
    ok = same and nbytes == flat.numel() * 4 and torch.equal(back, full * 2) and mine.shape[0] == (4 if rank == 0 else 3)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    sys.path.insert(0, ROOT)
    import kbnet_b200
    for n in (0, 1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            rs = [kbnet_b200.dist.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(hi - lo for lo, hi in rs) - min(hi - lo for lo, hi in rs) <= 1


def test_broadcast_and_sharding_world_size_2_gloo():
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]
