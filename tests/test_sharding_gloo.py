"""CPU-only world_size-2 test of the multi-GPU host logic (block sharding + the single gather),
on the gloo backend."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import __graft_entry__ as entry


def test_unit_range_partitions(pkg):
    from mpecdsa_b200 import sharding
    for total in (0, 1, 7, 131072):
        for world in (1, 2, 3, 8):
            spans = [sharding.unit_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.session_range(65536, 3, 8) == (3 * 8192 * 2, 4 * 8192 * 2)
    with pytest.raises(ValueError):
        sharding.unit_range(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    entry.load_package()
    from mpecdsa_b200 import sharding
    lo, hi = sharding.unit_range(1001, rank, world)
    rec = torch.tensor([rank, lo, hi, sum(range(lo, hi))], dtype=torch.int64)
    allrec = sharding.gather_records(rec, world)
    q.put((rank, allrec.tolist()))
    dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0] == out[1]
    rows = out[0]
    assert [r[0] for r in rows] == [0, 1]
    assert rows[0][2] == rows[1][1] and rows[1][2] == 1001
    assert rows[0][3] + rows[1][3] == sum(range(1001))
