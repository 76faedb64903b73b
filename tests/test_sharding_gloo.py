"""CPU: the N>1 host logic (scene ownership + record all-gather) under gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from video_3d_llm_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_indices(n_items)
        local = torch.tensor([[i, 10 * i, 100 + i] for i in mine], dtype=torch.int32).reshape(len(mine), 3)
        full = sharding.all_gather_records(local, n_items)
        raw = sharding.all_gather_records(local[:1])
        q.put((rank, mine, full.tolist(), list(raw.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [6, 5])
def test_round_robin_all_gather_world2(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [[i, 10 * i, 100 + i] for i in range(n_items)]
    owned = []
    for rank, mine, full, raw_shape in res:
        assert mine == list(range(rank, n_items, 2))
        assert full == expect  # every rank ends with all scenes in global order
        assert raw_shape == [2, 3]
        owned += mine
    assert sorted(owned) == list(range(n_items))


def test_single_process_is_identity():
    rec = torch.arange(12, dtype=torch.int32).reshape(4, 3)
    assert torch.equal(sharding.all_gather_records(rec), rec)
    assert sharding.shard_indices(7, 1, 3) == [1, 4]
    assert sharding.padded_count(7, 3) == 3
    with pytest.raises(ValueError):
        sharding.shard_indices(4, 3, 2)
