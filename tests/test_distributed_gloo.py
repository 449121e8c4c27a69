"""World-size-2 run of the multi-GPU host logic on the CPU (gloo): block sharding and the replication helper that moves the key
buffers from rank 0 to every rank.  On the GPU box the same helpers run over NCCL on the library's device buffers."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_blocks, out_q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dmod = importlib.import_module("tfhe-aes-2_b200.distributed")
    # "keys": rank 0 holds the data, the others hold garbage of the same shape
    rng = np.random.default_rng(123)
    truth = [rng.integers(0, 256, n, dtype=np.uint8) for n in (1 << 20, 12345, 7)]
    bufs = [torch.from_numpy(t.copy()) if rank == 0 else torch.zeros(t.size, dtype=torch.uint8) for t in truth]
    dmod.broadcast_tensors(bufs, src=0, chunk_bytes=1 << 18)
    ok = all(np.array_equal(b.numpy(), t) for b, t in zip(bufs, truth))
    lo, hi = dmod.shard_range(n_blocks, rank, world)
    # every rank "encrypts" its shard (here: a checksum per block); rank 0 gathers to check coverage and order
    mine = torch.tensor([(i * 2654435761) % 1000003 for i in range(lo, hi)], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.numel()], dtype=torch.int64))
    padded = torch.zeros(max(int(s) for s in sizes), dtype=torch.int64)
    padded[:mine.numel()] = mine
    gathered = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded)
    if rank == 0:
        full = torch.cat([g[:int(s)] for g, s in zip(gathered, sizes)]).tolist()
        out_q.put((ok, full, [int(s) for s in sizes]))
    else:
        out_q.put((ok, None, None))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_blocks", [10, 1024, 1])
def test_two_rank_sharding_and_key_replication(n_blocks):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_blocks, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] for r in res), "broadcast did not reproduce rank 0's buffers"
    full, sizes = next((r[1], r[2]) for r in res if r[1] is not None)
    assert full == [(i * 2654435761) % 1000003 for i in range(n_blocks)]      # every block exactly once, in counter order
    assert sum(sizes) == n_blocks and max(sizes) - min(sizes) <= 1


def test_shard_range_properties():
    sys.path.insert(0, ROOT)
    dmod = importlib.import_module("tfhe-aes-2_b200.distributed")
    for n in (0, 1, 7, 10, 128, 1024, 1025):
        for w in (1, 2, 4, 8):
            parts = [dmod.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    assert [dmod.shard_range(10, r, 8) for r in range(8)] == [(0, 2), (2, 4), (4, 5), (5, 6), (6, 7), (7, 8), (8, 9), (9, 10)]
