"""CPU tier: the N>1 path of bench.py's sharding logic over gloo, world_size 2 (images are independent units;
the only collective is the broadcast of the shared Huffman/quant table blob)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import jpegdec_b200 as J
    from tests import common as T
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blob = torch.zeros(J.TABLE_BLOB_BYTES, dtype=torch.uint8)
    if rank == 0:
        b = np.zeros(J.TABLE_BLOB_BYTES, np.uint8)
        assert J.lib().JPEGB200_exportTables(T.image("tulips"), len(T.image("tulips")), b.ctypes.data)
        blob.copy_(torch.from_numpy(b))
    dist.broadcast(blob, src=0)
    # every rank rebuilds the tables of its own images and must agree with the broadcast blob (shared tables)
    mine = np.zeros(J.TABLE_BLOB_BYTES, np.uint8)
    J.lib().JPEGB200_exportTables(T.image("tulips"), len(T.image("tulips")), mine.ctypes.data)
    same = bool((torch.from_numpy(mine)[:16 + 12800] == blob[:16 + 12800]).all())
    # shard: image i -> rank i % world ; weak scaling bookkeeping (max over ranks of the time, sum of the work)
    n_total = 10
    my_images = [i for i in range(n_total) if i % world == rank]
    t = torch.tensor([0.5 + 0.25 * rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([len(my_images)], dtype=torch.int64)
    dist.all_reduce(cnt)
    q.put((rank, same, float(t.item()), int(cnt.item()), my_images))
    dist.destroy_process_group()


def test_two_rank_shard_and_table_broadcast():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res)                  # blob identical on both ranks
    assert res[0][2] == res[1][2] == 0.75          # max over ranks
    assert res[0][3] == res[1][3] == 10            # every image decoded exactly once
    assert sorted(res[0][4] + res[1][4]) == list(range(10))
