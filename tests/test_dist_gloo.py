"""world_size-2 gloo test of the frame sharding / result gather used by bench.py --gpus N."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from sassd_b200 import dist as D
    r, w, _ = D.init_from_env(backend="gloo")
    mine = D.shard_frames(10, r, w)
    # a fake per-frame result: frame id in every slot, count = frame id % 5
    det = torch.stack([torch.full((4, 9), float(f)) for f in mine])
    nd = torch.tensor([f % 5 for f in mine], dtype=torch.int32)
    det_all, nd_all = D.gather_detections(det, nd)
    g_det, g_nd = D.interleave(det_all, nd_all)
    t = D.max_over_ranks(1.0 + r, torch.device("cpu"))
    D.barrier()
    q.put((r, mine, g_det[:, 0, 0].tolist(), g_nd.tolist(), t))
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == [0, 2, 4, 6, 8] and res[1][1] == [1, 3, 5, 7, 9]
    for r in res:
        assert r[2] == [float(i) for i in range(10)]       # global frame order restored on every rank
        assert r[3] == [i % 5 for i in range(10)]
        assert r[4] == 2.0                                 # max over ranks


def test_single_process_passthrough():
    from sassd_b200 import dist as D
    det, nd = torch.zeros(3, 4, 9), torch.zeros(3, dtype=torch.int32)
    a, b = D.gather_detections(det, nd)
    assert a.shape == (1, 3, 4, 9) and b.shape == (1, 3)
    assert D.shard_frames(7, 1, 3) == [1, 4]
