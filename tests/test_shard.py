"""Multi-GPU host logic on the CPU: streams are sharded with no data-path collective; ranks only exchange
per-rank sample counts for the whole-job throughput.  world_size-2 gloo run of the same planner + walker."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

import conftest
from conftest import load_vector


def test_lpt_shards_are_balanced_and_complete():
    import minimp3_b200 as m
    sizes = [391105, 130613, 2352, 133120, 53498, 15673, 13374, 166661, 95760, 26645]*13
    for n in (1, 2, 4, 8):
        shard_of = m.shard_streams(sizes, n)
        assert len(shard_of) == len(sizes) and set(shard_of) == set(range(n))
        loads = [sum(s for s, k in zip(sizes, shard_of) if k == r) for r in range(n)]
        assert max(loads) - min(loads) <= max(sizes)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, names, result_path):
    sys.path.insert(0, conftest.ROOT)
    sys.path.insert(0, os.path.join(conftest.ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import harness
    import minimp3_b200 as m
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    datas = [load_vector(n) for n in names]
    shard_of = m.shard_streams([len(d) for d in datas], world)
    mine = [i for i, s in enumerate(shard_of) if s == rank]
    samples = sum(harness.walk(datas[i], raw=True)["info"]["samples_decoded"] for i in mine)
    t = torch.tensor([samples, len(mine)], dtype=torch.int64)
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)          # bookkeeping only: no PCM ever crosses ranks
    if rank == 0:
        np.save(result_path, np.stack([g.numpy() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_walk(tmp_path):
    import harness
    names = ["l3-si_block", "l3-compl", "l3-he_mode", "M2L3_noise", "l3-si_huff", "l3-he_free", "l3-si"]
    total = sum(harness.walk(load_vector(n), raw=True)["info"]["samples_decoded"] for n in names)
    out = str(tmp_path / "g.npy")
    mp.spawn(_worker, args=(2, _free_port(), names, out), nprocs=2, join=True)
    g = np.load(out)
    assert g[:, 0].sum() == total and g[:, 1].sum() == len(names) and (g[:, 1] > 0).all()
