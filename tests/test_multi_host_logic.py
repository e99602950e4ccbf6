"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path that needs no device -- the marker
block partition (Bayes::set_block_of_markers, bayes.cpp:903-925, as the engines of a 2-GPU run compute it from
vranks) tiles the markers, and the handle exchange helper (api.Engine.exchange_buffers) imports every peer's
handles, and only the peers', in rank order."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_of(Mt, R, world, rank):
    """What gmrm_create derives for engine `rank`: virtual ranks [rank*R/world, (rank+1)*R/world) and their markers."""
    from oracle import oracle_py as O
    Vl = R // world
    S0, _, _ = O.block_of_markers(Mt, R, rank * Vl)
    Sl, Ml, _ = O.block_of_markers(Mt, R, rank * Vl + Vl - 1)
    return S0, Sl + Ml - S0


def worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gmrm_b200 import api

    class FakeEngine:                       # the exchange helper only needs these three members
        class cfg:
            world_rank = rank
        imported = []

        def export_buffers(self):
            return bytes([rank]) * 320

        def import_buffers(self, r, h):
            self.imported.append((r, h))

    def gather(x):
        lst = [None] * world
        dist.all_gather_object(lst, x)
        return lst

    fe = FakeEngine()
    api.Engine.exchange_buffers(fe, gather)
    ok = [r for r, _ in fe.imported] == [r for r in range(world) if r != rank] and all(h == bytes([r]) * 320 for r, h in fe.imported)
    Mt, R = 1203, 16
    shards = gather(shard_of(Mt, R, world, rank))
    tiled = shards[0][0] == 0 and all(shards[i][0] + shards[i][1] == shards[i + 1][0] for i in range(world - 1)) and \
        shards[-1][0] + shards[-1][1] == Mt
    out[rank] = bool(ok and tiled)
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_host_logic():
    world = 2
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=worker, args=(r, world, 29641, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        assert all(p.exitcode == 0 for p in procs)
        assert dict(out) == {0: True, 1: True}
