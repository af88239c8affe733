"""CPU tier, world_size 2 over gloo: the data-parallel gradient hook (engine.GradAllReduce) sums the flat
gradient bucket over ranks and reports the world size -- the host logic of the N>1 learner."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    from xingtian_b200 import capi, engine

    class FakeNet(object):
        grads = torch.full((1000,), float(rank + 1))
    net = FakeNet()
    hook = engine.GradAllReduce(net)
    ws = hook._cb(None, None, 0, None)                       # query form
    rc = hook._cb(None, C.c_void_p(net.grads.data_ptr()), 1000, None)
    ok = ws == world and rc == 0 and bool(torch.all(net.grads == sum(range(1, world + 1))))
    # per-rank env sharding used by bench.py: disjoint seeds -> different rollouts
    hook.close()
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_grad_allreduce_hook_gloo(tmp_path):
    from xingtian_b200 import build
    build.build()
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "ok%d" % r)).read() == "1"
