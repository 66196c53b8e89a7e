"""CPU, world_size 2 over gloo: the host side of the multi-GPU path (parth_b200/sharded.py) --
the node -> rank plan and the pack / all-gather / unpack of label slices -- with numpy arrays
standing in for the device label store."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_is_balanced_and_deterministic():
    from parth_b200 import sharded
    rng = np.random.default_rng(0)
    nodes = rng.permutation(511)[:97]
    sizes = rng.integers(1, 40000, len(nodes))
    for world in (1, 2, 4, 8):
        owner = sharded.plan_lpt(nodes, sizes, world)
        assert owner.min() >= 0 and owner.max() < world
        loads = np.array([sizes[owner == r].sum() for r in range(world)])
        assert loads.max() - loads.min() <= sizes.max()          # LPT bound
        assert np.array_equal(owner, sharded.plan_lpt(nodes, sizes, world))
    # ties: lower node id first, lower rank first
    assert sharded.plan_lpt([9, 3, 5], [7, 7, 7], 2).tolist() == [0, 0, 1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from parth_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)                          # same inputs on every rank
        n_nodes = 64
        node_ptr = np.concatenate([[0], np.cumsum(rng.integers(0, 50, n_nodes))])
        dirty = np.sort(rng.choice(n_nodes, 23, replace=False))
        sizes = node_ptr[dirty + 1] - node_ptr[dirty]
        owner = sharded.plan_lpt(dirty, sizes, world)
        # "device" label store: the true result is a seeded permutation per node; a rank only knows its own nodes
        truth = np.concatenate([np.random.default_rng(1000 + x).permutation(node_ptr[x + 1] - node_ptr[x])
                                for x in range(n_nodes)]).astype(np.int32)
        labels = np.full(len(truth), -1, np.int32)
        clean = np.setdiff1d(np.arange(n_nodes), dirty)
        for x in clean:
            labels[node_ptr[x]:node_ptr[x + 1]] = truth[node_ptr[x]:node_ptr[x + 1]]
        for x, o in zip(dirty, owner):
            if o == rank:
                labels[node_ptr[x]:node_ptr[x + 1]] = truth[node_ptr[x]:node_ptr[x + 1]]

        def pack(ids, t):
            t.copy_(torch.from_numpy(np.concatenate([labels[node_ptr[x]:node_ptr[x + 1]] for x in ids])))

        def unpack(ids, t):
            a = t.numpy()
            o = 0
            for x in ids:
                n = node_ptr[x + 1] - node_ptr[x]
                labels[node_ptr[x]:node_ptr[x + 1]] = a[o:o + n]
                o += n

        nbytes = sharded.exchange(dirty, owner, sizes, rank, world, pack, unpack, torch.device("cpu"))
        q.put((rank, bool(np.array_equal(labels, truth)), int(nbytes), int(sizes.sum()) * 4))
    finally:
        dist.destroy_process_group()


def test_exchange_over_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, nbytes, expect in res:
        assert ok, f"rank {rank}: labels differ after the exchange"
        assert nbytes == expect
