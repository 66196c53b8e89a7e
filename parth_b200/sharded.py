"""Multi-GPU driver of the reuse update: one process per GPU, every rank holds the same graph and
separator tree, the dirty nodes of an update are the unit that shards.

The library (parth_b200_set_shard) deals the nodes that need re-ordering to the ranks and orders
only this rank's share; this module moves the results: each rank packs the label slices of its
nodes into one buffer, the buffers are all-gathered (NCCL over NVLink on the GPU box; gloo in the
CPU tests), every rank unpacks the other ranks' slices, then the permutation is expanded.  Only
dirty slices travel -- never the whole permutation.

`plan_lpt` / `exchange` are plain host logic (no CUDA), tested with world_size 2 over gloo.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import api


def plan_lpt(nodes, sizes, world):
    """largest node first onto the least loaded rank (ties: lower node id, lower rank); the same
    rule as my_share() in csrc/stage_reorder.cu, so host tests can predict the device plan"""
    nodes = np.asarray(nodes, np.int64)
    sizes = np.asarray(sizes, np.int64)
    order = sorted(range(len(nodes)), key=lambda i: (-int(sizes[i]), int(nodes[i])))
    load = [0] * world
    owner = np.zeros(len(nodes), np.int32)
    for i in order:
        best = min(range(world), key=lambda r: (load[r], r))
        owner[i] = best
        load[best] += int(sizes[i])
    return owner


def segments(owner, sizes, world):
    """per rank: indices of its nodes (plan order) and the length of its packed buffer"""
    idx = [np.nonzero(np.asarray(owner) == r)[0] for r in range(world)]
    lens = [int(np.asarray(sizes)[i].sum()) for i in idx]
    return idx, lens


def exchange(nodes, owner, sizes, rank, world, pack, unpack, device, group=None):
    """all-gather the packed label slices.  pack(node_ids, tensor) fills `tensor` with the
    concatenated slices of this rank's nodes; unpack(node_ids, tensor) consumes another rank's."""
    idx, lens = segments(owner, sizes, world)
    width = max(max(lens), 1)
    buf = torch.zeros(world * width, dtype=torch.int32, device=device)
    mine = buf[rank * width:(rank + 1) * width]
    if lens[rank] > 0:
        pack(np.asarray(nodes)[idx[rank]], mine[:lens[rank]])
    if world > 1:
        dist.all_gather_into_tensor(buf, mine.clone(), group=group)
    for r in range(world):
        if r != rank and lens[r] > 0:
            unpack(np.asarray(nodes)[idx[r]], buf[r * width:r * width + lens[r]])
    return sum(lens) * 4  # bytes of label slices that exist in this update


class ShardedParth:
    """ParthAPI + the collective step.  Use exactly like api.ParthAPI; every rank must make the
    same calls with the same inputs."""

    def __init__(self, device, rank=None, world=None, group=None):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.group = group
        self.device = torch.device("cuda", device)
        self.g = api.ParthAPI(device)
        self.g.setShard(self.rank, self.world)
        self.stream = torch.cuda.ExternalStream(self.g.stream(), device=self.device)
        self.exchanged_bytes = 0

    def _exchange(self):
        nodes, owner, sizes = self.g.exchange_plan()
        with torch.cuda.stream(self.stream):
            self.exchanged_bytes = exchange(
                nodes, owner, sizes, self.rank, self.world,
                lambda ids, t: self.g.pack_labels_dev(ids, t.data_ptr()),
                lambda ids, t: self.g.unpack_labels_dev(ids, t.data_ptr()),
                self.device, self.group)
            self.stream.synchronize()

    def computePermutationDev(self, dperm_ptr, dim=3):
        rc = self.g.computePermutationDev(dperm_ptr, dim)
        if rc == api.NEEDS_EXCHANGE:
            self._exchange()
            self.g.finishPermutationDev(dperm_ptr, dim)
            rc = 0
        return rc

    def computePermutation(self, dim=3, out=None):
        rc, perm = self.g.computePermutation(dim, out=out)
        if rc == api.NEEDS_EXCHANGE:
            self._exchange()
            perm = self.g.finishPermutation(dim, out=out)
            rc = 0
        return rc, perm
