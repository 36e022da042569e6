"""Element partitioning across ranks (one process per GPU) and the interface exchange of the
explicit loop.

Assembly needs no collective: every rank assembles its own elements.  In the explicit loop the
internal force of a dof shared by two partitions is the sum of both partial products, so between
`Explicit.step_begin()` (U update, E = K_local U) and `Explicit.step_end()` the interface entries
of E are summed across ranks with `torch.distributed` (NCCL on GPUs, gloo in the CPU tests).
Strip partitions have at most two neighbours, so the exchange is a pair of isend/irecv of the
packed interface values (tens of KB: latency-bound)."""
from __future__ import annotations

import numpy as np


def partition_elements(nelem, nparts):
    """Contiguous element ranges (FinEtools block meshes are generated strip by strip)."""
    b = np.linspace(0, nelem, nparts + 1).astype(np.int64)
    return [(int(b[r]), int(b[r + 1])) for r in range(nparts)]


def local_mesh(conn, xyz, lo, hi):
    """Restrict a global mesh to elements [lo, hi): local 1-based connectivity, the local
    coordinates and the local->global node map."""
    c = np.asarray(conn)[lo:hi]
    nodes = np.unique(c)  # sorted global 1-based node ids
    lut = np.zeros(int(nodes.max()) + 1, dtype=np.int64)
    lut[nodes] = np.arange(1, nodes.size + 1)
    return lut[c], np.asarray(xyz)[nodes - 1], nodes


def shared_nodes(conn, ranges):
    """For every pair of partitions the global node ids (1-based, sorted) they share."""
    sets = [np.unique(np.asarray(conn)[lo:hi]) for lo, hi in ranges]
    out = {}
    for a in range(len(sets)):
        for b in range(a + 1, len(sets)):
            s = np.intersect1d(sets[a], sets[b])
            if s.size:
                out[(a, b)] = s
    return out


class InterfaceExchange:
    """Sums interface entries of a per-rank vector with the neighbouring ranks.
    `links`: list of (peer_rank, local_indices) -- for a given pair the two ranks must list the
    shared dofs in the same (global) order."""

    def __init__(self, links, device):
        import torch

        self.links = [(int(p), torch.as_tensor(np.asarray(ix, dtype=np.int64), device=device)) for p, ix in links if len(ix)]
        self.send = [torch.empty(ix.numel(), dtype=torch.float64, device=device) for _, ix in self.links]
        self.recv = [torch.empty(ix.numel(), dtype=torch.float64, device=device) for _, ix in self.links]

    def exchange_sum(self, vec):
        import torch
        import torch.distributed as dist

        if not self.links:
            return vec
        ops = []
        for (peer, ix), s, r in zip(self.links, self.send, self.recv):
            torch.index_select(vec, 0, ix, out=s)
            ops.append(dist.P2POp(dist.isend, s, peer))
            ops.append(dist.P2POp(dist.irecv, r, peer))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for (peer, ix), r in zip(self.links, self.recv):
            vec.index_add_(0, ix, r)
        return vec


class DevicePointer:
    """Zero-copy torch view of library-owned device memory (`torch.as_tensor(DevicePointer(...))`)."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def strip_links(rank, world, lo_dofs, hi_dofs):
    """Links of a 1-D strip partition: `lo_dofs` are this rank's local dof indices (0-based) on the
    edge shared with rank-1, `hi_dofs` those shared with rank+1, both ordered along the edge."""
    links = []
    if rank > 0:
        links.append((rank - 1, lo_dofs))
    if rank < world - 1:
        links.append((rank + 1, hi_dofs))
    return links
