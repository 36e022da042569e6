"""world_size-2 gloo test of the multi-rank explicit path's host logic (partition, interface
maps, exchange): K_local U summed over the interface equals the global K U.  CPU only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fe_external as fx
from oracle import shells as osh


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fsb200  # noqa: F401  (loads the package; no GPU call is made)
    from fsb200 import partition as pt

    # global problem (every rank builds it only to check the answer)
    xy, conn = fx.t3block(1.0, 0.8, 6, 8)
    xyz = fx.xyz3(xy)
    xyz[:, 2] = 0.1 * np.sin(3 * xyz[:, 0])
    Dps, Dt = osh.shell_material_stiffness(fx.moduli_iso(70e9, 0.3))
    nrm, val = osh.t3ff_associategeometry(xyz, conn)
    d = fx.DofField(xyz.shape[0]).numberdofs()
    na = d.nalldofs
    Ke = osh.t3ff_stiffness_elmats(xyz, conn, nrm, val, Dps, Dt, 0.01)
    Kglob = fx.csc_to_scipy(*fx.assemble_matrix("sparse", Ke, d.gatherdofnums(conn), na), na, na)
    rng = np.random.default_rng(7)
    U = rng.standard_normal(na)
    # this rank's partition, local numbering
    ranges = pt.partition_elements(conn.shape[0], world)
    lo, hi = ranges[rank]
    lconn, lxyz, nodes = pt.local_mesh(conn, xyz, lo, hi)
    ld = fx.DofField(lxyz.shape[0]).numberdofs()
    Kl = fx.csc_to_scipy(*fx.assemble_matrix("sparse", Ke[lo:hi], ld.gatherdofnums(lconn), ld.nalldofs), ld.nalldofs, ld.nalldofs)
    gdof_of_local = d.dofnums[nodes - 1].ravel() - 1  # local dof k (node-major) -> global dof
    Ul = U[gdof_of_local]
    El = torch.from_numpy(Kl @ Ul)
    shared = pt.shared_nodes(conn, ranges)
    links = []
    for (a, b), s in shared.items():
        if rank in (a, b):
            peer = b if rank == a else a
            lnodes = np.searchsorted(nodes, s)  # local 0-based node ids, in global order
            links.append((peer, (lnodes[:, None] * 6 + np.arange(6)[None, :]).ravel()))
    ex = pt.InterfaceExchange(links, "cpu")
    ex.exchange_sum(El)
    ref = (Kglob @ U)[gdof_of_local]
    err = float(np.linalg.norm(El.numpy() - ref) / np.linalg.norm(ref))
    q.put((rank, err, len(links)))
    dist.barrier()
    dist.destroy_process_group()


def test_interface_exchange_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    for rank, err, nlinks in res:
        assert nlinks == 1
        assert err < 1e-13, (rank, err)
