"""Host logic of the multi-GPU gathering of assembled blocks (option A), on CPU with gloo:
every rank assembles (here: with the oracle) the elements touching the nodes whose columns it owns, on its
local mesh with order-preserving local dof numbers; the gathered column blocks must BE the global matrix --
colptr / rowval bit-exact, values equal (the same element contributions in the same order)."""
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


def _problem(permute):
    xy, conn = fx.t3block(1.0, 0.8, 7, 9)
    xyz = fx.xyz3(xy)
    xyz[:, 2] = 0.1 * np.sin(3 * xyz[:, 0]) + 0.05 * xyz[:, 1] ** 2
    # an orphan node (no element): it still owns six empty columns of the global matrix
    xyz = np.vstack([xyz, [[2.0, 2.0, 0.0]]])
    Dps, Dt = osh.shell_material_stiffness(fx.moduli_iso(70e9, 0.3))
    nrm, val = osh.t3ff_associategeometry(xyz, conn)
    d = fx.DofField(xyz.shape[0])
    edge = np.flatnonzero(xyz[:, 0] < 1e-9)
    for comp in (1, 2, 3, 5):
        d.setebc(edge, comp)
    d.setebc(np.array([17, 33]), 6)
    perm = np.random.default_rng(3).permutation(xyz.shape[0]) if permute else None
    d.numberdofs(perm)
    Ke = osh.t3ff_stiffness_elmats(xyz, conn, nrm, val, Dps, Dt, 0.01)
    return xyz, conn, d, Ke


def _worker(rank, world, port, kind, permute, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fsb200 import partition as pt

    xyz, conn, d, Ke = _problem(permute)
    nfree = int((~d.is_fixed).sum())
    na = d.nalldofs
    cpg, rvg, nzg = fx.assemble_matrix(kind, Ke, d.gatherdofnums(conn), na, nfree)
    plan = pt.ColumnBlockPlan(conn, d.dofnums, nfree, kind, rank, world)
    # local assembly with the oracle on the local mesh / local dof numbers
    ldofs = plan.dofnums[plan.conn - 1].reshape(plan.conn.shape[0], -1)
    cp, rv, nz = fx.assemble_matrix(kind, plan.restrict_elements(Ke), ldofs, plan.nall, plan.nfree)
    c0, c1 = plan.lcol_lo, plan.lcol_hi
    s0, s1 = cp[c0] - 1, cp[c1] - 1

    def fill(cnt, rows, vals):
        cnt.copy_(torch.from_numpy(np.diff(cp[c0 : c1 + 1])))
        rows.copy_(torch.from_numpy(plan.loc2glob[rv[s0:s1] - 1]))
        vals.copy_(torch.from_numpy(nz[s0:s1]))

    colptr, rowval, nzval = pt.gather_blocks(fill, c1 - c0, s1 - s0, "cpu")
    ok_pattern = np.array_equal(colptr.numpy(), cpg) and np.array_equal(rowval.numpy(), rvg)
    ok_values = np.array_equal(nzval.numpy(), nzg)
    # vectors: owned entries of a per-rank vector over the local dofs
    F = np.arange(1.0, (nfree if kind.startswith("ffblock") else na) + 1)
    Fl = torch.from_numpy(F[plan.loc2glob[: plan.local_ncols()] - 1].copy())
    Fg = pt.gather_vector(Fl, plan, "cpu")
    ok_vec = np.array_equal(Fg.numpy(), F)
    q.put((rank, ok_pattern, ok_values, ok_vec, len(plan.elems), conn.shape[0], plan.col_hi - plan.col_lo))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kind,permute", [(2, "ffblock", False), (3, "sparse", True), (3, "ffblock", True), (2, "diag", False)])
def test_gathered_column_blocks_are_the_global_matrix(world, kind, permute):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, kind, permute, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    ncols = 0
    for rank, okp, okv, okvec, nel, nel_glob, nc in res:
        assert okp, (rank, "pattern differs")
        assert okv, (rank, "values differ")
        assert okvec, (rank, "vector differs")
        assert nc > 0
        if not permute:
            assert nel < nel_glob  # a strip plus its interface elements, not the whole mesh
        ncols += nc


def test_plan_bounds_fall_on_node_boundaries():
    from fsb200 import partition as pt

    xyz, conn, d, _ = _problem(False)
    nfree = int((~d.is_fixed).sum())
    for world in (2, 3, 5):
        plans = [pt.ColumnBlockPlan(conn, d.dofnums, nfree, "ffblock", r, world) for r in range(world)]
        assert plans[0].col_lo == 0 and plans[-1].col_hi == nfree
        for a, b in zip(plans[:-1], plans[1:]):
            assert a.col_hi == b.col_lo
            # the node holding the first owned column of b has no free dof below it
            node = np.argwhere(d.dofnums == b.col_lo + 1)[0][0]
            fr = d.dofnums[node][d.dofnums[node] <= nfree]
            assert fr.min() == b.col_lo + 1
        for p in plans:
            assert np.all(np.diff(p.loc2glob) > 0)
            assert p.dofnums.min() == 1 and p.dofnums.max() == p.nall


def test_plan_with_more_ranks_than_node_runs():
    """Degenerate split: 2 triangles (4 nodes, some dofs fixed) over 7 ranks -- some ranks own nothing; the owned ranges
    still tile the columns and every element is assembled by the ranks that need it."""
    from fsb200 import partition as pt

    conn = np.array([[1, 2, 3], [2, 4, 3]])
    d = fx.DofField(4)
    d.setebc([0], 1)
    d.setebc([0], 2)
    d.numberdofs()
    nfree = int((~d.is_fixed).sum())
    plans = [pt.ColumnBlockPlan(conn, d.dofnums, nfree, "ffblock", r, 7) for r in range(7)]
    assert plans[0].col_lo == 0 and plans[-1].col_hi == nfree
    assert all(a.col_hi == b.col_lo for a, b in zip(plans[:-1], plans[1:]))
    assert any(p.empty for p in plans)
    for p in plans:
        if p.empty:
            assert len(p.elems) == 0
        else:
            assert len(p.elems) >= 1 and p.lcol_hi - p.lcol_lo == p.col_hi - p.col_lo
