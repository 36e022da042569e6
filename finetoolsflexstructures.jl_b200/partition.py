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


# ---- gathering assembled blocks into ONE global CSC (option A) ---------------
#
# Rank r owns a contiguous range of COLUMNS of the global matrix (aligned on node boundaries) and
# assembles every element that touches a node holding one of those columns, on a local mesh whose
# dofs are renumbered by an order-preserving map.  Its owned columns are then complete -- interface
# elements are computed by both neighbours instead of exchanging partial sums -- rows stay ascending
# after mapping back, and the global CSC is the concatenation of the blocks: every rank writes its block
# into its slice of the global arrays and the slices are exchanged in place in one batch of point-to-point
# transfers (NCCL over NVLink on GPUs, gloo in the CPU tests).

_FF_KINDS = ("ffblock", "ffblock_diag")


class ColumnBlockPlan:
    """What rank `rank` of `world` assembles and which columns of the global matrix it owns.

    Attributes (all host numpy arrays):
      col_lo, col_hi    owned columns [col_lo, col_hi) of the global matrix, 0-based
      elems             global 0-based ids of the elements assembled here (ascending)
      nodes             global 1-based ids of the local nodes (ascending)
      conn              local 1-based connectivity, (len(elems), nnpe)
      dofnums           local dof numbers (len(nodes), 6), 1-based; order-preserving renumbering of the global ones
      nfree, nall       local counts
      loc2glob          (nall,) global 1-based dof of local dof k+1  (the row map of fsgpu_result_block)
      lcol_lo, lcol_hi  the owned columns in local numbering, 0-based
    """

    def __init__(self, conn, dofnums, nfree, kind, rank, world):
        conn = np.asarray(conn, dtype=np.int64)
        dofnums = np.asarray(dofnums, dtype=np.int64)
        nn = dofnums.shape[0]
        nallg = dofnums.size
        self.rank, self.world, self.kind = int(rank), int(world), kind
        self.nfree_global, self.nall_global = int(nfree), int(nallg)
        self.ncols_global = int(nfree) if kind in _FF_KINDS else int(nallg)
        node_of_dof = np.empty(nallg, dtype=np.int64)
        node_of_dof[dofnums.ravel() - 1] = np.repeat(np.arange(nn), dofnums.shape[1])
        self._bounds = self._column_bounds(dofnums, node_of_dof, int(nfree), self.ncols_global, world)
        self.col_lo, self.col_hi = int(self._bounds[rank]), int(self._bounds[rank + 1])
        owned = np.unique(node_of_dof[self.col_lo : self.col_hi])  # 0-based nodes holding an owned column
        mark = np.zeros(nn + 1, dtype=bool)
        mark[owned + 1] = True
        self.elems = np.flatnonzero(mark[conn].any(axis=1)) if conn.size else np.zeros(0, dtype=np.int64)
        # nodes without elements still own (empty) columns
        self.nodes = np.unique(np.concatenate([conn[self.elems].ravel(), owned + 1]))
        lut = np.zeros(nn + 1, dtype=np.int64)
        lut[self.nodes] = np.arange(1, self.nodes.size + 1)
        self.conn = lut[conn[self.elems]]
        g = dofnums[self.nodes - 1]
        self.loc2glob = np.sort(g.ravel())
        self.dofnums = np.asfortranarray(np.searchsorted(self.loc2glob, g) + 1)
        self.nall = int(self.loc2glob.size)
        self.nfree = int(np.searchsorted(self.loc2glob, nfree, side="right"))
        self.lcol_lo = int(np.searchsorted(self.loc2glob, self.col_lo + 1))
        self.lcol_hi = int(np.searchsorted(self.loc2glob, self.col_hi + 1))
        assert self.lcol_hi - self.lcol_lo == self.col_hi - self.col_lo
        self.empty = self.col_hi == self.col_lo  # more ranks than node runs: nothing to assemble, nothing to contribute

    @staticmethod
    def _column_bounds(dofnums, node_of_dof, nfree, ncols, world):
        """Equal column counts, each interior boundary moved down to the first column of the run (free or
        prescribed dofs of one node, consecutive by `numberdofs!`) it falls into."""
        b = np.linspace(0, ncols, world + 1).astype(np.int64)
        for k in range(1, world):
            c = int(b[k])
            if c <= 0 or c >= ncols:
                continue
            d = dofnums[node_of_dof[c]]
            same = d[(d <= nfree) == (c + 1 <= nfree)]  # the node's dofs of the same class as column c
            lo = c + 1
            while lo - 1 in same:  # walk down the consecutive run
                lo -= 1
            b[k] = lo - 1
        return np.maximum.accumulate(b)

    def restrict_elements(self, a):
        """Per-element data (first axis = elements) of the elements assembled here."""
        return np.asarray(a)[self.elems]

    def restrict_nodes(self, a):
        """Nodal data (first axis = nodes) of the local nodes, e.g. geom0.values, nodal normals, u1, Rfield1."""
        return np.asarray(a)[self.nodes - 1]

    def local_ncols(self):
        return self.nfree if self.kind in _FF_KINDS else self.nall


def gather_blocks(fill, ncols_b, nnz_b, device, group=None):
    """Concatenate per-rank column blocks into the global CSC on every rank.

    `fill(colcount, rowval, nzval)` writes this rank's block (Int64 counts per column, Int64 1-based global rows,
    Float64 values) into the three tensors it is given -- slices of the global arrays, so nothing is copied
    before the collective.  Returns device tensors (colptr [ncols+1] Int64 1-based, rowval, nzval)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = torch.tensor([int(ncols_b), int(nnz_b)], dtype=torch.int64, device=device)
    allsz = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(allsz, sizes, group=group)
    allsz = torch.stack(allsz).cpu().numpy()
    coff = np.concatenate([[0], np.cumsum(allsz[:, 0])])
    zoff = np.concatenate([[0], np.cumsum(allsz[:, 1])])
    colptr = torch.empty(int(coff[-1]) + 1, dtype=torch.int64, device=device)
    rowval = torch.empty(int(zoff[-1]), dtype=torch.int64, device=device)
    nzval = torch.empty(int(zoff[-1]), dtype=torch.float64, device=device)
    counts = colptr[1:]
    fill(counts[coff[rank] : coff[rank + 1]], rowval[zoff[rank] : zoff[rank + 1]], nzval[zoff[rank] : zoff[rank + 1]])
    # every rank sends its slices to all peers and receives theirs in place, as ONE batch: all transfers are in
    # flight together (both directions of every NVLink; NVSwitch gives each pair full bandwidth)
    grank = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    ops = []
    for buf, off in ((counts, coff), (rowval, zoff), (nzval, zoff)):
        mine = buf[off[rank] : off[rank + 1]]
        for r in range(world):
            if r == rank:
                continue
            if mine.numel():
                ops.append(dist.P2POp(dist.isend, mine, grank(r), group))
            if off[r + 1] > off[r]:
                ops.append(dist.P2POp(dist.irecv, buf[off[r] : off[r + 1]], grank(r), group))
    if ops:
        for wk in dist.batch_isend_irecv(ops):
            wk.wait()
    colptr[0] = 1
    counts.cumsum_(0)
    counts += 1
    return colptr, rowval, nzval


def gather_matrix(ctx, plan, device, group=None, to_host=True):
    """`makematrix!` of an element-partitioned assembly: the GLOBAL SparseMatrixCSC from the per-rank results
    of contexts that assembled `plan`'s local meshes (every rank gets the whole matrix: the solver hand-off)."""
    import torch

    from .context import SparseMatrixCSC

    if plan.kind not in ("sparse", "diag") + _FF_KINDS:
        raise ValueError("gather_matrix: targets sparse / ffblock / diag (symmetrise a gathered 'sparse' matrix on the host instead)")
    row_map = torch.as_tensor(plan.loc2glob, device=device)
    if row_map.is_cuda:
        torch.cuda.current_stream(row_map.device).synchronize()  # the library works on its own stream
    nnz_b = 0 if plan.empty else ctx.result_block(plan.lcol_lo, plan.lcol_hi)

    def fill(cnt, rv, nz):
        if plan.empty:
            return
        ctx.result_block(plan.lcol_lo, plan.lcol_hi, row_map, cnt if cnt.numel() else None, rv if rv.numel() else None, nz if nz.numel() else None)

    colptr, rowval, nzval = gather_blocks(fill, plan.col_hi - plan.col_lo, nnz_b, device, group)
    n = plan.ncols_global
    if not to_host:
        return n, colptr, rowval, nzval
    return SparseMatrixCSC(n, n, colptr.cpu().numpy(), rowval.cpu().numpy(), nzval.cpu().numpy())


def gather_vector(local_vec, plan, device, group=None):
    """Global vector (SysvecAssembler: nalldofs, or FBlock: nfree -- as `plan.kind` says) from per-rank vectors
    over the local dofs: every rank contributes the entries of the dofs it owns."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    b = plan._bounds
    out = torch.empty(plan.ncols_global, dtype=torch.float64, device=device)
    out[b[rank] : b[rank + 1]] = local_vec[plan.lcol_lo : plan.lcol_hi]
    grank = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    mine = out[b[rank] : b[rank + 1]]
    ops = []
    for r in range(world):
        if r == rank:
            continue
        if mine.numel():
            ops.append(dist.P2POp(dist.isend, mine, grank(r), group))
        if b[r + 1] > b[r]:
            ops.append(dist.P2POp(dist.irecv, out[b[r] : b[r + 1]], grank(r), group))
    if ops:
        for wk in dist.batch_isend_irecv(ops):
            wk.wait()
    return out


# ---- row-partitioned explicit loop (fsgpu_explicit_create_dist) -----------------------------------
#
# The exchange itself lives in the library (peer-mapped windows written by the step kernel); the host only
# carries the opaque address blobs between the ranks once.


def connect_ranks(ex, group=None):
    """All-gather the export blobs of a row-partitioned `Explicit` over torch.distributed (any backend) and
    connect: one process per GPU."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(ex.export()).to(dev)
    allb = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allb, mine, group=group)
    ex.connect([b.cpu().numpy() for b in allb])


def connect_local(exs):
    """Contexts of ONE process (several GPUs, or several partitions on one GPU in the tests): every collective call
    -- connect included -- must be issued concurrently, one host thread per rank."""
    blobs = [e.export() for e in exs]
    run_collective([lambda e=e: e.connect(blobs) for e in exs])


def run_collective(calls):
    """Run one callable per rank concurrently (ctypes releases the GIL during library calls); returns their results."""
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=len(calls)) as pool:
        futs = [pool.submit(c) for c in calls]
        errs = [(r, f.exception()) for r, f in enumerate(futs)]
        errs = [(r, e) for r, e in errs if e is not None]
        if errs:  # a rank that fails makes its peers time out: report every rank's own error
            raise RuntimeError("; ".join(f"rank {r}: {e!r}" for r, e in errs)) from errs[0][1]
        return [f.result() for f in futs]


def rcm_permutation(conn, nnodes):
    """Reverse Cuthill-McKee node order of the mesh graph (`symrcm` of plate_expl_examples.jl:119-121): the
    permutation handed to `numberdofs!` so that contiguous dof ranges are compact strips of the mesh."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee

    c = np.asarray(conn, dtype=np.int64) - 1
    nn = c.shape[1]
    i = np.repeat(c, nn, axis=1).ravel()
    j = np.tile(c, (1, nn)).ravel()
    g = sp.csr_matrix((np.ones(i.size, dtype=np.int8), (i, j)), shape=(nnodes, nnodes))
    return np.asarray(reverse_cuthill_mckee(g, symmetric_mode=True), dtype=np.int64)
