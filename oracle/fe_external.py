"""Restated EXTERNAL semantics (FinEtools 8.2.5 / FinEtoolsDeforLinear 3.0.6 /
Julia SparseArrays) that the reference hot path calls but does not contain.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

None of this code exists under /root/reference; only the call sites do
(SURVEY.md App. A).  Each function names the call site it serves.
All index arrays are Int64 and 1-based, as on the Julia side.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------
# small matrix utilities (call sites: src/FEMMShellT3FFModule.jl:378-380,
# src/RotUtilModule.jl:35)
# ----------------------------------------------------------------------------


def skewmat(a):
    """Skew matrix of a (...,3) vector array."""
    a = np.asarray(a, dtype=np.float64)
    S = np.zeros(a.shape[:-1] + (3, 3))
    S[..., 0, 1] = -a[..., 2]
    S[..., 0, 2] = a[..., 1]
    S[..., 1, 0] = a[..., 2]
    S[..., 1, 2] = -a[..., 0]
    S[..., 2, 0] = -a[..., 1]
    S[..., 2, 1] = a[..., 0]
    return S


def rotmat3(a):
    """`rotmat3!(R, a)`: rotation matrix of the rotation vector `a` (Rodrigues):
    R = cos|a| (I - n n') + sin|a| skew(n) + n n',  n = a/|a|;  identity for a = 0.
    Batched over leading axes."""
    a = np.asarray(a, dtype=np.float64)
    na = np.sqrt(np.sum(a * a, axis=-1))
    safe = np.where(na > 0.0, na, 1.0)
    n = a / safe[..., None]
    ca = np.cos(na)[..., None, None]
    sa = np.sin(na)[..., None, None]
    nn = n[..., :, None] * n[..., None, :]
    eye = np.broadcast_to(np.eye(3), nn.shape)
    R = ca * (eye - nn) + sa * skewmat(n) + nn
    R = np.where((na > 0.0)[..., None, None], R, eye)
    return R


# ----------------------------------------------------------------------------
# mesh generators (App. A.5; call sites test/test_shell_statics.jl:27,
# test/test_q4rs_shell_statics.jl:27)
# ----------------------------------------------------------------------------


def _block_nodes(L, W, nL, nW):
    xs = np.linspace(0.0, L, nL + 1)
    ys = np.linspace(0.0, W, nW + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")  # x fastest
    return np.column_stack([X.ravel(), Y.ravel()])


def t3block(L, W, nL, nW):
    """`T3block(L, W, nL, nW, :a)`: nodes x-fastest; cell (i,j), f=(j-1)(nL+1)+i gives
    triangles [f, f+1, f+nL+1] and [f+1, f+nL+2, f+nL+1].  Loop order i outer, j inner."""
    xy = _block_nodes(L, W, nL, nW)
    conn = np.empty((2 * nL * nW, 3), dtype=np.int64)
    k = 0
    for i in range(1, nL + 1):
        for j in range(1, nW + 1):
            f = (j - 1) * (nL + 1) + i
            conn[k] = (f, f + 1, f + nL + 1)
            conn[k + 1] = (f + 1, f + nL + 2, f + nL + 1)
            k += 2
    return xy, conn


def q4block(L, W, nL, nW):
    """`Q4block`: same node order, quads [f, f+1, f+nL+2, f+nL+1]."""
    xy = _block_nodes(L, W, nL, nW)
    conn = np.empty((nL * nW, 4), dtype=np.int64)
    k = 0
    for i in range(1, nL + 1):
        for j in range(1, nW + 1):
            f = (j - 1) * (nL + 1) + i
            conn[k] = (f, f + 1, f + nL + 2, f + nL + 1)
            k += 1
    return xy, conn


def xyz3(xy):
    xy = np.asarray(xy, dtype=np.float64)
    if xy.shape[1] == 3:
        return xy.copy()
    return np.column_stack([xy, np.zeros(xy.shape[0])])


def selectnode_box(xyz, box, inflate=0.0):
    """`selectnode(fens; box=..., inflate=...)` -> 0-based node indices."""
    b = np.asarray(box, dtype=np.float64).reshape(-1, 2)
    m = np.ones(xyz.shape[0], dtype=bool)
    for d in range(b.shape[0]):
        m &= (xyz[:, d] >= b[d, 0] - inflate) & (xyz[:, d] <= b[d, 1] + inflate)
    return np.nonzero(m)[0]


# ----------------------------------------------------------------------------
# fields / dof numbering (App. A.1; call sites `numberdofs!`, `gatherdofnums!`
# src/FEMMShellT3FFModule.jl:732)
# ----------------------------------------------------------------------------


class DofField:
    """Minimal NodalField for the generalized-displacement field `dchi` (nnodes x 6)."""

    def __init__(self, nnodes, ndof=6):
        self.values = np.zeros((nnodes, ndof))
        self.is_fixed = np.zeros((nnodes, ndof), dtype=bool)
        self.dofnums = np.zeros((nnodes, ndof), dtype=np.int64)
        self.nfreedofs = 0

    def setebc(self, nodes, comp):
        """`setebc!(f, nodes, true, comp)` with comp 1-based, zero prescribed value."""
        self.is_fixed[np.asarray(nodes, dtype=np.int64), comp - 1] = True

    def numberdofs(self, perm=None):
        """`numberdofs!(f[, perm])`: free dofs first, node-major in visiting order,
        prescribed dofs continue from nfree+1 in the same visiting order."""
        nn, nd = self.is_fixed.shape
        order = np.arange(nn) if perm is None else np.asarray(perm, dtype=np.int64)
        fixed = self.is_fixed[order]  # visiting order
        free_flat = (~fixed).ravel()
        nfree = int(free_flat.sum())
        nums = np.empty(nn * nd, dtype=np.int64)
        nums[free_flat] = np.arange(1, nfree + 1)
        nums[~free_flat] = np.arange(nfree + 1, nn * nd + 1)
        self.dofnums[order] = nums.reshape(nn, nd)
        self.nfreedofs = nfree
        return self

    @property
    def nalldofs(self):
        return self.dofnums.size

    def gatherdofnums(self, conn):
        """(nelem, nnpe) 1-based conn -> (nelem, nnpe*6) dof numbers, node-major."""
        return self.dofnums[np.asarray(conn) - 1].reshape(len(conn), -1)

    def scattersysvec(self, u):
        free = self.dofnums <= self.nfreedofs
        self.values[free] = u[self.dofnums[free] - 1]
        return self


# ----------------------------------------------------------------------------
# materials (App. A.6; call sites `tangentmoduli!`
# src/FEMMShellT3FFModule.jl:328, src/CompositeLayupModule.jl:78)
# ----------------------------------------------------------------------------


def moduli_iso(E, nu):
    """MatDeforElastIso 3-D tangent moduli, strain order [xx,yy,zz,xy,xz,yz]."""
    lam = E * nu / (1 + nu) / (1 - 2 * nu)
    mu = E / 2.0 / (1 + nu)
    m1 = np.array([1.0, 1, 1, 0, 0, 0])
    D = lam * np.outer(m1, m1) + 2 * mu * np.eye(6)
    for k in (3, 4, 5):
        D[k, k] = mu
    return D


def moduli_ortho(E1, E2, E3, nu12, nu13, nu23, G12, G13, G23):
    """MatDeforElastOrtho: D = inverse of the 6x6 compliance."""
    C = np.zeros((6, 6))
    C[0, 0] = 1 / E1
    C[0, 1] = C[1, 0] = -nu12 / E1
    C[0, 2] = C[2, 0] = -nu13 / E1
    C[1, 1] = 1 / E2
    C[1, 2] = C[2, 1] = -nu23 / E2
    C[2, 2] = 1 / E3
    C[3, 3] = 1 / G12
    C[4, 4] = 1 / G13
    C[5, 5] = 1 / G23
    return np.linalg.inv(C)


# ----------------------------------------------------------------------------
# Q4 shape functions and rules (App. A.4; call sites `integrationdata`
# src/FEMMShellQ4RSModule.jl:480,894)
# ----------------------------------------------------------------------------


def q4_shape(xi, eta):
    N = 0.25 * np.array(
        [(1 - xi) * (1 - eta), (1 + xi) * (1 - eta), (1 + xi) * (1 + eta), (1 - xi) * (1 + eta)]
    )
    dN = 0.25 * np.array(
        [
            [-(1 - eta), -(1 - xi)],
            [(1 - eta), -(1 + xi)],
            [(1 + eta), (1 + xi)],
            [-(1 + eta), (1 - xi)],
        ]
    )
    return N, dN


def gauss_rule_2x2():
    """GaussRule(2, 2): tensor product, first coordinate outer loop."""
    g = 0.577350269189626
    pc = np.array([[-g, -g], [-g, g], [g, -g], [g, g]])
    w = np.ones(4)
    return pc, w


def gauss_rule_1x1():
    return np.array([[0.0, 0.0]]), np.array([4.0])


def simpson13_rule_2d():
    """Simpson13Rule(2): 3x3 tensor product, 1-D weights (1/3, 4/3, 1/3)."""
    p1 = np.array([-1.0, 0.0, 1.0])
    w1 = np.array([1.0, 4.0, 1.0]) / 3.0
    pc, w = [], []
    for i in range(3):
        for j in range(3):
            pc.append([p1[i], p1[j]])
            w.append(w1[i] * w1[j])
    return np.array(pc), np.array(w)


def nodal_rule_q4():
    """NodalTensorProductRule(2): points at the 4 nodes in node order, unit weights."""
    return np.array([[-1.0, -1], [1, -1], [1, 1], [-1, 1]]), np.ones(4)


# ----------------------------------------------------------------------------
# assemblers (App. A.2; call sites startassembly!/assemble!/makematrix!
# src/FEMMShellT3FFModule.jl:664,733,735; src/AssemblyModule.jl:28-53)
# ----------------------------------------------------------------------------


def coo_full(elmats, dofnums):
    """SysmatAssemblerSparse.assemble!: per element, `for j in cols, for i in rows`
    append (dof_row[i], dof_col[j], mat[i, j]) -- all entries incl. zeros."""
    elmats = np.asarray(elmats)
    ne, n, _ = elmats.shape
    dn = np.asarray(dofnums, dtype=np.int64).reshape(ne, n)
    I = np.broadcast_to(dn[:, None, :], (ne, n, n))  # [e, j, i] = dn[e, i]
    J = np.broadcast_to(dn[:, :, None], (ne, n, n))  # [e, j, i] = dn[e, j]
    V = np.transpose(elmats, (0, 2, 1))  # [e, j, i] = mat[i, j]
    return I.reshape(-1).copy(), J.reshape(-1).copy(), V.reshape(-1).copy()


def coo_symm(elmats, dofnums):
    """SysmatAssemblerSparseSymm.assemble!: only local i >= j triples, j outer."""
    elmats = np.asarray(elmats)
    ne, n, _ = elmats.shape
    dn = np.asarray(dofnums, dtype=np.int64).reshape(ne, n)
    jj, ii = np.nonzero(np.triu(np.ones((n, n), dtype=bool)))  # rows of this = j, cols = i>=j
    I = dn[:, ii]
    J = dn[:, jj]
    V = elmats[:, ii, jj]
    return I.reshape(-1).copy(), J.reshape(-1).copy(), V.reshape(-1).copy()


def coo_diag(elmats, dofnums):
    """SysmatAssemblerSparseDiag.assemble!: only (d_j, d_j, mat[j, j])."""
    elmats = np.asarray(elmats)
    ne, n, _ = elmats.shape
    dn = np.asarray(dofnums, dtype=np.int64).reshape(ne, n)
    V = np.einsum("ejj->ej", elmats)
    return dn.reshape(-1).copy(), dn.reshape(-1).copy(), V.reshape(-1).copy()


def sparse_csc(I, J, V, m, n):
    """Julia `sparse(I, J, V, m, n)`: CSC, row indices ascending within a column,
    duplicates combined with `+`, numerical zeros RETAINED as stored entries.
    Returns 1-based (colptr[n+1], rowval[nnz], nzval[nnz])."""
    I = np.asarray(I, dtype=np.int64)
    J = np.asarray(J, dtype=np.int64)
    V = np.asarray(V, dtype=np.float64)
    if I.size and (I.min() < 1 or I.max() > m or J.min() < 1 or J.max() > n):
        raise ValueError("degree of freedom out of range")
    key = (J - 1) * np.int64(m) + (I - 1)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    vs = V[order]
    if ks.size == 0:
        return np.ones(n + 1, dtype=np.int64), np.zeros(0, np.int64), np.zeros(0)
    first = np.ones(ks.size, dtype=bool)
    first[1:] = ks[1:] != ks[:-1]
    starts = np.nonzero(first)[0]
    nzval = np.add.reduceat(vs, starts)
    uk = ks[starts]
    rowval = uk % m + 1
    cols = uk // m
    counts = np.bincount(cols, minlength=n)
    colptr = np.ones(n + 1, dtype=np.int64)
    colptr[1:] = 1 + np.cumsum(counts)
    return colptr, rowval.astype(np.int64), nzval


def csc_to_scipy(colptr, rowval, nzval, m, n):
    import scipy.sparse as sp

    return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(m, n))


def csc_block_ff(colptr, rowval, nzval, nfr, nfc):
    """`matrix_blocked(A, nfr, nfc)[:ff]` = A[1:nfr, 1:nfc]; stored zeros are kept."""
    cp = colptr[: nfc + 1]
    lo, hi = cp[0] - 1, cp[-1] - 1
    rv = rowval[lo:hi]
    nz = nzval[lo:hi]
    keep = rv <= nfr
    colid = np.repeat(np.arange(nfc), np.diff(cp))
    counts = np.bincount(colid[keep], minlength=nfc)
    ncp = np.ones(nfc + 1, dtype=np.int64)
    ncp[1:] = 1 + np.cumsum(counts)
    return ncp, rv[keep].copy(), nz[keep].copy()


def csc_symm_finish(colptr, rowval, nzval, n):
    """SysmatAssemblerSparseSymm.makematrix!: S = S + transpose(S); S[j,j] *= 0.5.
    Sparse `+` (Julia `map`, zero-preserving path) DROPS results that are exactly
    0.0 (either sign), so the stored pattern is value dependent; `S[j,j] *= 0.5`
    on a non-stored diagonal entry stores nothing.  [doc -- parity unpinned]"""
    I, J, V = findnz_csc(colptr, rowval, nzval)
    cp, rv, nz = sparse_csc(np.concatenate([I, J]), np.concatenate([J, I]),
                            np.concatenate([V, V]), n, n)
    colid = np.repeat(np.arange(1, n + 1, dtype=np.int64), np.diff(cp))
    keep = nz != 0.0
    cp2, rv2, nz2 = sparse_csc(rv[keep], colid[keep], nz[keep], n, n)
    colid2 = np.repeat(np.arange(1, n + 1, dtype=np.int64), np.diff(cp2))
    nz2 = np.where(rv2 == colid2, nz2 * 0.5, nz2)
    return cp2, rv2, nz2


def findnz_csc(colptr, rowval, nzval):
    """`findnz(S)`: column-major order triples, stored zeros kept."""
    n = len(colptr) - 1
    J = np.repeat(np.arange(1, n + 1, dtype=np.int64), np.diff(colptr))
    return rowval.copy(), J, nzval.copy()


def sparsecsr(I, J, V, m, n):
    """`SparseMatricesCSR.sparsecsr(I, J, V, m, n)` -> 1-based (rowptr, colval, nzval);
    implemented (as the package does) as the CSC build of the transpose."""
    rowptr, colval, nz = sparse_csc(J, I, V, n, m)
    return rowptr, colval, nz


def assemble_matrix(kind, elmats, dofnums, nall, nfree=None):
    """Emulate startassembly!/assemble!/makematrix! for an assembler `kind`:
    'sparse' | 'symm' | 'diag' | 'ffblock' | 'ffblock_diag' | 'csrsymm'."""
    if kind == "sparse":
        return sparse_csc(*coo_full(elmats, dofnums), nall, nall)
    if kind == "symm":
        cp, rv, nz = sparse_csc(*coo_symm(elmats, dofnums), nall, nall)
        return csc_symm_finish(cp, rv, nz, nall)
    if kind == "diag":
        return sparse_csc(*coo_diag(elmats, dofnums), nall, nall)
    if kind == "ffblock":
        cp, rv, nz = sparse_csc(*coo_full(elmats, dofnums), nall, nall)
        return csc_block_ff(cp, rv, nz, nfree, nfree)
    if kind == "ffblock_diag":
        cp, rv, nz = sparse_csc(*coo_diag(elmats, dofnums), nall, nall)
        return csc_block_ff(cp, rv, nz, nfree, nfree)
    if kind == "csrsymm":
        cp, rv, nz = sparse_csc(*coo_full(elmats, dofnums), nall, nall)
        return sparsecsr(*findnz_csc(cp, rv, nz), nall, nall)
    raise ValueError(kind)


def assemble_vector(elvecs, dofnums, nall, nfree=None):
    """SysvecAssembler (nfree None) / SysvecAssemblerFBlock(nfree): F[d] += v for
    1 <= d <= n, sequentially in element order."""
    n = nall if nfree is None else nfree
    F = np.zeros(n)
    d = np.asarray(dofnums, dtype=np.int64).reshape(-1)
    v = np.asarray(elvecs, dtype=np.float64).reshape(-1)
    keep = (d >= 1) & (d <= n)
    np.add.at(F, d[keep] - 1, v[keep])
    return F


# ----------------------------------------------------------------------------
# consistent loads used only to reproduce the reference's known-answer tests
# (`distribloads`, test/test_shell_statics.jl:82-84)
# ----------------------------------------------------------------------------


def distribloads_t3(xyz, conn, f6):
    """Uniform force intensity on T3 surface: each node gets f * A / 3."""
    c = np.asarray(conn) - 1
    e1 = xyz[c[:, 1]] - xyz[c[:, 0]]
    e2 = xyz[c[:, 2]] - xyz[c[:, 0]]
    A = 0.5 * np.linalg.norm(np.cross(e1, e2), axis=1)
    nn = xyz.shape[0]
    F = np.zeros((nn, 6))
    for a in range(3):
        np.add.at(F, c[:, a], (A / 3.0)[:, None] * np.asarray(f6)[None, :])
    return F


def distribloads_q4(xyz, conn, f6):
    """Uniform force intensity on Q4 surface with GaussRule(2,2)."""
    c = np.asarray(conn) - 1
    pc, w = gauss_rule_2x2()
    nn = xyz.shape[0]
    F = np.zeros((nn, 6))
    X = xyz[c]  # (ne,4,3)
    for q in range(4):
        N, dN = q4_shape(*pc[q])
        J = np.einsum("eai,ak->eik", X, dN)  # (ne,3,2)
        Jac = np.linalg.norm(np.cross(J[:, :, 0], J[:, :, 1]), axis=1)
        for a in range(4):
            np.add.at(F, c[:, a], (N[a] * Jac * w[q])[:, None] * np.asarray(f6)[None, :])
    return F


def solve_blocked(K_csc, F_nodal, dchi: DofField):
    """`solve_blocked!(dchi, K, F)` with zero prescribed values: K_ff u_f = F_f."""
    import scipy.sparse.linalg as spla

    nf = dchi.nfreedofs
    Fv = np.zeros(dchi.nalldofs)
    Fv[dchi.dofnums.ravel() - 1] = F_nodal.ravel()
    Kff = K_csc[:nf, :nf].tocsc()
    u = spla.spsolve(Kff, Fv[:nf])
    dchi.scattersysvec(u)
    return dchi


def mergenodes(xyz, conn, tolerance):
    """FinEtools `mergenodes(fens, fes, tolerance)` (MeshModificationModule, un-vendored FinEtools 8.2.5): nodes closer than
    `tolerance` are fused onto the lowest-numbered one of the cluster, the node set is compacted in order and the
    connectivity renumbered.  Restated from the documented behaviour; used on meshes with exactly coincident MPC nodes
    (test/test_shell_statics.jl:598), where every fusing criterion gives the same result."""
    from scipy.spatial import cKDTree

    rep = np.arange(xyz.shape[0])
    for a, b in sorted(cKDTree(xyz).query_pairs(tolerance)):
        rep[max(a, b)] = rep[min(a, b)]
    keep = np.unique(rep)
    remap = -np.ones(xyz.shape[0], dtype=np.int64)
    remap[keep] = np.arange(len(keep))
    return xyz[keep], remap[rep[np.asarray(conn) - 1]] + 1


def field_from_integpoints_invdist(xyz, conn, loc, values):
    """Nodal field from integration-point values, FinEtools `fieldfromintegpoints(...; nodevalmethod = :invdistance)`
    (FEMMBaseModule, un-vendored FinEtools 8.2.5; the reference's shells feed it through their `inspectintegpoints`,
    src/FEMMShellT3FFModule.jl:850-962): every integration point adds value / (d + dmin / 1e9) to the nodes of its element,
    d the SQUARED distance node - point, dmin the smallest positive one of the element; the node value is the weighted mean.
    Restated from the published algorithm; that the weight is the squared distance is pinned by the goldens of
    test/test_shell_statics.jl:690-728 (the plain distance misses them by up to 80 %).
    conn (ne, nn) 1-based, loc (ne, npts, 3), values (ne, npts, ncomp) -> (nnodes, ncomp)."""
    c = np.asarray(conn) - 1
    loc = np.asarray(loc, dtype=np.float64).reshape(c.shape[0], -1, 3)
    values = np.asarray(values, dtype=np.float64).reshape(c.shape[0], loc.shape[1], -1)
    num = np.zeros((xyz.shape[0], values.shape[2]))
    den = np.zeros(xyz.shape[0])
    for q in range(loc.shape[1]):
        d = ((xyz[c] - loc[:, q, None, :]) ** 2).sum(axis=2)
        dmin = np.where(d > 0, d, np.inf).min(axis=1) / 1.0e9
        w = 1.0 / (d + dmin[:, None])
        for a in range(c.shape[1]):
            np.add.at(num, c[:, a], w[:, a, None] * values[:, q, :])
            np.add.at(den, c[:, a], w[:, a])
    out = np.zeros_like(num)
    nz = den > 0
    out[nz] = num[nz] / den[nz, None]
    return out


def q4_to_t3(conn4):
    """FinEtools `Q4toT3(fens, fes)` (MeshModificationModule, un-vendored FinEtools 8.2.5), default orientation: quad
    (1, 2, 3, 4) -> triangles (1, 2, 3) and (1, 3, 4), two per quad in quad order.  Pinned by the Raasch hook goldens of
    test/test_shell_statics.jl:226-231 (the alternate diagonal misses them by 3e-6 .. 1e-5, this one matches to 1e-10)."""
    c = np.asarray(conn4, dtype=np.int64)
    out = np.empty((2 * c.shape[0], 3), dtype=np.int64)
    out[0::2] = c[:, [0, 1, 2]]
    out[1::2] = c[:, [0, 2, 3]]
    return out


def boundary_edges(conn):
    """FinEtools `meshboundary` of a surface mesh: the element edges that belong to exactly one element, as node pairs."""
    c = np.asarray(conn, dtype=np.int64)
    n = c.shape[1]
    e = np.concatenate([np.stack([c[:, i], c[:, (i + 1) % n]], axis=1) for i in range(n)])
    key = np.sort(e, axis=1)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    return e[cnt[inv.ravel()] == 1]
