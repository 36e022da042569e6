"""Deterministic synthetic workloads of the named sizes (BASELINE.json configs, SURVEY
section 8(d)); vectorised so the 1M-4M element meshes build in a few seconds.
Connectivity order and node numbering follow FinEtools' T3block / Q4block (App. A.5)."""
from __future__ import annotations

import numpy as np


def _block_nodes(L, W, nL, nW):
    xs = np.linspace(0.0, L, nL + 1)
    ys = np.linspace(0.0, W, nW + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    return np.column_stack([X.ravel(), Y.ravel(), np.zeros(X.size)])


def _cells(nL, nW):
    i = np.repeat(np.arange(1, nL + 1, dtype=np.int64), nW)  # i outer
    j = np.tile(np.arange(1, nW + 1, dtype=np.int64), nL)  # j inner
    return (j - 1) * (nL + 1) + i


def q4block(L, W, nL, nW):
    f = _cells(nL, nW)
    return _block_nodes(L, W, nL, nW), np.column_stack([f, f + 1, f + nL + 2, f + nL + 1])


def t3block(L, W, nL, nW):
    f = _cells(nL, nW)
    conn = np.empty((2 * f.size, 3), dtype=np.int64)
    conn[0::2] = np.column_stack([f, f + 1, f + nL + 1])
    conn[1::2] = np.column_stack([f + 1, f + nL + 2, f + nL + 1])
    return _block_nodes(L, W, nL, nW), conn


def number_dofs(is_fixed, perm=None):
    nn, nd = is_fixed.shape
    order = np.arange(nn) if perm is None else perm
    free = ~is_fixed[order].ravel()
    nfree = int(free.sum())
    nums = np.empty(nn * nd, dtype=np.int64)
    nums[free] = np.arange(1, nfree + 1)
    nums[~free] = np.arange(nfree + 1, nn * nd + 1)
    dof = np.empty((nn, nd), dtype=np.int64)
    dof[order] = nums.reshape(nn, nd)
    return np.asfortranarray(dof), nfree


def c2_q4rs_plate(n=1000, seed=0):
    """C2: Q4RS square plate, n x n quads (1M at n = 1000), L = 10, E = 30e6, nu = 0.3, t = 0.1,
    interior nodes perturbed in-plane by <= L/n/5, z = 0.05 sin(pi x/L) sin(pi y/L); hard simple
    support (examples/shells/statics/homogeneous/plates/simply_supp_square_plate_udl_examples.jl:565-590)."""
    L = 10.0
    xyz, conn = q4block(L, L, n, n)
    rng = np.random.Generator(np.random.PCG64(seed))
    tol = L / n / 100
    bx = (np.abs(xyz[:, 0]) < tol) | (np.abs(xyz[:, 0] - L) < tol)
    by = (np.abs(xyz[:, 1]) < tol) | (np.abs(xyz[:, 1] - L) < tol)
    interior = ~(bx | by)
    shift = L / n / 5
    d = 2 * (rng.random((xyz.shape[0], 2)) - 0.5) * shift
    xyz[interior, 0] += d[interior, 0]
    xyz[interior, 1] += d[interior, 1]
    xyz[:, 2] = 0.05 * np.sin(np.pi * xyz[:, 0] / L) * np.sin(np.pi * xyz[:, 1] / L)
    fixed = np.zeros((xyz.shape[0], 6), dtype=bool)
    for c in (1, 2, 3, 4, 6):
        fixed[bx, c - 1] = True
    for c in (1, 2, 3, 5, 6):
        fixed[by, c - 1] = True
    dof, nfree = number_dofs(fixed)
    return dict(name=f"C2 Q4RS plate {n}x{n}", kind="q4", xyz=np.asfortranarray(xyz), conn=conn, dofnums=dof, nfree=nfree,
                E=30e6, nu=0.3, rho=1.0, thickness=0.1)


def c4_t3ff_panel(nx=2000, ny=1000):
    """C4: T3FF shallow cylindrical panel 2 x 1, z = R (cos(x/R) - 1), R = 5, aluminium
    (examples/shells/dynamics/homogeneous/explicit/plate_expl_examples.jl:35-46), clamped boundary."""
    Lx, Ly, R = 2.0, 1.0, 5.0
    xyz, conn = t3block(Lx, Ly, nx, ny)
    x = xyz[:, 0].copy()
    xyz[:, 0] = R * np.sin(x / R)
    xyz[:, 2] = R * (np.cos(x / R) - 1)
    tol = 1e-9
    b = (x < tol) | (x > Lx - tol) | (xyz[:, 1] < tol) | (xyz[:, 1] > Ly - tol)
    fixed = np.zeros((xyz.shape[0], 6), dtype=bool)
    fixed[b, :] = True
    dof, nfree = number_dofs(fixed)
    return dict(name=f"C4 T3FF panel {nx}x{ny}x2", kind="t3", xyz=np.asfortranarray(xyz), conn=conn, dofnums=dof, nfree=nfree,
                E=68e9, nu=0.33, rho=2660.0, thickness=1e-3)


def c4_strip(rank, world, nx=2000, ny=1000, Ly=1.0):
    """Element-partitioned C4: the global panel is `world` strips of ny cell rows stacked in y;
    this returns rank `rank`'s strip in local numbering plus the local free-dof indices (0-based)
    of the bottom / top interface rows, ordered along x (identical order on both sides)."""
    Lx, R = 2.0, 5.0
    xyz, conn = t3block(Lx, Ly, nx, ny)
    x = xyz[:, 0].copy()
    xyz[:, 0] = R * np.sin(x / R)
    xyz[:, 1] += rank * Ly
    xyz[:, 2] = R * (np.cos(x / R) - 1)
    tol = 1e-9
    yl = xyz[:, 1] - rank * Ly
    b = (x < tol) | (x > Lx - tol)
    if rank == 0:
        b |= yl < tol
    if rank == world - 1:
        b |= yl > Ly - tol
    fixed = np.zeros((xyz.shape[0], 6), dtype=bool)
    fixed[b, :] = True
    dof, nfree = number_dofs(fixed)
    bottom = np.arange(0, nx + 1)
    top = ny * (nx + 1) + np.arange(0, nx + 1)

    def free_dofs(nodes):
        d = dof[nodes].ravel()
        return (d[d <= nfree] - 1).astype(np.int64)

    return dict(name=f"C4 T3FF strip {rank}/{world} {nx}x{ny}x2", kind="t3", xyz=np.asfortranarray(xyz), conn=conn, dofnums=dof,
                nfree=nfree, E=68e9, nu=0.33, rho=2660.0, thickness=1e-3,
                lo_dofs=free_dofs(bottom) if rank > 0 else np.zeros(0, np.int64),
                hi_dofs=free_dofs(top) if rank < world - 1 else np.zeros(0, np.int64),
                lo_nodes=bottom if rank > 0 else np.zeros(0, np.int64), hi_nodes=top if rank < world - 1 else np.zeros(0, np.int64))


def c1_small_t3(n=64):
    """CPU-runnable small case (stand-in for C1's size class): curved T3 panel."""
    w = c4_t3ff_panel(n, n)
    w["name"] = f"small T3FF panel {n}x{n}x2"
    return w


def cylindrical_csys(locs):
    """Layup csys of the cylinder examples (`cylindrical!`, examples/shells/dynamics/homogeneous/explicit/
    clamp_cyl_expl_examples.jl:62-68): e1 = axial (z), e3 = radial (outward), e2 = e3 x e1; (n, 3) -> (n, 3, 3)."""
    locs = np.asarray(locs, dtype=np.float64)
    er = np.column_stack([locs[:, 0], locs[:, 1], np.zeros(len(locs))])
    er /= np.linalg.norm(er, axis=1)[:, None]
    ez = np.tile([0.0, 0.0, 1.0], (len(locs), 1))
    return np.stack([ez, np.cross(er, ez), er], axis=2)


def c3_t3ffcomp_cylinder(ncirc=1000, nlen=1000):
    """C3: laminated cylinder R = 0.1, L = 0.8 (examples/shells/dynamics/homogeneous/explicit/
    clamp_cyl_expl_examples.jl:59-84), T3block(360 deg, L, ncirc, nlen) wrapped with the seam merged
    -> 2*ncirc*nlen triangles; 4 plies [0/90/90/0] of t/4, t = R/100, lamina of
    test/test_composite_shell_statics.jl:21-27, rho = 1500; layup csys = cylindrical (axial, hoop,
    radial) evaluated per element centroid; one end clamped."""
    R, L = 0.1, 0.8
    xy, conn = t3block(2 * np.pi, L, ncirc, nlen)
    nrow = ncirc + 1
    # merge the seam: node (i = ncirc, j) -> node (i = 0, j); renumber compactly
    ids = np.arange(xy.shape[0])
    i = ids % nrow
    j = ids // nrow
    keep = i < ncirc
    new = np.where(keep, j * ncirc + i, j * ncirc)  # 0-based new ids
    conn = new[conn - 1] + 1
    a = xy[keep, 0]
    z = xy[keep, 1]
    xyz = np.column_stack([R * np.cos(a), R * np.sin(a), z])
    fixed = np.zeros((xyz.shape[0], 6), dtype=bool)
    fixed[z < 1e-12, :] = True
    dof, nfree = number_dofs(fixed)
    # cylindrical csys at element centroids: e1 = axial (z), e3 = radial (outward), e2 = e3 x e1
    c = xyz[conn - 1].mean(axis=1)
    er = np.column_stack([c[:, 0], c[:, 1], np.zeros(len(c))])
    er /= np.linalg.norm(er, axis=1)[:, None]
    ez = np.tile([0.0, 0.0, 1.0], (len(c), 1))
    e2 = np.cross(er, ez)
    cs = np.stack([ez, e2, er], axis=2)  # columns = basis vectors
    # nodal normals are radial for this geometry (the layup csys normal)
    nrm = np.column_stack([xyz[:, 0], xyz[:, 1], np.zeros(xyz.shape[0])]) / R
    return dict(name=f"C3 T3FFComp cylinder {ncirc}x{nlen}x2", kind="t3", xyz=np.asfortranarray(xyz), conn=conn, dofnums=dof,
                nfree=nfree, csmat=cs, normals=np.asfortranarray(nrm), thickness=R / 100,
                lamina=(1500.0, 133860e6, 7706e6, 0.301, 4306e6, 4306e6, 2760e6), angles=(0.0, 90.0, 90.0, 0.0))


def c5_beam_lattice(ncell=69, seed=0):
    """C5: cubic lattice of ncell^3 cells: (ncell+1)^3 nodes, 3*ncell*(ncell+1)^2 members; rectangular
    section b = h = 0.02*cell (Bernoulli for x/y members, Timoshenko 5/6 for z members), x1x2 = z for x/y
    members, x for z members; E = 71240, nu = 0.31, rho = 5e-9 (test/test_beam_modal.jl:19-21); displaced
    state u1 ~ U(-1,1)*0.01*cell, Rfield1 = exp of rotation vectors ~ U(-1,1)*0.05."""
    n1 = ncell + 1
    cell = 1.0
    g = np.arange(n1)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    nid = (X * n1 + Y) * n1 + Z + 1  # 1-based
    xyz = np.column_stack([X.ravel(), Y.ravel(), Z.ravel()]).astype(np.float64) * cell
    cx = np.column_stack([nid[:-1, :, :].ravel(), nid[1:, :, :].ravel()])
    cy = np.column_stack([nid[:, :-1, :].ravel(), nid[:, 1:, :].ravel()])
    cz = np.column_stack([nid[:, :, :-1].ravel(), nid[:, :, 1:].ravel()])
    conn = np.concatenate([cx, cy, cz]).astype(np.int64)
    ne = conn.shape[0]
    b = 0.02 * cell
    A = np.full(ne, b * b)
    I2 = np.full(ne, b**4 / 12)
    I3 = I2.copy()
    I1 = I2 + I3
    J = np.full(ne, 0.141 * b**4)
    A2s = np.full(ne, np.inf)
    A3s = np.full(ne, np.inf)
    zmem = np.arange(ne) >= len(cx) + len(cy)
    A2s[zmem] = A3s[zmem] = 5.0 / 6.0 * b * b
    x1x2 = np.tile([0.0, 0.0, 1.0], (ne, 1))
    x1x2[zmem] = [1.0, 0.0, 0.0]
    rng = np.random.Generator(np.random.PCG64(seed))
    u1 = (rng.random((xyz.shape[0], 3)) * 2 - 1) * 0.01 * cell
    rv = (rng.random((xyz.shape[0], 3)) * 2 - 1) * 0.05
    th = np.linalg.norm(rv, axis=1)
    k = rv / th[:, None]
    K = np.zeros((len(th), 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -k[:, 2], k[:, 1], k[:, 2], -k[:, 0], -k[:, 1], k[:, 0]
    Rm = np.eye(3)[None] + np.sin(th)[:, None, None] * K + (1 - np.cos(th))[:, None, None] * (K @ K)
    Rf = np.asfortranarray(Rm.transpose(0, 2, 1).reshape(-1, 9))  # each row a column-major 3x3
    fixed = np.zeros((xyz.shape[0], 6), dtype=bool)
    fixed[xyz[:, 2] < 1e-12, :] = True
    dof, nfree = number_dofs(fixed)
    return dict(name=f"C5 beam lattice {ncell}^3", kind="l2", xyz=np.asfortranarray(xyz), conn=conn, dofnums=dof, nfree=nfree,
                sections=dict(A=A, I1=I1, I2=I2, I3=I3, J=J, A2s=A2s, A3s=A3s, x1x2=x1x2), u1=np.asfortranarray(u1), Rfield1=Rf,
                E=71240.0, nu=0.31, rho=5e-9)


def t3refine(xyz, conn):
    """`T3refine`: split every triangle into four through the edge midpoints (shared per edge)."""
    c = conn - 1
    edges = np.concatenate([c[:, [0, 1]], c[:, [1, 2]], c[:, [2, 0]]])
    es = np.sort(edges, axis=1)
    key = es[:, 0].astype(np.int64) * (xyz.shape[0] + 1) + es[:, 1]
    uk, inv = np.unique(key, return_inverse=True)
    a, b = uk // (xyz.shape[0] + 1), uk % (xyz.shape[0] + 1)
    mid = 0.5 * (xyz[a] + xyz[b])
    nn = xyz.shape[0]
    ne = c.shape[0]
    m01, m12, m20 = nn + inv[:ne], nn + inv[ne : 2 * ne], nn + inv[2 * ne :]
    new = np.concatenate([
        np.column_stack([c[:, 0], m01, m20]),
        np.column_stack([m01, c[:, 1], m12]),
        np.column_stack([m20, m12, c[:, 2]]),
        np.column_stack([m01, m12, m20]),
    ])
    return np.vstack([xyz, mid]), new + 1


def c1_double_cell_box(nref=4):
    """C1: double-cell box of examples/shells/dynamics/homogeneous/free_vibration/dcbs_vibration_examples.jl:21-59
    (6-node cross-section, 7 L2 segments, Q4extrudeL2 with 5 layers, Q4toT3, T3refine x nref; nref = 4 gives
    17 920 T3 / 8 991 nodes), E = 207 GPa, nu = 0.3, rho = 7850, t = 12.7 mm, clamped at z = 0.  The junction
    nodes of the middle wall have invalid nodal normals."""
    L0, B0 = 2.54, 1.27
    X = np.array([[0.0, 0, 0], [B0, 0, 0], [B0, B0, 0], [B0, 2 * B0, 0], [0, 2 * B0, 0], [0, B0, 0]])
    segs = np.array([[1, 2], [6, 3], [5, 4], [2, 3], [3, 4], [5, 6], [6, 1]])
    nl, nn1 = 5, 6
    xyz = np.vstack([X + np.array([0, 0, k * L0 / nl]) for k in range(nl + 1)])
    quads = []
    for k in range(1, nl + 1):
        for a, b in segs:
            quads.append([a + (k - 1) * nn1, b + (k - 1) * nn1, b + k * nn1, a + k * nn1])
    q = np.array(quads, dtype=np.int64)
    conn = np.concatenate([q[:, [0, 1, 2]], q[:, [0, 2, 3]]])  # Q4toT3
    for _ in range(nref):
        xyz, conn = t3refine(xyz, conn)
    fixed = np.zeros((xyz.shape[0], 6), dtype=bool)
    fixed[np.abs(xyz[:, 2]) < B0 / max(nref, 1) / 1000, :] = True
    dof, nfree = number_dofs(fixed)
    return dict(name=f"C1 double-cell box, T3refine x{nref}", kind="t3", xyz=np.asfortranarray(xyz), conn=conn.astype(np.int64), dofnums=dof,
                nfree=nfree, fixed=fixed, E=207e9, nu=0.3, rho=7850.0, thickness=12.7e-3)
