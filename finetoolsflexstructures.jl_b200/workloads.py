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
                hi_dofs=free_dofs(top) if rank < world - 1 else np.zeros(0, np.int64))


def c1_small_t3(n=64):
    """CPU-runnable small case (stand-in for C1's size class): curved T3 panel."""
    w = c4_t3ff_panel(n, n)
    w["name"] = f"small T3FF panel {n}x{n}x2"
    return w
