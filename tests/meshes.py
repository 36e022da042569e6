"""Small deterministic test meshes (shared by the CPU and GPU tests).  Test infrastructure."""
import numpy as np

from oracle import fe_external as fx


def wavy(xyz, crease=4.0, xc=0.5):
    xyz = xyz.copy()
    xyz[:, 2] = 0.3 * np.sin(2 * xyz[:, 0]) + 0.1 * xyz[:, 1] ** 2 + np.where(xyz[:, 0] > xc, (xyz[:, 0] - xc) * crease, 0)
    return xyz


def shell_mesh(kind, n=6, seed=1, crease=4.0, perturb=0.02):
    """kind 't3' | 'q4'; curved, creased (some invalid nodal normals), perturbed."""
    rng = np.random.default_rng(seed)
    gen = fx.t3block if kind == "t3" else fx.q4block
    xy, conn = gen(1.0, 1.0, n, n)
    xyz = wavy(fx.xyz3(xy), crease, xc=round(0.5 * n) / n)
    xyz[:, :2] += rng.uniform(-1, 1, (xyz.shape[0], 2)) * perturb / n * 6
    return xyz, conn


def clamp_edge_dofs(xyz, n_extra_fixed=3, seed=2):
    """DofField with the x=0 edge clamped plus a few scattered single-dof supports."""
    rng = np.random.default_rng(seed)
    d = fx.DofField(xyz.shape[0])
    edge = np.nonzero(np.abs(xyz[:, 0]) < 1e-9 + 0.02)[0]
    edge = np.nonzero(xyz[:, 0] <= xyz[:, 0].min() + 1e-12)[0] if len(edge) == 0 else edge
    for c in range(1, 7):
        d.setebc(edge, c)
    others = rng.choice(xyz.shape[0], size=n_extra_fixed, replace=False)
    for k, nd in enumerate(others):
        d.setebc([nd], 1 + (k % 6))
    d.numberdofs()
    return d


def beam_lattice(ne=40, seed=3):
    rng = np.random.default_rng(seed)
    xyz = np.cumsum(rng.uniform(0.2, 1.0, (ne + 1, 3)), axis=0)
    conn = np.column_stack([np.arange(1, ne + 1), np.arange(2, ne + 2)]).astype(np.int64)
    u1 = rng.uniform(-1, 1, (ne + 1, 3)) * 0.02
    R1 = fx.rotmat3(rng.uniform(-1, 1, (ne + 1, 3)) * 0.1).transpose(0, 2, 1).reshape(-1, 9)
    one = np.ones(ne)
    even = np.arange(ne) % 2 == 0
    sec = dict(
        A=0.01 * one * rng.uniform(1, 2, ne),
        I1=2e-5 * one,
        I2=1e-5 * one * rng.uniform(1, 2, ne),
        I3=1.5e-5 * one,
        J=1.7e-5 * one,
        A2s=np.where(even, np.inf, 0.008),
        A3s=np.where(even, np.inf, 0.007),
        x1x2=rng.uniform(-1, 1, (ne, 3)),
    )
    return xyz, conn, u1, R1, sec
