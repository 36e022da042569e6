"""Oracle: corotational beam (FEMMCorotBeam) -- restoring force, material and
geometric stiffness, mass; rotation-field update.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
Follows src/FEMMCorotBeamModule.jl, src/FESetL2BeamModule.jl:108-127,
src/RotUtilModule.jl:29-42 (reference v3.6.4).  Batched over elements.
"""
from __future__ import annotations

import numpy as np

from . import fe_external as fx


def _cross(a, b):
    return np.stack(
        [
            -a[:, 2] * b[:, 1] + a[:, 1] * b[:, 2],
            a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
            -a[:, 1] * b[:, 0] + a[:, 0] * b[:, 1],
        ],
        axis=1,
    )


def _frame(x, x1x2):
    """Frame with axis 1 along the chord, axis 3 = e1 x x1x2 (normalised), axis 2 = e3 x e1.
    src/FESetL2BeamModule.jl:108-127, src/FEMMCorotBeamModule.jl:155-180."""
    e1 = x[:, 1] - x[:, 0]
    L = np.sqrt(np.sum(e1**2, axis=1))
    e1 = e1 / L[:, None]
    e3 = _cross(e1, x1x2)
    e3 = e3 / np.sqrt(np.sum(e3**2, axis=1))[:, None]
    e2 = _cross(e3, e1)
    return L, np.stack([e1, e2, e3], axis=2)


def local_frame_and_def(x0, x1x2, xt, RI, RJ):
    """Lt, Ft (ne,3,3), dN (ne,6), L0.  src/FEMMCorotBeamModule.jl:208-243."""
    L0, F0 = _frame(x0, x1x2)
    FtI = np.einsum("eij,ejk->eik", RI, F0)
    FtJ = np.einsum("eij,ejk->eik", RJ, F0)
    Lt, Ft = _frame(xt, FtI[:, :, 1] + FtJ[:, :, 1])
    LI = np.einsum("eji,ejk->eik", Ft, FtI)
    LJ = np.einsum("eji,ejk->eik", Ft, FtJ)
    dN = np.zeros((x0.shape[0], 6))
    dN[:, 0] = Lt - L0
    dN[:, 5] = (LJ[:, 2, 1] / LJ[:, 1, 1] - LI[:, 2, 1] / LI[:, 1, 1] - LJ[:, 1, 2] / LJ[:, 2, 2] + LI[:, 1, 2] / LI[:, 2, 2]) / 2
    TH2I = -LI[:, 2, 0] / LI[:, 0, 0]
    TH2J = -LJ[:, 2, 0] / LJ[:, 0, 0]
    TH3I = LI[:, 1, 0] / LI[:, 0, 0]
    TH3J = LJ[:, 1, 0] / LJ[:, 0, 0]
    dN[:, 1] = TH3I - TH3J
    dN[:, 2] = TH3I + TH3J
    dN[:, 3] = -TH2I + TH2J
    dN[:, 4] = -TH2I - TH2J
    return Lt, Ft, dN, L0


def local_cartesian_to_natural(L):
    """aN (ne,6,12).  src/FEMMCorotBeamModule.jl:265-290."""
    aN = np.zeros((L.shape[0], 6, 12))
    aN[:, 0, 0] = -1
    aN[:, 0, 6] = +1
    aN[:, 1, 5] = +1
    aN[:, 1, 11] = -1
    aN[:, 2, 1] = 2 / L
    aN[:, 2, 5] = +1
    aN[:, 2, 7] = -2 / L
    aN[:, 2, 11] = +1
    aN[:, 3, 4] = -1
    aN[:, 3, 10] = +1
    aN[:, 4, 2] = 2 / L
    aN[:, 4, 4] = -1
    aN[:, 4, 8] = -2 / L
    aN[:, 4, 10] = -1
    aN[:, 5, 3] = -1
    aN[:, 5, 9] = +1
    return aN


def natural_stiffness(E, G, A, I2, I3, J, A2s, A3s, L):
    """DN diagonal (ne,6); Bernoulli iff A2s == Inf || A3s == Inf, else Timoshenko.
    src/FEMMCorotBeamModule.jl:729-779."""
    DN = np.zeros((L.shape[0], 6))
    bern = np.isinf(A2s) | np.isinf(A3s)
    DN[:, 0] = E * A / L
    DN[:, 1] = E * I3 / L
    DN[:, 3] = E * I2 / L
    DN[:, 5] = G * J / L
    with np.errstate(divide="ignore", invalid="ignore"):
        Phi3 = 12 * E * I3 / (G * A2s * L**2)
        Phi2 = 12 * E * I2 / (G * A3s * L**2)
    DN[:, 2] = np.where(bern, 3 * E * I3 / L, 3 * E * I3 / L / (1 + Phi3))
    DN[:, 4] = np.where(bern, 3 * E * I2 / L, 3 * E * I2 / L / (1 + Phi2))
    return DN


def _Te(Ft):
    ne = Ft.shape[0]
    Te = np.zeros((ne, 12, 12))
    for k in range(4):
        Te[:, 3 * k : 3 * k + 3, 3 * k : 3 * k + 3] = Ft
    return Te


def local_geometric_stiffness(PN, L):
    """Krenk's local geometric stiffness (ne,12,12).  src/FEMMCorotBeamModule.jl:321-382."""
    ne = L.shape[0]
    N = PN[:, 0]
    S2 = -2 * PN[:, 2] / L
    S3 = -2 * PN[:, 4] / L
    M1 = PN[:, 5]
    M2I = PN[:, 3] + PN[:, 4]
    M2J = PN[:, 3] - PN[:, 4]
    M3I = -(PN[:, 1] + PN[:, 2])
    M3J = -(PN[:, 1] - PN[:, 2])
    z = np.zeros(ne)
    SM = np.zeros((ne, 12, 12))

    def put(r, c, rows):
        SM[:, r : r + 3, c : c + 3] = np.stack([np.stack(rw, axis=1) for rw in rows], axis=1)

    put(0, 0, [[z, -S2 / L, -S3 / L], [-S2 / L, N / L, z], [-S3 / L, z, N / L]])
    put(0, 3, [[z, z, z], [-M2I / L, M1 / L, z], [-M3I / L, z, M1 / L]])
    put(0, 6, [[z, S2 / L, S3 / L], [S2 / L, -N / L, z], [S3 / L, z, -N / L]])
    put(0, 9, [[z, z, z], [M2J / L, -M1 / L, z], [M3J / L, z, -M1 / L]])
    put(3, 3, [[z, M3I / 2, -M2I / 2], [M3I / 2, z, z], [-M2I / 2, z, z]])
    put(3, 6, [[z, M2I / L, M3I / L], [z, -M1 / L, z], [z, z, -M1 / L]])
    put(3, 9, [[z, z, z], [z, z, M1 / 2], [z, -M1 / 2, z]])
    put(6, 6, [[z, -S2 / L, -S3 / L], [-S2 / L, N / L, z], [-S3 / L, z, N / L]])
    put(6, 9, [[z, z, z], [-M2J / L, M1 / L, z], [-M3J / L, z, M1 / L]])
    put(9, 9, [[z, -M3J / 2, M2J / 2], [-M3J / 2, z, z], [M2J / 2, z, z]])
    il = np.tril_indices(12, -1)
    SM[:, il[0], il[1]] = SM[:, il[1], il[0]]
    return SM


def _gather(xyz, conn, u1, Rfield1):
    c = np.asarray(conn) - 1
    x0 = xyz[c]
    x1 = x0 + u1[c]
    R = Rfield1[c].reshape(-1, 2, 3, 3).transpose(0, 1, 3, 2)  # rows are column-major 3x3
    return x0, x1, R[:, 0], R[:, 1]


def beam_restoringforce_elvecs(xyz, conn, u1, Rfield1, sec, E, nu):
    """elvec = Te * (-aN' * DN * dN), (ne,12).  `sec`: dict of per-element arrays
    A, I2, I3, J, A2s, A3s, x1x2 (ne,3).  src/FEMMCorotBeamModule.jl:1112-1159."""
    G = E / 2 / (1 + nu)
    x0, x1, RI, RJ = _gather(xyz, conn, u1, Rfield1)
    L1, Ft, dN, _ = local_frame_and_def(x0, sec["x1x2"], x1, RI, RJ)
    DN = natural_stiffness(E, G, sec["A"], sec["I2"], sec["I3"], sec["J"], sec["A2s"], sec["A3s"], L1)
    PN = DN * dN
    aN = local_cartesian_to_natural(L1)
    LF = np.einsum("eki,ek->ei", aN, PN)
    return np.einsum("eij,ej->ei", _Te(Ft), -LF)


def beam_stiffness_elmats(xyz, conn, u1, Rfield1, sec, E, nu):
    """Te (aN' DN aN) Te'.  src/FEMMCorotBeamModule.jl:972-1023."""
    G = E / 2 / (1 + nu)
    x0, x1, RI, RJ = _gather(xyz, conn, u1, Rfield1)
    L1, Ft, dN, _ = local_frame_and_def(x0, sec["x1x2"], x1, RI, RJ)
    DN = natural_stiffness(E, G, sec["A"], sec["I2"], sec["I3"], sec["J"], sec["A2s"], sec["A3s"], L1)
    aN = local_cartesian_to_natural(L1)
    SM = np.einsum("eki,ek,ekj->eij", aN, DN, aN)
    Te = _Te(Ft)
    return np.einsum("eij,ejk,elk->eil", Te, SM, Te)


def beam_geostiffness_elmats(xyz, conn, u1, Rfield1, sec, E, nu):
    """Te K_G(PN, L1) Te'.  src/FEMMCorotBeamModule.jl:1042-1094."""
    G = E / 2 / (1 + nu)
    x0, x1, RI, RJ = _gather(xyz, conn, u1, Rfield1)
    L1, Ft, dN, _ = local_frame_and_def(x0, sec["x1x2"], x1, RI, RJ)
    DN = natural_stiffness(E, G, sec["A"], sec["I2"], sec["I3"], sec["J"], sec["A2s"], sec["A3s"], L1)
    PN = DN * dN
    SM = local_geometric_stiffness(PN, L1)
    Te = _Te(Ft)
    return np.einsum("eij,ejk,elk->eil", Te, SM, Te)


def local_mass(A, I1, I2, I3, rho, L, mass_type):
    """4 mass forms (0 consistent no rot. inertia, 1 consistent with, 2 lumped no,
    3 lumped with).  src/FEMMCorotBeamModule.jl:384-554."""
    ne = L.shape[0]
    MM = np.zeros((ne, 12, 12))
    if mass_type in (0, 1):
        c1 = rho * A * L
        ent = {
            (0, 0): 1 / 3, (0, 6): 1 / 6, (1, 1): 13 / 35, (1, 5): 11 * L / 210, (1, 7): 9 / 70,
            (1, 11): -13 * L / 420, (2, 2): 13 / 35, (2, 4): -11 * L / 210, (2, 8): 9 / 70,
            (2, 10): 13 * L / 420, (3, 3): I1 / 3 / A, (3, 9): I1 / 6 / A, (4, 4): L**2 / 105,
            (4, 8): -13 * L / 420, (4, 10): -(L**2) / 140, (5, 5): L**2 / 105, (5, 7): 13 * L / 420,
            (5, 11): -(L**2) / 140, (6, 6): 1 / 3, (7, 7): 13 / 35, (7, 11): -11 * L / 210,
            (8, 8): 13 / 35, (8, 10): 11 * L / 210, (9, 9): I1 / 3 / A, (10, 10): L**2 / 105,
            (11, 11): L**2 / 105,
        }
        for (i, j), v in ent.items():
            MM[:, i, j] = c1 * v
        if mass_type == 1:
            c2 = rho / L
            ent2 = {
                (1, 1): 6 / 5 * I2, (1, 5): L / 10 * I2, (1, 7): -6 / 5 * I2, (1, 11): L / 10 * I2,
                (2, 2): 6 / 5 * I3, (2, 4): -L / 10 * I3, (2, 8): -6 / 5 * I3, (2, 10): -L / 10 * I3,
                (4, 4): 2 * L**2 / 15 * I3, (4, 8): L / 10 * I3, (4, 10): -(L**2) / 30 * I3,
                (5, 5): 2 * L**2 / 15 * I2, (5, 7): -L / 10 * I2, (5, 11): -(L**2) / 30 * I2,
                (7, 7): 6 / 5 * I2, (7, 11): -L / 10 * I2, (8, 8): 6 / 5 * I3, (8, 10): L / 10 * I3,
                (10, 10): 2 * L**2 / 15 * I3, (11, 11): 2 * L**2 / 15 * I2,
            }
            for (i, j), v in ent2.items():
                MM[:, i, j] += c2 * v
        il = np.tril_indices(12, -1)
        MM[:, il[0], il[1]] = MM[:, il[1], il[0]]
    else:
        CA = A * rho * L / 2.0
        d = [CA, CA, CA, 0 * CA, 0 * CA, 0 * CA]
        if mass_type == 3:
            d = [CA, CA, CA, rho * I1 * L / 2.0, rho * I2 * L / 2.0, rho * I3 * L / 2.0]
        for k in range(6):
            MM[:, k, k] = d[k]
            MM[:, k + 6, k + 6] = d[k]
    return MM


def beam_mass_elmats(xyz, conn, u1, Rfield1, sec, rho, mass_type=1):
    """Te M_local(L0) Te'.  src/FEMMCorotBeamModule.jl:810-868."""
    x0, x1, RI, RJ = _gather(xyz, conn, u1, Rfield1)
    L1, Ft, dN, L0 = local_frame_and_def(x0, sec["x1x2"], x1, RI, RJ)
    MM = local_mass(sec["A"], sec["I1"], sec["I2"], sec["I3"], rho, L0, mass_type)
    Te = _Te(Ft)
    return np.einsum("eij,ejk,elk->eil", Te, MM, Te)


def initial_Rfield(nnodes):
    """src/RotUtilModule.jl:16-22."""
    R = np.zeros((nnodes, 9))
    R[:, 0] = R[:, 4] = R[:, 8] = 1.0
    return R


def update_rotation_field(Rfield, dchi_values):
    """R <- exp(dtheta) R per node; rows are column-major 3x3.  src/RotUtilModule.jl:29-42."""
    R = Rfield.reshape(-1, 3, 3).transpose(0, 2, 1)
    Rd = fx.rotmat3(dchi_values[:, 3:6])
    Ru = np.einsum("nij,njk->nik", Rd, R)
    return Ru.transpose(0, 2, 1).reshape(-1, 9).copy()


def beam_gyroscopic_elmats(xyz, conn, u1, Rfield1, v1, sec, rho, mass_type=1):
    """Quadratic-inertial-term (gyroscopic) matrix Ge = Omega~ M - M Omega~.
    src/FEMMCorotBeamModule.jl:883-952."""
    c = np.asarray(conn) - 1
    x0, x1, RI, RJ = _gather(xyz, conn, u1, Rfield1)
    L1, Ft, dN, L0 = local_frame_and_def(x0, sec["x1x2"], x1, RI, RJ)
    MM = local_mass(sec["A"], sec["I1"], sec["I2"], sec["I3"], rho, L0, mass_type)
    Te = _Te(Ft)
    M = np.einsum("eij,ejk,elk->eil", Te, MM, Te)
    ev = v1[c]  # (ne, 2, 6)
    evf = np.concatenate([np.einsum("enk,ekj->enj", ev[:, :, 0:3], Ft), np.einsum("enk,ekj->enj", ev[:, :, 3:6], Ft)], axis=2)
    Om = (
        ((evf[:, 0, 3] + evf[:, 1, 3]) / 2)[:, None] * Ft[:, :, 0]
        + ((evf[:, 0, 2] - evf[:, 1, 2]) / L1)[:, None] * Ft[:, :, 1]
        + ((evf[:, 1, 1] - evf[:, 0, 1]) / L1)[:, None] * Ft[:, :, 2]
    )
    OS = fx.skewmat(Om)
    Ot = _Te(OS)
    return np.einsum("eij,ejk->eik", Ot, M) - np.einsum("eij,ejk->eik", M, Ot)


def beam_distribloads_elvecs(xyz, conn, u1, Rfield1, sec, force):
    """Uniform global force per unit length (3,) or (ne,3) -> element vectors (ne,12).
    src/FEMMCorotBeamModule.jl:1186-1247."""
    x0, x1, RI, RJ = _gather(xyz, conn, u1, Rfield1)
    L1, Ft, dN, L0 = local_frame_and_def(x0, sec["x1x2"], x1, RI, RJ)
    f = np.broadcast_to(np.asarray(force, dtype=np.float64), (x0.shape[0], 3))
    Lf = np.einsum("eji,ej->ei", Ft, f)
    ev = np.zeros((x0.shape[0], 12))
    for k in range(3):
        ev[:, k] = Lf[:, k] * L0 / 2
        ev[:, 6 + k] = Lf[:, k] * L0 / 2
    ev[:, 4] = -Lf[:, 2] * L0**2 / 12
    ev[:, 5] = +Lf[:, 1] * L0**2 / 12
    ev[:, 10] = +Lf[:, 2] * L0**2 / 12
    ev[:, 11] = -Lf[:, 1] * L0**2 / 12
    return np.einsum("eij,ej->ei", _Te(Ft), ev)
