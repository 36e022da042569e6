"""Oracle: T3FF / Q4RS shells and their composite variants (element matrices,
lumped mass, nodal normals).  Batched NumPy float64 restatement that follows the
reference step by step (dense transformation matrices, QtEQ products), NOT the
restructured arithmetic of the CUDA kernels.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

All arrays are batched over elements on axis 0.  `conn` is 1-based.
"""
from __future__ import annotations

import numpy as np

from . import fe_external as fx

# ---------------------------------------------------------------------------
# shared helpers
# ---------------------------------------------------------------------------


def e_g(J):
    """Element triad from the tangents J (ne,3,2).
    src/FEMMShellT3FFModule.jl:283-305, src/FEMMShellQ4RSModule.jl:248-263."""
    E = np.zeros(J.shape[:-2] + (3, 3))
    e1 = J[..., :, 0] / np.sqrt(np.sum(J[..., :, 0] ** 2, axis=-1))[..., None]
    j2 = J[..., :, 1]
    e3 = np.stack(
        [
            -e1[..., 2] * j2[..., 1] + e1[..., 1] * j2[..., 2],
            e1[..., 2] * j2[..., 0] - e1[..., 0] * j2[..., 2],
            -e1[..., 1] * j2[..., 0] + e1[..., 0] * j2[..., 1],
        ],
        axis=-1,
    )
    e3 = e3 / np.sqrt(np.sum(e3**2, axis=-1))[..., None]
    e2 = np.stack(
        [
            -e3[..., 2] * e1[..., 1] + e3[..., 1] * e1[..., 2],
            e3[..., 2] * e1[..., 0] - e3[..., 0] * e1[..., 2],
            -e3[..., 1] * e1[..., 0] + e3[..., 0] * e1[..., 1],
        ],
        axis=-1,
    )
    E[..., :, 0] = e1
    E[..., :, 1] = e2
    E[..., :, 2] = e3
    return E


def shell_material_stiffness(D6):
    """Plane-stress and transverse-shear reductions of the 3-D moduli.
    src/FEMMShellT3FFModule.jl:323-342 (same in Q4RS :265-284, Ply :75-95)."""
    Dps = np.zeros((3, 3))
    Dps[0:2, 0:2] = D6[0:2, 0:2] - np.outer(D6[0:2, 2], D6[2, 0:2]) / D6[2, 2]
    ix = [0, 1, 3]
    for i in range(3):
        Dps[2, i] = Dps[i, 2] = D6[3, ix[i]]
    Dt = np.zeros((2, 2))
    Dt[0, 0] = D6[4, 4]
    Dt[1, 1] = D6[5, 5]
    return Dps, Dt


def nodal_triads_e(E_G, normals, normal_valid, conn):
    """A_Es (ne,nn,3,3), nvalid (ne,nn).  Rotation VECTOR is e3 x n_k, i.e. the
    angle is sin(theta) (Appendix B.1).  src/FEMMShellT3FFModule.jl:355-388."""
    c = np.asarray(conn) - 1
    nk = normals[c]  # (ne,nn,3)
    nvalid = normal_valid[c].astype(bool)
    nk_e = np.einsum("eji,ekj->eki", E_G, nk)  # E_G' * nk
    e3 = np.zeros_like(nk_e)
    e3[..., 2] = 1.0
    nk_e = np.where(nvalid[..., None], nk_e, e3)
    r = np.stack([-nk_e[..., 1], nk_e[..., 0], np.zeros_like(nk_e[..., 0])], axis=-1)
    nr = np.sqrt(np.sum(r * r, axis=-1))
    A = fx.rotmat3(r)
    eye = np.broadcast_to(np.eye(3), A.shape)
    A = np.where((nr > 1.0e-12)[..., None, None], A, eye)
    return A, nvalid


def transfmat_g_to_a(A_Es, E_G):
    """Block-diagonal global->nodal T.  src/FEMMShellT3FFModule.jl:398-419."""
    ne, nn = A_Es.shape[:2]
    T = np.zeros((ne, 6 * nn, 6 * nn))
    for i in range(nn):
        blk = np.einsum("eji,ekj->eik", A_Es[:, i], E_G)  # A' * E_G'
        o = 6 * i
        T[:, o : o + 3, o : o + 3] = blk
        T[:, o + 3 : o + 6, o + 3 : o + 6] = blk
    return T


def transfmat_a_to_e(A_Es, gradN_e):
    """Nodal->element T with the drilling-consistency coupling rows.
    src/FEMMShellT3FFModule.jl:421-463."""
    ne, nn = A_Es.shape[:2]
    T = np.zeros((ne, 6 * nn, 6 * nn))
    for i in range(nn):
        ro = 6 * i
        A = A_Es[:, i]
        A33 = A[:, 2, 2]
        T[:, ro : ro + 3, ro : ro + 3] = A
        for cl in range(2):
            for rw in range(2):
                T[:, ro + 3 + rw, ro + 3 + cl] = A[:, rw, cl] - (1 / A33) * A[:, rw, 2] * A[:, cl, 2]
        m1 = (1 / A33) * A[:, 0, 2]
        m2 = (1 / A33) * A[:, 1, 2]
        for j in range(nn):
            co = 6 * j
            for k in range(3):
                a3 = 1 / 2 * (A[:, 1, k] * gradN_e[:, j, 0] - A[:, 0, k] * gradN_e[:, j, 1])
                T[:, ro + 3, co + k] += m1 * a3
                T[:, ro + 4, co + k] += m2 * a3
    return T


def add_btdb_ut_only(elmat, B, c, D):
    """`add_btdb_ut_only!`: upper triangle of elmat += c * B' D B (App. A.3)."""
    c = np.asarray(c)
    DB = np.einsum("ij,ejk->eik", D, B) if D.ndim == 2 else np.einsum("eij,ejk->eik", D, B)
    K = np.einsum("eki,ekj->eij", B, DB) * c[:, None, None]
    iu = np.triu_indices(elmat.shape[-1])
    elmat[:, iu[0], iu[1]] += K[:, iu[0], iu[1]]
    return elmat


def add_b1tdb2(elmat, B1, B2, c, D):
    """`add_b1tdb2!`: full elmat += c * B1' D B2 (App. A.3)."""
    c = np.asarray(c)
    DB = np.einsum("eij,ejk->eik", D, B2)
    elmat += np.einsum("eki,ekj->eij", B1, DB) * c[:, None, None]
    return elmat


def complete_lt(elmat):
    """`complete_lt!`: mirror the upper triangle into the lower."""
    n = elmat.shape[-1]
    il = np.tril_indices(n, -1)
    elmat[:, il[0], il[1]] = elmat[:, il[1], il[0]]
    return elmat


def qteq(E, Q):
    """TransformerQtEQ: Q' (E Q).  src/TransformerModule.jl:34-42."""
    return np.einsum("eki,ekj->eij", Q, np.einsum("eij,ejk->eik", E, Q))


def stab_lyly(alpha):
    return lambda t, h: t**2 / (t**2 + alpha * h**2)


T3_DEFAULT_ALPHA = 5 / 12 / 1.5  # src/FEMMShellT3FFModule.jl:36
Q4_DEFAULT_ALPHA = 0.1  # src/FEMMShellQ4RSModule.jl:32

# ---------------------------------------------------------------------------
# layup -> element angle and stress/strain rotation matrices
# ---------------------------------------------------------------------------


def layup2element_angle(E_G, lcsmat):
    """src/TransformerModule.jl:92-105."""
    M = np.einsum("eji,ejk->eik", E_G, lcsmat)
    M2 = M[:, 0:2, 0:2].copy()
    M2[:, :, 0] /= np.sqrt(np.sum(M2[:, :, 0] ** 2, axis=1))[:, None]
    M2[:, :, 1] /= np.sqrt(np.sum(M2[:, :, 1] ** 2, axis=1))[:, None]
    m = (M2[:, 0, 0] + M2[:, 1, 1]) / 2
    n = (M2[:, 0, 1] - M2[:, 1, 0]) / 2
    sn = np.where(n >= 0.0, 1.0, -1.0)
    n = sn * np.sqrt(1 - m**2)
    return m, n


def plane_stress_Tinv(m, n):
    """src/CompositeLayupModule.jl:410-421 (batched or scalar m, n)."""
    m = np.asarray(m, dtype=np.float64)
    n = np.asarray(n, dtype=np.float64)
    T = np.zeros(m.shape + (3, 3))
    T[..., 0, 0] = m**2
    T[..., 0, 1] = n**2
    T[..., 0, 2] = 2 * (m * n)
    T[..., 1, 0] = n**2
    T[..., 1, 1] = m**2
    T[..., 1, 2] = -2 * (m * n)
    T[..., 2, 0] = -(m * n)
    T[..., 2, 1] = m * n
    T[..., 2, 2] = m**2 - n**2
    return T


def plane_stress_Tbar(m, n):
    """Tbar = Tinv^T.  src/CompositeLayupModule.jl:364-378."""
    return np.swapaxes(plane_stress_Tinv(m, n), -1, -2)


def plane_stress_T(angle):
    """src/CompositeLayupModule.jl:438-443."""
    return plane_stress_Tinv(np.cos(-angle), np.sin(-angle))


def transverse_shear_T(m, n):
    """src/CompositeLayupModule.jl:463-471."""
    m = np.asarray(m, dtype=np.float64)
    n = np.asarray(n, dtype=np.float64)
    T = np.zeros(m.shape + (2, 2))
    T[..., 0, 0] = m
    T[..., 0, 1] = -n
    T[..., 1, 0] = n
    T[..., 1, 1] = m
    return T


# ---------------------------------------------------------------------------
# T3FF
# ---------------------------------------------------------------------------


def t3_geometry(X):
    """X (ne,3,3) node coordinates (rows = nodes).  Returns J0, E_G, ecoords_e,
    gradN_e, Ae.  src/FEMMShellT3FFModule.jl:269-314,465-484."""
    J0 = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=-1)
    E_G = e_g(J0)
    ec = np.zeros((X.shape[0], 3, 2))
    ec[:, 1, 0] = np.einsum("ei,ei->e", J0[:, :, 0], E_G[:, :, 0])
    ec[:, 1, 1] = np.einsum("ei,ei->e", J0[:, :, 0], E_G[:, :, 1])
    ec[:, 2, 0] = np.einsum("ei,ei->e", J0[:, :, 1], E_G[:, :, 0])
    ec[:, 2, 1] = np.einsum("ei,ei->e", J0[:, :, 1], E_G[:, :, 1])
    a = ec[:, 1, 0] - ec[:, 0, 0]
    b = ec[:, 1, 1] - ec[:, 0, 1]
    c = ec[:, 2, 0] - ec[:, 0, 0]
    d = ec[:, 2, 1] - ec[:, 0, 1]
    J = a * d - b * c
    g = np.zeros((X.shape[0], 3, 2))
    g[:, 0, 0] = (b - d) / J
    g[:, 1, 0] = d / J
    g[:, 2, 0] = -b / J
    g[:, 0, 1] = (c - a) / J
    g[:, 1, 1] = -c / J
    g[:, 2, 1] = a / J
    return J0, E_G, ec, g, J / 2


def _bm(gradN):
    """src/FEMMShellT3FFModule.jl:539-548, src/FEMMShellQ4RSModule.jl:548-558."""
    ne, nn = gradN.shape[:2]
    B = np.zeros((ne, 3, 6 * nn))
    for i in range(nn):
        B[:, 0, 6 * i + 0] = gradN[:, i, 0]
        B[:, 1, 6 * i + 1] = gradN[:, i, 1]
        B[:, 2, 6 * i + 0] = gradN[:, i, 1]
        B[:, 2, 6 * i + 1] = gradN[:, i, 0]
    return B


def _bb(gradN):
    """src/FEMMShellT3FFModule.jl:550-561, src/FEMMShellQ4RSModule.jl:568-578."""
    ne, nn = gradN.shape[:2]
    B = np.zeros((ne, 3, 6 * nn))
    for i in range(nn):
        B[:, 0, 6 * i + 4] = gradN[:, i, 0]
        B[:, 1, 6 * i + 3] = -gradN[:, i, 1]
        B[:, 2, 6 * i + 3] = -gradN[:, i, 0]
        B[:, 2, 6 * i + 4] = gradN[:, i, 1]
    return B


def _t3_add_bs_o(Bs, ec, Ae, ordering):
    """DSG shear B for one node ordering.  src/FEMMShellT3FFModule.jl:486-526."""
    s, p, q = ordering
    a = ec[:, p, 0] - ec[:, s, 0]
    b = ec[:, p, 1] - ec[:, s, 1]
    c = ec[:, q, 0] - ec[:, s, 0]
    d = ec[:, q, 1] - ec[:, s, 1]
    m = 1 / 2 / Ae
    co = s * 6
    Bs[:, 0, co + 2] += m * (b - d)
    Bs[:, 0, co + 4] += m * Ae
    Bs[:, 1, co + 2] += m * (c - a)
    Bs[:, 1, co + 3] += m * (-Ae)
    co = p * 6
    Bs[:, 0, co + 2] += m * d
    Bs[:, 0, co + 3] += m * (-b * d / 2)
    Bs[:, 0, co + 4] += m * (a * d / 2)
    Bs[:, 1, co + 2] += m * (-c)
    Bs[:, 1, co + 3] += m * (b * c / 2)
    Bs[:, 1, co + 4] += m * (-a * c / 2)
    co = q * 6
    Bs[:, 0, co + 2] += m * (-b)
    Bs[:, 0, co + 3] += m * (b * d / 2)
    Bs[:, 0, co + 4] += m * (-b * c / 2)
    Bs[:, 1, co + 2] += m * a
    Bs[:, 1, co + 3] += m * (-a * d / 2)
    Bs[:, 1, co + 4] += m * (a * c / 2)
    return Bs


_T3_ORDERINGS = [(0, 1, 2), (1, 2, 0), (2, 0, 1)]


def _t3_bs(ec, Ae):
    """src/FEMMShellT3FFModule.jl:528-537."""
    Bs = np.zeros((ec.shape[0], 2, 18))
    for o in _T3_ORDERINGS:
        _t3_add_bs_o(Bs, ec, Ae, o)
    Bs *= 1 / 3
    return Bs


def _t3_finish(elmat, E_G, gradN_e, normals, normal_valid, conn, drilling_stiffness_scale):
    """Transforms + drilling.  src/FEMMShellT3FFModule.jl:708-730."""
    complete_lt(elmat)
    A_Es, nvalid = nodal_triads_e(E_G, normals, normal_valid, conn)
    T = transfmat_a_to_e(A_Es, gradN_e)
    elmat = qteq(elmat, T)
    kavg = (
        np.mean(
            np.stack([elmat[:, k, k] for k in (3, 9, 15, 4, 10, 16)], axis=0),
            axis=0,
        )
        * drilling_stiffness_scale
    )
    for k in range(3):
        elmat[:, 6 * k + 5, 6 * k + 5] += np.where(nvalid[:, k], kavg, 0.0)
    T = transfmat_g_to_a(A_Es, E_G)
    return qteq(elmat, T)


def t3ff_stiffness_elmats(
    xyz,
    conn,
    normals,
    normal_valid,
    Dps,
    Dt,
    thickness,
    stab_fun=None,
    drilling_stiffness_scale=1.0,
    transv_shear_formulation=0,
):
    """Element stiffness matrices (ne,18,18) of FEMMShellT3FF.
    `Dt` is the raw transverse-shear matrix; the 5/6 factor is applied here.
    `thickness` scalar or per-element array.  src/FEMMShellT3FFModule.jl:635-736."""
    stab_fun = stab_fun or stab_lyly(T3_DEFAULT_ALPHA)
    c = np.asarray(conn) - 1
    ne = c.shape[0]
    X = xyz[c]
    J0, E_G, ec, g, Ae = t3_geometry(X)
    t = np.broadcast_to(np.asarray(thickness, dtype=np.float64), (ne,))
    Dt = Dt * (5 / 6)
    elmat = np.zeros((ne, 18, 18))
    add_btdb_ut_only(elmat, _bm(g), t * Ae, Dps)
    add_btdb_ut_only(elmat, _bb(g), (t**3) / 12 * Ae, Dps)
    h = np.sqrt(2 * Ae)
    if transv_shear_formulation == 1:
        for o in _T3_ORDERINGS:
            Bs = np.zeros((ne, 2, 18))
            _t3_add_bs_o(Bs, ec, Ae, o)
            add_btdb_ut_only(elmat, Bs, t * stab_fun(t, h) * Ae / 3, Dt)
    else:
        add_btdb_ut_only(elmat, _t3_bs(ec, Ae), t * stab_fun(t, h) * Ae, Dt)
    return _t3_finish(elmat, E_G, g, normals, normal_valid, conn, drilling_stiffness_scale)


def lumped_elmats(tmass, rmass, nn):
    """Full element matrices with explicit zeros, diagonal = (t,t,t,r,r,r) per node.
    src/FEMMShellT3FFModule.jl:793-805."""
    ne = tmass.shape[0]
    M = np.zeros((ne, 6 * nn, 6 * nn))
    for k in range(nn):
        for d in range(3):
            M[:, 6 * k + d, 6 * k + d] += tmass
        for d in range(3, 6):
            M[:, 6 * k + d, 6 * k + d] += rmass
    return M


def t3ff_mass_elmats(xyz, conn, rho, thickness):
    """src/FEMMShellT3FFModule.jl:757-811."""
    c = np.asarray(conn) - 1
    X = xyz[c]
    _, _, _, _, Ae = t3_geometry(X)
    t = np.broadcast_to(np.asarray(thickness, dtype=np.float64), (c.shape[0],))
    tmass = rho * (t * Ae) / 3
    rmass = rho * (t**3 / 12 * Ae) / 3
    return lumped_elmats(tmass, rmass, 3)


def t3ff_associategeometry(xyz, conn, threshold_angle=30.0, normal_dir=None, normals0=None):
    """Nodal normals + validity.  `normal_dir` (3,) overrides the element normal
    with a fixed direction (a cartesian layup csys in the composite FEMM).
    src/FEMMShellT3FFModule.jl:570-616; Comp: src/FEMMShellT3FFCompModule.jl:489-543."""
    c = np.asarray(conn) - 1
    nn = xyz.shape[0]
    X = xyz[c]
    J0 = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=-1)
    E3 = e_g(J0)[:, :, 2]
    nd_ = None if normal_dir is None else np.asarray(normal_dir, float)
    normals = np.zeros((nn, 3)) if normals0 is None else normals0.copy()
    for k in range(3):
        # per (element, node) directions: a general csys evaluated at the node (`_compute_nodal_normal!`, Comp :203-207,509)
        contrib = E3 if nd_ is None else (nd_[:, k] if nd_.ndim == 3 else np.broadcast_to(nd_, E3.shape))
        np.add.at(normals, c[:, k], contrib)
    nrm = np.sqrt(np.sum(normals**2, axis=1))
    normals = np.where((nrm > 0)[:, None], normals / np.where(nrm > 0, nrm, 1.0)[:, None], normals)
    ntol = 1 - np.sqrt(1 - np.sin(threshold_angle / 180 * np.pi) ** 2)
    valid = np.ones(nn, dtype=bool)
    for k in range(3):
        nd = np.einsum("ei,ei->e", normals[c[:, k]], E3)
        bad = nd < 1 - ntol
        valid[c[bad, k]] = False
    return normals, valid


# ---------------------------------------------------------------------------
# T3FF composite
# ---------------------------------------------------------------------------


def t3ffcomp_stiffness_elmats(
    xyz,
    conn,
    normals,
    normal_valid,
    A,
    B,
    D,
    H,
    layup_thickness,
    lcsmat,
    stab_fun=None,
    drilling_stiffness_scale=1.0,
    transv_shear_formulation=0,
):
    """Element stiffness of FEMMShellT3FFComp for one layup group.
    `lcsmat` (ne,3,3) or (3,3): layup csys matrix at the element centroid.
    src/FEMMShellT3FFCompModule.jl:561-689."""
    stab_fun = stab_fun or stab_lyly(T3_DEFAULT_ALPHA)
    c = np.asarray(conn) - 1
    ne = c.shape[0]
    X = xyz[c]
    J0, E_G, ec, g, Ae = t3_geometry(X)
    lcs = np.broadcast_to(np.asarray(lcsmat, dtype=np.float64), (ne, 3, 3))
    m, n = layup2element_angle(E_G, lcs)
    Tps = plane_stress_Tbar(m, n)
    Tts = transverse_shear_T(m, n)
    bc = lambda M: np.broadcast_to(M, (ne,) + M.shape).copy()
    sA, sB, sD, sH = qteq(bc(A), Tps), qteq(bc(B), Tps), qteq(bc(D), Tps), qteq(bc(H), Tts)
    t = layup_thickness
    elmat = np.zeros((ne, 18, 18))
    Bm, Bb = _bm(g), _bb(g)
    add_btdb_ut_only(elmat, Bm, Ae, sA)
    add_btdb_ut_only(elmat, Bb, Ae, sD)
    add_b1tdb2(elmat, Bm, Bb, Ae, sB)
    add_b1tdb2(elmat, Bb, Bm, Ae, sB)
    h = np.sqrt(2 * Ae)
    if transv_shear_formulation == 1:
        for o in _T3_ORDERINGS:
            Bs = np.zeros((ne, 2, 18))
            _t3_add_bs_o(Bs, ec, Ae, o)
            add_btdb_ut_only(elmat, Bs, stab_fun(t, h) * Ae / 3, sH)
    else:
        add_btdb_ut_only(elmat, _t3_bs(ec, Ae), stab_fun(t, h) * Ae, sH)
    return _t3_finish(elmat, E_G, g, normals, normal_valid, conn, drilling_stiffness_scale)


def t3ffcomp_mass_elmats(xyz, conn, mass_density, moi_density):
    """src/FEMMShellT3FFCompModule.jl:710-770."""
    c = np.asarray(conn) - 1
    _, _, _, _, Ae = t3_geometry(xyz[c])
    return lumped_elmats(mass_density * (Ae / 3), moi_density * (Ae / 3), 3)


# ---------------------------------------------------------------------------
# Q4RS
# ---------------------------------------------------------------------------


def _q4_local_derivatives(J, E_G, dNp):
    """gradN_e = E2' J (J'J)^-1 gradNparams.  src/FEMMShellQ4RSModule.jl:305-323."""
    G = np.einsum("eki,ekj->eij", J, J)
    detG = G[:, 0, 0] * G[:, 1, 1] - G[:, 0, 1] * G[:, 1, 0]
    if np.any(np.isclose(detG, 0.0)):
        raise ZeroDivisionError("Singular metric matrix in _gradN_e!")
    invG = np.linalg.inv(G)
    tmp = np.einsum("eij,aj->eai", invG, dNp)  # (ne,4,2)
    g3 = np.einsum("eij,eaj->eai", J, tmp)  # (ne,4,3)
    out = np.zeros((J.shape[0], 4, 2))
    out[:, :, 0] = np.einsum("ei,eai->ea", E_G[:, :, 0], g3)
    out[:, :, 1] = np.einsum("ei,eai->ea", E_G[:, :, 1], g3)
    return out


def _q4_ecoords_e(X, E_G):
    """Centroid-relative projected coordinates.  src/FEMMShellQ4RSModule.jl:588-608."""
    cen = X.sum(axis=1) / 4
    return np.einsum("eam,emk->eak", X - cen[:, None, :], E_G[:, :, 0:2])


def _q4_mitc_bs(ec, r, s):
    """MITC4 (Bathe-Dvorkin 1985) shear strain-displacement matrix, built from the
    DEFINITION in the reference's derivation comment (src/FEMMShellQ4RSModule.jl:
    628-752): the tying strains g_rz, g_sz are linear functionals of (W, Tx, Ty);
    each B column is that functional evaluated on a unit dof.  (The reference's
    24 closed-form entries, :778-802, are the SymPy expansion of the same thing.)"""
    ne = ec.shape[0]
    X, Y = ec[:, :, 0], ec[:, :, 1]
    X1, X2, X3, X4 = X[:, 0], X[:, 1], X[:, 2], X[:, 3]
    Y1, Y2, Y3, Y4 = Y[:, 0], Y[:, 1], Y[:, 2], Y[:, 3]
    J11 = X1 * (s - 1) / 4 - X2 * (s - 1) / 4 + X3 * (s + 1) / 4 - X4 * (s + 1) / 4
    J21 = Y1 * (s - 1) / 4 - Y2 * (s - 1) / 4 + Y3 * (s + 1) / 4 - Y4 * (s + 1) / 4
    J12 = X1 * (r - 1) / 4 - X2 * (r + 1) / 4 + X3 * (r + 1) / 4 - X4 * (r - 1) / 4
    J22 = Y1 * (r - 1) / 4 - Y2 * (r + 1) / 4 + Y3 * (r + 1) / 4 - Y4 * (r - 1) / 4
    Aa = np.sqrt(J11**2 + J21**2)
    Bb = np.sqrt(J12**2 + J22**2)
    ca, sa, cb, sb = J11 / Aa, J21 / Aa, J12 / Bb, J22 / Bb
    detJ = J11 * J22 - J12 * J21
    Ax, Ay = X1 - X2 - X3 + X4, Y1 - Y2 - Y3 + Y4
    Bx, By = X1 - X2 + X3 - X4, Y1 - Y2 + Y3 - Y4
    Cx, Cy = X1 + X2 - X3 - X4, Y1 + Y2 - Y3 - Y4
    SC = np.sqrt((Cx + r * Bx) ** 2 + (Cy + r * By) ** 2) / 8 / detJ
    SA = np.sqrt((Ax + s * Bx) ** 2 + (Ay + s * By) ** 2) / 8 / detJ

    def edge(W, Tx, Ty, a, b):
        return (W[a] - W[b]) / 2 + (X[:, a] - X[:, b]) / 4 * (Ty[a] + Ty[b]) - (Y[:, a] - Y[:, b]) / 4 * (
            Tx[a] + Tx[b]
        )

    Bs = np.zeros((ne, 2, 24))
    for node in range(4):
        for comp in range(3):  # W, Tx, Ty -> dof columns 3, 4, 5 (1-based) of the node
            W, Tx, Ty = np.zeros(4), np.zeros(4), np.zeros(4)
            (W, Tx, Ty)[comp][node] = 1.0
            grz = SC * ((1 + s) * edge(W, Tx, Ty, 0, 1) + (1 - s) * edge(W, Tx, Ty, 3, 2))
            gsz = SA * ((1 + r) * edge(W, Tx, Ty, 0, 3) + (1 - r) * edge(W, Tx, Ty, 1, 2))
            gxz = -(grz * sb + gsz * (-sa))
            gyz = -(grz * (-cb) + gsz * ca)
            col = 6 * node + 2 + comp
            Bs[:, 0, col] = gxz
            Bs[:, 1, col] = gyz
    return Bs


def q4_diameter(X):
    """max distance node 1 -> others (NOT the true diameter; App. B.4).
    src/FEMMShellQ4RSModule.jl:861-870."""
    d = np.sum((X[:, 1:, :] - X[:, 0:1, :]) ** 2, axis=2)
    return np.sqrt(np.max(d, axis=1))


def _q4_add_drilling(elmat, normals, normal_valid, conn, scale):
    """src/FEMMShellQ4RSModule.jl:807-859."""
    if scale == 0.0:
        return elmat
    c = np.asarray(conn) - 1
    ne = c.shape[0]
    nv = normals[c]  # (ne,4,3)
    nn_ = np.sqrt(np.sum(nv**2, axis=2))
    ok = normal_valid[c].astype(bool) & (nn_ != 0.0)
    tang = np.zeros((ne, 4))
    for k in range(4):
        n = nv[:, k] / np.where(nn_[:, k] != 0, nn_[:, k], 1.0)[:, None]
        P = np.eye(3)[None] - n[:, :, None] * n[:, None, :]
        Krr = elmat[:, 6 * k + 3 : 6 * k + 6, 6 * k + 3 : 6 * k + 6]
        Kt = np.einsum("eij,ejk,ekl->eil", P, Krr, P)
        tang[:, k] = np.maximum(0.0, np.einsum("eii->e", Kt) / 2.0)
    cnt = ok.sum(axis=1)
    kavg = np.where(cnt > 0, (tang * ok).sum(axis=1) / np.maximum(cnt, 1), 0.0) * scale
    for k in range(4):
        Ke = kavg[:, None, None] * (nv[:, k, :, None] * nv[:, k, None, :])
        Ke = np.where((ok[:, k] & (kavg != 0.0))[:, None, None], Ke, 0.0)
        elmat[:, 6 * k + 3 : 6 * k + 6, 6 * k + 3 : 6 * k + 6] += Ke
    return elmat


def _q4_gp_setup(X, normals, normal_valid, conn, xi, eta):
    N, dNp = fx.q4_shape(xi, eta)
    J = np.einsum("eai,ak->eik", X, dNp)  # locjac!: J = x' * gradNparams
    Jac = np.sqrt(np.sum(np.cross(J[:, :, 0], J[:, :, 1]) ** 2, axis=1))  # Jacobiansurface
    E_G = e_g(J)
    ec = _q4_ecoords_e(X, E_G)
    g = _q4_local_derivatives(J, E_G, dNp)
    A_Es, nvalid = nodal_triads_e(E_G, normals, normal_valid, conn)
    Tga = transfmat_g_to_a(A_Es, E_G)
    Tae = transfmat_a_to_e(A_Es, g)
    T = np.einsum("eij,ejk->eik", Tae, Tga)
    return N, J, Jac, E_G, ec, g, T


def q4rs_stiffness_elmats(
    xyz,
    conn,
    normals,
    normal_valid,
    Dps,
    Dt,
    thickness,
    rule=None,
    stab_fun=None,
    drilling_stiffness_scale=1.0,
):
    """Element stiffness matrices (ne,24,24) of FEMMShellQ4RS.
    `thickness`: scalar, (ne,) or (ne,npts).  src/FEMMShellQ4RSModule.jl:877-947."""
    stab_fun = stab_fun or stab_lyly(Q4_DEFAULT_ALPHA)
    pc, w = rule if rule is not None else fx.gauss_rule_2x2()
    c = np.asarray(conn) - 1
    ne = c.shape[0]
    X = xyz[c]
    Dt = Dt * (5 / 6)
    h = q4_diameter(X)
    tt = np.asarray(thickness, dtype=np.float64)
    elmat = np.zeros((ne, 24, 24))
    for j in range(len(w)):
        _, _, Jac, E_G, ec, g, T = _q4_gp_setup(X, normals, normal_valid, conn, *pc[j])
        t = tt[:, j] if tt.ndim == 2 else np.broadcast_to(tt, (ne,))
        Bm = np.einsum("eij,ejk->eik", _bm(g), T)
        add_btdb_ut_only(elmat, Bm, t * Jac * w[j], Dps)
        Bb = np.einsum("eij,ejk->eik", _bb(g), T)
        add_btdb_ut_only(elmat, Bb, (t**3 / 12.0) * Jac * w[j], Dps)
        Bs = np.einsum("eij,ejk->eik", _q4_mitc_bs(ec, *pc[j]), T)
        add_btdb_ut_only(elmat, Bs, t * stab_fun(t, h) * Jac * w[j], Dt)
    complete_lt(elmat)
    return _q4_add_drilling(elmat, normals, normal_valid, conn, drilling_stiffness_scale)


def q4rs_mass_elmats(xyz, conn, rho, thickness, rule=None):
    """src/FEMMShellQ4RSModule.jl:968-1022."""
    pc, w = rule if rule is not None else fx.gauss_rule_2x2()
    c = np.asarray(conn) - 1
    ne = c.shape[0]
    X = xyz[c]
    tt = np.asarray(thickness, dtype=np.float64)
    tmass = np.zeros(ne)
    rmass = np.zeros(ne)
    for j in range(len(w)):
        _, dNp = fx.q4_shape(*pc[j])
        J = np.einsum("eai,ak->eik", X, dNp)
        Jac = np.sqrt(np.sum(np.cross(J[:, :, 0], J[:, :, 1]) ** 2, axis=1))
        t = tt[:, j] if tt.ndim == 2 else np.broadcast_to(tt, (ne,))
        tmass += rho * t * Jac * w[j]
        rmass += rho * t**3 / 12 * Jac * w[j]
    return lumped_elmats(tmass / 4, rmass / 4, 4)


def q4rs_associategeometry(xyz, conn, threshold_angle=30.0, normal_dir=None):
    """Jacobian-weighted nodal normals; always reset first.
    src/FEMMShellQ4RSModule.jl:472-525; Comp: src/FEMMShellQ4RSCompModule.jl:447-511."""
    c = np.asarray(conn) - 1
    nn = xyz.shape[0]
    X = xyz[c]
    pcn, _ = fx.nodal_rule_q4()
    normals = np.zeros((nn, 3))
    enormals = []
    for j in range(4):
        _, dNp = fx.q4_shape(*pcn[j])
        J = np.einsum("eai,ak->eik", X, dNp)
        Jac = np.sqrt(np.sum(np.cross(J[:, :, 0], J[:, :, 1]) ** 2, axis=1))
        if normal_dir is None:
            n = e_g(J)[:, :, 2]
        else:
            nd_ = np.asarray(normal_dir, float)
            n = nd_[:, j] if nd_.ndim == 3 else np.broadcast_to(nd_, (len(c), 3))
        enormals.append(n)
        np.add.at(normals, c[:, j], Jac[:, None] * n)
    nrm = np.sqrt(np.sum(normals**2, axis=1))
    normals = np.where((nrm > 0)[:, None], normals / np.where(nrm > 0, nrm, 1.0)[:, None], normals)
    ntol = 1 - np.sqrt(1 - np.sin(threshold_angle / 180 * np.pi) ** 2)
    valid = np.ones(nn, dtype=bool)
    for j in range(4):
        nd = np.einsum("ei,ei->e", normals[c[:, j]], enormals[j])
        valid[c[nd < 1 - ntol, j]] = False
    return normals, valid


# ---------------------------------------------------------------------------
# Q4RS composite
# ---------------------------------------------------------------------------


def q4rscomp_stiffness_elmats(
    xyz,
    conn,
    normals,
    normal_valid,
    A,
    B,
    D,
    H,
    layup_thickness,
    lcsmat,
    rule=None,
    stab_fun=None,
    drilling_stiffness_scale=1.0,
):
    """Element stiffness of FEMMShellQ4RSComp for one layup group.
    `lcsmat`: (3,3), (ne,3,3) or (ne,npts,3,3) layup csys matrices (the reference
    evaluates the csys callback per Gauss point with the shape-function vector as
    the "location", App. B.9).  src/FEMMShellQ4RSCompModule.jl:861-958."""
    stab_fun = stab_fun or stab_lyly(Q4_DEFAULT_ALPHA)
    pc, w = rule if rule is not None else fx.gauss_rule_2x2()
    c = np.asarray(conn) - 1
    ne = c.shape[0]
    X = xyz[c]
    h = q4_diameter(X)
    t = layup_thickness
    lcs = np.asarray(lcsmat, dtype=np.float64)
    bc = lambda M: np.broadcast_to(M, (ne,) + M.shape).copy()
    elmat = np.zeros((ne, 24, 24))
    for j in range(len(w)):
        _, _, Jac, E_G, ec, g, T = _q4_gp_setup(X, normals, normal_valid, conn, *pc[j])
        l = lcs[:, j] if lcs.ndim == 4 else np.broadcast_to(lcs, (ne, 3, 3))
        m, n = layup2element_angle(E_G, l)
        Tps = plane_stress_Tbar(m, n)
        Tts = transverse_shear_T(m, n)
        sA, sB, sD, sH = qteq(bc(A), Tps), qteq(bc(B), Tps), qteq(bc(D), Tps), qteq(bc(H), Tts)
        Bm = np.einsum("eij,ejk->eik", _bm(g), T)
        Bb = np.einsum("eij,ejk->eik", _bb(g), T)
        add_btdb_ut_only(elmat, Bm, Jac * w[j], sA)
        add_btdb_ut_only(elmat, Bb, Jac * w[j], sD)
        add_b1tdb2(elmat, Bm, Bb, Jac * w[j], sB)
        add_b1tdb2(elmat, Bb, Bm, Jac * w[j], sB)
        Bs = np.einsum("eij,ejk->eik", _q4_mitc_bs(ec, *pc[j]), T)
        add_btdb_ut_only(elmat, Bs, stab_fun(t, h) * Jac * w[j], sH)
    complete_lt(elmat)
    return _q4_add_drilling(elmat, normals, normal_valid, conn, drilling_stiffness_scale)


def q4rscomp_mass_elmats(xyz, conn, mass_density, moi_density, rule=None):
    """src/FEMMShellQ4RSCompModule.jl:979-1038."""
    pc, w = rule if rule is not None else fx.gauss_rule_2x2()
    c = np.asarray(conn) - 1
    X = xyz[c]
    tmass = np.zeros(c.shape[0])
    rmass = np.zeros(c.shape[0])
    for j in range(len(w)):
        _, dNp = fx.q4_shape(*pc[j])
        J = np.einsum("eai,ak->eik", X, dNp)
        Jac = np.sqrt(np.sum(np.cross(J[:, :, 0], J[:, :, 1]) ** 2, axis=1))
        tmass += mass_density * Jac * w[j]
        rmass += moi_density * Jac * w[j]
    return lumped_elmats(tmass / 4, rmass / 4, 4)


# ---------------------------------------------------------------------------
# inspectintegpoints: stress resultants per element / integration point
# ---------------------------------------------------------------------------
BENDING_MOMENT, TRANSVERSE_SHEAR, MEMBRANE_FORCE = 1, 2, 3


def _rotate_out(quant, vec, E_G, ocs):
    """Rotate a resultant into the output csys (src/FEMMShellT3FFModule.jl:926-956)."""
    ne = E_G.shape[0]
    l = E_G if ocs is None else np.broadcast_to(np.asarray(ocs, dtype=np.float64), (ne, 3, 3))
    m, n = layup2element_angle(E_G, l)
    o2 = np.zeros((ne, 2, 2))
    o2[:, 0, 0] = o2[:, 1, 1] = m
    o2[:, 0, 1] = n
    o2[:, 1, 0] = -n
    out = np.zeros((ne, 3))
    if quant == TRANSVERSE_SHEAR:
        fo = np.einsum("eji,ej->ei", o2, vec)
        out[:, 0:2] = fo
    else:
        M = np.zeros((ne, 2, 2))
        M[:, 0, 0], M[:, 1, 1], M[:, 0, 1], M[:, 1, 0] = vec[:, 0], vec[:, 1], vec[:, 2], vec[:, 2]
        mo = np.einsum("eji,ejk,ekl->eil", o2, M, o2)
        out[:, 0], out[:, 1], out[:, 2] = mo[:, 0, 0], mo[:, 1, 1], mo[:, 0, 1]
    return out


def t3ff_resultants(xyz, conn, normals, normal_valid, Dps, Dt, thickness, u, quant, ocs=None, stab_fun=None):
    """(ne,3) resultants of FEMMShellT3FF.  `u` (nnodes,6).  src/FEMMShellT3FFModule.jl:850-962."""
    stab_fun = stab_fun or stab_lyly(T3_DEFAULT_ALPHA)
    c = np.asarray(conn) - 1
    ne = c.shape[0]
    X = xyz[c]
    J0, E_G, ec, g, Ae = t3_geometry(X)
    t = np.broadcast_to(np.asarray(thickness, dtype=np.float64), (ne,))
    edisp = u[c].reshape(ne, 18)
    A_Es, nvalid = nodal_triads_e(E_G, normals, normal_valid, conn)
    edisp_n = np.einsum("eij,ej->ei", transfmat_g_to_a(A_Es, E_G), edisp)
    edisp_e = np.einsum("eij,ej->ei", transfmat_a_to_e(A_Es, g), edisp_n)
    if quant == BENDING_MOMENT:
        vec = ((t**3) / 12)[:, None] * np.einsum("ij,ej->ei", Dps, np.einsum("eij,ej->ei", _bb(g), edisp_e))
    elif quant == MEMBRANE_FORCE:
        vec = t[:, None] * np.einsum("ij,ej->ei", Dps, np.einsum("eij,ej->ei", _bm(g), edisp_e))
    else:
        h = np.sqrt(2 * Ae)
        shr = np.einsum("eij,ej->ei", _t3_bs(ec, Ae), edisp_e)
        vec = (t * stab_fun(t, h))[:, None] * np.einsum("ij,ej->ei", Dt * (5 / 6), shr)
    return _rotate_out(quant, vec, E_G, ocs)


def q4rs_resultants(xyz, conn, normals, normal_valid, Dps, Dt, thickness, u, quant, ocs=None, rule=None, stab_fun=None):
    """(ne,npts,3) resultants of FEMMShellQ4RS.  src/FEMMShellQ4RSModule.jl:1061-1170."""
    stab_fun = stab_fun or stab_lyly(Q4_DEFAULT_ALPHA)
    pc, w = rule if rule is not None else fx.gauss_rule_2x2()
    c = np.asarray(conn) - 1
    ne = c.shape[0]
    X = xyz[c]
    h = q4_diameter(X)
    t = np.broadcast_to(np.asarray(thickness, dtype=np.float64), (ne,))
    edisp = u[c].reshape(ne, 24)
    out = np.zeros((ne, len(w), 3))
    for j in range(len(w)):
        _, _, Jac, E_G, ec, g, T = _q4_gp_setup(X, normals, normal_valid, conn, *pc[j])
        if quant == BENDING_MOMENT:
            B = np.einsum("eij,ejk->eik", _bb(g), T)
            vec = ((t**3) / 12)[:, None] * np.einsum("ij,ej->ei", Dps, np.einsum("eij,ej->ei", B, edisp))
        elif quant == MEMBRANE_FORCE:
            B = np.einsum("eij,ejk->eik", _bm(g), T)
            vec = t[:, None] * np.einsum("ij,ej->ei", Dps, np.einsum("eij,ej->ei", B, edisp))
        else:
            B = np.einsum("eij,ejk->eik", _q4_mitc_bs(ec, *pc[j]), T)
            vec = (t * stab_fun(t, h))[:, None] * np.einsum("ij,ej->ei", Dt * (5 / 6), np.einsum("eij,ej->ei", B, edisp))
        out[:, j, :] = _rotate_out(quant, vec, E_G, ocs)
    return out


def _laminate_resultant(quant, sA, sB, sD, sH, memstr, kurv, shrstr, stab):
    """mom = sB eps + sD kappa; frc = sA eps + sB kappa; shear = stab_fun sH gamma
    (src/FEMMShellT3FFCompModule.jl:916-937)."""
    if quant == BENDING_MOMENT:
        return np.einsum("eij,ej->ei", sB, memstr) + np.einsum("eij,ej->ei", sD, kurv)
    if quant == MEMBRANE_FORCE:
        return np.einsum("eij,ej->ei", sA, memstr) + np.einsum("eij,ej->ei", sB, kurv)
    return stab[:, None] * np.einsum("eij,ej->ei", sH, shrstr)


def t3ffcomp_resultants(xyz, conn, normals, normal_valid, A, B, D, H, layup_thickness, lcsmat, u, quant, ocs=None, stab_fun=None):
    """(ne,3) resultants of FEMMShellT3FFComp (one layup group); the default output csys is the layup's
    (`outputcsys = self.layup_groups[1][1].csys`).  src/FEMMShellT3FFCompModule.jl:809-943."""
    stab_fun = stab_fun or stab_lyly(T3_DEFAULT_ALPHA)
    c = np.asarray(conn) - 1
    ne = c.shape[0]
    X = xyz[c]
    J0, E_G, ec, g, Ae = t3_geometry(X)
    edisp = u[c].reshape(ne, 18)
    A_Es, nvalid = nodal_triads_e(E_G, normals, normal_valid, conn)
    edisp_n = np.einsum("eij,ej->ei", transfmat_g_to_a(A_Es, E_G), edisp)
    edisp_e = np.einsum("eij,ej->ei", transfmat_a_to_e(A_Es, g), edisp_n)
    lcs = np.broadcast_to(np.asarray(lcsmat, dtype=np.float64), (ne, 3, 3))
    m, n = layup2element_angle(E_G, lcs)
    Tps, Tts = plane_stress_Tbar(m, n), transverse_shear_T(m, n)
    bc = lambda M: np.broadcast_to(M, (ne,) + M.shape).copy()
    sA, sB, sD, sH = qteq(bc(A), Tps), qteq(bc(B), Tps), qteq(bc(D), Tps), qteq(bc(H), Tts)
    kurv = np.einsum("eij,ej->ei", _bb(g), edisp_e)
    memstr = np.einsum("eij,ej->ei", _bm(g), edisp_e)
    shr = np.einsum("eij,ej->ei", _t3_bs(ec, Ae), edisp_e)
    t = np.full(ne, float(layup_thickness))
    vec = _laminate_resultant(quant, sA, sB, sD, sH, memstr, kurv, shr, stab_fun(t, np.sqrt(2 * Ae)))
    return _rotate_out(quant, vec, E_G, lcs if ocs is None else ocs)


def q4rscomp_resultants(xyz, conn, normals, normal_valid, A, B, D, H, layup_thickness, lcsmat, u, quant, ocs=None, rule=None, stab_fun=None):
    """(ne,npts,3) resultants of FEMMShellQ4RSComp (one layup group).  As written in the reference the strains
    are B T (T u): `edisp_e = T edisp` (:1163) and the B matrices already carry T (:1181-1184, :1198) -- SURVEY
    App. B.9; restated as is.  src/FEMMShellQ4RSCompModule.jl:1061-1210."""
    stab_fun = stab_fun or stab_lyly(Q4_DEFAULT_ALPHA)
    pc, w = rule if rule is not None else fx.gauss_rule_2x2()
    c = np.asarray(conn) - 1
    ne = c.shape[0]
    X = xyz[c]
    h = q4_diameter(X)
    t = np.full(ne, float(layup_thickness))
    edisp = u[c].reshape(ne, 24)
    lcs = np.asarray(lcsmat, dtype=np.float64)
    bc = lambda M: np.broadcast_to(M, (ne,) + M.shape).copy()
    out = np.zeros((ne, len(w), 3))
    for j in range(len(w)):
        _, _, Jac, E_G, ec, g, T = _q4_gp_setup(X, normals, normal_valid, conn, *pc[j])
        l = lcs[:, j] if lcs.ndim == 4 else np.broadcast_to(lcs, (ne, 3, 3))
        m, n = layup2element_angle(E_G, l)
        Tps, Tts = plane_stress_Tbar(m, n), transverse_shear_T(m, n)
        sA, sB, sD, sH = qteq(bc(A), Tps), qteq(bc(B), Tps), qteq(bc(D), Tps), qteq(bc(H), Tts)
        edisp_e = np.einsum("eij,ej->ei", T, edisp)
        Bm = np.einsum("eij,ejk->eik", _bm(g), T)
        Bb = np.einsum("eij,ejk->eik", _bb(g), T)
        Bs = np.einsum("eij,ejk->eik", _q4_mitc_bs(ec, *pc[j]), T)
        kurv = np.einsum("eij,ej->ei", Bb, edisp_e)
        memstr = np.einsum("eij,ej->ei", Bm, edisp_e)
        shr = np.einsum("eij,ej->ei", Bs, edisp_e)
        vec = _laminate_resultant(quant, sA, sB, sD, sH, memstr, kurv, shr, stab_fun(t, h))
        if ocs is None:
            o = l
        else:
            o = np.asarray(ocs, dtype=np.float64)
            o = o[:, j] if o.ndim == 4 else o
        out[:, j, :] = _rotate_out(quant, vec, E_G, o)
    return out
