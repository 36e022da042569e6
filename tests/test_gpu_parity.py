"""GPU parity tests: libfsgpu (through the C ABI / the reference-facing Python mirror)
against the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): sparsity patterns / colptr / rowval bit-exact;
element matrices, assembled values, force vectors <= 1e-12 relative Frobenius norm (FP64).
"""
import os

import numpy as np
import pytest

from oracle import beam as obeam
from oracle import explicit as oexp
from oracle import fe_external as fx
from oracle import layup as oly
from oracle import shells as osh
from tests import meshes

pytestmark = pytest.mark.gpu

TOL = 1e-12


def relfro(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


@pytest.fixture(scope="module")
def fs():
    import fsb200

    return fsb200


E_, NU_, T_, RHO_ = 200e9, 0.3, 0.01, 7850.0


def _iso():
    return osh.shell_material_stiffness(fx.moduli_iso(E_, NU_))


def _layup():
    D6 = oly.lamina_moduli(133860e6, 7706e6, 0.301, 4306e6, 4306e6, 2760e6)
    plies = [oly.Ply(f"p{k}", D6, 0.0025, a, 1500.0) for k, a in enumerate((0, 90, 45, -30))]
    lay = oly.CompositeLayup("test", plies)
    th = np.deg2rad(20.0)
    cs = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    return lay, cs


def _fs_layup(fs, cs):
    f = fs.femm
    mat = f.lamina_material(1500.0, 133860e6, 7706e6, 0.301, 4306e6, 4306e6, 2760e6)
    plies = [f.Ply(f"p{k}", mat, 0.0025, a) for k, a in enumerate((0, 90, 45, -30))]
    return f.CompositeLayup("test", plies, cs)


def _make_femm(fs, kind, conn, comp=False, cs=None):
    f = fs.femm
    if kind == "t3":
        idom = f.IntegDomain(conn, None, T_)
        return f.FEMMShellT3FFComp(idom, _fs_layup(fs, cs)) if comp else f.FEMMShellT3FF(idom, f.MatDeforElastIso(E_, NU_, RHO_))
    idom = f.IntegDomain(conn, f.GaussRule2x2(), T_)
    return f.FEMMShellQ4RSComp(idom, _fs_layup(fs, cs)) if comp else f.FEMMShellQ4RS(idom, f.MatDeforElastIso(E_, NU_, RHO_))


def _oracle_normals(kind, xyz, conn, fixed=None):
    if kind == "t3":
        return osh.t3ff_associategeometry(xyz, conn, normal_dir=fixed)
    return osh.q4rs_associategeometry(xyz, conn, normal_dir=fixed)


def _oracle_K(kind, comp, xyz, conn, normals, valid, cs=None, drill=1.0, sheark=0):
    Dps, Dt = _iso()
    if comp:
        lay, _ = _layup()
        A, B, D = lay.laminate_stiffnesses()
        H = lay.laminate_transverse_stiffness()
        if kind == "t3":
            return osh.t3ffcomp_stiffness_elmats(xyz, conn, normals, valid, A, B, D, H, lay.thickness, cs, drilling_stiffness_scale=drill, transv_shear_formulation=sheark)
        return osh.q4rscomp_stiffness_elmats(xyz, conn, normals, valid, A, B, D, H, lay.thickness, cs, drilling_stiffness_scale=drill)
    if kind == "t3":
        return osh.t3ff_stiffness_elmats(xyz, conn, normals, valid, Dps, Dt, T_, drilling_stiffness_scale=drill, transv_shear_formulation=sheark)
    return osh.q4rs_stiffness_elmats(xyz, conn, normals, valid, Dps, Dt, T_, drilling_stiffness_scale=drill)


def _oracle_M(kind, comp, xyz, conn):
    if comp:
        lay, _ = _layup()
        md, mi = lay.laminate_inertia()
        return (osh.t3ffcomp_mass_elmats if kind == "t3" else osh.q4rscomp_mass_elmats)(xyz, conn, md, mi)
    return (osh.t3ff_mass_elmats if kind == "t3" else osh.q4rs_mass_elmats)(xyz, conn, RHO_, T_)


# ---------------------------------------------------------------------------------------
# nodal normals
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["t3", "q4"])
def test_associategeometry(fs, kind):
    xyz, conn = meshes.shell_mesh(kind, n=9)
    femm = _make_femm(fs, kind, conn)
    geom0 = fs.femm.NodalField(xyz)
    fs.femm.associategeometry(femm, geom0)
    n_o, v_o = _oracle_normals(kind, xyz, conn)
    assert (~v_o).sum() > 0, "test mesh must contain invalid normals"
    assert np.array_equal(femm._normal_valid, v_o)
    assert np.abs(femm._normals - n_o).max() < 1e-13


# ---------------------------------------------------------------------------------------
# raw element matrices
# ---------------------------------------------------------------------------------------
def test_t3ffcomp_with_csys_callback(fs):
    """FEMMShellT3FFComp with a cylindrical layup csys CALLBACK: nodal normals from the csys evaluated at the nodes,
    laminate rotation from the csys evaluated at the centroids (src/FEMMShellT3FFCompModule.jl:509,617)."""
    from fsb200 import workloads as wl

    f = fs.femm
    w = wl.c3_t3ffcomp_cylinder(24, 6)
    xyz, conn = np.asarray(w["xyz"]), w["conn"]
    lay, _ = _layup()
    femm = f.FEMMShellT3FFComp(f.IntegDomain(conn, None, T_), _fs_layup(fs, wl.cylindrical_csys))
    geom0 = f.NodalField(xyz)
    f.associategeometry(femm, geom0)
    dirs = wl.cylindrical_csys(xyz[conn - 1].reshape(-1, 3))[:, :, 2].reshape(conn.shape[0], 3, 3)
    no, vo = osh.t3ff_associategeometry(xyz, conn, normal_dir=dirs)
    assert np.abs(femm._normals - no).max() < 1e-14 and np.array_equal(femm._normal_valid, vo)
    lcs = wl.cylindrical_csys(xyz[conn - 1].mean(axis=1))
    A, B, D = lay.laminate_stiffnesses()
    H = lay.laminate_transverse_stiffness()
    Ko = osh.t3ffcomp_stiffness_elmats(xyz, conn, no, vo, A, B, D, H, lay.thickness, lcs)
    femm._sync_mesh(geom0)
    femm._sync_stab()
    Kg = femm.ctx.element_matrices(13, 0, femm._params())
    assert relfro(Kg, Ko) < TOL


@pytest.mark.parametrize("kind", ["t3", "q4"])
def test_associategeometry_general_csys(fs, kind):
    """Nodal normals from a csys evaluated per element and node (cylindrical layup csys of a laminated cylinder,
    examples/shells/dynamics/homogeneous/explicit/clamp_cyl_expl_examples.jl:62-68): fsgpu_associategeometry_dirs."""
    rng = np.random.default_rng(4)
    gen = fx.t3block if kind == "t3" else fx.q4block
    xy, conn = gen(2 * np.pi * 0.9, 1.0, 14, 5)
    R = 0.5
    xyz = np.column_stack([R * np.cos(xy[:, 0]), R * np.sin(xy[:, 0]), xy[:, 1]])
    xyz += rng.uniform(-1, 1, xyz.shape) * 0.01
    X = xyz[conn - 1]  # (angle, z) parametrisation: the element normals point outward
    dirs = X.copy()
    dirs[:, :, 2] = 0.0
    dirs /= np.linalg.norm(dirs, axis=2, keepdims=True)  # radial direction at every node of every element
    ctx = fs.Context()
    ctx.set_mesh(conn, xyz)
    ctx.associategeometry_dirs(dirs, 30.0)
    n, v = ctx.get_normals()
    no, vo = _oracle_normals(kind, xyz, conn, fixed=dirs)
    assert np.abs(n - no).max() < 1e-14
    assert np.array_equal(v, vo)
    assert v.sum() > v.size // 2


@pytest.mark.parametrize("kind,comp,sheark", [("t3", False, 0), ("t3", False, 1), ("t3", True, 0), ("t3", True, 1), ("q4", False, 0), ("q4", True, 0)])
def test_element_stiffness(fs, kind, comp, sheark):
    xyz, conn = meshes.shell_mesh(kind, n=7)
    lay_cs = _layup()[1] if comp else None
    normals, valid = _oracle_normals(kind, xyz, conn)
    femm = _make_femm(fs, kind, conn, comp, lay_cs)
    femm.drilling_stiffness_scale = 0.8
    femm.transv_shear_formulation = sheark
    geom0 = fs.femm.NodalField(xyz)
    femm._sync_mesh(geom0)
    femm.ctx.set_normals(normals, valid)
    Kg = femm.ctx.element_matrices(femm._kind(), 0, femm._params())
    Ko = _oracle_K(kind, comp, xyz, conn, normals, valid, lay_cs, drill=0.8, sheark=sheark)
    worst = max(relfro(Kg[e], Ko[e]) for e in range(conn.shape[0]))
    assert worst < TOL, worst


@pytest.mark.parametrize("comp", [False, True])
@pytest.mark.parametrize("rule", ["simpson13", "gauss1"])
def test_q4_other_integration_rules(fs, comp, rule):
    """9-point Simpson rule (chunks of 4 points: the CHUNKED kernel) and a 1-point rule (idle point lanes):
    element matrices, assembled stiffness and lumped mass against the oracle."""
    xyz, conn = meshes.shell_mesh("q4", n=6)
    lay, cs = _layup()
    f = fs.femm
    orule = fx.simpson13_rule_2d() if rule == "simpson13" else fx.gauss_rule_1x1()
    grule = f.Simpson13Rule2() if rule == "simpson13" else f.GaussRule1x1()
    assert np.array_equal(orule[0], grule[0]) and np.array_equal(orule[1], grule[1])
    idom = f.IntegDomain(conn, grule, T_)
    femm = f.FEMMShellQ4RSComp(idom, _fs_layup(fs, cs)) if comp else f.FEMMShellQ4RS(idom, f.MatDeforElastIso(E_, NU_, RHO_))
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6))).numberdofs()
    f.associategeometry(femm, geom0)
    normals, valid = _oracle_normals("q4", xyz, conn, fixed=cs[:, 2] if comp else None)
    if comp:
        A, B, D = lay.laminate_stiffnesses()
        H = lay.laminate_transverse_stiffness()
        Ko = osh.q4rscomp_stiffness_elmats(xyz, conn, normals, valid, A, B, D, H, lay.thickness, cs, rule=orule)
        md, mi = lay.laminate_inertia()
        Mo = osh.q4rscomp_mass_elmats(xyz, conn, md, mi, rule=orule)
    else:
        Dps, Dt = _iso()
        Ko = osh.q4rs_stiffness_elmats(xyz, conn, normals, valid, Dps, Dt, T_, rule=orule)
        Mo = osh.q4rs_mass_elmats(xyz, conn, RHO_, T_, rule=orule)
    Kg = femm.ctx.element_matrices(femm._kind(), 0, femm._params())
    worst = max(relfro(Kg[e], Ko[e]) for e in range(conn.shape[0]))
    assert worst < TOL, worst
    u0, R0 = f.NodalField(np.zeros((xyz.shape[0], 3))), f.initial_Rfield(xyz.shape[0])
    od = fx.DofField(xyz.shape[0]).numberdofs()
    dn = od.gatherdofnums(conn)
    K = f.stiffness(femm, f.SysmatAssemblerSparse(), geom0, u0, R0, dchi)
    _check_matrix(K, fx.assemble_matrix("sparse", Ko, dn, od.nalldofs), od.nalldofs)
    M = f.mass(femm, f.SysmatAssemblerSparseDiag(), geom0, dchi)
    _check_matrix(M, fx.assemble_matrix("diag", Mo, dn, od.nalldofs), od.nalldofs)


@pytest.mark.parametrize("kind,comp", [("t3", False), ("t3", True), ("q4", False), ("q4", True)])
def test_element_mass(fs, kind, comp):
    xyz, conn = meshes.shell_mesh(kind, n=5)
    lay_cs = _layup()[1] if comp else None
    normals, valid = _oracle_normals(kind, xyz, conn)
    femm = _make_femm(fs, kind, conn, comp, lay_cs)
    femm._sync_mesh(fs.femm.NodalField(xyz))
    femm.ctx.set_normals(normals, valid)
    Mg = femm.ctx.element_matrices(femm._kind(), 1, femm._params())
    Mo = _oracle_M(kind, comp, xyz, conn)
    assert relfro(Mg, Mo) < TOL


# ---------------------------------------------------------------------------------------
# inspectintegpoints (batched resultants)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["t3", "q4"])
def test_resultants(fs, kind):
    f = fs.femm
    xyz, conn = meshes.shell_mesh(kind, n=7)
    femm = _make_femm(fs, kind, conn)
    geom0 = f.NodalField(xyz)
    f.associategeometry(femm, geom0)
    normals, valid = _oracle_normals(kind, xyz, conn)
    Dps, Dt = _iso()
    rng = np.random.default_rng(13)
    u = rng.standard_normal((xyz.shape[0], 6)) * 1e-3
    th = np.deg2rad(25.0)
    ocs = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    ofun = osh.t3ff_resultants if kind == "t3" else osh.q4rs_resultants
    for q, name in ((1, "moment"), (2, "shear"), (3, "membrane")):
        got = f.inspectintegpoints(femm, geom0, f.NodalField(u), None, name, outputcsys=ocs)
        ref = ofun(xyz, conn, normals, valid, Dps, Dt, T_, u, q, ocs=ocs)
        ref = ref.reshape(got.shape)
        assert relfro(got, ref) < 1e-11, (name, relfro(got, ref))
        # default output csys = element triad: the reference's sqrt(1 - m^2) with m ~ 1 amplifies
        # round-off to ~1e-8, so this case is only comparable to that level
        got = f.inspectintegpoints(femm, geom0, f.NodalField(u), None, name)
        ref = ofun(xyz, conn, normals, valid, Dps, Dt, T_, u, q).reshape(got.shape)
        assert relfro(got, ref) < 1e-6


@pytest.mark.parametrize("kind", ["t3", "q4"])
def test_resultants_laminated(fs, kind):
    """inspectintegpoints of FEMMShellT3FFComp / FEMMShellQ4RSComp (the latter with the reference's double
    application of T): default output csys = layup csys, and an explicit one."""
    f = fs.femm
    xyz, conn = meshes.shell_mesh(kind, n=7)
    lay, cs = _layup()
    femm = _make_femm(fs, kind, conn, comp=True, cs=cs)
    geom0 = f.NodalField(xyz)
    f.associategeometry(femm, geom0)
    normals, valid = _oracle_normals(kind, xyz, conn, fixed=cs[:, 2])
    A, B, D = lay.laminate_stiffnesses()
    H = lay.laminate_transverse_stiffness()
    rng = np.random.default_rng(17)
    u = rng.standard_normal((xyz.shape[0], 6)) * 1e-3
    th = np.deg2rad(-35.0)
    ocs = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    ofun = osh.t3ffcomp_resultants if kind == "t3" else osh.q4rscomp_resultants
    for q, name in ((1, "moment"), (2, "shear"), (3, "membrane")):
        for o in (None, ocs):
            got = f.inspectintegpoints(femm, geom0, f.NodalField(u), None, name, outputcsys=o)
            ref = ofun(xyz, conn, normals, valid, A, B, D, H, lay.thickness, cs, u, q, ocs=o).reshape(got.shape)
            assert relfro(got, ref) < 1e-11, (name, o is None, relfro(got, ref))
    # felist selects elements
    sel = np.array([3, 1, 7])
    got = f.inspectintegpoints(femm, geom0, f.NodalField(u), sel, "moment")
    ref = ofun(xyz, conn, normals, valid, A, B, D, H, lay.thickness, cs, u, 1)
    assert relfro(got, ref.reshape(conn.shape[0], -1, 3)[sel - 1]) < 1e-11


# ---------------------------------------------------------------------------------------
# assembled matrices, every assembler target: pattern bit-exact, values 1e-12
# ---------------------------------------------------------------------------------------
ASM = {"sparse": "SysmatAssemblerSparse", "symm": "SysmatAssemblerSparseSymm", "diag": "SysmatAssemblerSparseDiag", "ffblock": "ffblock", "ffblock_diag": "ffblock_diag", "csrsymm": "SysmatAssemblerSparseCSRSymm"}


def _assembler(fs, name):
    f = fs.femm
    if name == "ffblock":
        return f.SysmatAssemblerFFBlock()
    if name == "ffblock_diag":
        return f.SysmatAssemblerFFBlock(f.SysmatAssemblerSparseDiag())
    return getattr(f, ASM[name])()


def _check_matrix(S, ref, nrows):
    cp, rv, nz = ref
    assert S.colptr.dtype == np.int64 and S.rowval.dtype == np.int64
    assert np.array_equal(S.colptr, cp), "colptr differs"
    assert np.array_equal(S.rowval, rv), "rowval differs"
    assert relfro(S.nzval, nz) < TOL, relfro(S.nzval, nz)
    assert S.m == nrows and S.n == nrows


@pytest.mark.parametrize("kind", ["t3", "q4"])
@pytest.mark.parametrize("asm", ["sparse", "symm", "diag", "ffblock", "ffblock_diag", "csrsymm"])
def test_assembled_stiffness_and_mass(fs, kind, asm):
    xyz, conn = meshes.shell_mesh(kind, n=8)
    od = meshes.clamp_edge_dofs(xyz)
    femm = _make_femm(fs, kind, conn)
    f = fs.femm
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6)))
    dchi.is_fixed[:] = od.is_fixed
    dchi.numberdofs()
    assert np.array_equal(dchi.dofnums, od.dofnums)
    f.associategeometry(femm, geom0)
    normals, valid = _oracle_normals(kind, xyz, conn)
    u0 = f.NodalField(np.zeros((xyz.shape[0], 3)))
    R0 = f.initial_Rfield(xyz.shape[0])
    K = f.stiffness(femm, _assembler(fs, asm), geom0, u0, R0, dchi)
    M = f.mass(femm, _assembler(fs, asm), geom0, dchi)
    dn = od.gatherdofnums(conn)
    Ko = _oracle_K(kind, False, xyz, conn, normals, valid)
    Mo = _oracle_M(kind, False, xyz, conn)
    n = od.nfreedofs if asm.startswith("ffblock") else od.nalldofs
    _check_matrix(K, fx.assemble_matrix(asm, Ko, dn, od.nalldofs, od.nfreedofs), n)
    _check_matrix(M, fx.assemble_matrix(asm, Mo, dn, od.nalldofs, od.nfreedofs), n)
    if asm == "symm":
        A = K.to_scipy()
        assert abs(A - A.T).max() == 0.0, "SparseSymm result must be exactly symmetric"


@pytest.mark.parametrize("kind,comp", [("t3", False), ("t3", True), ("q4", False)])
@pytest.mark.parametrize("variant", ["shuffled", "duplicated", "reversed"])
def test_element_order_and_duplicates(fs, kind, comp, variant):
    """The fast-path kernels merge, inside a warp, the blocks of different elements that land on the same
    matrix block.  The result must not depend on which elements share a warp: random element order (hardly any
    merging), reversed order, and repeated elements (the same three nodes twice in one warp, also with the
    opposite edge orientation) are all checked against the oracle."""
    xyz, conn = meshes.shell_mesh(kind, n=9)
    rng = np.random.default_rng(5)
    if variant == "shuffled":
        conn = conn[rng.permutation(conn.shape[0])]
    elif variant == "reversed":
        conn = conn[::-1].copy()
    else:
        # every 7th element twice in a row; T3 copies start from another node (same triangle, rotated connectivity)
        rep = conn[::7]
        if kind == "t3":
            rep = np.roll(rep, 1, axis=1)
        parts = []
        for k in range(0, conn.shape[0], 7):
            parts.append(conn[k:k + 1])
            parts.append(rep[k // 7:k // 7 + 1])
            parts.append(conn[k + 1:k + 7])
        conn = np.ascontiguousarray(np.concatenate(parts))
    lay, cs_all = _layup()
    cs = cs_all if comp else None
    od = meshes.clamp_edge_dofs(xyz)
    f = fs.femm
    femm = _make_femm(fs, kind, conn, comp, cs)
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6)))
    dchi.is_fixed[:] = od.is_fixed
    dchi.numberdofs()
    f.associategeometry(femm, geom0)
    normals, valid = _oracle_normals(kind, xyz, conn, fixed=cs[:, 2] if comp else None)
    u0, R0 = f.NodalField(np.zeros((xyz.shape[0], 3))), f.initial_Rfield(xyz.shape[0])
    Ko = _oracle_K(kind, comp, xyz, conn, normals, valid, cs)
    dn = od.gatherdofnums(conn)
    for asm in ("ffblock", "sparse"):
        K = f.stiffness(femm, _assembler(fs, asm), geom0, u0, R0, dchi)
        n = od.nfreedofs if asm == "ffblock" else od.nalldofs
        _check_matrix(K, fx.assemble_matrix(asm, Ko, dn, od.nalldofs, od.nfreedofs), n)


def test_symm_drops_exact_zeros_flat_plate(fs):
    """Flat axis-aligned plate: membrane/bending cross terms are exact zeros that
    SysmatAssemblerSparseSymm drops (value-dependent pattern)."""
    xy, conn = fx.t3block(2.0, 1.0, 6, 4)
    xyz = fx.xyz3(xy)
    f = fs.femm
    femm = _make_femm(fs, "t3", conn)
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6))).numberdofs()
    f.associategeometry(femm, geom0)
    u0, R0 = f.NodalField(np.zeros((xyz.shape[0], 3))), f.initial_Rfield(xyz.shape[0])
    K = f.stiffness(femm, geom0, u0, R0, dchi)  # default assembler = SparseSymm
    normals, valid = _oracle_normals("t3", xyz, conn)
    Ko = _oracle_K("t3", False, xyz, conn, normals, valid)
    od = fx.DofField(xyz.shape[0]).numberdofs()
    cp, rv, nz = fx.assemble_matrix("symm", Ko, od.gatherdofnums(conn), od.nalldofs)
    full = fx.assemble_matrix("sparse", Ko, od.gatherdofnums(conn), od.nalldofs)
    assert len(rv) < len(full[1]), "the flat plate must have droppable zeros"
    assert len(K.rowval) < len(full[1]), "the GPU path must drop exact zeros too"
    n = od.nalldofs
    Kg, Kr = K.to_scipy(), fx.csc_to_scipy(cp, rv, nz, n, n)
    assert relfro(Kg.toarray(), Kr.toarray()) < TOL
    assert abs(Kg - Kg.T).max() == 0.0
    # Which entries cancel to an EXACT zero depends on the floating-point operation order
    # (the reference's own pattern depends on its BLAS); patterns may therefore differ, but
    # only in entries at round-off level.
    import scipy.sparse as sp

    Pg = sp.csc_matrix((np.ones_like(Kg.data), Kg.indices, Kg.indptr), shape=Kg.shape)
    Pr = sp.csc_matrix((np.ones_like(Kr.data), Kr.indices, Kr.indptr), shape=Kr.shape)
    diff = (Pg - Pr).tocoo()
    sel = diff.data != 0
    vals = np.abs(np.asarray((Kg + Kr)[diff.row[sel], diff.col[sel]])).ravel()
    assert sel.sum() < 0.05 * len(rv)
    assert vals.size == 0 or vals.max() < 1e-12 * np.abs(nz).max()


@pytest.mark.parametrize("kind", ["t3", "q4"])
def test_assembled_composite(fs, kind):
    xyz, conn = meshes.shell_mesh(kind, n=6)
    lay, cs = _layup()
    f = fs.femm
    femm = _make_femm(fs, kind, conn, True, cs)
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6))).numberdofs()
    f.associategeometry(femm, geom0)
    normals, valid = _oracle_normals(kind, xyz, conn, fixed=cs[:, 2])
    assert np.array_equal(femm._normal_valid, valid)
    u0, R0 = f.NodalField(np.zeros((xyz.shape[0], 3))), f.initial_Rfield(xyz.shape[0])
    K = f.stiffness(femm, f.SysmatAssemblerSparse(), geom0, u0, R0, dchi)
    M = f.mass(femm, f.SysmatAssemblerSparseDiag(), geom0, dchi)
    od = fx.DofField(xyz.shape[0]).numberdofs()
    dn = od.gatherdofnums(conn)
    _check_matrix(K, fx.assemble_matrix("sparse", _oracle_K(kind, True, xyz, conn, normals, valid, cs), dn, od.nalldofs), od.nalldofs)
    _check_matrix(M, fx.assemble_matrix("diag", _oracle_M(kind, True, xyz, conn), dn, od.nalldofs), od.nalldofs)

@pytest.mark.parametrize("comp", [False, True])
def test_deterministic_tile_path(fs, comp):
    """fsgpu_set_deterministic: the owner-computes T3 kernel (no atomics) gives the same pattern, values
    within 1e-12 of the oracle, and bitwise identical values from run to run."""
    if os.environ.get("FSGPU_FORCE_GENERIC"):
        pytest.skip("the tile kernel needs the run-structured (fast) addressing")
    xyz, conn = meshes.shell_mesh("t3", n=12)
    lay, cs = _layup()
    od = meshes.clamp_edge_dofs(xyz)
    f = fs.femm
    femm = _make_femm(fs, "t3", conn, comp, cs)
    femm.ctx.set_deterministic(True)
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6)))
    dchi.is_fixed[:] = od.is_fixed
    dchi.numberdofs()
    f.associategeometry(femm, geom0)
    normals, valid = _oracle_normals("t3", xyz, conn, fixed=cs[:, 2] if comp else None)
    u0, R0 = f.NodalField(np.zeros((xyz.shape[0], 3))), f.initial_Rfield(xyz.shape[0])
    for asm in ("ffblock", "sparse"):
        K1 = f.stiffness(femm, _assembler(fs, asm), geom0, u0, R0, dchi)
        assert femm.ctx.scatter_path == 2, "tile kernel did not engage"
        K2 = f.stiffness(femm, _assembler(fs, asm), geom0, u0, R0, dchi)
        assert np.array_equal(K1.nzval, K2.nzval), "deterministic path must be bitwise reproducible"
        Ko = _oracle_K("t3", comp, xyz, conn, normals, valid, cs)
        n = od.nfreedofs if asm == "ffblock" else od.nalldofs
        _check_matrix(K1, fx.assemble_matrix(asm, Ko, od.gatherdofnums(conn), od.nalldofs, od.nfreedofs), n)
    femm.ctx.set_deterministic(False)
    K3 = f.stiffness(femm, _assembler(fs, "ffblock"), geom0, u0, R0, dchi)
    assert femm.ctx.scatter_path == 1


def test_rcm_like_permuted_numbering(fs):
    """numberdofs!(dchi, perm): dofs of a node are no longer monotone in the node index."""
    xyz, conn = meshes.shell_mesh("t3", n=7)
    rng = np.random.default_rng(5)
    perm = rng.permutation(xyz.shape[0])
    od = meshes.clamp_edge_dofs(xyz)
    od.numberdofs(perm)
    f = fs.femm
    femm = _make_femm(fs, "t3", conn)
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6)))
    dchi.is_fixed[:] = od.is_fixed
    dchi.numberdofs(perm)
    f.associategeometry(femm, geom0)
    normals, valid = _oracle_normals("t3", xyz, conn)
    u0, R0 = f.NodalField(np.zeros((xyz.shape[0], 3))), f.initial_Rfield(xyz.shape[0])
    K = f.stiffness(femm, f.SysmatAssemblerFFBlock(), geom0, u0, R0, dchi)
    Ko = _oracle_K("t3", False, xyz, conn, normals, valid)
    _check_matrix(K, fx.assemble_matrix("ffblock", Ko, od.gatherdofnums(conn), od.nalldofs, od.nfreedofs), od.nfreedofs)


# ---------------------------------------------------------------------------------------
# column blocks for multi-GPU gathering; the collective itself: tests/test_gpu_multi.py
# ---------------------------------------------------------------------------------------
def _local_femm(fs, kind, plan, xyz, normals, valid):
    f = fs.femm
    femm = _make_femm(fs, kind, plan.conn)
    geom = f.NodalField(plan.restrict_nodes(xyz))
    # the nodal normals come from the GLOBAL mesh (halo nodes see elements this rank does not assemble)
    femm._normals, femm._normal_valid = np.asfortranarray(plan.restrict_nodes(normals)), plan.restrict_nodes(valid)
    femm._associatedgeometry = True
    dchi = f.NodalField.__new__(f.NodalField)
    dchi.values, dchi.dofnums, dchi._nfree = None, plan.dofnums, plan.nfree
    return femm, geom, dchi


@pytest.mark.parametrize("kind,asm,world,permute", [("t3", "ffblock", 3, False), ("q4", "sparse", 2, True), ("t3", "diag", 2, False)])
def test_column_blocks_concatenate_to_global_matrix(fs, kind, asm, world, permute):
    """Every 'rank' (here: one after the other on one GPU) assembles its plan's local mesh; the blocks that
    fsgpu_result_block writes, concatenated, are the single-context global matrix: pattern bit-exact."""
    import torch

    f, pt = fs.femm, fs.partition
    xyz, conn = meshes.shell_mesh(kind, n=9)
    od = meshes.clamp_edge_dofs(xyz)
    perm = np.random.default_rng(11).permutation(xyz.shape[0]) if permute else None
    od.numberdofs(perm)
    femm = _make_femm(fs, kind, conn)
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6)))
    dchi.is_fixed[:] = od.is_fixed
    dchi.numberdofs(perm)
    f.associategeometry(femm, geom0)
    Kg = f.stiffness(femm, _assembler(fs, asm), geom0, None, None, dchi)
    dev = torch.device("cuda", 0)
    cnts, rows, vals = [], [], []
    for r in range(world):
        plan = pt.ColumnBlockPlan(conn, dchi.dofnums, od.nfreedofs, asm, r, world)
        lf, lg, ld = _local_femm(fs, kind, plan, xyz, femm._normals, femm._normal_valid)
        f.stiffness(lf, _assembler(fs, asm), lg, None, None, ld)
        nb = lf.ctx.result_block(plan.lcol_lo, plan.lcol_hi)
        cnt = torch.empty(plan.col_hi - plan.col_lo, dtype=torch.int64, device=dev)
        rv = torch.empty(nb, dtype=torch.int64, device=dev)
        nz = torch.empty(nb, dtype=torch.float64, device=dev)
        rm = torch.as_tensor(plan.loc2glob, device=dev)
        torch.cuda.synchronize()
        assert lf.ctx.result_block(plan.lcol_lo, plan.lcol_hi, rm, cnt, rv, nz) == nb
        cnts.append(cnt.cpu().numpy()), rows.append(rv.cpu().numpy()), vals.append(nz.cpu().numpy())
        if world > 1 and not permute:
            assert len(plan.elems) < conn.shape[0]
    colptr = np.concatenate([[1], 1 + np.cumsum(np.concatenate(cnts))])
    assert np.array_equal(colptr, Kg.colptr)
    assert np.array_equal(np.concatenate(rows), Kg.rowval)
    assert relfro(np.concatenate(vals), Kg.nzval) < TOL


def test_result_block_argument_checks(fs):
    f = fs.femm
    xyz, conn = meshes.shell_mesh("t3", n=4)
    femm = _make_femm(fs, "t3", conn)
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6))).numberdofs()
    f.associategeometry(femm, geom0)
    K = f.stiffness(femm, f.SysmatAssemblerSparse(), geom0, None, None, dchi)
    assert femm.ctx.result_block(0, K.n) == K.nzval.size
    assert femm.ctx.result_block(5, 5) == 0
    with pytest.raises(fs.FsgpuError):
        femm.ctx.result_block(3, K.n + 1)
    f.stiffness(femm, f.SysmatAssemblerSparseSymm(), geom0, None, None, dchi)
    with pytest.raises(fs.FsgpuError):
        femm.ctx.result_block(0, 6)


# ---------------------------------------------------------------------------------------
# corotational beam
# ---------------------------------------------------------------------------------------
EB, NUB, RHOB = 71240.0, 0.31, 5e-9


def _beam(fs):
    xyz, conn, u1, R1, sec = meshes.beam_lattice()
    f = fs.femm
    secs = f.FESetL2Beam(sec["A"], sec["I1"], sec["I2"], sec["I3"], sec["J"], sec["A2s"], sec["A3s"], sec["x1x2"])
    femm = f.FEMMCorotBeam(f.IntegDomain(conn), f.MatDeforElastIso(EB, NUB, RHOB), secs)
    od = fx.DofField(xyz.shape[0])
    for c in range(1, 7):
        od.setebc([0], c)
    od.numberdofs()
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6)))
    dchi.is_fixed[:] = od.is_fixed
    dchi.numberdofs()
    return f, femm, xyz, conn, u1, R1, sec, od, dchi


def test_beam_element_level(fs):
    f, femm, xyz, conn, u1, R1, sec, od, dchi = _beam(fs)
    femm._sync_mesh(f.NodalField(xyz))
    femm.ctx.set_state(u1, R1)
    for op, ref in ((0, obeam.beam_stiffness_elmats(xyz, conn, u1, R1, sec, EB, NUB)), (2, obeam.beam_geostiffness_elmats(xyz, conn, u1, R1, sec, EB, NUB))):
        got = femm.ctx.element_matrices(2, op, femm._params())
        assert max(relfro(got[e], ref[e]) for e in range(len(conn))) < TOL
    for mt in range(4):
        got = femm.ctx.element_matrices(2, 1, femm._params(mt))
        ref = obeam.beam_mass_elmats(xyz, conn, u1, R1, sec, RHOB, mt)
        assert max(relfro(got[e], ref[e]) for e in range(len(conn))) < TOL
    ev = femm.ctx.element_vectors(femm._params())
    ref = obeam.beam_restoringforce_elvecs(xyz, conn, u1, R1, sec, EB, NUB)
    assert max(relfro(ev[e], ref[e]) for e in range(len(conn))) < TOL


def test_beam_operators(fs):
    f, femm, xyz, conn, u1, R1, sec, od, dchi = _beam(fs)
    geom0, uf, Rf = f.NodalField(xyz), f.NodalField(u1), f.NodalField(R1)
    dn = od.gatherdofnums(conn)
    K = f.stiffness(femm, f.SysmatAssemblerFFBlock(), geom0, uf, Rf, dchi)
    _check_matrix(K, fx.assemble_matrix("ffblock", obeam.beam_stiffness_elmats(xyz, conn, u1, R1, sec, EB, NUB), dn, od.nalldofs, od.nfreedofs), od.nfreedofs)
    Kg = f.geostiffness(femm, f.SysmatAssemblerSparse(), geom0, uf, Rf, dchi)
    _check_matrix(Kg, fx.assemble_matrix("sparse", obeam.beam_geostiffness_elmats(xyz, conn, u1, R1, sec, EB, NUB), dn, od.nalldofs), od.nalldofs)
    M = f.mass(femm, geom0, uf, Rf, dchi, mass_type=1)
    _check_matrix(M, fx.assemble_matrix("symm", obeam.beam_mass_elmats(xyz, conn, u1, R1, sec, RHOB, 1), dn, od.nalldofs), od.nalldofs)
    ev = obeam.beam_restoringforce_elvecs(xyz, conn, u1, R1, sec, EB, NUB)
    Fr = f.restoringforce(femm, f.SysvecAssemblerFBlock(), geom0, uf, Rf, dchi)
    assert relfro(Fr, fx.assemble_vector(ev, dn, od.nalldofs, od.nfreedofs)) < TOL
    Fa = f.restoringforce(femm, geom0, uf, Rf, dchi)
    assert relfro(Fa, fx.assemble_vector(ev, dn, od.nalldofs)) < TOL


def test_beam_gyroscopic_and_distribloads(fs):
    f, femm, xyz, conn, u1, R1, sec, od, dchi = _beam(fs)
    geom0, uf, Rf = f.NodalField(xyz), f.NodalField(u1), f.NodalField(R1)
    rng = np.random.default_rng(21)
    v1 = rng.standard_normal((xyz.shape[0], 6))
    dn = od.gatherdofnums(conn)
    for mt in (1, 3):
        Ge = obeam.beam_gyroscopic_elmats(xyz, conn, u1, R1, v1, sec, RHOB, mt)
        femm._sync_mesh(geom0)
        femm.ctx.set_state(u1, R1)
        femm.ctx.set_velocity(v1)
        got = femm.ctx.element_matrices(2, 3, femm._params(mt))
        # Ge = Omega~ M - M Omega~ is a commutator: the two products nearly cancel, so round-off is
        # relative to |Omega~ M| (>> |Ge|); measure the error on that scale
        Me = obeam.beam_mass_elmats(xyz, conn, u1, R1, sec, RHOB, mt)
        scale = [np.linalg.norm(Me[e]) * np.linalg.norm(v1) for e in range(len(conn))]
        assert max(np.linalg.norm(got[e] - Ge[e]) / scale[e] for e in range(len(conn))) < TOL
        assert max(relfro(got[e], Ge[e]) for e in range(len(conn))) < 1e-10
        G = f.gyroscopic(femm, f.SysmatAssemblerSparse(), geom0, uf, Rf, f.NodalField(v1), dchi, mass_type=mt)
        cp, rv, nz = fx.assemble_matrix("sparse", Ge, dn, od.nalldofs)
        assert np.array_equal(G.colptr, cp) and np.array_equal(G.rowval, rv)
        assert relfro(G.nzval, nz) < 1e-10
    q = np.array([0.3, -1.2, 2.5])
    ev = obeam.beam_distribloads_elvecs(xyz, conn, u1, R1, sec, q)
    F = f.distribloads_global(femm, f.SysvecAssemblerFBlock(), geom0, uf, Rf, dchi, q)
    assert relfro(F, fx.assemble_vector(ev, dn, od.nalldofs, od.nfreedofs)) < TOL
    qe = rng.standard_normal((len(conn), 3))
    ev = obeam.beam_distribloads_elvecs(xyz, conn, u1, R1, sec, qe)
    F = f.distribloads_global(femm, geom0, uf, Rf, dchi, qe)
    assert relfro(F, fx.assemble_vector(ev, dn, od.nalldofs)) < TOL


def test_update_rotation_field(fs):
    f, femm, xyz, conn, u1, R1, sec, od, dchi = _beam(fs)
    femm._sync_mesh(f.NodalField(xyz))
    rng = np.random.default_rng(9)
    dchi.values[:] = rng.uniform(-1, 1, dchi.values.shape) * 0.3
    dchi.values[3, 3:] = 0.0  # zero rotation vector branch
    Rf = f.NodalField(R1)
    f.update_rotation_field(femm, Rf, dchi)
    assert np.abs(Rf.values - obeam.update_rotation_field(R1, dchi.values)).max() < 1e-14


# ---------------------------------------------------------------------------------------
# fetch paths
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("asm", ["ffblock", "sparse", "csrsymm"])
def test_fetch_narrow_rows_match_wide(fs, asm, monkeypatch):
    """fsgpu_fetch_matrix: large results move the row indices as int32 and widen them with host threads
    (multi-chunk ring, values on a second stream); the arrays must equal the device-widened path bit for bit."""
    xyz, conn = meshes.shell_mesh("q4", n=170)  # nnz ~ 9.3 M: more than one ring chunk (8 Mi entries)
    od = meshes.clamp_edge_dofs(xyz)
    f = fs.femm
    femm = _make_femm(fs, "q4", conn)
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6)))
    dchi.is_fixed[:] = od.is_fixed
    dchi.numberdofs()
    f.associategeometry(femm, geom0)
    u0, R0 = f.NodalField(np.zeros((xyz.shape[0], 3))), f.initial_Rfield(xyz.shape[0])
    monkeypatch.setenv("FSGPU_FETCH_NARROW_MIN", "-1")
    Kw = f.stiffness(femm, _assembler(fs, asm), geom0, u0, R0, dchi)
    for nth, mode, chunk in (("3", "entries", None), ("16", "entries", None), ("16", "runs", None), ("5", "runs", "100003")):
        monkeypatch.setenv("FSGPU_FETCH_NARROW_MIN", "1")
        monkeypatch.setenv("FSGPU_HOST_THREADS", nth)
        monkeypatch.setenv("FSGPU_FETCH_MODE", mode)  # int32 entries | run-length (position, first row) pairs
        if chunk:
            monkeypatch.setenv("FSGPU_FETCH_CHUNK_UNITS", chunk)  # many ring chunks, runs straddling chunk ends
        Kn = femm.ctx.fetch_matrix()
        assert np.array_equal(Kn.colptr, Kw.colptr)
        assert np.array_equal(Kn.rowval, Kw.rowval), (nth, mode, chunk)
        assert np.array_equal(Kn.nzval, Kw.nzval)
    assert Kw.rowval.size > (1 << 23), "mesh too small to exercise a second ring chunk"


# ---------------------------------------------------------------------------------------
# COO -> CSC, error paths
# ---------------------------------------------------------------------------------------
def test_coo_to_csc(fs):
    rng = np.random.default_rng(11)
    m, n, nt = 57, 43, 5000
    I = rng.integers(1, m + 1, nt)
    J = rng.integers(1, n + 1, nt)
    V = rng.standard_normal(nt)
    V[::7] = 0.0  # explicit zeros are kept
    ctx = fs.Context()
    S = ctx.coo_to_csc(I, J, V, m, n)
    cp, rv, nz = fx.sparse_csc(I, J, V, m, n)
    assert np.array_equal(S.colptr, cp) and np.array_equal(S.rowval, rv)
    assert np.abs(S.nzval - nz).max() < 1e-13
    E = ctx.coo_to_csc([], [], [], 5, 4)
    assert np.array_equal(E.colptr, np.ones(5, dtype=np.int64)) and E.rowval.size == 0


def test_reference_assembler_equivalence_vector(fs):
    """test/test_utilities.jl:12-47 restated: two dense blocks into a 7x7 through COO->CSC."""
    m1 = np.array([[0.24406, 0.599773, 0.833404, 0.0420141], [0.786024, 0.00206713, 0.995379, 0.780298], [0.845816, 0.198459, 0.355149, 0.224996]])
    m2 = np.array([[0.146618, 0.53471, 0.614342, 0.737833], [0.479719, 0.41354, 0.00760941, 0.836455], [0.254868, 0.476189, 0.460794, 0.00919633], [0.159064, 0.261821, 0.317078, 0.77646], [0.643538, 0.429817, 0.59788, 0.958909]])
    I, J, V = [], [], []
    for mat, d in ((m1.T @ m1, [5, 2, 1, 4]), (m2.T @ m2, [2, 3, 1, 5])):
        for j in range(4):
            for i in range(4):
                I.append(d[i]); J.append(d[j]); V.append(mat[i, j])
    S = fs.Context().coo_to_csc(I, J, V, 7, 7).to_scipy().toarray()
    ref = np.zeros((7, 7))
    for i, j, v in zip(I, J, V):
        ref[i - 1, j - 1] += v
    assert relfro(S, ref) < 1e-14


def test_error_paths(fs):
    f = fs.femm
    xyz, conn = meshes.shell_mesh("t3", n=3)
    femm = _make_femm(fs, "t3", conn)
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6))).numberdofs()
    u0, R0 = f.NodalField(np.zeros((xyz.shape[0], 3))), f.initial_Rfield(xyz.shape[0])
    with pytest.raises(fs.FsgpuError) as ei:  # @assert self._associatedgeometry
        f.stiffness(femm, geom0, u0, R0, dchi)
    assert ei.value.code == 3
    ctx = fs.Context()
    ctx.set_mesh(conn, xyz)
    bad = dchi.dofnums.copy()
    bad[0, 0] = dchi.dofnums.size + 1
    with pytest.raises(fs.FsgpuError) as ei:  # FinEtools assemble!: dof > size
        ctx.set_dofnums(bad, dchi.dofnums.size)
    assert ei.value.code == 4
    badconn = conn.copy()
    badconn[0, 0] = xyz.shape[0] + 5
    with pytest.raises(fs.FsgpuError):
        ctx.set_mesh(badconn, xyz)
    # singular metric: all four nodes coincident -> det(J'J) is an exact zero
    xq = np.zeros((4, 3))
    cq = np.array([[1, 2, 3, 4]])
    fq = _make_femm(fs, "q4", cq)
    gq = f.NodalField(xq)
    fq._sync_mesh(gq)
    fq._normals, fq._normal_valid = np.tile([0, 0, 1.0], (4, 1)), np.ones(4, bool)
    fq.ctx.set_normals(fq._normals, fq._normal_valid)
    fq._associatedgeometry = True
    dq = f.NodalField(np.zeros((4, 6))).numberdofs()
    with pytest.raises(fs.FsgpuError) as ei:
        f.stiffness(fq, gq, u0, R0, dq)
    assert ei.value.code == 5


def test_empty_mesh(fs):
    ctx = fs.Context()
    ctx.set_mesh(np.zeros((0, 3), dtype=np.int64), np.zeros((4, 3)))
    ctx.set_dofnums(np.arange(1, 25).reshape(4, 6), 24)
    nr, nc, nnz = ctx.symbolic(0)
    assert (nr, nc, nnz) == (24, 24, 0)


# ---------------------------------------------------------------------------------------
# explicit central differences
# ---------------------------------------------------------------------------------------
def test_explicit_loop(fs):
    import scipy.sparse as sp

    f = fs.femm
    xy, conn = fx.t3block(1.0, 0.6, 12, 8)
    xyz = fx.xyz3(xy)
    xyz[:, 2] = 0.05 * np.sin(3 * xyz[:, 0])
    od = meshes.clamp_edge_dofs(xyz, n_extra_fixed=0)
    femm = _make_femm(fs, "t3", conn)
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6)))
    dchi.is_fixed[:] = od.is_fixed
    dchi.numberdofs()
    f.associategeometry(femm, geom0)
    u0, R0 = f.NodalField(np.zeros((xyz.shape[0], 3))), f.initial_Rfield(xyz.shape[0])
    K = f.stiffness(femm, f.SysmatAssemblerFFBlock(), geom0, u0, R0, dchi)
    femm.ctx.shell_mass_diag(femm._params(), 3, nfree_only=True)
    nf = od.nfreedofs
    Md = femm.ctx.fetch_vector(nf)
    # oracle K, M
    normals, valid = _oracle_normals("t3", xyz, conn)
    dn = od.gatherdofnums(conn)
    cp, rv, nz = fx.assemble_matrix("ffblock", _oracle_K("t3", False, xyz, conn, normals, valid), dn, od.nalldofs, nf)
    Ko = fx.csc_to_scipy(cp, rv, nz, nf, nf).tocsr()
    cpm, rvm, nzm = fx.assemble_matrix("ffblock_diag", _oracle_M("t3", False, xyz, conn), dn, od.nalldofs, nf)
    Mo = np.zeros(nf)
    Mo[rvm - 1] = nzm
    assert relfro(Md, Mo) < TOL
    ex = fs.Explicit(femm.ctx, c_scale=2 * 0.02 * 2 * np.pi * 1000, dt=0.0)  # dt set below
    lam = ex.omega_max_sq(40)
    lam_o = oexp.pwr_largest(Ko, Mo, 200)
    assert abs(lam - lam_o) / lam_o < 0.05
    ex.close()
    dt = 0.9 * 2 / np.sqrt(lam_o)
    cs = 2 * 0.02 * 2 * np.pi * 1000
    ex = fs.Explicit(femm.ctx, c_scale=cs, dt=dt)
    rng = np.random.default_rng(4)
    F0 = rng.standard_normal(nf) * 10.0
    nsteps = 100
    fsc = np.sin(np.arange(1, nsteps + 1) * dt * 2 * np.pi * 2000.0)
    ex.set_load(F0)
    ex.start(0.0)
    ex.step(nsteps, fsc)
    U, V, A = ex.get_state()
    tt = lambda t: F0 * np.sin(t * 2 * np.pi * 2000.0)
    Uo, Vo, Ao = oexp.cd_loop(Mo, Ko, cs, np.zeros(nf), np.zeros(nf), nsteps, dt, tt)
    assert relfro(U, Uo) < 1e-9 and relfro(V, Vo) < 1e-9
    # spmv and kinetic energy
    x = rng.standard_normal(nf)
    assert relfro(ex.spmv(x), Ko @ x) < TOL
    assert abs(ex.kinetic_energy() - 0.5 * np.dot(Vo * Mo, Vo)) <= 1e-9 * abs(0.5 * np.dot(Vo * Mo, Vo))
    # host-CSR construction path gives the same result
    Kcsr = Ko
    ex2 = fs.Explicit(fs.Context(), K=(Kcsr.indptr.astype(np.int64) + 1, Kcsr.indices.astype(np.int64) + 1, Kcsr.data), mdiag=Mo, c_scale=cs, dt=dt)
    ex2.set_load(F0)
    ex2.start(0.0)
    ex2.step(nsteps, fsc)
    assert relfro(ex2.get_state()[0], Uo) < 1e-9
    # the fused step precomputes the next step's displacements: the state must not depend on how the steps
    # are grouped into calls, nor on mixing the fused step with the two-halves (multi-rank) form
    ex3 = fs.Explicit(femm.ctx, c_scale=cs, dt=dt)
    ex3.set_load(F0)
    ex3.start(0.0)
    ex3.step(37, fsc[:37])
    assert relfro(ex3.get_state()[0], oexp.cd_loop(Mo, Ko, cs, np.zeros(nf), np.zeros(nf), 37, dt, tt)[0]) < 1e-9
    for k in range(37, 40):
        ex3.step_begin()
        ex3.step_end(fsc[k])
    ex3.step(1, fsc[40:41])
    ex3.step(nsteps - 41, fsc[41:])
    U3, V3, _ = ex3.get_state()
    assert relfro(U3, Uo) < 1e-9 and relfro(V3, Vo) < 1e-9


# ---------------------------------------------------------------------------------------
# BASELINE configs[0]: modal check -- frequencies from GPU-assembled K, M vs oracle-assembled (1e-9)
# and vs the reference's FV12 goldens (test/test_shell_dynamics.jl:113-133)
# ---------------------------------------------------------------------------------------
def test_modal_check_fv12(fs):
    import scipy.linalg as sla

    f = fs.femm
    E, nu, rho, th, L, n = 200e3 * 1e6, 0.3, 8000.0, 0.05, 10.0, 8
    xy, conn = fx.t3block(L, L, n, n)
    xyz = fx.xyz3(xy - L / 2)
    femm = f.FEMMShellT3FF(f.IntegDomain(conn, None, th), f.MatDeforElastIso(E, nu, rho), stab_alpha=0.2)
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6))).numberdofs()
    f.associategeometry(femm, geom0)
    u0, R0 = f.NodalField(np.zeros((xyz.shape[0], 3))), f.initial_Rfield(xyz.shape[0])
    K = f.stiffness(femm, geom0, u0, R0, dchi).to_scipy().toarray()  # default SysmatAssemblerSparseSymm
    M = f.mass(femm, geom0, dchi).to_scipy().toarray()
    sh = (0.5 * 2 * np.pi) ** 2
    fs_g = np.real(np.sqrt((sla.eigh(K + sh * M, M, eigvals_only=True)[:14] - sh).astype(complex))) / (2 * np.pi)
    Dps, Dt = osh.shell_material_stiffness(fx.moduli_iso(E, nu))
    nrm, val = osh.t3ff_associategeometry(xyz, conn)
    od = fx.DofField(xyz.shape[0]).numberdofs()
    dn, na = od.gatherdofnums(conn), od.nalldofs
    Ko = fx.csc_to_scipy(*fx.assemble_matrix("symm", osh.t3ff_stiffness_elmats(xyz, conn, nrm, val, Dps, Dt, th, stab_fun=osh.stab_lyly(0.2)), dn, na), na, na).toarray()
    Mo = fx.csc_to_scipy(*fx.assemble_matrix("symm", osh.t3ff_mass_elmats(xyz, conn, rho, th), dn, na), na, na).toarray()
    fs_o = np.real(np.sqrt((sla.eigh(Ko + sh * Mo, Mo, eigvals_only=True)[:14] - sh).astype(complex))) / (2 * np.pi)
    assert np.max(np.abs(fs_g[6:] - fs_o[6:]) / fs_o[6:]) < 1e-9
    ref = [1.572130183778014, 2.2424585076387427, 2.8079394352847316, 3.883763676656034, 4.039123204140305, 6.787320617260535, 6.920636670319986, 7.127888889722697]
    assert np.max(np.abs(fs_g[6:] - ref) / np.array(ref)) < 1e-6  # the reference test's own tolerance


def test_full_size_properties_t3(fs):
    """Size-independent checks on a larger mesh (the oracle is too slow there): symmetry of the
    assembled K, rigid-body translations in the null space, lumped mass sums to rho*t*area."""
    f = fs.femm
    from fsb200 import workloads as wl

    w = wl.c4_t3ff_panel(400, 200)
    femm = f.FEMMShellT3FF(f.IntegDomain(w["conn"], None, w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]))
    geom0 = f.NodalField(w["xyz"])
    dchi = f.NodalField(np.zeros((w["xyz"].shape[0], 6))).numberdofs()  # free-free
    f.associategeometry(femm, geom0)
    K = f.stiffness(femm, f.SysmatAssemblerSparse(), geom0, None, None, dchi).to_scipy()
    scale = abs(K).max()
    assert abs(K - K.T).max() < 1e-9 * scale
    for d in range(3):
        u = np.zeros(dchi.dofnums.shape)
        u[:, d] = 1.0
        v = np.zeros(dchi.dofnums.size)
        v[dchi.dofnums.ravel() - 1] = u.ravel()
        assert np.abs(K @ v).max() < 1e-7 * scale
    femm.ctx.shell_mass_diag(femm._params(), 3, nfree_only=False)
    Md = femm.ctx.fetch_vector(dchi.dofnums.size)
    c = w["conn"] - 1
    X = np.asarray(w["xyz"])
    area = 0.5 * np.linalg.norm(np.cross(X[c[:, 1]] - X[c[:, 0]], X[c[:, 2]] - X[c[:, 0]]), axis=1).sum()
    tm = Md[dchi.dofnums[:, 0] - 1].sum()
    assert abs(tm - w["rho"] * w["thickness"] * area) < 1e-10 * tm


def test_c1_double_cell_box_modal_check(fs):
    """BASELINE configs[0] (the README example, README.md:90-102): double-cell box, T3refine x4,
    17 920 T3 / 8 991 nodes, branched shell with invalid nodal normals at the T-junctions.
    The 4 lowest frequencies from GPU-assembled K, M equal those from oracle-assembled K, M to 1e-9
    (same eigensolver on both), and the README log to ~1e-6 (its digits predate the current
    reference revision)."""
    import scipy.sparse.linalg as spla
    from fsb200 import workloads as wl

    f = fs.femm
    w = wl.c1_double_cell_box(4)
    xyz, conn = np.ascontiguousarray(w["xyz"]), w["conn"]
    assert conn.shape[0] == 17920 and xyz.shape[0] == 8991
    femm = f.FEMMShellT3FF(f.IntegDomain(conn, None, w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]))
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6)))
    dchi.is_fixed[:] = w["fixed"]
    dchi.numberdofs()
    f.associategeometry(femm, geom0)
    nrm, val = osh.t3ff_associategeometry(xyz, conn)
    assert np.array_equal(femm._normal_valid, val) and (~val).sum() > 100
    assert np.abs(femm._normals - nrm).max() < 1e-12
    nf = f.nfreedofs(dchi)
    K = f.stiffness(femm, f.SysmatAssemblerFFBlock(), geom0, None, None, dchi)
    M = f.mass(femm, f.SysmatAssemblerFFBlock(f.SysmatAssemblerSparseDiag()), geom0, dchi)
    Dps, Dt = osh.shell_material_stiffness(fx.moduli_iso(w["E"], w["nu"]))
    od = fx.DofField(xyz.shape[0])
    od.is_fixed[:] = w["fixed"]
    od.numberdofs()
    dn, na = od.gatherdofnums(conn), od.nalldofs
    rK = fx.assemble_matrix("ffblock", osh.t3ff_stiffness_elmats(xyz, conn, nrm, val, Dps, Dt, w["thickness"]), dn, na, nf)
    rM = fx.assemble_matrix("ffblock_diag", osh.t3ff_mass_elmats(xyz, conn, w["rho"], w["thickness"]), dn, na, nf)
    _check_matrix(K, rK, nf)
    _check_matrix(M, rM, nf)
    fg = np.sort(np.sqrt(spla.eigsh(K.to_scipy(), 4, M.to_scipy(), sigma=0.0, which="LM", return_eigenvectors=False))) / (2 * np.pi)
    fo = np.sort(np.sqrt(spla.eigsh(fx.csc_to_scipy(*rK, nf, nf), 4, fx.csc_to_scipy(*rM, nf, nf), sigma=0.0, which="LM", return_eigenvectors=False))) / (2 * np.pi)
    assert np.max(np.abs(fg - fo) / fo) < 1e-9
    readme = np.array([20.301524870325565, 25.533290848730623, 28.914284995255777, 30.620822302876647])
    assert np.max(np.abs(fg - readme) / readme) < 1e-5


def test_full_size_c2_properties(fs):
    """BASELINE configs[1] at FULL size (1M Q4RS elements): size-independent properties -- the
    stored pattern equals the 9-node stencil count, K is symmetric to round-off, rigid translations
    are in the null space, the numeric phase is repeatable."""
    from fsb200 import workloads as wl

    f = fs.femm
    w = wl.c2_q4rs_plate(1000)
    femm = f.FEMMShellQ4RS(f.IntegDomain(w["conn"], f.GaussRule2x2(), w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]))
    geom0 = f.NodalField.__new__(f.NodalField)
    geom0.values = w["xyz"]
    dchi = f.NodalField.__new__(f.NodalField)
    nn = w["xyz"].shape[0]
    dchi.values, dchi.dofnums, dchi._nfree = None, np.asfortranarray(np.arange(1, 6 * nn + 1, dtype=np.int64).reshape(nn, 6)), 6 * nn
    f.associategeometry(femm, geom0)
    femm._startassembly(f.SysmatAssemblerSparse(), dchi)
    femm._sync_stab()
    nr, nc, nnz = femm.ctx.result_size() if False else (6 * nn, 6 * nn, None)
    femm.ctx.shell_op("q4rs_stiffness", femm._params())
    m, n, nnz = femm.ctx.result_size()
    # structured n x n quads: node-pair count = sum over nodes of the stencil size = (3n+1)^2 ... counted directly
    N = 1000
    pairs = (3 * (N + 1) - 2) ** 2  # sum of (1-D stencil sizes)^2: 1-D sizes are 2 at the ends, 3 inside
    assert nnz == 36 * pairs
    import ctypes as C

    import torch

    cp, rv, nz = C.c_void_p(), C.c_void_p(), C.c_void_p()
    fs._lib.check(fs._lib.lib.fsgpu_result_device(femm.ctx._h, C.byref(cp), C.byref(rv), C.byref(nz)))
    from fsb200.partition import DevicePointer

    vals = torch.as_tensor(DevicePointer(nz.value, nnz), device="cuda")
    colptr = torch.as_tensor(DevicePointer(cp.value, n + 1, "<i4"), device="cuda").long()
    rowval = torch.as_tensor(DevicePointer(rv.value, nnz, "<i4"), device="cuda").long()
    A = torch.sparse_csr_tensor(colptr, rowval, vals, size=(n, m))  # CSR view of the CSC arrays = K^T
    scale = float(vals.abs().max())
    for d in range(3):
        v = torch.zeros(n, dtype=torch.float64, device="cuda")
        v[d::6] = 1.0
        assert float((A @ v).abs().max()) < 1e-7 * scale  # K^T t = 0 for a rigid translation t
    x = torch.randn(n, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    y = torch.randn(n, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    # symmetry: y' K^T x == x' K^T y
    a, b = float(y @ (A @ x)), float(x @ (A @ y))
    assert abs(a - b) < 1e-10 * max(abs(a), abs(b), scale)
    s1 = float(vals.sum())
    femm.ctx.shell_op("q4rs_stiffness", femm._params())
    vals2 = torch.as_tensor(DevicePointer(nz.value, nnz), device="cuda")
    assert abs(float(vals2.sum()) - s1) <= 1e-9 * abs(s1) + 1e-6 * scale


# ---------------------------------------------------------------------------------------
# committed fixtures (tests/golden/oracle_fixtures.npz): the GPU path against frozen numbers
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["t3", "q4"])
def test_against_committed_fixtures(fs, kind):
    fxt = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_fixtures.npz"))
    f = fs.femm
    xyz, conn = meshes.shell_mesh(kind, n=3)
    femm = _make_femm(fs, kind, conn)
    geom0 = f.NodalField(xyz)
    f.associategeometry(femm, geom0)
    assert np.abs(femm._normals - fxt[f"{kind}_normals"]).max() < 1e-14
    assert np.array_equal(femm._normal_valid, fxt[f"{kind}_valid"])
    femm._sync_stab()
    Kg = femm.ctx.element_matrices(femm._kind(), 0, femm._params())
    assert relfro(Kg, fxt[f"{kind}_K"]) < TOL
    od = meshes.clamp_edge_dofs(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6)))
    dchi.is_fixed[:] = od.is_fixed
    dchi.numberdofs()
    K = f.stiffness(femm, f.SysmatAssemblerFFBlock(), geom0, None, None, dchi)
    assert np.array_equal(K.colptr, fxt[f"{kind}_colptr"]) and np.array_equal(K.rowval, fxt[f"{kind}_rowval"])
    assert relfro(K.nzval, fxt[f"{kind}_nzval"]) < TOL
    u = np.random.default_rng(5).standard_normal((xyz.shape[0], 6)) * 1e-3
    th = np.deg2rad(20.0)
    cs = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    got = f.inspectintegpoints(femm, geom0, f.NodalField(u), None, "moment", outputcsys=cs)
    assert relfro(got, fxt[f"{kind}_moment"].reshape(got.shape)) < 1e-11
