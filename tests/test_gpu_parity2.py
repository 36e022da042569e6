"""GPU parity tests, part 2 (round 2): product paths the first suite did not reach -- per-element / per-point
thickness, the sampled stabilisation factor, T3FF's accumulating associategeometry!, Q4RSComp's per-point layup csys
with the shape-function "location" quirk, permuted numbering for Q4RS and the beam, load-factor tables of the
explicit loop, one triangle of the result, and 100k-element slices of the actual bench meshes (C2, C3, C4, C5)
against the C port / the NumPy oracle.  Tolerances as in test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import beam as obeam
from oracle import explicit as oexp
from oracle import fe_external as fx
from oracle import layup as oly
from oracle import shells as osh
from tests import meshes
from tests.test_gpu_parity import E_, NU_, RHO_, T_, TOL, _check_matrix, _fs_layup, _iso, _layup, _make_femm, _oracle_K, _oracle_normals, relfro

pytestmark = pytest.mark.gpu

P = lambda a: a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def fs():
    import fsb200

    return fsb200


@pytest.fixture(scope="module")
def refport():
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "cport")
    subprocess.run(["make", "-C", d, "-s", "-B"], check=True)
    lib = C.CDLL(os.path.join(d, "librefport.so"))
    lib.ref_coo_to_csc.restype = C.c_int64
    return lib


def _fields(f, xyz, od=None, perm=None):
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6)))
    if od is not None:
        dchi.is_fixed[:] = od.is_fixed
    dchi.numberdofs(perm) if perm is not None else dchi.numberdofs()
    u0, R0 = f.NodalField(np.zeros((xyz.shape[0], 3))), f.initial_Rfield(xyz.shape[0])
    return geom0, dchi, u0, R0


# ---------------------------------------------------------------------------------------
# thickness arrays (src/FEMMShellT3FFModule.jl:676 `t = self.integdomain.otherdimension(centroid, fes.conn[i], ...)`;
# src/FEMMShellQ4RSModule.jl:921 per integration point)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,per_point", [("t3", False), ("q4", False), ("q4", True)])
def test_thickness_arrays(fs, kind, per_point):
    f = fs.femm
    xyz, conn = meshes.shell_mesh(kind, n=7)
    ne = conn.shape[0]
    rng = np.random.default_rng(11)
    t = T_ * rng.uniform(0.5, 2.0, (ne, 4) if per_point else ne)
    normals, valid = _oracle_normals(kind, xyz, conn)
    Dps, Dt = _iso()
    idom = f.IntegDomain(conn, None if kind == "t3" else f.GaussRule2x2(), t)
    femm = (f.FEMMShellT3FF if kind == "t3" else f.FEMMShellQ4RS)(idom, f.MatDeforElastIso(E_, NU_, RHO_))
    geom0, dchi, u0, R0 = _fields(f, xyz, meshes.clamp_edge_dofs(xyz))
    f.associategeometry(femm, geom0)
    femm._sync_stab()
    if kind == "t3":
        Ko, Mo = osh.t3ff_stiffness_elmats(xyz, conn, normals, valid, Dps, Dt, t), osh.t3ff_mass_elmats(xyz, conn, RHO_, t)
    else:
        Ko, Mo = osh.q4rs_stiffness_elmats(xyz, conn, normals, valid, Dps, Dt, t), osh.q4rs_mass_elmats(xyz, conn, RHO_, t)
    Kg = femm.ctx.element_matrices(femm._kind(), 0, femm._params())
    assert max(relfro(Kg[e], Ko[e]) for e in range(ne)) < TOL
    Mg = femm.ctx.element_matrices(femm._kind(), 1, femm._params())
    assert relfro(Mg, Mo) < TOL
    od = meshes.clamp_edge_dofs(xyz)
    K = f.stiffness(femm, f.SysmatAssemblerFFBlock(), geom0, u0, R0, dchi)
    _check_matrix(K, fx.assemble_matrix("ffblock", Ko, od.gatherdofnums(conn), od.nalldofs, od.nfreedofs), od.nfreedofs)


# ---------------------------------------------------------------------------------------
# stab_fun outside the t^2/(t^2 + alpha h^2) family: sampled per element into fsgpu_set_stab_factor
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,comp", [("t3", False), ("q4", False), ("t3", True), ("q4", True)])
def test_sampled_stab_factor(fs, kind, comp):
    f = fs.femm
    xyz, conn = meshes.shell_mesh(kind, n=6)
    lay, cs = _layup()
    stab = lambda t, h: t**2 / (t**2 + 0.3 * h**2 + 0.05 * t * h)
    femm = _make_femm(fs, kind, conn, comp, cs)
    femm.stab_fun = stab
    geom0, dchi, u0, R0 = _fields(f, xyz)
    f.associategeometry(femm, geom0)
    normals, valid = _oracle_normals(kind, xyz, conn, fixed=cs[:, 2] if comp else None)
    Dps, Dt = _iso()
    if comp:
        A, B, D = lay.laminate_stiffnesses()
        H = lay.laminate_transverse_stiffness()
        fn = osh.t3ffcomp_stiffness_elmats if kind == "t3" else osh.q4rscomp_stiffness_elmats
        Ko = fn(xyz, conn, normals, valid, A, B, D, H, lay.thickness, cs, stab_fun=stab)
    else:
        fn = osh.t3ff_stiffness_elmats if kind == "t3" else osh.q4rs_stiffness_elmats
        Ko = fn(xyz, conn, normals, valid, Dps, Dt, T_, stab_fun=stab)
    K = f.stiffness(femm, f.SysmatAssemblerSparse(), geom0, u0, R0, dchi)
    od = fx.DofField(xyz.shape[0]).numberdofs()
    _check_matrix(K, fx.assemble_matrix("sparse", Ko, od.gatherdofnums(conn), od.nalldofs), od.nalldofs)
    Kg = femm.ctx.element_matrices(femm._kind(), 0, femm._params())
    assert max(relfro(Kg[e], Ko[e]) for e in range(conn.shape[0])) < TOL


# ---------------------------------------------------------------------------------------
# T3FF associategeometry! never resets (src/FEMMShellT3FFModule.jl:575-587): a second call
# accumulates onto the unit normals of the first and can only turn more nodes invalid
# ---------------------------------------------------------------------------------------
def test_t3ff_associategeometry_twice(fs):
    f = fs.femm
    xyz, conn = meshes.shell_mesh("t3", n=9)
    femm = _make_femm(fs, "t3", conn)
    geom0 = f.NodalField(xyz)
    f.associategeometry(femm, geom0)
    n1, v1 = femm._normals.copy(), femm._normal_valid.copy()
    o1, ov1 = osh.t3ff_associategeometry(xyz, conn)
    assert np.abs(n1 - o1).max() < 1e-13 and np.array_equal(v1, ov1)
    f.associategeometry(femm, geom0)
    o2, ov2 = osh.t3ff_associategeometry(xyz, conn, normals0=o1)
    assert np.abs(femm._normals - o2).max() < 1e-13
    assert np.array_equal(femm._normal_valid, ov1 & ov2)
    # (unit normal + the same element normals again: the direction cannot change, "harmless after normalisation")
    assert np.abs(femm._normals - n1).max() < 1e-13 and not ov1.all()
    # the Q4RS FEMM resets (src/FEMMShellQ4RSModule.jl:483-484): a second call reproduces the first
    xyz4, conn4 = meshes.shell_mesh("q4", n=9)
    fq = _make_femm(fs, "q4", conn4)
    g4 = f.NodalField(xyz4)
    f.associategeometry(fq, g4)
    a = fq._normals.copy()
    f.associategeometry(fq, g4)
    assert np.abs(a - fq._normals).max() < 1e-14  # (the accumulation uses atomics: not bitwise)


# ---------------------------------------------------------------------------------------
# Q4RSComp: layup csys per element AND integration point, `updatecsmat!(layup.csys, Ns[j], J, -1, 0)`
# (src/FEMMShellQ4RSCompModule.jl:929): the location handed to the csys is the vector of shape-function values
# ---------------------------------------------------------------------------------------
def _csys_from_tangents(XYZ, tangents, feid, qpid):
    """A csys callback that uses BOTH arguments the reference passes: e1 along the first tangent rotated about the
    surface normal by an angle that depends on the 'location' (for Q4RSComp: the shape-function values)."""
    t1, t2 = tangents[:, :, 0], tangents[:, :, 1]
    e3 = np.cross(t1, t2)
    e3 /= np.linalg.norm(e3, axis=1, keepdims=True)
    a = t1 / np.linalg.norm(t1, axis=1, keepdims=True)
    b = np.cross(e3, a)
    th = 0.3 + 1.1 * XYZ[:, 0] - 0.7 * XYZ[:, 1] + 0.4 * XYZ[:, 2]
    e1 = np.cos(th)[:, None] * a + np.sin(th)[:, None] * b
    e2 = np.cross(e3, e1)
    return np.stack([e1, e2, e3], axis=-1)


@pytest.mark.parametrize("rule", ["gauss2x2", "simpson13"])
def test_q4rscomp_per_point_csys_callback(fs, rule):
    f = fs.femm
    xyz, conn = meshes.shell_mesh("q4", n=6, crease=0.0)
    lay, _ = _layup()
    orule = fx.gauss_rule_2x2() if rule == "gauss2x2" else fx.simpson13_rule_2d()
    grule = f.GaussRule2x2() if rule == "gauss2x2" else f.Simpson13Rule2()
    femm = f.FEMMShellQ4RSComp(f.IntegDomain(conn, grule, T_), _fs_layup(fs, _csys_from_tangents))
    geom0, dchi, u0, R0 = _fields(f, xyz)
    f.associategeometry(femm, geom0)
    # independent evaluation of what the reference does
    X = xyz[conn - 1]
    ne = conn.shape[0]
    pcn, _ = fx.nodal_rule_q4()
    dirs = np.zeros((ne, 4, 3))
    for j in range(4):
        _, dNp = fx.q4_shape(*pcn[j])
        J = np.einsum("eai,ak->eik", X, dNp)
        dirs[:, j] = _csys_from_tangents(X[:, j], J, None, None)[:, :, 2]  # evaluated AT THE NODE (:471)
    no, vo = osh.q4rs_associategeometry(xyz, conn, normal_dir=dirs)
    assert np.abs(femm._normals - no).max() < 1e-13 and np.array_equal(femm._normal_valid, vo)
    pc, w = orule
    lcs = np.zeros((ne, len(w), 3, 3))
    for j in range(len(w)):
        N, dNp = fx.q4_shape(*pc[j])
        J = np.einsum("eai,ak->eik", X, dNp)
        lcs[:, j] = _csys_from_tangents(np.broadcast_to(np.ravel(N), (ne, 4)), J, None, None)  # location = Ns[j]  (:929)
    A, B, D = lay.laminate_stiffnesses()
    H = lay.laminate_transverse_stiffness()
    Ko = osh.q4rscomp_stiffness_elmats(xyz, conn, no, vo, A, B, D, H, lay.thickness, lcs, rule=orule)
    femm._sync_stab()
    Kg = femm.ctx.element_matrices(femm._kind(), 0, femm._params())
    assert max(relfro(Kg[e], Ko[e]) for e in range(ne)) < TOL
    # the location quirk matters: the centroid as location gives a different matrix
    lcs_c = np.repeat(_csys_from_tangents(X.mean(axis=1), np.einsum("eai,ak->eik", X, fx.q4_shape(0.0, 0.0)[1]), None, None)[:, None], len(w), axis=1)
    Kc = osh.q4rscomp_stiffness_elmats(xyz, conn, no, vo, A, B, D, H, lay.thickness, lcs_c, rule=orule)
    assert relfro(Kc, Ko) > 1e-3
    K = f.stiffness(femm, f.SysmatAssemblerSparse(), geom0, u0, R0, dchi)
    od = fx.DofField(xyz.shape[0]).numberdofs()
    _check_matrix(K, fx.assemble_matrix("sparse", Ko, od.gatherdofnums(conn), od.nalldofs), od.nalldofs)


def test_two_layup_groups_with_different_csys(fs):
    """Each layup group uses its own csys and thickness (src/FEMMShellT3FFCompModule.jl:501-509,600-620)."""
    f = fs.femm
    xyz, conn = meshes.shell_mesh("t3", n=6, crease=0.0)
    ne = conn.shape[0]
    lay, cs1 = _layup()
    th = np.deg2rad(-55.0)
    cs2 = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    g1, g2 = np.arange(1, ne // 2 + 1), np.arange(ne // 2 + 1, ne + 1)
    femm = f.FEMMShellT3FFComp(f.IntegDomain(conn, None, T_), [(_fs_layup(fs, cs1), g1), (_fs_layup(fs, cs2), g2)])
    geom0, dchi, u0, R0 = _fields(f, xyz)
    f.associategeometry(femm, geom0)
    lcs = np.zeros((ne, 3, 3))
    lcs[g1 - 1], lcs[g2 - 1] = cs1, cs2
    no, vo = osh.t3ff_associategeometry(xyz, conn, normal_dir=np.repeat(lcs[:, None, :, 2], 3, axis=1))
    assert np.abs(femm._normals - no).max() < 1e-13 and np.array_equal(femm._normal_valid, vo)
    A, B, D = lay.laminate_stiffnesses()
    H = lay.laminate_transverse_stiffness()
    Ko = osh.t3ffcomp_stiffness_elmats(xyz, conn, no, vo, A, B, D, H, lay.thickness, lcs)
    femm._sync_stab()
    Kg = femm.ctx.element_matrices(femm._kind(), 0, femm._params())
    assert max(relfro(Kg[e], Ko[e]) for e in range(ne)) < TOL


# ---------------------------------------------------------------------------------------
# numberdofs!(dchi, perm) for Q4RS and the beam (the first suite permutes only T3FF)
# ---------------------------------------------------------------------------------------
def test_permuted_numbering_q4(fs):
    f = fs.femm
    xyz, conn = meshes.shell_mesh("q4", n=7)
    perm = np.random.default_rng(6).permutation(xyz.shape[0])
    od = meshes.clamp_edge_dofs(xyz)
    od.numberdofs(perm)
    femm = _make_femm(fs, "q4", conn)
    geom0, dchi, u0, R0 = _fields(f, xyz, od, perm)
    assert np.array_equal(dchi.dofnums, od.dofnums)
    f.associategeometry(femm, geom0)
    normals, valid = _oracle_normals("q4", xyz, conn)
    Dps, Dt = _iso()
    Ko = osh.q4rs_stiffness_elmats(xyz, conn, normals, valid, Dps, Dt, T_)
    for asm, a in (("ffblock", f.SysmatAssemblerFFBlock()), ("sparse", f.SysmatAssemblerSparse())):
        K = f.stiffness(femm, a, geom0, u0, R0, dchi)
        n = od.nfreedofs if asm == "ffblock" else od.nalldofs
        _check_matrix(K, fx.assemble_matrix(asm, Ko, od.gatherdofnums(conn), od.nalldofs, od.nfreedofs), n)


def test_permuted_numbering_beam(fs):
    f = fs.femm
    xyz, conn, u1, R1, sec = meshes.beam_lattice()
    EB, NUB = 71240.0, 0.31
    perm = np.random.default_rng(7).permutation(xyz.shape[0])
    od = fx.DofField(xyz.shape[0])
    for c in range(1, 7):
        od.setebc([0], c)
    od.setebc([17], 3)
    od.numberdofs(perm)
    secs = f.FESetL2Beam(sec["A"], sec["I1"], sec["I2"], sec["I3"], sec["J"], sec["A2s"], sec["A3s"], sec["x1x2"])
    femm = f.FEMMCorotBeam(f.IntegDomain(conn), f.MatDeforElastIso(EB, NUB, 5e-9), secs)
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6)))
    dchi.is_fixed[:] = od.is_fixed
    dchi.numberdofs(perm)
    uf, Rf = f.NodalField(u1), f.NodalField(R1)
    dn = od.gatherdofnums(conn)
    K = f.stiffness(femm, f.SysmatAssemblerFFBlock(), geom0, uf, Rf, dchi)
    _check_matrix(K, fx.assemble_matrix("ffblock", obeam.beam_stiffness_elmats(xyz, conn, u1, R1, sec, EB, NUB), dn, od.nalldofs, od.nfreedofs), od.nfreedofs)
    Kg = f.geostiffness(femm, f.SysmatAssemblerSparse(), geom0, uf, Rf, dchi)
    _check_matrix(Kg, fx.assemble_matrix("sparse", obeam.beam_geostiffness_elmats(xyz, conn, u1, R1, sec, EB, NUB), dn, od.nalldofs), od.nalldofs)
    Fr = f.restoringforce(femm, f.SysvecAssemblerFBlock(), geom0, uf, Rf, dchi)
    ev = obeam.beam_restoringforce_elvecs(xyz, conn, u1, R1, sec, EB, NUB)
    assert relfro(Fr, fx.assemble_vector(ev, dn, od.nalldofs, od.nfreedofs)) < TOL


# ---------------------------------------------------------------------------------------
# one triangle of the result (fsgpu_fetch_matrix_uplo)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("asm", ["ffblock", "sparse", "symm"])
@pytest.mark.parametrize("narrow", [False, True])
def test_fetch_triangle(fs, asm, narrow, monkeypatch):
    import scipy.sparse as sp

    if narrow:  # force the pinned-ring / host-expansion path and many pieces on a small matrix
        monkeypatch.setenv("FSGPU_FETCH_NARROW_MIN", "0")
        monkeypatch.setenv("FSGPU_FETCH_CHUNK_UNITS", "64")
    f = fs.femm
    xyz, conn = meshes.shell_mesh("q4", n=8)
    od = meshes.clamp_edge_dofs(xyz)
    femm = _make_femm(fs, "q4", conn)
    geom0, dchi, u0, R0 = _fields(f, xyz, od)
    f.associategeometry(femm, geom0)
    mk = {"ffblock": f.SysmatAssemblerFFBlock, "sparse": f.SysmatAssemblerSparse, "symm": f.SysmatAssemblerSparseSymm}[asm]
    f.stiffness(femm, mk(), geom0, u0, R0, dchi)
    Kf = femm.ctx.fetch_matrix()  # full matrix and both triangles of the SAME device result: bitwise comparable
    Kc = Kf.to_scipy().tocsc()
    cp, rv, nz = Kf.colptr - 1, Kf.rowval - 1, Kf.nzval
    col = np.repeat(np.arange(Kf.n), np.diff(cp))
    for uplo in ("L", "U"):
        keep = rv >= col if uplo == "L" else rv <= col
        T = femm.ctx.fetch_matrix_uplo(uplo)
        assert np.array_equal(T.rowval - 1, rv[keep]), "explicit zeros are kept by the triangle as by the full fetch"
        assert np.array_equal(T.nzval, nz[keep])
        assert np.array_equal(np.diff(T.colptr), np.bincount(col[keep], minlength=Kf.n))
        assert femm.ctx.result_size_uplo(uplo) == int(keep.sum())
    # through the operator API: the assembler carries the triangle request
    a = mk()
    a.uplo = "L"
    T = f.stiffness(femm, a, geom0, u0, R0, dchi)
    ref = sp.tril(Kc, format="csc")
    assert abs(T.to_scipy() - ref).max() < 1e-12 * abs(ref).max()


# ---------------------------------------------------------------------------------------
# T3 emission plan on meshes that are not manifold strips: a fan of 14 triangles around one node (the ten elements of
# the first warp all contribute to one diagonal block: four descriptors of <= 3 contributors), an edge shared by four
# triangles (more than two contributors per edge block), rotated connectivities, a partly filled second warp
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("sheark", [0, 1])
def test_t3_plan_fans_and_nonmanifold_edges(fs, sheark):
    f = fs.femm
    nf = 14
    ang = np.linspace(0.0, 2 * np.pi, nf, endpoint=False)
    ring = np.column_stack([np.cos(ang), np.sin(ang), 0.25 * np.cos(3 * ang)])
    xyz = np.vstack([[0.0, 0.0, 0.4], ring, [[2.0, 0.1, 0.9], [2.0, 0.2, -0.8], [1.9, 1.0, 0.1], [2.1, -1.0, 0.0]]])
    fan = [[1, 2 + k, 2 + (k + 1) % nf] if k % 3 else [2 + (k + 1) % nf, 1, 2 + k] for k in range(nf)]  # rotated connectivities
    a, b = 2, nf + 2  # ring node 1 and the first extra node: an edge with four triangles on it
    tj = [[a, b, nf + 3], [b, a, nf + 4], [a, b, nf + 5], [b, a, 3]]
    conn = np.array(fan[:10] + tj[:2] + fan[10:] + tj[2:], dtype=np.int64)
    assert conn.shape[0] == 18 and conn.max() == xyz.shape[0] and conn.min() == 1
    od = meshes.clamp_edge_dofs(xyz, n_extra_fixed=4)
    femm = _make_femm(fs, "t3", conn)
    femm.transv_shear_formulation = sheark
    geom0, dchi, u0, R0 = _fields(f, xyz, od)
    f.associategeometry(femm, geom0)
    normals, valid = _oracle_normals("t3", xyz, conn)
    Ko = _oracle_K("t3", False, xyz, conn, normals, valid, sheark=sheark)
    for asm, a_ in (("ffblock", f.SysmatAssemblerFFBlock()), ("sparse", f.SysmatAssemblerSparse())):
        K = f.stiffness(femm, a_, geom0, u0, R0, dchi)
        assert femm.ctx.scatter_path == 1  # the run-structured path with the plan-driven emission
        n = od.nfreedofs if asm == "ffblock" else od.nalldofs
        _check_matrix(K, fx.assemble_matrix(asm, Ko, od.gatherdofnums(conn), od.nalldofs, od.nfreedofs), n)


# ---------------------------------------------------------------------------------------
# the reference's own NAFEMS LE5 test (Z-section with creases: invalid nodal normals at half of the nodes), end to end:
# GPU-assembled stiffness, host solve, the reference's golden numbers (test/test_shell_statics.jl:530-531)
# ---------------------------------------------------------------------------------------
def test_le5_z_section_golden_through_the_gpu(fs):
    from tests.test_oracle_goldens import LE5_GOLDEN, le5_problem

    f = fs.femm
    xyz, conn, E, nu, th, d, F = le5_problem()
    femm = f.FEMMShellT3FF(f.IntegDomain(conn, None, th), f.MatDeforElastIso(E, nu, 1.0), stab_alpha=0.2)
    geom0, dchi, u0, R0 = _fields(f, xyz, d)
    f.associategeometry(femm, geom0)
    assert (~np.asarray(femm._normal_valid, dtype=bool)).sum() == 18
    K = f.stiffness(femm, geom0, u0, R0, dchi)  # default assembler: SysmatAssemblerSparseSymm
    fx.solve_blocked(K.to_scipy().tocsc(), F, d)
    uz = d.values[:, 2]
    assert abs(uz.min() - LE5_GOLDEN[0]) < 1e-9 * abs(LE5_GOLDEN[0])
    assert abs(uz.max() - LE5_GOLDEN[1]) < 1e-9 * abs(LE5_GOLDEN[1])


# ---------------------------------------------------------------------------------------
# the reference's resultants test (irregular barrel vault, cylindrical output csys, nodal fields by FinEtools'
# inverse-distance rule; test/test_shell_statics.jl:577-728) through the GPU path: stiffness, host solve, batched
# inspectintegpoints, the reference's 16 (min, max) numbers with the reference's tolerance
# ---------------------------------------------------------------------------------------
def test_barrelvault_resultant_fields_through_the_gpu(fs):
    from tests.test_oracle_goldens import barrelvault_resultants_problem, check_barrelvault_fields

    f = fs.femm
    P = barrelvault_resultants_problem()
    xyz, conn, d = P["xyz"], P["conn"], P["dof"]
    femm = f.FEMMShellT3FF(f.IntegDomain(conn, None, P["th"]), f.MatDeforElastIso(P["E"], P["nu"], 1.0), stab_alpha=0.2)
    femm.drilling_stiffness_scale = 0.1
    geom0, dchi, u0, R0 = _fields(f, xyz, d)
    f.associategeometry(femm, geom0)
    K = f.stiffness(femm, geom0, u0, R0, dchi)
    fx.solve_blocked(K.to_scipy().tocsc(), fx.distribloads_t3(xyz, conn, [-0.625, 0, 0, 0, 0, 0]), d)
    assert np.abs(d.values - P["u"]).max() < 1e-9 * np.abs(P["u"]).max()  # the oracle's solution of the same problem
    names = {osh.BENDING_MOMENT: "moment", osh.MEMBRANE_FORCE: "membrane", osh.TRANSVERSE_SHEAR: "shear"}
    u = f.NodalField(d.values.copy())
    check_barrelvault_fields(P, lambda q: f.inspectintegpoints(femm, geom0, u, None, names[q], outputcsys=P["ocs"])[:, 0, :])
    # the mirror of FinEtools' fieldfromintegpoints / elemfieldfromintegpoints on top of the batched resultants
    from tests.test_oracle_goldens import BARRELVAULT_FIELDS

    for q, gold in BARRELVAULT_FIELDS.items():
        fld = f.fieldfromintegpoints(femm, geom0, u, names[q], list(range(1, len(gold) + 1)), outputcsys=P["ocs"]).values
        for k, (lo, hi) in enumerate(gold):
            assert abs(fld[:, k].min() - lo) <= 0.01 * abs(lo) and abs(fld[:, k].max() - hi) <= 0.01 * abs(hi)
        # the device averaging against the oracle's restatement of the rule on the same device resultants
        res = f.inspectintegpoints(femm, geom0, u, None, names[q], outputcsys=P["ocs"])
        host = fx.field_from_integpoints_invdist(xyz, conn, P["cen"][:, None, :], res)
        assert np.abs(fld - host[:, : len(gold)]).max() <= 1e-12 * np.abs(host).max()
        ef = f.elemfieldfromintegpoints(femm, geom0, u, names[q], 1, outputcsys=P["ocs"])
        assert ef.shape == (conn.shape[0], 1)


@pytest.mark.parametrize("kind", ["t3", "q4"])
def test_nodal_field_device_matches_the_numpy_rule(fs, kind):
    """fsgpu_shell_nodal_field (inverse squared distance from the centroid / the integration points) against the oracle's
    restatement of the rule applied to the device's own resultants."""
    f = fs.femm
    xyz, conn = meshes.shell_mesh(kind, n=7)
    femm = _make_femm(fs, kind, conn)
    geom0 = f.NodalField(xyz)
    f.associategeometry(femm, geom0)
    u = f.NodalField(np.random.default_rng(17).standard_normal((xyz.shape[0], 6)) * 1e-3)
    th = np.deg2rad(25.0)
    ocs = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    X = xyz[conn - 1]
    if kind == "t3":
        loc = X.mean(axis=1)[:, None, :]  # the reference's inspectintegpoints hands the centroid to the inspector
    else:
        pc, _ = fx.gauss_rule_2x2()
        loc = np.stack([np.einsum("a,eai->ei", fx.q4_shape(*p)[0], X) for p in pc], axis=1)
    for name in ("moment", "shear", "membrane"):
        fld = f.fieldfromintegpoints(femm, geom0, u, name, [1, 2, 3], outputcsys=ocs).values
        res = f.inspectintegpoints(femm, geom0, u, None, name, outputcsys=ocs)
        host = fx.field_from_integpoints_invdist(xyz, conn, loc, res)
        assert fld.shape == host.shape and np.abs(fld - host).max() <= 1e-12 * np.abs(host).max(), name


# ---------------------------------------------------------------------------------------
# pageable destinations (a Julia `Vector`): values and colptr travel through the pinned staging ring
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("narrow", [False, True])
def test_fetch_into_pageable_arrays_through_the_ring(fs, narrow, monkeypatch):
    import torch

    monkeypatch.setenv("FSGPU_VALUE_RING_WORDS", "4096")  # many pieces, all three staging buffers reused
    monkeypatch.setenv("FSGPU_FETCH_NARROW_MIN", "0" if narrow else "-1")
    f = fs.femm
    xyz, conn = meshes.shell_mesh("q4", n=64)
    od = meshes.clamp_edge_dofs(xyz)
    femm = _make_femm(fs, "q4", conn)
    geom0, dchi, u0, R0 = _fields(f, xyz, od)
    f.associategeometry(femm, geom0)
    f.stiffness(femm, f.SysmatAssemblerSparse(), geom0, u0, R0, dchi)
    m, n, nnz = femm.ctx.result_size()
    assert nnz > (1 << 20) and n > 20000
    pinned = [torch.empty(k, dtype=dt, pin_memory=True) for k, dt in ((n + 1, torch.int64), (nnz, torch.int64), (nnz, torch.float64))]
    Kp = femm.ctx.fetch_matrix(out=tuple(t.numpy() for t in pinned))  # pinned: direct DMA
    Kh = femm.ctx.fetch_matrix()  # numpy arrays: pageable, the SAME device result
    assert np.array_equal(Kh.colptr, Kp.colptr) and np.array_equal(Kh.rowval, Kp.rowval)
    assert np.array_equal(Kh.nzval, Kp.nzval)
    v = np.full(nnz, np.nan)
    femm.ctx.fetch_values(v)  # values only (pattern reused)
    assert np.array_equal(v, Kp.nzval)


# ---------------------------------------------------------------------------------------
# explicit loop: arbitrary load-factor table (force!(F, t) of plate_expl_examples.jl:86 sampled per step)
# ---------------------------------------------------------------------------------------
def test_explicit_load_factor_table(fs):
    f = fs.femm
    xy, conn = fx.t3block(1.0, 0.6, 10, 6)
    xyz = fx.xyz3(xy)
    xyz[:, 2] = 0.04 * np.cos(2 * xyz[:, 0]) * xyz[:, 1]
    od = meshes.clamp_edge_dofs(xyz, n_extra_fixed=0)
    femm = _make_femm(fs, "t3", conn)
    geom0, dchi, u0, R0 = _fields(f, xyz, od)
    f.associategeometry(femm, geom0)
    K = f.stiffness(femm, f.SysmatAssemblerFFBlock(), geom0, u0, R0, dchi)
    femm.ctx.shell_mass_diag(femm._params(), 3, nfree_only=True)
    nf = od.nfreedofs
    Mo = femm.ctx.fetch_vector(nf)
    Ko = K.to_scipy().tocsr()
    lam = oexp.pwr_largest(Ko, Mo, 200)
    dt = 0.8 * 2 / np.sqrt(lam)
    cs = 350.0
    rng = np.random.default_rng(8)
    F0 = rng.standard_normal(nf)
    nsteps = 120
    tab = np.where(np.arange(nsteps) < 40, np.linspace(0, 3, nsteps), 0.0) + 0.2 * rng.standard_normal(nsteps)  # ramp, drop, noise
    ex = fs.Explicit(femm.ctx, c_scale=cs, dt=dt)
    ex.set_load(F0)
    ex.start(float(0.0))
    ex.step(50, tab[:50])
    ex.step(nsteps - 50, tab[50:])
    U, V, A = ex.get_state()
    k = {"i": -1}

    def force(t):  # call 0: the initial acceleration (factor of ex.start); call s >= 1: step s
        v = F0 * (0.0 if k["i"] < 0 else tab[k["i"]])
        k["i"] += 1
        return v

    Uo, Vo, Ao = oexp.cd_loop(Mo, Ko, cs, np.zeros(nf), np.zeros(nf), nsteps, dt, force)
    assert relfro(U, Uo) < 1e-9 and relfro(V, Vo) < 1e-9 and relfro(A, Ao) < 1e-9
    ex.close()


# ---------------------------------------------------------------------------------------
# slices of the ACTUAL bench meshes (full-size node sets, first 100k elements) against the C port of the
# reference algorithm (C2, C4) and the NumPy oracle (C3, C5): element matrices and the assembled CSC
# ---------------------------------------------------------------------------------------
NSLICE = 100_000


def _ws_fields(f, w):
    geom0 = f.NodalField.__new__(f.NodalField)
    geom0.values = w["xyz"]
    dchi = f.NodalField.__new__(f.NodalField)
    dchi.values, dchi.dofnums, dchi._nfree = None, w["dofnums"], w["nfree"]
    return geom0, dchi


def _cport_elmats_and_csc(refport, nn, conn, xyz, normals, valid, dofnums, nfree, E, nu, t, alpha):
    Dps, Dt = osh.shell_material_stiffness(fx.moduli_iso(E, nu))
    Dps, Dt = np.ascontiguousarray(Dps), np.ascontiguousarray(Dt)
    pc, wt = fx.gauss_rule_2x2()
    pc = np.ascontiguousarray(pc)
    n, ne, nnodes = 6 * nn, conn.shape[0], xyz.shape[0]
    connC = np.ascontiguousarray(conn)
    xyzF, nF, v8, dF = np.asfortranarray(xyz), np.asfortranarray(normals), np.ascontiguousarray(valid.astype(np.uint8)), np.asfortranarray(dofnums)
    out = np.zeros((ne, n, n))
    refport.ref_shell_stiffness_elmats(nn, C.c_int64(ne), P(connC), C.c_int64(nnodes), P(xyzF), P(nF), P(v8), P(Dps), P(Dt), C.c_double(t),
                                       C.c_double(alpha), C.c_double(1.0), 4, P(pc), P(wt), P(out))
    nt = ne * n * n
    I, J, V = np.zeros(nt, np.int64), np.zeros(nt, np.int64), np.zeros(nt)
    refport.ref_shell_stiffness_coo(nn, C.c_int64(ne), P(connC), C.c_int64(nnodes), P(xyzF), P(nF), P(v8), P(dF), P(Dps), P(Dt), C.c_double(t),
                                    C.c_double(alpha), C.c_double(1.0), 4, P(pc), P(wt), os.cpu_count() or 1, P(I), P(J), P(V))
    nall = dofnums.size
    cp, rv, nz = np.zeros(nfree + 1, np.int64), np.zeros(nt, np.int64), np.zeros(nt)
    nnz = refport.ref_coo_to_csc(C.c_int64(nt), P(I), P(J), P(V), C.c_int64(nall), C.c_int64(nall), C.c_int64(nfree), C.c_int64(nfree), P(cp), P(rv), P(nz))
    return out.transpose(0, 2, 1), cp, rv[:nnz], nz[:nnz]


@pytest.mark.parametrize("cfg", ["C2", "C4"])
def test_bench_mesh_slice_against_c_port(fs, refport, cfg):
    from fsb200 import workloads as wl

    f = fs.femm
    if cfg == "C2":
        w, nn, alpha = wl.c2_q4rs_plate(1000), 4, 0.1
    else:
        w, nn, alpha = wl.c4_t3ff_panel(2000, 1000), 3, osh.T3_DEFAULT_ALPHA
    mat = f.MatDeforElastIso(w["E"], w["nu"], w["rho"])
    geom0, dchi = _ws_fields(f, w)
    # nodal normals of the FULL mesh (device), then operators on the first NSLICE elements of it
    full = (f.FEMMShellQ4RS if nn == 4 else f.FEMMShellT3FF)(f.IntegDomain(w["conn"], f.GaussRule2x2() if nn == 4 else None, w["thickness"]), mat)
    f.associategeometry(full, geom0)
    normals, valid = full._normals, full._normal_valid
    full.ctx.close()
    conn = np.ascontiguousarray(w["conn"][:NSLICE])
    femm = (f.FEMMShellQ4RS if nn == 4 else f.FEMMShellT3FF)(f.IntegDomain(conn, f.GaussRule2x2() if nn == 4 else None, w["thickness"]), mat)
    femm._normals, femm._normal_valid, femm._associatedgeometry = normals, valid, True
    K = f.stiffness(femm, f.SysmatAssemblerFFBlock(), geom0, None, None, dchi)
    Kg = femm.ctx.element_matrices(femm._kind(), 0, femm._params())
    Ke, cp, rv, nz = _cport_elmats_and_csc(refport, nn, conn, np.asarray(w["xyz"]), np.asarray(normals), np.asarray(valid), w["dofnums"], w["nfree"],
                                           w["E"], w["nu"], w["thickness"], alpha)
    num = np.sqrt(np.einsum("eij,eij->e", Kg - Ke, Kg - Ke))
    den = np.sqrt(np.einsum("eij,eij->e", Ke, Ke))
    assert (num / den).max() < TOL, (num / den).max()
    assert np.array_equal(K.colptr, cp) and np.array_equal(K.rowval, rv), "pattern of the bench-mesh slice"
    assert relfro(K.nzval, nz) < TOL
    femm.ctx.close()


def test_bench_mesh_slice_c3(fs):
    from fsb200 import workloads as wl

    f = fs.femm
    w = wl.c3_t3ffcomp_cylinder(1000, 1000)
    lam = w["lamina"]
    t = w["thickness"]
    mat = f.lamina_material(*lam)
    plies = [f.Ply(f"p{k}", mat, t / 4, a) for k, a in enumerate(w["angles"])]
    conn = np.ascontiguousarray(w["conn"][:NSLICE])
    femm = f.FEMMShellT3FFComp(f.IntegDomain(conn, None, t), f.CompositeLayup("C3", plies, wl.cylindrical_csys))
    geom0, dchi = _ws_fields(f, w)
    xyz = np.asarray(w["xyz"])
    femm._sync_mesh(geom0)
    femm._normals, femm._normal_valid, femm._associatedgeometry = np.asfortranarray(w["normals"]), np.ones(xyz.shape[0], bool), True
    femm.ctx.set_normals(femm._normals, femm._normal_valid)
    femm._sync_stab()
    Kg = femm.ctx.element_matrices(femm._kind(), 0, femm._params())
    D6 = oly.lamina_moduli(*lam[1:])
    lay = oly.CompositeLayup("c3", [oly.Ply(f"p{k}", D6, t / 4, a, lam[0]) for k, a in enumerate(w["angles"])])
    A, B, D = lay.laminate_stiffnesses()
    H = lay.laminate_transverse_stiffness()
    lcs = wl.cylindrical_csys(xyz[conn - 1].mean(axis=1))
    Ko = osh.t3ffcomp_stiffness_elmats(xyz, conn, np.asarray(w["normals"]), np.ones(xyz.shape[0], bool), A, B, D, H, lay.thickness, lcs)
    num = np.sqrt(np.einsum("eij,eij->e", Kg - Ko, Kg - Ko))
    den = np.sqrt(np.einsum("eij,eij->e", Ko, Ko))
    assert (num / den).max() < TOL, (num / den).max()
    md, mi = lay.laminate_inertia()
    Mg = femm.ctx.element_matrices(femm._kind(), 1, femm._params())
    assert relfro(Mg, osh.t3ffcomp_mass_elmats(xyz, conn, md, mi)) < TOL
    femm.ctx.close()


def test_bench_mesh_slice_c5(fs):
    from fsb200 import workloads as wl

    f = fs.femm
    w = wl.c5_beam_lattice(69)
    sl = slice(0, NSLICE)
    sc = {k: (v[sl] if k != "x1x2" else v[sl]) for k, v in w["sections"].items()}
    conn = np.ascontiguousarray(w["conn"][sl])
    secs = f.FESetL2Beam(sc["A"], sc["I1"], sc["I2"], sc["I3"], sc["J"], sc["A2s"], sc["A3s"], sc["x1x2"])
    femm = f.FEMMCorotBeam(f.IntegDomain(conn), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]), secs)
    geom0, dchi = _ws_fields(f, w)
    xyz, u1, R1 = np.asarray(w["xyz"]), np.asarray(w["u1"]), np.asarray(w["Rfield1"])
    femm._sync_mesh(geom0)
    femm.ctx.set_state(u1, R1)
    for op, ref in ((0, obeam.beam_stiffness_elmats(xyz, conn, u1, R1, sc, w["E"], w["nu"])), (2, obeam.beam_geostiffness_elmats(xyz, conn, u1, R1, sc, w["E"], w["nu"]))):
        got = femm.ctx.element_matrices(2, op, femm._params())
        # 1e-12 on the relative Frobenius norm of the whole slice (the north star's measure); single elements whose
        # natural deformations nearly cancel are conditioned worse than that in the reference arithmetic itself
        assert relfro(got, ref) < TOL
        num = np.sqrt(np.einsum("eij,eij->e", got - ref, got - ref))
        den = np.sqrt(np.einsum("eij,eij->e", ref, ref))
        rel = num / np.where(den > 0, den, 1.0)
        assert np.median(rel) < 1e-14 and rel.max() < 1e-10, (np.median(rel), rel.max())
    ev = femm.ctx.element_vectors(femm._params())
    ref = obeam.beam_restoringforce_elvecs(xyz, conn, u1, R1, sc, w["E"], w["nu"])
    assert relfro(ev, ref) < TOL
    femm.ctx.close()


# ---------------------------------------------------------------------------------------
# built-in csys kinds evaluated on the device (fsgpu_associategeometry_csys, fsgpu_set_layup_csys)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,csname", [("t3", "cylindrical"), ("t3", "spherical"), ("t3", "normal_axis"), ("q4", "cylindrical"), ("q4", "normal_axis")])
def test_device_csys_kinds(fs, kind, csname):
    f = fs.femm
    rng = np.random.default_rng(21)
    gen = fx.t3block if kind == "t3" else fx.q4block
    xy, conn = gen(1.2, 0.9, 9, 6)
    if csname == "spherical":  # a spherical cap around the z axis, away from the pole
        th, ph = 0.5 + xy[:, 0], 0.4 + xy[:, 1]
        xyz = 0.7 * np.column_stack([np.sin(ph) * np.cos(th), np.sin(ph) * np.sin(th), np.cos(ph)]) + np.array([0.1, -0.2, 0.3])
        cs = f.CSysKind.spherical(axis=(0.0, 0.0, 1.0), origin=(0.1, -0.2, 0.3))
    else:  # a cylindrical panel around an inclined axis through a shifted origin
        ax = np.array([0.2, 1.0, -0.1])
        ax /= np.linalg.norm(ax)
        p1 = np.cross(ax, [0.0, 0.0, 1.0])
        p1 /= np.linalg.norm(p1)
        p2 = np.cross(ax, p1)
        org = np.array([0.3, -0.1, 0.2])
        a = 0.3 + xy[:, 0]
        xyz = org + 0.5 * (np.cos(a)[:, None] * p1 + np.sin(a)[:, None] * p2) + xy[:, 1][:, None] * ax
        cs = f.CSysKind.cylindrical(axis=ax, origin=org) if csname == "cylindrical" else f.CSysKind.normal_axis(axis=ax)
    xyz = xyz + rng.uniform(-1, 1, xyz.shape) * 2e-3
    lay, _ = _layup()
    idom = f.IntegDomain(conn, None if kind == "t3" else f.GaussRule2x2(), T_)
    femm = (f.FEMMShellT3FFComp if kind == "t3" else f.FEMMShellQ4RSComp)(idom, _fs_layup(fs, cs))
    geom0, dchi, u0, R0 = _fields(f, xyz)
    f.associategeometry(femm, geom0)
    X = xyz[conn - 1]
    ne, nn = conn.shape
    dirs = np.zeros((ne, nn, 3))
    for k in range(nn):
        if nn == 3:
            J = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=-1)
        else:
            J = np.einsum("eai,ak->eik", X, fx.q4_shape(*fx.nodal_rule_q4()[0][k])[1])
        dirs[:, k] = cs(X[:, k], J)[:, :, 2]
    no, vo = (osh.t3ff_associategeometry if kind == "t3" else osh.q4rs_associategeometry)(xyz, conn, normal_dir=dirs)
    assert np.abs(femm._normals - no).max() < 1e-13 and np.array_equal(femm._normal_valid, vo)
    A, B, D = lay.laminate_stiffnesses()
    H = lay.laminate_transverse_stiffness()
    if kind == "t3":
        J0 = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=-1)
        lcs = cs(X.mean(axis=1), J0)
        Ko = osh.t3ffcomp_stiffness_elmats(xyz, conn, no, vo, A, B, D, H, lay.thickness, lcs)
    else:
        if csname == "cylindrical":
            # the reference hands Q4RSComp's layup csys the shape-function values as the location (App. B.9): a
            # position-based csys is meaningless there (the normals above ARE position-based and are checked);
            # normal_axis (tangent-based) is the usable kind for the stiffness and is tested
            return
        pc, w = fx.gauss_rule_2x2()
        lcs = np.zeros((ne, len(w), 3, 3))
        for j in range(len(w)):
            N, dNp = fx.q4_shape(*pc[j])
            lcs[:, j] = cs(np.broadcast_to(np.ravel(N), (ne, 4)), np.einsum("eai,ak->eik", X, dNp))
        Ko = osh.q4rscomp_stiffness_elmats(xyz, conn, no, vo, A, B, D, H, lay.thickness, lcs)
    femm._sync_stab()
    Kg = femm.ctx.element_matrices(femm._kind(), 0, femm._params())
    assert max(relfro(Kg[e], Ko[e]) for e in range(ne)) < TOL


# ---------------------------------------------------------------------------------------
# COO -> CSC: duplicates are combined left to right in input order, as Julia's sparse() does -- bit for bit
# ---------------------------------------------------------------------------------------
def test_coo_to_csc_bitwise_against_c_port(fs, refport):
    rng = np.random.default_rng(31)
    m, n, nt = 5000, 3000, 400_000
    I = rng.integers(1, m + 1, nt).astype(np.int64)
    J = rng.integers(1, n + 1, nt).astype(np.int64)
    # heavy duplication on a few entries (long segments) and values spanning 30 orders of magnitude: any other
    # association of the additions changes the last bits
    hot = rng.integers(0, nt, 60_000)
    I[hot], J[hot] = I[hot[0]] % 7 + 1, J[hot[0]] % 5 + 1
    V = rng.standard_normal(nt) * 10.0 ** rng.uniform(-15, 15, nt)
    ctx = fs.Context()
    S = ctx.coo_to_csc(I, J, V, m, n)
    cp, rv, nz = S.colptr, S.rowval, S.nzval
    a = [C.c_int64(nt), P(I), P(J), P(V), C.c_int64(m), C.c_int64(n), C.c_int64(m), C.c_int64(n)]
    nnz = refport.ref_coo_to_csc(*a, None, None, None)
    rcp, rrv, rnz = np.zeros(n + 1, np.int64), np.zeros(nnz, np.int64), np.zeros(nnz)
    refport.ref_coo_to_csc(*a, P(rcp), P(rrv), P(rnz))
    assert np.array_equal(cp, rcp) and np.array_equal(rv, rrv)
    assert np.array_equal(nz, rnz), "values must be bitwise those of the sequential left-to-right sums"
    assert np.array_equal(nz, ctx.coo_to_csc(I, J, V, m, n).nzval)  # and reproducible from run to run


# ---------------------------------------------------------------------------------------
# SysmatAssemblerSparseSymm (the default assembler): `S + S'` drops results that are exactly 0.0, so the stored
# pattern depends on the VALUES.  What can be guaranteed, and is asserted here:
#  * entries all of whose element contributions are exactly zero (membrane/bending coupling of a flat plate, the
#    off-diagonals of a lumped mass) are dropped by the GPU path exactly as by the reference semantics;
#  * on a mesh without exact symmetries the patterns are identical;
#  * the two patterns can differ ONLY in entries that are sums of >= 2 non-zero element contributions cancelling by
#    mesh symmetry -- whether such a sum is exactly 0.0 or 1e-17 of the largest entry depends on the order of the
#    floating-point operations (in the reference: on its BLAS), never on anything structural.
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["t3", "q4"])
def test_symm_pattern_characterised(fs, kind):
    import scipy.sparse as sp

    f = fs.femm
    for flat in (True, False):
        if flat:
            xy, conn = (fx.t3block if kind == "t3" else fx.q4block)(2.0, 1.0, 7, 5)
            xyz = fx.xyz3(xy)
        else:
            xyz, conn = meshes.shell_mesh(kind, n=6)
        femm = _make_femm(fs, kind, conn)
        geom0, dchi, u0, R0 = _fields(f, xyz)
        f.associategeometry(femm, geom0)
        K = f.stiffness(femm, geom0, u0, R0, dchi)  # default assembler = SparseSymm
        normals, valid = _oracle_normals(kind, xyz, conn)
        Dps, Dt = _iso()
        Ko = (osh.t3ff_stiffness_elmats if kind == "t3" else osh.q4rs_stiffness_elmats)(xyz, conn, normals, valid, Dps, Dt, T_)
        od = fx.DofField(xyz.shape[0]).numberdofs()
        dn = od.gatherdofnums(conn)
        n = od.nalldofs
        cp, rv, nz = fx.assemble_matrix("symm", Ko, dn, n)
        Kg, Kr = K.to_scipy().tocsc(), fx.csc_to_scipy(cp, rv, nz, n, n).tocsc()
        assert relfro(Kg.toarray(), Kr.toarray()) < TOL and abs(Kg - Kg.T).max() == 0.0
        I, J, V = fx.coo_full(Ko, dn)
        nzc = sp.coo_matrix(((V != 0).astype(np.float64), (I - 1, J - 1)), shape=(n, n)).tocsc()  # non-zero contributions per entry
        Pg = sp.csc_matrix((np.ones_like(Kg.data), Kg.indices, Kg.indptr), shape=(n, n))
        Pr = sp.csc_matrix((np.ones_like(Kr.data), Kr.indices, Kr.indptr), shape=(n, n))
        # (1) structural zeros: never stored by either
        struct_zero = sp.coo_matrix((np.ones(len(V)), (I - 1, J - 1)), shape=(n, n)).tocsc()
        struct_zero.data[:] = 1.0
        only_zero = (struct_zero - (nzc > 0).astype(np.float64)).tocsc()
        only_zero.eliminate_zeros()
        assert Pg.multiply(only_zero).nnz == 0 and Pr.multiply(only_zero).nnz == 0
        if flat:
            assert only_zero.nnz > 0 and Kg.nnz < struct_zero.nnz, "the flat plate must have droppable zeros"
        d = (Pg - Pr).tocoo()
        sel = d.data != 0
        if not flat:
            assert sel.sum() == 0, "no exact symmetries: identical patterns"
            continue
        # (2) every difference is a cancelling multi-element sum at round-off level
        rr, cc = d.row[sel], d.col[sel]
        if len(rr):
            assert np.asarray(nzc[rr, cc]).ravel().min() >= 2
            big = max(np.abs(Kg.data).max(), np.abs(Kr.data).max())
            assert np.abs(np.asarray((Kg + Kr)[rr, cc])).ravel().max() <= 1e-15 * big


# ---------------------------------------------------------------------------------------
# fsgpu_set_deterministic for Q4RS / Q4RSComp / the beam / T3FF with AVERAGE_K (no tile kernel): the order-fixed
# gather path -- same pattern, values within 1e-12 of the oracle, bitwise identical from run to run
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["q4", "q4comp", "t3_sheark", "beam"])
def test_deterministic_gather_path(fs, case):
    if os.environ.get("FSGPU_FORCE_GENERIC"):
        pytest.skip("the gather path needs the run-structured (fast) addressing")
    f = fs.femm
    if case == "beam":
        xyz, conn, u1, R1, sec = meshes.beam_lattice()
        EB, NUB = 71240.0, 0.31
        secs = f.FESetL2Beam(sec["A"], sec["I1"], sec["I2"], sec["I3"], sec["J"], sec["A2s"], sec["A3s"], sec["x1x2"])
        femm = f.FEMMCorotBeam(f.IntegDomain(conn), f.MatDeforElastIso(EB, NUB, 5e-9), secs)
        od = fx.DofField(xyz.shape[0])
        for c in range(1, 7):
            od.setebc([0], c)
        od.numberdofs()
        geom0, dchi, _, _ = _fields(f, xyz, od)
        uf, Rf = f.NodalField(u1), f.NodalField(R1)
        femm.ctx.set_deterministic(True)
        ops = [(lambda a: f.stiffness(femm, a, geom0, uf, Rf, dchi), obeam.beam_stiffness_elmats(xyz, conn, u1, R1, sec, EB, NUB)),
               (lambda a: f.geostiffness(femm, a, geom0, uf, Rf, dchi), obeam.beam_geostiffness_elmats(xyz, conn, u1, R1, sec, EB, NUB))]
    else:
        kind = "t3" if case.startswith("t3") else "q4"
        comp = case.endswith("comp")
        xyz, conn = meshes.shell_mesh(kind, n=9)
        lay, cs = _layup()
        od = meshes.clamp_edge_dofs(xyz)
        femm = _make_femm(fs, kind, conn, comp, cs)
        sheark = 1 if case == "t3_sheark" else 0
        femm.transv_shear_formulation = sheark
        femm.ctx.set_deterministic(True)
        geom0, dchi, u0, R0 = _fields(f, xyz, od)
        f.associategeometry(femm, geom0)
        normals, valid = _oracle_normals(kind, xyz, conn, fixed=cs[:, 2] if comp else None)
        from tests.test_gpu_parity import _oracle_K

        ops = [(lambda a: f.stiffness(femm, a, geom0, u0, R0, dchi), _oracle_K(kind, comp, xyz, conn, normals, valid, cs, sheark=sheark))]
    dn = od.gatherdofnums(conn)
    for op, Ke in ops:
        for asm, mk in (("ffblock", f.SysmatAssemblerFFBlock), ("sparse", f.SysmatAssemblerSparse)):
            K1 = op(mk())
            assert femm.ctx.scatter_path == 3, "gather path did not engage"
            K2 = op(mk())
            assert np.array_equal(K1.nzval, K2.nzval), "order-fixed assembly must be bitwise reproducible"
            n = od.nfreedofs if asm == "ffblock" else od.nalldofs
            _check_matrix(K1, fx.assemble_matrix(asm, Ke, dn, od.nalldofs, od.nfreedofs), n)
    femm.ctx.set_deterministic(False)
    op, Ke = ops[0]
    op(f.SysmatAssemblerFFBlock())
    assert femm.ctx.scatter_path == 1


# ---------------------------------------------------------------------------------------
# explicit loop with the reference's two closures: a general force!(F, t) (spatial distribution changing every step)
# and peek(step, U, V, t) every nbtw steps (plate_expl_examples.jl:86,92)
# ---------------------------------------------------------------------------------------
def test_explicit_force_and_peek_closures(fs):
    f = fs.femm
    xy, conn = fx.t3block(1.0, 0.6, 8, 5)
    xyz = fx.xyz3(xy)
    xyz[:, 2] = 0.03 * np.sin(3 * xyz[:, 0])
    od = meshes.clamp_edge_dofs(xyz, n_extra_fixed=0)
    femm = _make_femm(fs, "t3", conn)
    geom0, dchi, u0, R0 = _fields(f, xyz, od)
    f.associategeometry(femm, geom0)
    K = f.stiffness(femm, f.SysmatAssemblerFFBlock(), geom0, u0, R0, dchi)
    femm.ctx.shell_mass_diag(femm._params(), 3, nfree_only=True)
    nf = od.nfreedofs
    Mo = femm.ctx.fetch_vector(nf)
    Ko = K.to_scipy().tocsr()
    dt = 0.8 * 2 / np.sqrt(oexp.pwr_largest(Ko, Mo, 200))
    cs = 500.0
    rng = np.random.default_rng(9)
    Fa, Fb = rng.standard_normal(nf), rng.standard_normal(nf)
    force = lambda t: Fa * np.sin(4.0e4 * t) + Fb * (t * 1.0e5) ** 2  # not separable: the pattern itself moves
    nsteps, nbtw = 60, 15
    seen, seen_o = [], []
    ex = fs.Explicit(femm.ctx, c_scale=cs, dt=dt)
    U, V, A = ex.run(nsteps, dt, force=force, peek=lambda s, U, V, t: seen.append((s, t, U.copy(), V.copy())), nbtw=nbtw)
    Uo, Vo, Ao = oexp.cd_loop(Mo, Ko, cs, np.zeros(nf), np.zeros(nf), nsteps, dt, force,
                              peek=lambda s, U, V, t: seen_o.append((s, t, U.copy(), V.copy())) if s % nbtw == 0 else None)
    assert relfro(U, Uo) < 1e-9 and relfro(V, Vo) < 1e-9 and relfro(A, Ao) < 1e-9
    assert [s for s, *_ in seen] == [s for s, *_ in seen_o] == [0, 15, 30, 45, 60]
    for (s, t, Ug, Vg), (so, to, Ur, Vr) in zip(seen[1:], seen_o[1:]):
        assert abs(t - to) < 1e-15 and relfro(Ug, Ur) < 1e-9 and relfro(Vg, Vr) < 1e-9
    ex.close()
    # separable form on the device, peeked
    ex = fs.Explicit(femm.ctx, c_scale=cs, dt=dt)
    sc = lambda t: np.cos(3.0e4 * t)
    U2, V2, _ = ex.run(nsteps, dt, F0=Fa, fscale=sc, peek=lambda *a: None, nbtw=7)
    U2o, V2o, _ = oexp.cd_loop(Mo, Ko, cs, np.zeros(nf), np.zeros(nf), nsteps, dt, lambda t: Fa * sc(t))
    assert relfro(U2, U2o) < 1e-9 and relfro(V2, V2o) < 1e-9
    ex.close()



def test_explicit_create_rejects_malformed_csr(fs):
    """fsgpu_explicit_create validates the CSR on the device (a bad index would be an out-of-bounds read in the step kernel)."""
    n = 50
    rp = np.arange(1, 2 * n + 2, 2, dtype=np.int64)  # two entries per row
    cv = np.tile(np.array([1, 2], dtype=np.int64), n)
    nz = np.ones(2 * n)
    M = np.ones(n)
    ctx = fs.Context()
    ex = fs.Explicit(ctx, K=(rp, cv, nz), mdiag=M, c_scale=0.0, dt=1e-3)  # well formed
    ex.close()
    bad_cv = cv.copy()
    bad_cv[17] = n + 5
    with pytest.raises(fs._lib.FsgpuError) as ei:
        fs.Explicit(ctx, K=(rp, bad_cv, nz), mdiag=M, c_scale=0.0, dt=1e-3)
    assert ei.value.code == fs._lib.ERR_ARG
    bad_rp = rp.copy()
    bad_rp[10] = bad_rp[12]
    with pytest.raises(fs._lib.FsgpuError):
        fs.Explicit(ctx, K=(bad_rp, cv, nz), mdiag=M, c_scale=0.0, dt=1e-3)
