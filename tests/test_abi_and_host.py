"""CPU-side checks: libfsgpu.so loads and exports every symbol of include/fsgpu.h, fails loudly
without a GPU; the C port of the reference algorithm and the kernels' element math (host build of
csrc/fsgpu_math.cuh) agree with the NumPy oracle; the host mirror's layup integration agrees too."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import beam as obeam
from oracle import fe_external as fx
from oracle import layup as oly
from oracle import shells as osh
from tests import meshes
from tests.conftest import ROOT, has_gpu

P = lambda a: a.ctypes.data_as(C.c_void_p)


def relfro(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_library_exports_every_declared_symbol():
    import fsb200

    hdr = open(os.path.join(ROOT, "include", "fsgpu.h")).read()
    declared = set(re.findall(r"\b(fsgpu_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib = C.CDLL(fsb200.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(fsb200.EXPORTED_SYMBOLS), declared ^ set(fsb200.EXPORTED_SYMBOLS)


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    import fsb200

    with pytest.raises(fsb200.FsgpuError) as ei:
        fsb200.Context()
    assert "no CPU fallback" in str(ei.value)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "finetoolsflexstructures.jl_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".jl", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("see oracle/", ""), fn


@pytest.fixture(scope="module")
def refport():
    d = os.path.join(ROOT, "oracle", "cport")
    subprocess.run(["make", "-C", d, "-s"], check=True)
    lib = C.CDLL(os.path.join(d, "librefport.so"))
    lib.ref_coo_to_csc.restype = C.c_int64
    return lib


@pytest.mark.parametrize("kind", ["t3", "q4"])
def test_c_port_matches_numpy_oracle(refport, kind):
    xyz, conn = meshes.shell_mesh(kind, n=8)
    nn = 3 if kind == "t3" else 4
    n = 6 * nn
    Dps, Dt = osh.shell_material_stiffness(fx.moduli_iso(200e9, 0.3))
    pc, w = fx.gauss_rule_2x2()
    if kind == "t3":
        nrm, val = osh.t3ff_associategeometry(xyz, conn)
        Ko = osh.t3ff_stiffness_elmats(xyz, conn, nrm, val, Dps, Dt, 0.01, drilling_stiffness_scale=0.8)
        alpha = osh.T3_DEFAULT_ALPHA
    else:
        nrm, val = osh.q4rs_associategeometry(xyz, conn)
        Ko = osh.q4rs_stiffness_elmats(xyz, conn, nrm, val, Dps, Dt, 0.01, drilling_stiffness_scale=0.8)
        alpha = 0.1
    od = meshes.clamp_edge_dofs(xyz)
    xyzF, nF, v8 = np.asfortranarray(xyz), np.asfortranarray(nrm), val.astype(np.uint8)
    connC, dF = np.ascontiguousarray(conn), np.asfortranarray(od.dofnums)
    Dps, Dt, pc = np.ascontiguousarray(Dps), np.ascontiguousarray(Dt), np.ascontiguousarray(pc)
    out = np.zeros((len(conn), n, n))
    refport.ref_shell_stiffness_elmats(nn, C.c_int64(len(conn)), P(connC), C.c_int64(len(xyz)), P(xyzF), P(nF), P(v8), P(Dps), P(Dt),
                                       C.c_double(0.01), C.c_double(alpha), C.c_double(0.8), 4, P(pc), P(w), P(out))
    got = out.transpose(0, 2, 1)
    assert max(relfro(got[e], Ko[e]) for e in range(len(conn))) < 1e-13
    nt = len(conn) * n * n
    I, J, V = np.zeros(nt, np.int64), np.zeros(nt, np.int64), np.zeros(nt)
    refport.ref_shell_stiffness_coo(nn, C.c_int64(len(conn)), P(connC), C.c_int64(len(xyz)), P(xyzF), P(nF), P(v8), P(dF), P(Dps), P(Dt),
                                    C.c_double(0.01), C.c_double(alpha), C.c_double(0.8), 4, P(pc), P(w), 2, P(I), P(J), P(V))
    Io, Jo, Vo = fx.coo_full(Ko, od.gatherdofnums(conn))
    assert np.array_equal(I, Io) and np.array_equal(J, Jo), "COO emission order"
    na, nf = od.nalldofs, od.nfreedofs
    a = [C.c_int64(nt), P(I), P(J), P(V), C.c_int64(na), C.c_int64(na), C.c_int64(nf), C.c_int64(nf)]
    nnz = refport.ref_coo_to_csc(*a, None, None, None)
    cp, rv, nz = np.zeros(nf + 1, np.int64), np.zeros(nnz, np.int64), np.zeros(nnz)
    refport.ref_coo_to_csc(*a, P(cp), P(rv), P(nz))
    rcp, rrv, rnz = fx.assemble_matrix("ffblock", Ko, od.gatherdofnums(conn), na, nf)
    assert np.array_equal(cp, rcp) and np.array_equal(rv, rrv) and relfro(nz, rnz) < 1e-13


def test_c_port_explicit_matches_oracle(refport):
    import scipy.sparse as sp

    from oracle import explicit as oexp

    rng = np.random.default_rng(0)
    n = 200
    A = sp.random(n, n, 0.05, random_state=1, format="csr")
    K = (A + A.T + sp.identity(n) * 5).tocsr()
    K.sort_indices()
    M = rng.uniform(1, 2, n)
    F0 = rng.standard_normal(n)
    dt, cs, ns = 0.05, 0.3, 50
    fsc = np.sin(np.arange(1, ns + 1) * dt)
    Uo, Vo, Ao = oexp.cd_loop(M, K, cs, np.zeros(n), np.zeros(n), ns, dt, lambda t: F0 * np.sin(t))
    U, V = np.zeros(n), np.zeros(n)
    A0 = (1.0 / (M + dt / 2 * cs * M)) * (F0 * 0.0)
    rp, cv = (K.indptr + 1).astype(np.int64), (K.indices + 1).astype(np.int64)
    refport.ref_explicit_steps(C.c_int64(n), P(rp), P(cv), P(K.data), P(M), C.c_double(cs), C.c_double(dt), P(F0), P(fsc), C.c_int64(ns), P(U), P(V), P(A0), 1)
    assert relfro(U, Uo) < 1e-12 and relfro(V, Vo) < 1e-12


@pytest.fixture(scope="module")
def hostmath(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hm") / "libhostmath.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "hostmath", "harness.cpp")], check=True)
    return C.CDLL(so)


def _layup():
    D6 = oly.lamina_moduli(133860e6, 7706e6, 0.301, 4306e6, 4306e6, 2760e6)
    lay = oly.CompositeLayup("x", [oly.Ply(f"p{k}", D6, 0.0025, a, 1500.0) for k, a in enumerate((0, 90, 45, -30))])
    A, B, D = lay.laminate_stiffnesses()
    H = lay.laminate_transverse_stiffness()
    md, mi = lay.laminate_inertia()
    th = np.deg2rad(20.0)
    cs = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    return lay, A, B, D, H, np.concatenate([A.ravel(), B.ravel(), D.ravel(), H.ravel(), [lay.thickness, md, mi]]), cs


@pytest.mark.parametrize("comp", [False, True])
@pytest.mark.parametrize("sheark", [0, 1])
def test_kernel_math_t3(hostmath, comp, sheark):
    """The element math the CUDA kernels are built from (csrc/fsgpu_math.cuh), compiled for the host."""
    xyz, conn = meshes.shell_mesh("t3", n=6)
    nrm, val = osh.t3ff_associategeometry(xyz, conn)
    assert (~val).any()
    Dps, Dt = osh.shell_material_stiffness(fx.moduli_iso(200e9, 0.3))
    lay, A, B, D, H, group, cs = _layup()
    if comp:
        Ko = osh.t3ffcomp_stiffness_elmats(xyz, conn, nrm, val, A, B, D, H, lay.thickness, cs, drilling_stiffness_scale=0.7, transv_shear_formulation=sheark)
    else:
        Ko = osh.t3ff_stiffness_elmats(xyz, conn, nrm, val, Dps, Dt, 0.01, drilling_stiffness_scale=0.7, transv_shear_formulation=sheark)
    Dpsc, Dt56, csc = np.ascontiguousarray(Dps), np.ascontiguousarray(Dt * 5 / 6), np.ascontiguousarray(cs)
    worst = 0.0
    for e in range(len(conn)):
        c = conn[e] - 1
        X, N, V = np.ascontiguousarray(xyz[c]), np.ascontiguousarray(nrm[c]), np.ascontiguousarray(val[c].astype(np.uint8))
        out = np.zeros((18, 18))
        hostmath.hm_t3_elmat(P(X), P(N), P(V), P(Dpsc), P(Dt56), C.c_double(0.01), C.c_double(osh.T3_DEFAULT_ALPHA), C.c_double(0.7), sheark,
                             P(group) if comp else None, P(csc) if comp else None, P(out))
        worst = max(worst, relfro(out.T, Ko[e]))
    assert worst < 1e-13, worst


@pytest.mark.parametrize("comp", [False, True])
def test_kernel_math_q4(hostmath, comp):
    xyz, conn = meshes.shell_mesh("q4", n=6)
    nrm, val = osh.q4rs_associategeometry(xyz, conn)
    assert (~val).any()
    Dps, Dt = osh.shell_material_stiffness(fx.moduli_iso(200e9, 0.3))
    lay, A, B, D, H, group, cs = _layup()
    pc, w = fx.gauss_rule_2x2()
    if comp:
        Ko = osh.q4rscomp_stiffness_elmats(xyz, conn, nrm, val, A, B, D, H, lay.thickness, cs, drilling_stiffness_scale=0.9)
    else:
        Ko = osh.q4rs_stiffness_elmats(xyz, conn, nrm, val, Dps, Dt, 0.01, drilling_stiffness_scale=0.9)
    Dpsc, Dt56 = np.ascontiguousarray(Dps), np.ascontiguousarray(Dt * 5 / 6)
    xi, eta, css, tt = np.ascontiguousarray(pc[:, 0]), np.ascontiguousarray(pc[:, 1]), np.ascontiguousarray(np.tile(cs.ravel(), 4)), np.full(4, 0.01)
    worst = 0.0
    for e in range(len(conn)):
        c = conn[e] - 1
        X, N, V = np.ascontiguousarray(xyz[c]), np.ascontiguousarray(nrm[c]), np.ascontiguousarray(val[c].astype(np.uint8))
        out = np.zeros((24, 24))
        hostmath.hm_q4_elmat(P(X), P(N), P(V), P(Dpsc), P(Dt56), P(tt), C.c_double(0.1), C.c_double(0.9), 4, P(xi), P(eta), P(w),
                             P(group) if comp else None, P(css) if comp else None, P(out))
        worst = max(worst, relfro(out.T, Ko[e]))
    assert worst < 1e-13, worst


def test_kernel_math_beam(hostmath):
    xyz, conn, u1, R1, sec = meshes.beam_lattice(20)
    E, nu, rho = 71240.0, 0.31, 5e-9
    refs = {(0, 1): obeam.beam_stiffness_elmats(xyz, conn, u1, R1, sec, E, nu), (2, 1): obeam.beam_geostiffness_elmats(xyz, conn, u1, R1, sec, E, nu),
            (3, 1): obeam.beam_restoringforce_elvecs(xyz, conn, u1, R1, sec, E, nu)}
    for mt in range(4):
        refs[(1, mt)] = obeam.beam_mass_elmats(xyz, conn, u1, R1, sec, rho, mt)
    for (op, mt), ref in refs.items():
        for e in range(len(conn)):
            c = conn[e] - 1
            secv = np.array([sec[k][e] for k in ("A", "I1", "I2", "I3", "J", "A2s", "A3s")] + list(sec["x1x2"][e]))
            out = np.zeros(12) if op == 3 else np.zeros((12, 12))
            hostmath.hm_beam(P(np.ascontiguousarray(xyz[c])), P(np.ascontiguousarray(u1[c])), P(np.ascontiguousarray(R1[c[0]])), P(np.ascontiguousarray(R1[c[1]])),
                             P(secv), C.c_double(E), C.c_double(nu), C.c_double(rho), mt, op, P(out))
            assert relfro(out if op == 3 else out.T, ref[e]) < 1e-13


def test_host_mirror_layup_record_matches_oracle():
    import fsb200

    f = fsb200.femm
    lay, A, B, D, H, group, cs = _layup()
    mat = f.lamina_material(1500.0, 133860e6, 7706e6, 0.301, 4306e6, 4306e6, 2760e6)
    rec = f.CompositeLayup("x", [f.Ply(f"p{k}", mat, 0.0025, a) for k, a in enumerate((0, 90, 45, -30))], cs).group_record()
    assert np.abs(rec - group).max() <= 1e-15 * np.abs(group).max()
    d = f.NodalField(np.zeros((5, 6)))
    d.setebc([1, 3], 2)
    d.numberdofs([4, 2, 0, 1, 3])
    od = fx.DofField(5)
    od.setebc([1, 3], 2)
    od.numberdofs([4, 2, 0, 1, 3])
    assert np.array_equal(d.dofnums, od.dofnums) and f.nfreedofs(d) == od.nfreedofs


def test_workload_generators_match_finetools_block_meshes():
    import fsb200
    from fsb200 import workloads as wl

    for nL, nW in ((5, 4), (3, 7)):
        xy, c = fx.q4block(3.0, 2.0, nL, nW)
        X, Cq = wl.q4block(3.0, 2.0, nL, nW)
        assert np.array_equal(c, Cq) and np.allclose(X[:, :2], xy)
        xy, c = fx.t3block(3.0, 2.0, nL, nW)
        X, Ct = wl.t3block(3.0, 2.0, nL, nW)
        assert np.array_equal(c, Ct)

