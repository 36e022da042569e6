"""Diagnostic: SparseSymm pattern of the GPU path vs the oracle on flat / curved meshes: which entries differ."""
import sys
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, ".")
import fsb200
from oracle import fe_external as fx, shells as osh
from tests.test_gpu_parity import _make_femm, _oracle_K, _oracle_normals
from tests import meshes

f = fsb200.femm
for name in ("flat_t3", "flat_q4", "flat_t3_rot", "curved_t3", "curved_q4"):
    kind = "t3" if "t3" in name else "q4"
    if name.startswith("flat"):
        xy, conn = (fx.t3block if kind == "t3" else fx.q4block)(2.0, 1.0, 6, 4)
        xyz = fx.xyz3(xy)
        if name.endswith("rot"):
            c, s = np.cos(0.3), np.sin(0.3)
            xyz = xyz @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]]).T
    else:
        xyz, conn = meshes.shell_mesh(kind, n=6)
    femm = _make_femm(fsb200, kind, conn)
    geom0 = f.NodalField(xyz)
    dchi = f.NodalField(np.zeros((xyz.shape[0], 6))).numberdofs()
    f.associategeometry(femm, geom0)
    u0, R0 = f.NodalField(np.zeros((xyz.shape[0], 3))), f.initial_Rfield(xyz.shape[0])
    K = f.stiffness(femm, geom0, u0, R0, dchi)
    normals, valid = _oracle_normals(kind, xyz, conn)
    Ko = _oracle_K(kind, False, xyz, conn, normals, valid)
    od = fx.DofField(xyz.shape[0]).numberdofs()
    cp, rv, nz = fx.assemble_matrix("symm", Ko, od.gatherdofnums(conn), od.nalldofs)
    full = fx.assemble_matrix("sparse", Ko, od.gatherdofnums(conn), od.nalldofs)
    n = od.nalldofs
    Kg, Kr = K.to_scipy().tocsc(), fx.csc_to_scipy(cp, rv, nz, n, n).tocsc()
    Pg = sp.csc_matrix((np.ones_like(Kg.data), Kg.indices, Kg.indptr), shape=Kg.shape)
    Pr = sp.csc_matrix((np.ones_like(Kr.data), Kr.indices, Kr.indptr), shape=Kr.shape)
    d = (Pg - Pr).tocoo()
    gpu_only = d.data > 0
    ref_only = d.data < 0
    mx = np.abs(nz).max()
    print(f"{name}: full nnz {len(full[1])}, oracle symm nnz {len(rv)}, gpu symm nnz {Kg.nnz}; gpu-only {gpu_only.sum()}, oracle-only {ref_only.sum()}")
    for lab, sel, M in (("gpu-only", gpu_only, Kg), ("oracle-only", ref_only, Kr)):
        if sel.sum():
            v = np.abs(np.asarray(M[d.row[sel], d.col[sel]])).ravel()
            cats = {}
            for r, c in zip(d.row[sel] % 6, d.col[sel] % 6):
                cats[(int(r), int(c))] = cats.get((int(r), int(c)), 0) + 1
            print(f"   {lab}: |value|/max in [{v.min() / mx:.1e}, {v.max() / mx:.1e}]  dof-type pairs {dict(sorted(cats.items()))}")
