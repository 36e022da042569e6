"""Writes tests/golden/oracle_fixtures.npz: outputs of the CPU oracle on small seeded inputs, frozen AFTER the oracle was
pinned on the reference's known-answer tests (tests/test_oracle_goldens.py).  The oracle is the checker of every GPU parity
test; this fixture makes an accidental change of the checker itself visible on CPU.  Re-generate deliberately only:
    python tests/golden/make_fixtures.py
(The reference is Julia and cannot run in this image, so these are not reference outputs; the reference's own golden
numbers are the constants cited file:line in tests/test_oracle_goldens.py and indexed in tests/golden/README.md.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import beam as obeam  # noqa: E402
from oracle import fe_external as fx  # noqa: E402
from oracle import layup as oly  # noqa: E402
from oracle import shells as osh  # noqa: E402
from tests import meshes  # noqa: E402


def compute():
    out = {}
    Dps, Dt = osh.shell_material_stiffness(fx.moduli_iso(200e9, 0.3))
    D6 = oly.lamina_moduli(133860e6, 7706e6, 0.301, 4306e6, 4306e6, 2760e6)
    lay = oly.CompositeLayup("fx", [oly.Ply(f"p{k}", D6, 0.0025, a, 1500.0) for k, a in enumerate((0, 90, 45, -30))])
    th = np.deg2rad(20.0)
    cs = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    A, B, D = lay.laminate_stiffnesses()
    H = lay.laminate_transverse_stiffness()
    for kind in ("t3", "q4"):
        xyz, conn = meshes.shell_mesh(kind, n=3)
        ag = osh.t3ff_associategeometry if kind == "t3" else osh.q4rs_associategeometry
        nrm, val = ag(xyz, conn)
        out[f"{kind}_normals"], out[f"{kind}_valid"] = nrm, val
        if kind == "t3":
            out["t3_K"] = osh.t3ff_stiffness_elmats(xyz, conn, nrm, val, Dps, Dt, 0.01)
            out["t3_M"] = osh.t3ff_mass_elmats(xyz, conn, 7850.0, 0.01)
            nc, vc = ag(xyz, conn, normal_dir=cs[:, 2])
            out["t3comp_K"] = osh.t3ffcomp_stiffness_elmats(xyz, conn, nc, vc, A, B, D, H, lay.thickness, cs)
        else:
            out["q4_K"] = osh.q4rs_stiffness_elmats(xyz, conn, nrm, val, Dps, Dt, 0.01)
            out["q4_M"] = osh.q4rs_mass_elmats(xyz, conn, 7850.0, 0.01)
            nc, vc = ag(xyz, conn, normal_dir=cs[:, 2])
            out["q4comp_K"] = osh.q4rscomp_stiffness_elmats(xyz, conn, nc, vc, A, B, D, H, lay.thickness, cs)
        u = np.random.default_rng(5).standard_normal((xyz.shape[0], 6)) * 1e-3
        rf = osh.t3ff_resultants if kind == "t3" else osh.q4rs_resultants
        out[f"{kind}_moment"] = rf(xyz, conn, nrm, val, Dps, Dt, 0.01, u, 1, ocs=cs)
        d = meshes.clamp_edge_dofs(xyz)
        cp, rv, nz = fx.assemble_matrix("ffblock", out[f"{kind}_K"], d.gatherdofnums(conn), d.nalldofs, d.nfreedofs)
        out[f"{kind}_colptr"], out[f"{kind}_rowval"], out[f"{kind}_nzval"] = cp, rv, nz
    xyz, conn, u1, R1, sec = meshes.beam_lattice(6)
    E, nu, rho = 71240.0, 0.31, 5e-9
    out["beam_K"] = obeam.beam_stiffness_elmats(xyz, conn, u1, R1, sec, E, nu)
    out["beam_Kgeo"] = obeam.beam_geostiffness_elmats(xyz, conn, u1, R1, sec, E, nu)
    out["beam_F"] = obeam.beam_restoringforce_elvecs(xyz, conn, u1, R1, sec, E, nu)
    out["beam_M1"] = obeam.beam_mass_elmats(xyz, conn, u1, R1, sec, rho, 1)
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_fixtures.npz")
    np.savez_compressed(path, **compute())
    print("wrote", path, os.path.getsize(path), "bytes")
