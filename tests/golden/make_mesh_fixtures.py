"""Extracts the meshes of two of the reference's tests from their Abaqus input decks into small fixtures, as the reference's
`import_ABAQUS` + `compactnodes` deliver them (node coordinates in file order with the unconnected nodes removed, triangle
connectivities renumbered, 1-based):
  tests/golden/le5_mesh.npz          NAFEMS LE5 Z-section cantilever, test/test_shell_statics.jl:440-535 (nle5xf3c.inp)
  tests/golden/barrelvault_mesh.npz  irregular barrel vault of the resultants test, test/test_shell_statics.jl:577-764
                                     (barrelvault_stri3_irreg.inp; two coincident node pairs are left for `mergenodes`)
  tests/golden/raasch_meshes.npz     Raasch hook, S4 quads of the 1x9 / 3x18 / 5x36 / 10x72 decks (raasch_s4_*.inp),
                                     test/test_shell_statics.jl:140-236
Run in the build container (the reference tree is not available on the GPU box):  python tests/golden/make_mesh_fixtures.py"""
import os

import numpy as np

DECKS = {"le5_mesh.npz": "/root/reference/test/nle5xf3c.inp", "barrelvault_mesh.npz": "/root/reference/test/barrelvault_stri3_irreg.inp"}


def parse(path, nnpe=3):
    ids, xyz, conn = [], [], []
    mode = None
    for line in open(path):
        s = line.strip()
        if not s or s.startswith("**"):
            continue
        if s.startswith("*"):
            key = s.lower().split(",")[0].strip()
            mode = "node" if key == "*node" else ("elem" if key == "*element" else None)
            continue
        v = [x.strip() for x in s.split(",") if x.strip()]
        if mode == "node":
            ids.append(int(v[0]))
            xyz.append([float(x) for x in v[1:4]])
        elif mode == "elem":
            conn.append([int(x) for x in v[1:1 + nnpe]])
    return np.array(ids), np.array(xyz), np.array(conn, dtype=np.int64)


if __name__ == "__main__":
    for name, src in DECKS.items():
        ids, xyz, conn = parse(src)
        order = np.argsort(ids, kind="stable")  # import_ABAQUS stores node k at row k (ids are ascending in the deck)
        ids, xyz = ids[order], xyz[order]
        used = np.isin(ids, conn)  # findunconnnodes / compactnodes: connected nodes keep their relative order
        new = np.zeros(ids.max() + 1, dtype=np.int64)
        new[ids[used]] = np.arange(1, used.sum() + 1)
        out = os.path.join(os.path.dirname(os.path.abspath(__file__)), name)
        np.savez_compressed(out, xyz=xyz[used], conn=new[conn])
        print("wrote", out, xyz[used].shape, conn.shape)
    arrays = {}
    for m in ("1x9", "3x18", "5x36", "10x72"):
        ids, xyz, conn = parse(f"/root/reference/test/raasch_s4_{m}.inp", nnpe=4)
        order = np.argsort(ids, kind="stable")
        ids, xyz = ids[order], xyz[order]
        used = np.isin(ids, conn)
        new = np.zeros(ids.max() + 1, dtype=np.int64)
        new[ids[used]] = np.arange(1, used.sum() + 1)
        arrays[f"xyz_{m}"], arrays[f"conn_{m}"] = xyz[used], new[conn]
        print("raasch", m, xyz[used].shape, conn.shape)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "raasch_meshes.npz"), **arrays)
