"""Extracts the meshes of two of the reference's tests from their Abaqus input decks into small fixtures, as the reference's
`import_ABAQUS` + `compactnodes` deliver them (node coordinates in file order with the unconnected nodes removed, triangle
connectivities renumbered, 1-based):
  tests/golden/le5_mesh.npz          NAFEMS LE5 Z-section cantilever, test/test_shell_statics.jl:440-535 (nle5xf3c.inp)
  tests/golden/barrelvault_mesh.npz  irregular barrel vault of the resultants test, test/test_shell_statics.jl:577-764
                                     (barrelvault_stri3_irreg.inp; two coincident node pairs are left for `mergenodes`)
Run in the build container (the reference tree is not available on the GPU box):  python tests/golden/make_le5_fixture.py"""
import os

import numpy as np

DECKS = {"le5_mesh.npz": "/root/reference/test/nle5xf3c.inp", "barrelvault_mesh.npz": "/root/reference/test/barrelvault_stri3_irreg.inp"}


def parse(path):
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
            conn.append([int(x) for x in v[1:4]])
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
