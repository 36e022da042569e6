"""Extracts the mesh of the reference's NAFEMS LE5 test (Z-section cantilever, test/test_shell_statics.jl:440-535) from its
Abaqus input deck into a small fixture, tests/golden/le5_mesh.npz, as the reference's `import_ABAQUS` + `compactnodes`
deliver it: node coordinates in file order with the unconnected nodes removed, S3R connectivities renumbered (1-based).
Run in the build container (the reference tree is not available on the GPU box):  python tests/golden/make_le5_fixture.py"""
import os

import numpy as np

SRC = "/root/reference/test/nle5xf3c.inp"


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
    ids, xyz, conn = parse(SRC)
    order = np.argsort(ids, kind="stable")  # import_ABAQUS stores node k at row k (ids are ascending in the deck)
    ids, xyz = ids[order], xyz[order]
    used = np.isin(ids, conn)  # findunconnnodes / compactnodes: connected nodes keep their relative order
    new = np.zeros(ids.max() + 1, dtype=np.int64)
    new[ids[used]] = np.arange(1, used.sum() + 1)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "le5_mesh.npz")
    np.savez_compressed(out, xyz=xyz[used], conn=new[conn])
    print("wrote", out, xyz[used].shape, conn.shape)
