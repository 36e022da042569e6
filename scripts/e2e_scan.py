"""End-to-end C2 operator call (host arrays in, host CSC out) under different host-side settings of the fetch:
number of expanding threads (FSGPU_HOST_THREADS), row form (FSGPU_FETCH_MODE).  usage: python scripts/e2e_scan.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import fsb200
from bench import pin_copy, pinned
from fsb200 import workloads as wl

w = wl.c2_q4rs_plate(1000)
f = fsb200.femm
xyz_p, _a = pin_copy(w["xyz"])
conn_p, _b = pin_copy(np.ascontiguousarray(w["conn"]))
dof_p, _c = pin_copy(w["dofnums"])
femm = f.FEMMShellQ4RS(f.IntegDomain(conn_p, f.GaussRule2x2(), w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]))
geom0 = f.NodalField.__new__(f.NodalField)
geom0.values = xyz_p
f.associategeometry(femm, geom0)
d = f.NodalField.__new__(f.NodalField)
d.values, d.dofnums, d._nfree = None, dof_p, w["nfree"]
K = f.stiffness(femm, f.SysmatAssemblerFFBlock(), geom0, None, None, d)
nnz, nc = K.nzval.shape[0], K.colptr.shape[0]
_p = [pinned((nc,), np.int64), pinned((nnz,), np.int64), pinned((nnz,), np.float64)]  # (array, owning tensor)
out = tuple(a for a, _t in _p)


def step():
    global out
    femm.reset_uploads()
    return f.stiffness(femm, f.SysmatAssemblerFFBlock(), geom0, None, None, d, out=out)


def timed(label, n=4):
    step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{label:40s} e2e ms: min {min(ts):6.2f}  median {sorted(ts)[len(ts) // 2]:6.2f}", flush=True)


timed("default")
if "pageable" in sys.argv[1:]:
    pinned_out = out
    out = (np.empty(nc, np.int64), np.empty(nnz, np.int64), np.empty(nnz, np.float64))
    for a in out:
        a[...] = 0  # touch the pages
    timed("pageable destination (numpy arrays)")
    # pageable inputs as well: plain numpy copies of the mesh, dof numbers and (inside the FEMM) the normals
    geom0.values = np.array(xyz_p)
    d.dofnums = np.array(dof_p)
    femm.integdomain.conn = np.array(conn_p)
    timed("pageable destination and inputs")
    t0 = time.perf_counter()
    fsb200.context.check(fsb200.context.lib.fsgpu_fetch_matrix(femm.ctx._h, None, None, fsb200.context.ptr(out[2])))
    print(f"  values only, pageable: {(time.perf_counter() - t0) * 1e3:.1f} ms")
    t0 = time.perf_counter()
    fsb200.context.check(fsb200.context.lib.fsgpu_fetch_matrix(femm.ctx._h, None, fsb200.context.ptr(out[1]), None))
    print(f"  rows only, pageable: {(time.perf_counter() - t0) * 1e3:.1f} ms")
    assert np.array_equal(out[1], pinned_out[1]) and np.array_equal(out[0], pinned_out[0])
    assert np.abs(out[2] - pinned_out[2]).max() <= 1e-12 * np.abs(pinned_out[2]).max()  # RED order differs from run to run
    sys.exit(0)
for nth in (2, 4, 6, 8, 12, 16, 24):
    os.environ["FSGPU_HOST_THREADS"] = str(nth)
    timed(f"FSGPU_HOST_THREADS={nth}")
os.environ.pop("FSGPU_HOST_THREADS")
os.environ["FSGPU_FETCH_MODE"] = "entries"
timed("FSGPU_FETCH_MODE=entries")
os.environ.pop("FSGPU_FETCH_MODE")
for k, v in (("FSGPU_VALUE_PIECE_MB", "16"), ("FSGPU_VALUE_PIECE_MB", "256"), ("FSGPU_VALUE_INFLIGHT", "4")):
    os.environ[k] = v
    timed(f"{k}={v}")
    os.environ.pop(k)
