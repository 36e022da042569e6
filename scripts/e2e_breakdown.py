"""Phase-by-phase wall time of one end-to-end Q4RS stiffness operator call (C2, 1M elements)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import fsb200
from bench import pin_copy, pinned
from fsb200 import workloads as wl

w = wl.c2_q4rs_plate(int(sys.argv[1]) if len(sys.argv) > 1 else 1000)
f = fsb200.femm
xyz_p, _a = pin_copy(w["xyz"])
conn_p, _b = pin_copy(np.ascontiguousarray(w["conn"]))
dof_p, _c = pin_copy(w["dofnums"])
femm = f.FEMMShellQ4RS(f.IntegDomain(conn_p, f.GaussRule2x2(), w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]))
geom0 = f.NodalField.__new__(f.NodalField)
geom0.values = xyz_p
f.associategeometry(femm, geom0)
ctx = femm.ctx
L = fsb200._lib
nr, nc, nnz = 0, 0, 0


def T(name, fn):
    ctx.sync()
    t0 = time.perf_counter()
    r = fn()
    ctx.sync()
    print(f"  {name:28s} {(time.perf_counter() - t0) * 1e3:8.2f} ms")
    return r


for rep in range(3):
    print("rep", rep)
    T("set_mesh", lambda: ctx.set_mesh(conn_p, xyz_p))
    T("set_rule+thickness", lambda: (ctx.set_rule(*f.GaussRule2x2()), ctx.set_thickness(w["thickness"])))
    T("set_normals", lambda: ctx.set_normals(femm._normals, femm._normal_valid))
    T("set_dofnums", lambda: ctx.set_dofnums(dof_p, w["nfree"]))
    nr, nc, nnz = T("symbolic", lambda: ctx.symbolic(L.FFBLOCK))
    if rep == 0:
        cp_p, _d = pinned((nc + 1,), np.int64)
        rv_p, _e = pinned((nnz,), np.int64)
        nz_p, _g = pinned((nnz,), np.float64)
    T("set_stab_factor", lambda: femm._sync_stab())
    T("numeric", lambda: ctx.shell_op("q4rs_stiffness", femm._params()))
    T("fetch colptr", lambda: fsb200.context.check(fsb200.context.lib.fsgpu_fetch_matrix(ctx._h, fsb200.context.ptr(cp_p), None, None)))
    T("fetch rowval", lambda: fsb200.context.check(fsb200.context.lib.fsgpu_fetch_matrix(ctx._h, None, fsb200.context.ptr(rv_p), None)))
    T("fetch nzval", lambda: ctx.fetch_values(nz_p))

# the bench's e2e step, verbatim, on the default stream and on a torch stream
u0 = R0 = None


def e2e_step():
    g = f.NodalField.__new__(f.NodalField)
    g.values = xyz_p
    d = f.NodalField.__new__(f.NodalField)
    d.values, d.dofnums, d._nfree = None, dof_p, w["nfree"]
    femm.reset_uploads()
    return f.stiffness(femm, f.SysmatAssemblerFFBlock(), g, u0, R0, d, out=(cp_p, rv_p, nz_p))


import cProfile
import pstats

for label in ("default stream", "torch stream"):
    if label == "torch stream":
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
    e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    print(label, "e2e_step", (time.perf_counter() - t0) / 3 * 1e3, "ms")
pr = cProfile.Profile()
pr.enable()
e2e_step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
