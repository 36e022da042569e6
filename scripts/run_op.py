"""Run one operator a few times on a named synthetic workload (driver for ncu captures).
usage: python scripts/run_op.py {q4rs|t3ff|explicit} [n] [reps] [det]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fsb200
from fsb200 import workloads as wl

f = fsb200.femm
which = sys.argv[1] if len(sys.argv) > 1 else "q4rs"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
det = len(sys.argv) > 4 and sys.argv[4] == "det"
if which == "beam":
    w = wl.c5_beam_lattice(n if n < 200 else 69)
    sc = w["sections"]
    secs = f.FESetL2Beam(sc["A"], sc["I1"], sc["I2"], sc["I3"], sc["J"], sc["A2s"], sc["A3s"], sc["x1x2"])
    bf = f.FEMMCorotBeam(f.IntegDomain(w["conn"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]), secs)
    geom0 = f.NodalField.__new__(f.NodalField)
    geom0.values = w["xyz"]
    dchi = f.NodalField.__new__(f.NodalField)
    dchi.values, dchi.dofnums, dchi._nfree = None, w["dofnums"], w["nfree"]
    bf._sync_mesh(geom0)
    bf._startassembly(f.SysmatAssemblerFFBlock(), dchi)
    bf.ctx.set_state(w["u1"], w["Rfield1"])
    bp = bf._params()
    for op in ("stiffness", "geostiffness", "mass"):
        for _ in range(reps):
            bf.ctx.beam_op(op, bp)
            print("beam", op, "nelem", w["conn"].shape[0], "kernel ms", bf.ctx.last_kernel_ms, flush=True)
    sys.exit(0)
if which == "q4rs":
    w = wl.c2_q4rs_plate(n)
    femm = f.FEMMShellQ4RS(f.IntegDomain(w["conn"], f.GaussRule2x2(), w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]))
else:
    w = wl.c4_t3ff_panel(2 * n, n)
    femm = f.FEMMShellT3FF(f.IntegDomain(w["conn"], None, w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]))
geom0 = f.NodalField.__new__(f.NodalField)
geom0.values = w["xyz"]
dchi = f.NodalField.__new__(f.NodalField)
dchi.values, dchi.dofnums, dchi._nfree = None, w["dofnums"], w["nfree"]
if det:
    femm.ctx.set_deterministic(True)
f.associategeometry(femm, geom0)
femm._startassembly(f.SysmatAssemblerFFBlock(), dchi)
femm._sync_stab()
p = femm._params()
op = femm._opname + "_stiffness"
for _ in range(reps):
    femm.ctx.shell_op(op, p)
    print(which, "nelem", w["conn"].shape[0], "kernel ms", femm.ctx.last_kernel_ms, "path", femm.ctx.scatter_path, flush=True)
if which == "explicit":
    femm.ctx.shell_mass_diag(p, 3, nfree_only=True)
    ex = fsb200.Explicit(femm.ctx, c_scale=100.0, dt=1e-7)
    ex.set_load(np.ones(w["nfree"]))
    ex.start(1.0)
    import torch

    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ex.step(20)
        torch.cuda.synchronize()
        print("explicit ms/step", (time.perf_counter() - t0) / 20 * 1e3, flush=True)
