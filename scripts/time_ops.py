"""Step time (value-array clearing + element kernel + status read-back) and kernel time of the matrix operators
on the bench workloads.  usage: python scripts/time_ops.py [q4rs|t3ff|t3rcm|c3|beam ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fsb200
from fsb200 import partition as pt
from fsb200 import workloads as wl

f = fsb200.femm
import subprocess, threading
_smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "200"],
                        stdout=subprocess.PIPE, text=True)
_rows = []
threading.Thread(target=lambda: [_rows.append(l.strip()) for l in _smi.stdout], daemon=True).start()


def clocks():
    r = _rows[-8:]
    del _rows[:]
    return " | ".join(sorted(set(r)))



def field(values=None, dofnums=None, nfree=0):
    x = f.NodalField.__new__(f.NodalField)
    x.values, x.dofnums, x._nfree = values, dofnums, nfree
    return x


def timed(ctx, fn, reps=30):
    for _ in range(3):
        fn()
    ctx.sync()
    t0 = time.perf_counter()
    k = []
    for _ in range(reps):
        fn()
        k.append(ctx.last_kernel_ms)
    ctx.sync()
    return (time.perf_counter() - t0) / reps * 1e3, float(np.mean(k))


for which in sys.argv[1:] or ["q4rs", "t3ff", "t3rcm", "c3", "beam"]:
    if which == "beam":
        w = wl.c5_beam_lattice(69)
        sc = w["sections"]
        secs = f.FESetL2Beam(sc["A"], sc["I1"], sc["I2"], sc["I3"], sc["J"], sc["A2s"], sc["A3s"], sc["x1x2"])
        bf = f.FEMMCorotBeam(f.IntegDomain(w["conn"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]), secs)
        bf._sync_mesh(field(w["xyz"]))
        bf._startassembly(f.SysmatAssemblerFFBlock(), field(None, w["dofnums"], w["nfree"]))
        bf.ctx.set_state(w["u1"], w["Rfield1"])
        bp = bf._params()
        for op in ("stiffness", "geostiffness"):
            s, k = timed(bf.ctx, lambda: bf.ctx.beam_op(op, bp))
            print(f"beam {op}: step {s:.3f} ms kernel {k:.3f} ms path {bf.ctx.scatter_path}  [{clocks()}]", flush=True)
        bf.ctx.close()
        continue
    if which == "q4rs":
        w = wl.c2_q4rs_plate(1000)
        femm = f.FEMMShellQ4RS(f.IntegDomain(w["conn"], f.GaussRule2x2(), w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]))
    elif which == "c3":
        w = wl.c3_t3ffcomp_cylinder(1000, 1000)
        mat = f.lamina_material(*w["lamina"])
        t = w["thickness"]
        layup = f.CompositeLayup("C3", [f.Ply(f"p{k}", mat, t / 4, a) for k, a in enumerate(w["angles"])], wl.cylindrical_csys)
        femm = f.FEMMShellT3FFComp(f.IntegDomain(w["conn"], None, t), layup)
    else:
        w = wl.c4_t3ff_panel(2000, 1000)
        if which == "t3rcm":
            perm = pt.rcm_permutation(w["conn"], w["xyz"].shape[0])
            w["dofnums"], w["nfree"] = wl.number_dofs(w["dofnums"] > w["nfree"], perm)
        femm = f.FEMMShellT3FF(f.IntegDomain(w["conn"], None, w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]))
    f.associategeometry(femm, field(w["xyz"]))
    femm._startassembly(f.SysmatAssemblerFFBlock(), field(None, w["dofnums"], w["nfree"]))
    femm._sync_stab()
    p = femm._params()
    op = femm._opname + "_stiffness"
    s, k = timed(femm.ctx, lambda: femm.ctx.shell_op(op, p))
    print(f"{which}: step {s:.3f} ms kernel {k:.3f} ms  nelem {w['conn'].shape[0]} path {femm.ctx.scatter_path}  [{clocks()}]", flush=True)
    femm.ctx.close()
