"""Debug driver: two row-partitioned ranks as contexts of one process on one GPU (threads)."""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fsb200
from fsb200 import partition as pt
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_dist_explicit as T

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
f = fsb200.femm
w = T._global_problem(30, 22, True)
fg, gg, dg = T._femm(w, w["conn"], w["xyz"], w["dofnums"], w["nfree"])
f.stiffness(fg, f.SysmatAssemblerFFBlock(), gg, None, None, dg)
fg.ctx.shell_mass_diag(fg._params(), 3, nfree_only=True)
plans = [pt.ColumnBlockPlan(w["conn"], w["dofnums"], w["nfree"], "ffblock", r, world) for r in range(world)]
exs, keep = [], []
for r, plan in enumerate(plans):
    fr, gr, dr = T._femm(w, plan.conn, plan.restrict_nodes(w["xyz"]), plan.dofnums, plan.nfree, plan.restrict_nodes(fg._normals), plan.restrict_nodes(fg._normal_valid))
    f.stiffness(fr, f.SysmatAssemblerFFBlock(), gr, None, None, dr)
    fr.ctx.shell_mass_diag(fr._params(), 3, nfree_only=True)
    exs.append(fsb200.Explicit.create_dist(fr.ctx, r, world, plan.lcol_lo, plan.lcol_hi, plan.loc2glob[: plan.nfree], plan._bounds, c_scale=50.0, dt=2e-7))
    keep.append(fr)
    print("rank", r, "info", exs[-1].dist_info(), flush=True)
pt.connect_local(exs)
print("connected", flush=True)
b = plans[0]._bounds
xs = np.cos(np.arange(w["nfree"]) * 0.37)

def run(r):
    try:
        t0 = time.time()
        y = exs[r].spmv(xs[b[r]:b[r+1]])
        print(r, "spmv ok", time.time() - t0, flush=True)
        lam = exs[r].omega_max_sq(3)
        print(r, "omega ok", lam, flush=True)
        exs[r].start(1.0)
        exs[r].step(10)
        print(r, "step ok", flush=True)
        return y
    except Exception:
        print("rank", r, traceback.format_exc(), flush=True)
        raise

print(pt.run_collective([lambda r=r: run(r) for r in range(world)])[0][:3])
