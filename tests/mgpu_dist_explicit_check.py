"""Launched by torchrun (one rank per GPU) from tests/test_gpu_multi.py or by hand:
the row-partitioned explicit loop (fsgpu_explicit_create_dist: cudaIpc-mapped windows, halo entries written by the
step kernel over NVLink) must reproduce the single-GPU run of the same global, RCM-numbered mesh."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fsb200
from fsb200 import partition as pt
from fsb200 import workloads as wl

f = fsb200.femm
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
NX, NY, NSTEPS = int(os.environ.get("FS_NX", 120)), int(os.environ.get("FS_NY", 80)), 100

w = wl.c4_t3ff_panel(NX, NY)
perm = pt.rcm_permutation(w["conn"], w["xyz"].shape[0])
w["dofnums"], w["nfree"] = wl.number_dofs(w["dofnums"] > w["nfree"], perm)


def fields(xyz, dofnums, nfree):
    g = f.NodalField.__new__(f.NodalField)
    g.values = np.asfortranarray(xyz)
    d = f.NodalField.__new__(f.NodalField)
    d.values, d.dofnums, d._nfree = None, np.asfortranarray(dofnums), int(nfree)
    return g, d


mat = f.MatDeforElastIso(w["E"], w["nu"], w["rho"])
# single-GPU reference on the global mesh (every rank computes it: also the source of the nodal normals)
fg = f.FEMMShellT3FF(f.IntegDomain(w["conn"], None, w["thickness"]), mat, device=lr)
gg, dg = fields(w["xyz"], w["dofnums"], w["nfree"])
f.associategeometry(fg, gg)
f.stiffness(fg, f.SysmatAssemblerFFBlock(), gg, None, None, dg)
fg.ctx.shell_mass_diag(fg._params(), 3, nfree_only=True)
dt, cs = 2.0e-7, 50.0
F0 = np.zeros(w["nfree"])
wd = w["dofnums"][:, 2]
F0[wd[wd <= w["nfree"]] - 1] = 1.0
exg = fsb200.Explicit(fg.ctx, c_scale=cs, dt=dt)
lam_g = exg.omega_max_sq(10)
exg.set_load(F0)
exg.start(1.0)
exg.step(NSTEPS)
Ug = exg.get_state()[0]
ke_g = exg.kinetic_energy()
exg.close()

plan = pt.ColumnBlockPlan(w["conn"], w["dofnums"], w["nfree"], "ffblock", rank, world)
fr = f.FEMMShellT3FF(f.IntegDomain(plan.conn, None, w["thickness"]), mat, device=lr)
fr._normals, fr._normal_valid, fr._associatedgeometry = np.asfortranarray(plan.restrict_nodes(fg._normals)), plan.restrict_nodes(fg._normal_valid), True
gr, dr = fields(plan.restrict_nodes(w["xyz"]), plan.dofnums, plan.nfree)
f.stiffness(fr, f.SysmatAssemblerFFBlock(), gr, None, None, dr)
fr.ctx.shell_mass_diag(fr._params(), 3, nfree_only=True)
ex = fsb200.Explicit.create_dist(fr.ctx, rank, world, plan.lcol_lo, plan.lcol_hi, plan.loc2glob[: plan.nfree], plan._bounds, c_scale=cs, dt=dt)
pt.connect_ranks(ex)
b = plan._bounds
lam = ex.omega_max_sq(10)
ex.set_load(F0[b[rank] : b[rank + 1]])
ex.start(1.0)
ex.step(NSTEPS // 2)
ex.step(NSTEPS - NSTEPS // 2)
U = ex.get_state()[0]
ke = ex.kinetic_energy()
ref = Ug[b[rank] : b[rank + 1]]
err = np.abs(U - ref).max() / np.abs(Ug).max()
print(f"rank {rank}: dist_info={ex.dist_info()} rel.err U = {err:.3e}  omega^2 {lam:.12e} vs {lam_g:.12e}  KE {ke:.12e} vs {ke_g:.12e}", flush=True)
good = err < 1e-11 and abs(lam - lam_g) < 1e-10 * lam_g and abs(ke - ke_g) < 1e-10 * abs(ke_g)
ok = torch.tensor([1.0 if good else 0.0], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
ex.close()
dist.destroy_process_group()
sys.exit(0 if ok.item() == 1.0 else 1)
