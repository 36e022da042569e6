"""Launched by torchrun (>= 2 ranks, one GPU each) from tests/test_gpu_multi.py: element-partitioned
assembly with the owned-column plan, gathered over NCCL into ONE global CSC on every rank, must equal the
single-GPU assembly of the global mesh (colptr / rowval bit-exact, values <= 1e-12)."""
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
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
ok = True


def field(values=None, dofnums=None, nfree=0):
    x = f.NodalField.__new__(f.NodalField)
    x.values, x.dofnums, x._nfree = values, dofnums, nfree
    return x


def run(kind, asm_name, w):
    global ok
    asm = {"ffblock": f.SysmatAssemblerFFBlock, "sparse": f.SysmatAssemblerSparse}[asm_name]
    mat = f.MatDeforElastIso(w["E"], w["nu"], w["rho"])
    mk = (lambda c: f.FEMMShellT3FF(f.IntegDomain(c, None, w["thickness"]), mat, device=rank)) if kind == "t3" else (
        lambda c: f.FEMMShellQ4RS(f.IntegDomain(c, f.GaussRule2x2(), w["thickness"]), mat, device=rank))
    # single-GPU answer on the global mesh (every rank computes it: it is the checker)
    fg = mk(w["conn"])
    gg, dg = field(w["xyz"]), field(None, w["dofnums"], w["nfree"])
    f.associategeometry(fg, gg)
    Kg = f.stiffness(fg, asm(), gg, None, None, dg)
    fg.ctx.shell_mass_diag(fg._params(), 3 if kind == "t3" else 4, nfree_only=(asm_name == "ffblock"))
    Mg = fg.ctx.fetch_vector(Kg.n)
    # this rank's share
    plan = pt.ColumnBlockPlan(w["conn"], w["dofnums"], w["nfree"], asm_name, rank, world)
    fl = mk(plan.conn)
    fl._normals, fl._normal_valid = np.asfortranarray(plan.restrict_nodes(fg._normals)), plan.restrict_nodes(fg._normal_valid)
    fl._associatedgeometry = True
    gl, dl = field(np.asfortranarray(plan.restrict_nodes(w["xyz"]))), field(None, plan.dofnums, plan.nfree)
    f.stiffness(fl, asm(), gl, None, None, dl)
    K = pt.gather_matrix(fl.ctx, plan, dev)
    fl.ctx.shell_mass_diag(fl._params(), 3 if kind == "t3" else 4, nfree_only=(asm_name == "ffblock"))
    vp, vn = fl.ctx.vector_device()
    M = pt.gather_vector(torch.as_tensor(pt.DevicePointer(vp, vn), device=dev), plan, dev).cpu().numpy()
    e_pat = np.array_equal(K.colptr, Kg.colptr) and np.array_equal(K.rowval, Kg.rowval)
    e_val = np.linalg.norm(K.nzval - Kg.nzval) / np.linalg.norm(Kg.nzval)
    e_m = np.linalg.norm(M - Mg) / np.linalg.norm(Mg)
    print(f"rank {rank}: {kind}/{asm_name}: {len(plan.elems)}/{w['conn'].shape[0]} elements, columns [{plan.col_lo},{plan.col_hi}) of "
          f"{plan.ncols_global}; pattern {'ok' if e_pat else 'DIFFERS'}, values {e_val:.2e}, lumped mass {e_m:.2e}", flush=True)
    ok = ok and e_pat and e_val < 1e-12 and e_m < 1e-13 and len(plan.elems) < w["conn"].shape[0]


run("t3", "ffblock", wl.c4_t3ff_panel(60, 40))
run("q4", "sparse", wl.c2_q4rs_plate(36))
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
