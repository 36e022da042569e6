"""Launched by torchrun (2 ranks, one GPU each) from tests/test_gpu_multi.py: the element-
partitioned explicit loop (strip partition + NCCL interface exchange) must reproduce the
single-GPU run of the same global mesh."""
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
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
NX, NY, NSTEPS = 24, 10, 60


def setup(w, device):
    femm = f.FEMMShellT3FF(f.IntegDomain(w["conn"], None, w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]), device=device)
    g = f.NodalField.__new__(f.NodalField)
    g.values = w["xyz"]
    d = f.NodalField.__new__(f.NodalField)
    d.values, d.dofnums, d._nfree = None, w["dofnums"], w["nfree"]
    f.associategeometry(femm, g)
    return femm, g, d


# nodal normals must come from the GLOBAL mesh (interface nodes see elements of both strips):
# build the global mesh on every rank, compute normals there, hand the local slice to the strip.
wg = wl.c4_strip(0, 1, NX, NY * world, Ly=1.0 * world)  # global panel = `world` strips stacked
fg, gg, dg = setup(wg, rank)
nodes_per_row = NX + 1
lo_node = rank * NY * nodes_per_row
w = wl.c4_strip(rank, world, NX, NY)
nloc = w["xyz"].shape[0]
assert np.allclose(wg["xyz"][lo_node : lo_node + nloc], w["xyz"])
femm = f.FEMMShellT3FF(f.IntegDomain(w["conn"], None, w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]), device=rank)
g = f.NodalField.__new__(f.NodalField)
g.values = w["xyz"]
d = f.NodalField.__new__(f.NodalField)
d.values, d.dofnums, d._nfree = None, w["dofnums"], w["nfree"]
# partitioned associategeometry!: interface-node normal sums and validity flags combined over NCCL
node_links = pt.strip_links(rank, world, w["lo_nodes"], w["hi_nodes"])
f.associategeometry(femm, g, interface=(node_links, torch.device("cuda", rank)))
assert np.abs(femm._normals - fg._normals[lo_node : lo_node + nloc]).max() < 1e-14, "partitioned nodal normals differ from the global ones"
assert np.array_equal(femm._normal_valid, fg._normal_valid[lo_node : lo_node + nloc])

stream = torch.cuda.Stream()
femm.ctx.set_stream(stream.cuda_stream)
with torch.cuda.stream(stream):
    K = f.stiffness(femm, f.SysmatAssemblerFFBlock(), g, None, None, d)
    femm.ctx.shell_mass_diag(femm._params(), 3, nfree_only=True)
    nf = w["nfree"]
    ex_if = pt.InterfaceExchange(pt.strip_links(rank, world, w["lo_dofs"], w["hi_dofs"]), torch.device("cuda", rank))
    import ctypes as C

    vp, vn = C.c_void_p(), C.c_int64()
    fsb200._lib.check(fsb200._lib.lib.fsgpu_vector_device(femm.ctx._h, C.byref(vp), C.byref(vn)))
    Mt = torch.as_tensor(pt.DevicePointer(vp.value, vn.value), device="cuda")
    ex_if.exchange_sum(Mt)
    torch.cuda.synchronize()
    dt, cs = 2.0e-7, 50.0
    ex = fsb200.Explicit(femm.ctx, c_scale=cs, dt=dt)
    # load: unit z-force on every free w dof of the GLOBAL vector, restricted to this strip; interface
    # dofs get the full nodal value on both ranks (they are replicas of the same global dof)
    F0 = np.zeros(nf)
    wd = w["dofnums"][:, 2]
    F0[wd[wd <= nf] - 1] = 1.0
    ex.set_load(F0)
    ex.start(1.0)
    E = ex.device_state()[3]
    Et = torch.as_tensor(pt.DevicePointer(E, nf), device="cuda")
    for _ in range(NSTEPS):
        ex.step_begin()
        ex_if.exchange_sum(Et)
        ex.step_end(1.0)
    torch.cuda.synchronize()
    Uloc = ex.get_state()[0]

# single-GPU reference on the global mesh
Kg = f.stiffness(fg, f.SysmatAssemblerFFBlock(), gg, None, None, dg)
fg.ctx.shell_mass_diag(fg._params(), 3, nfree_only=True)
exg = fsb200.Explicit(fg.ctx, c_scale=cs, dt=dt)
F0g = np.zeros(wg["nfree"])
wdg = wg["dofnums"][:, 2]
F0g[wdg[wdg <= wg["nfree"]] - 1] = 1.0
exg.set_load(F0g)
exg.start(1.0)
exg.step(NSTEPS)
Ug = exg.get_state()[0]
# map local free dofs to global free dofs through the node numbering
dl = w["dofnums"]
dgl = wg["dofnums"][lo_node : lo_node + nloc]
free = dl <= nf
assert np.array_equal(free, dgl <= wg["nfree"])
err = np.linalg.norm(Uloc[dl[free] - 1] - Ug[dgl[free] - 1]) / np.linalg.norm(Ug[dgl[free] - 1])
print(f"rank {rank}: multi-GPU explicit vs single-GPU rel.err = {err:.3e}", flush=True)
ok = torch.tensor([1.0 if err < 1e-11 else 0.0], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if ok.item() == 1.0 else 1)
