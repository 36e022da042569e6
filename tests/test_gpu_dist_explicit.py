"""Row-partitioned explicit loop (fsgpu_explicit_create_dist: peer-mapped windows, halo entries pushed by the
step kernel) against the single-context run of the SAME global mesh.  The ranks here are contexts of one process
on ONE GPU, driven by one host thread each, so the driver's single-GPU box runs the whole exchange protocol
(flags, pushes, barriers, all-reduce); tests/mgpu_dist_explicit_check.py is the one-process-per-GPU form."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _global_problem(nx, ny, rcm=True):
    import fsb200
    from fsb200 import partition as pt
    from fsb200 import workloads as wl

    w = wl.c4_t3ff_panel(nx, ny)
    if rcm:  # numberdofs!(dchi, perm) with the RCM permutation (plate_expl_examples.jl:119-121,147)
        perm = pt.rcm_permutation(w["conn"], w["xyz"].shape[0])
        fixed = w["dofnums"] > w["nfree"]
        w["dofnums"], w["nfree"] = wl.number_dofs(fixed, perm)
    return w


def _femm(w, conn, xyz, dofnums, nfree, normals=None, valid=None):
    import fsb200

    f = fsb200.femm
    femm = f.FEMMShellT3FF(f.IntegDomain(conn, None, w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]), device=0)
    g = f.NodalField.__new__(f.NodalField)
    g.values = np.asfortranarray(xyz)
    d = f.NodalField.__new__(f.NodalField)
    d.values, d.dofnums, d._nfree = None, np.asfortranarray(dofnums), int(nfree)
    if normals is None:
        f.associategeometry(femm, g)
    else:
        femm._normals, femm._normal_valid, femm._associatedgeometry = np.asfortranarray(normals), valid, True
    return femm, g, d


def _load(w):
    F0 = np.zeros(w["nfree"])
    wd = w["dofnums"][:, 2]
    wd = wd[wd <= w["nfree"]] - 1
    F0[wd] = 1.0 + 0.25 * np.sin(np.arange(wd.size))
    return F0


@pytest.mark.parametrize("world,rcm", [(2, True), (3, True), (4, False)])
def test_row_partitioned_explicit_matches_single_context(world, rcm):
    import fsb200
    from fsb200 import partition as pt

    f = fsb200.femm
    nx, ny, nsteps = 30, 22, 80
    w = _global_problem(nx, ny, rcm)
    dt, cs = 2.0e-7, 50.0
    F0 = _load(w)
    fscale = 1.0 + 0.1 * np.cos(0.3 * np.arange(nsteps))

    # single context, global mesh
    fg, gg, dg = _femm(w, w["conn"], w["xyz"], w["dofnums"], w["nfree"])
    f.stiffness(fg, f.SysmatAssemblerFFBlock(), gg, None, None, dg)
    fg.ctx.shell_mass_diag(fg._params(), 3, nfree_only=True)
    exg = fsb200.Explicit(fg.ctx, c_scale=cs, dt=dt)
    lam_g = exg.omega_max_sq(12)
    exg.set_load(F0)
    exg.start(fscale[0])
    exg.step(nsteps // 2, fscale[: nsteps // 2])
    exg.step(nsteps - nsteps // 2, fscale[nsteps // 2 :])
    Ug, Vg, Ag = exg.get_state()
    ke_g = exg.kinetic_energy()
    xs = np.cos(np.arange(w["nfree"]) * 0.37)
    ys_g = exg.spmv(xs)

    # `world` contexts, each with the elements touching its own rows
    plans = [pt.ColumnBlockPlan(w["conn"], w["dofnums"], w["nfree"], "ffblock", r, world) for r in range(world)]
    exs, keep = [], []
    for r, plan in enumerate(plans):
        fr, gr, dr = _femm(w, plan.conn, plan.restrict_nodes(w["xyz"]), plan.dofnums, plan.nfree,
                           plan.restrict_nodes(fg._normals), plan.restrict_nodes(fg._normal_valid))
        f.stiffness(fr, f.SysmatAssemblerFFBlock(), gr, None, None, dr)
        fr.ctx.shell_mass_diag(fr._params(), 3, nfree_only=True)
        ex = fsb200.Explicit.create_dist(fr.ctx, r, world, plan.lcol_lo, plan.lcol_hi, plan.loc2glob[: plan.nfree], plan._bounds,
                                         c_scale=cs, dt=dt)
        exs.append(ex)
        keep.append(fr)
    pt.connect_local(exs)
    b = plans[0]._bounds
    info = [e.dist_info() for e in exs]
    assert sum(i[0] for i in info) == w["nfree"]
    assert all(i[1] > 0 and i[2] > 0 and i[4] >= 1 for i in info)  # every rank has halo entries and neighbours
    assert sum(i[1] for i in info) == sum(i[2] for i in info)  # every halo entry is pushed by exactly one owner

    def run(r):
        ex = exs[r]
        lam = ex.omega_max_sq(12)
        ex.set_load(F0[b[r] : b[r + 1]])
        ex.start(fscale[0])
        ex.step(nsteps // 2, fscale[: nsteps // 2])
        ex.step(nsteps - nsteps // 2, fscale[nsteps // 2 :])
        U, V, A = ex.get_state()
        ke = ex.kinetic_energy()
        y = ex.spmv(xs[b[r] : b[r + 1]])
        return lam, U, V, A, ke, y

    res = pt.run_collective([lambda r=r: run(r) for r in range(world)])
    U = np.concatenate([x[1] for x in res])
    V = np.concatenate([x[2] for x in res])
    ys = np.concatenate([x[5] for x in res])
    rel = lambda a, c: np.linalg.norm(a - c) / np.linalg.norm(c)
    assert rel(ys, ys_g) < 1e-13, rel(ys, ys_g)
    assert rel(U, Ug) < 1e-11 and rel(V, Vg) < 1e-11, (rel(U, Ug), rel(V, Vg))  # north star: <= 1e-9
    # global reductions: the same value (bit for bit) on every rank, equal to the single-context one
    assert len({x[0] for x in res}) == 1 and len({x[4] for x in res}) == 1
    assert abs(res[0][0] - lam_g) < 1e-10 * lam_g
    assert abs(res[0][4] - ke_g) < 1e-10 * abs(ke_g)
    pt.run_collective([e.close for e in exs])
    exg.close()


def test_row_partitioned_explicit_world_one_and_errors():
    import fsb200
    from fsb200 import partition as pt

    f = fsb200.femm
    w = _global_problem(12, 10, rcm=False)
    fg, gg, dg = _femm(w, w["conn"], w["xyz"], w["dofnums"], w["nfree"])
    f.stiffness(fg, f.SysmatAssemblerFFBlock(), gg, None, None, dg)
    fg.ctx.shell_mass_diag(fg._params(), 3, nfree_only=True)
    n = w["nfree"]
    ref = fsb200.Explicit(fg.ctx, c_scale=10.0, dt=1e-7)
    one = fsb200.Explicit.create_dist(fg.ctx, 0, 1, 0, n, None, [0, n], c_scale=10.0, dt=1e-7)
    F0 = _load(w)
    for e in (ref, one):
        e.set_load(F0)
        e.start(1.0)
        e.step(25)
    assert np.array_equal(ref.get_state()[0], one.get_state()[0])  # same kernel arithmetic, same order
    one.close()
    # bad row bounds / not connected
    with pytest.raises(fsb200.FsgpuError):
        fsb200.Explicit.create_dist(fg.ctx, 0, 2, 0, n, None, [0, n // 2, n])
    half = fsb200.Explicit.create_dist(fg.ctx, 0, 2, 0, 6 * (n // 12), None, [0, 6 * (n // 12), n])
    with pytest.raises(fsb200.FsgpuError):
        half.step(1)
    half.close()
    ref.close()
