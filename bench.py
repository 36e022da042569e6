#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric).

Workload at N=1: config[1] of BASELINE.json -- Q4RS homogeneous square plate, synthetic
1000 x 1000 quad mesh (1M elements), stiffness assembly to CSC (SysmatAssemblerFFBlock).
A "step" is one stiffness operator call over the whole mesh.

  value   element matrices assembled per second, device-resident inputs, pattern reused
          (numeric phase: value-array clear + element kernel + status read-back)
  e2e     the same operator through the reference-facing API with HOST (pinned) buffers:
          mesh/dofs/normals H2D + symbolic phase + numeric phase + CSC D2H, every step
  roofline / fp64 / cpu_baseline: see DESIGN.md "Measurement"

`--impl reference` times the CPU restatement of the reference algorithm (oracle/cport,
"port": no Julia in this image) on the box's host cores with all threads.

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
       torchrun ... bench.py --gpus N ...      (one rank per GPU, weak scaling)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic traffic / work per element of the dominant kernel
Q4_IN_BYTES = 32 + 97  # conn (4 x Int64 on the Julia side) + 1 node x (xyz 24 + normal 24 + valid 1 + dofnums 48)
Q4_FLOPS = 43.0e3


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        """rows delivered so far (GPU idle, sampler starting up) do not count"""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = self.rows[self.first:] or self.rows
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def load_workloads():
    """The pure-numpy workload generators of the package, imported by file: the CPU reference arm must not load
    libfsgpu.so (importing the package does)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("fsb200_workloads", os.path.join(ROOT, "finetoolsflexstructures.jl_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def dist_env():
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), ws


def pinned(shape, dtype, order="C"):
    import torch

    n = int(np.prod(shape))
    t = torch.empty(n, dtype={np.float64: torch.float64, np.int64: torch.int64, np.uint8: torch.uint8}[dtype], pin_memory=True)
    a = t.numpy().reshape(shape, order=order)
    return a, t


def pin_copy(a):
    order = "F" if a.flags.f_contiguous and not a.flags.c_contiguous else "C"
    out, keep = pinned(a.shape, a.dtype.type, order)
    out[...] = a
    return out, keep


# ---------------------------------------------------------------------------------------
# CPU reference arm (the oracle's C port of the reference algorithm)
# ---------------------------------------------------------------------------------------
def load_refport():
    d = os.path.join(ROOT, "oracle", "cport")
    so = os.path.join(d, "librefport.so")
    try:  # rebuild on this box (host CPU may differ from the build container)
        subprocess.run(["make", "-C", d, "-s", "-B"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        pass
    lib = C.CDLL(so)
    lib.ref_coo_to_csc.restype = C.c_int64
    return lib


_CPU_BUF = {}


def cpu_q4rs_assembly(w, nelem_sample, nthreads, normals, valid):
    """Times the reference algorithm on the first `nelem_sample` elements of the workload: the element loop with
    its COO append (OpenMP element-parallel; the reference's loop is serial) and ONE pass of the COO->CSC
    conversion (serial, like Julia's `sparse`).  All buffers are allocated and touched before the timers start.
    Returns elements/s for the whole and for the element loop alone."""
    from oracle import fe_external as fx
    from oracle import shells as osh

    lib = load_refport()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    conn = np.ascontiguousarray(w["conn"][:nelem_sample])
    nn = conn.shape[1]
    n = 6 * nn
    Dps, Dt = osh.shell_material_stiffness(fx.moduli_iso(w["E"], w["nu"]))
    Dps, Dt = np.ascontiguousarray(Dps), np.ascontiguousarray(Dt)
    pc, wt = fx.gauss_rule_2x2()
    pc = np.ascontiguousarray(pc)
    nt = nelem_sample * n * n
    nall, nfree = w["dofnums"].size, w["nfree"]
    key = (nt, nfree)
    if key not in _CPU_BUF:
        _CPU_BUF.clear()
        _CPU_BUF[key] = (np.zeros(nt, np.int64), np.zeros(nt, np.int64), np.zeros(nt), np.zeros(nfree + 1, np.int64),
                         np.zeros(nt, np.int64), np.zeros(nt))  # zeros: the pages are touched outside the timers
    I, J, V, cp, rv, nz = _CPU_BUF[key]
    v8 = np.ascontiguousarray(valid.astype(np.uint8))
    nF = np.asfortranarray(normals)
    nnodes = w["xyz"].shape[0]
    alpha = 0.1 if nn == 4 else 5 / 12 / 1.5
    t0 = time.perf_counter()
    lib.ref_shell_stiffness_coo(nn, C.c_int64(nelem_sample), P(conn), C.c_int64(nnodes), P(w["xyz"]), P(nF), P(v8),
                                P(w["dofnums"]), P(Dps), P(Dt), C.c_double(w["thickness"]), C.c_double(alpha), C.c_double(1.0), 4,
                                P(pc), P(wt), nthreads, P(I), P(J), P(V))
    t1 = time.perf_counter()
    # one pass: the outputs are sized for the worst case (every triple its own entry), so no counting pre-pass
    lib.ref_coo_to_csc(C.c_int64(nt), P(I), P(J), P(V), C.c_int64(nall), C.c_int64(nall), C.c_int64(nfree), C.c_int64(nfree), P(cp), P(rv), P(nz))
    t2 = time.perf_counter()
    return {"value": nelem_sample / (t2 - t0), "seconds": t2 - t0, "loop_s": t1 - t0, "sparse_s": t2 - t1,
            "loop_value": nelem_sample / (t1 - t0)}


def oracle_normals(w):
    """Nodal normals for the CPU arm (vectorised oracle; not timed)."""
    from oracle import shells as osh

    f = osh.q4rs_associategeometry if w["kind"] == "q4" else osh.t3ff_associategeometry
    return f(np.ascontiguousarray(w["xyz"]), w["conn"])


# ---------------------------------------------------------------------------------------
# Secondary workloads reported in the same JSON line ("other_workloads"): BASELINE configs[3]
# (explicit T3FF shell, 4M-element panel, element-partitioned) -- assembly rate and steps/s.
# ---------------------------------------------------------------------------------------
T3_IN_BYTES, T3_FLOPS = 73.0, 9.0e3
EXPL_FLOPS_PER_NNZ = 2.0


def explicit_c4(args, rank, local_rank, world, stream, hbm_peak, fp64_peak):
    """BASELINE configs[3]: ONE T3FF panel (4M elements at the default size), RCM-numbered as the reference example
    does (plate_expl_examples.jl:119-121,147), row-partitioned over the ranks (STRONG scaling): rank r owns a
    contiguous block of global rows and assembles every element that touches one of its nodes; the explicit loop
    exchanges halo displacements inside the step kernel (peer-mapped windows, no NCCL call on the step path)."""
    import torch
    import torch.distributed as dist

    import fsb200
    from fsb200 import partition as pt
    from fsb200 import workloads as wl

    f = fsb200.femm
    nx, ny = args.c4_nx, args.c4_nx // 2
    t0 = time.perf_counter()
    w = wl.c4_t3ff_panel(nx, ny)
    perm = pt.rcm_permutation(w["conn"], w["xyz"].shape[0])
    w["dofnums"], w["nfree"] = wl.number_dofs(w["dofnums"] > w["nfree"], perm)
    mesh_s = time.perf_counter() - t0
    nelem_g = w["conn"].shape[0]
    mat = f.MatDeforElastIso(w["E"], w["nu"], w["rho"])

    def field(values=None, dofnums=None, nfree=0):
        x = f.NodalField.__new__(f.NodalField)
        x.values, x.dofnums, x._nfree = values, dofnums, nfree
        return x

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nodal normals of the GLOBAL mesh on the device (halo nodes see elements this rank does not assemble)
    fg = f.FEMMShellT3FF(f.IntegDomain(w["conn"], None, w["thickness"]), mat, device=local_rank)
    fg.ctx.set_stream(stream.cuda_stream)
    geom_g = field(w["xyz"])
    f.associategeometry(fg, geom_g)
    ag_ms = None
    if world == 1:
        fg.ctx.associategeometry(fg.threshold_angle, None, True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            fg.ctx.associategeometry(fg.threshold_angle, None, True)
        barrier()
        ag_ms = (time.perf_counter() - t0) / 5 * 1e3
        plan = None
        femm, dchi = fg, field(None, w["dofnums"], w["nfree"])
        nelem = nelem_g
    else:
        t0 = time.perf_counter()
        plan = pt.ColumnBlockPlan(w["conn"], w["dofnums"], w["nfree"], "ffblock", rank, world)
        plan_s = time.perf_counter() - t0
        femm = f.FEMMShellT3FF(f.IntegDomain(plan.conn, None, w["thickness"]), mat, device=local_rank)
        femm.ctx.set_stream(stream.cuda_stream)
        femm._normals, femm._normal_valid = np.asfortranarray(plan.restrict_nodes(fg._normals)), plan.restrict_nodes(fg._normal_valid)
        femm._associatedgeometry = True
        fg.ctx.close()
        geom_l = field(np.asfortranarray(plan.restrict_nodes(w["xyz"])))
        dchi = field(None, plan.dofnums, plan.nfree)
        femm._sync_mesh(geom_l)
        nelem = int(len(plan.elems))
    t0 = time.perf_counter()
    femm._startassembly(f.SysmatAssemblerFFBlock(), dchi)
    femm.ctx.sync()
    sym_ms = (time.perf_counter() - t0) * 1e3
    femm._sync_stab()
    p = femm._params()

    # --- T3FF stiffness assembly (numeric phase, device resident; N > 1: the rank's share of the ONE mesh) ---
    for _ in range(3):
        femm.ctx.shell_op("t3ff_stiffness", p)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    e0.record(stream)
    nrep = max(3, args.steps // 2)
    for _ in range(nrep):
        femm.ctx.shell_op("t3ff_stiffness", p)
        kms.append(femm.ctx.last_kernel_ms)
    e1.record(stream)
    barrier()
    tm = torch.tensor([e0.elapsed_time(e1) / nrep, float(np.mean(kms))], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    asm_ms, kernel_ms = float(tm[0].item()), float(tm[1].item())
    nnz = femm.ctx.result_size()[2]
    t_fp64 = T3_FLOPS * nelem / (fp64_peak * 1e12) * 1e3
    alg = T3_IN_BYTES + 8.0 * nnz / nelem
    t_hbm = alg * nelem / (hbm_peak * 1e9) * 1e3
    asm = {"workload": f"T3FF stiffness -> CSC (FFBlock), ONE {nelem_g}-element panel over {world} rank(s) ({nelem} elements on this rank incl. interface elements)",
           "value": nelem_g / (asm_ms * 1e-3), "unit": "elements/s", "scaling": "strong", "ms_per_step": asm_ms, "kernel_ms": kernel_ms,
           "symbolic_ms": sym_ms, "nnz_this_rank": int(nnz), "associategeometry_ms": ag_ms,
           "roofline": {"bound": "fp64" if t_fp64 >= t_hbm else "hbm", "t_roof_ms": max(t_fp64, t_hbm), "frac": max(t_fp64, t_hbm) / asm_ms,
                        "frac_kernel_only": max(t_fp64, t_hbm) / kernel_ms, "basis": "this rank's elements, step = value-array clear + kernel",
                        "flops_per_element": T3_FLOPS, "peak_tflops": fp64_peak, "achieved_tflops": T3_FLOPS * nelem / (asm_ms * 1e-3) / 1e12,
                        "hbm": {"algorithmic_bytes_per_element": alg, "achieved_gbs": alg * nelem / (asm_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                "frac": t_hbm / asm_ms},
                        "kernel": "k_t3_stiffness<false,false,EmitRuns>"}}

    # --- explicit central differences: K = FF block (device resident), lumped M ---
    femm.ctx.shell_mass_diag(p, 3, nfree_only=True)
    cs = 100.0
    if world == 1:
        ex = fsb200.Explicit(femm.ctx, c_scale=cs, dt=0.0)
        b = np.array([0, w["nfree"]])
    else:
        ex = fsb200.Explicit.create_dist(femm.ctx, rank, world, plan.lcol_lo, plan.lcol_hi, plan.loc2glob[: plan.nfree], plan._bounds,
                                         c_scale=cs, dt=0.0)
        with torch.cuda.stream(stream):
            pt.connect_ranks(ex)
        b = plan._bounds
    lam = ex.omega_max_sq(args.power_its)  # row-partitioned: a global value, identical on every rank
    dt = 0.9 * 2 / np.sqrt(lam)
    ex.set_timestep(cs, dt)
    # load: a pressure patch on the nodes nearest the panel centre (w dofs), the same global vector on every rank
    F0g = np.zeros(w["nfree"])
    c = np.array([w["xyz"][:, 0].mean(), w["xyz"][:, 1].mean()])
    near = np.argsort((w["xyz"][:, 0] - c[0]) ** 2 + (w["xyz"][:, 1] - c[1]) ** 2)[: max(1, w["xyz"].shape[0] // 100)]
    wd = w["dofnums"][near, 2]
    F0g[wd[wd <= w["nfree"]] - 1] = 1.0
    ex.set_load(F0g[b[rank] : b[rank + 1]] if world > 1 else F0g)
    ex.start(1.0)
    nsteps = args.expl_steps
    ex.step(50)
    barrier()
    l0 = femm.ctx.launch_count
    e0.record(stream)
    ex.step(nsteps)
    e1.record(stream)
    barrier()
    launches = femm.ctx.launch_count - l0
    tm = torch.tensor([e0.elapsed_time(e1) / nsteps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    step_ms = float(tm.item())
    Uh = ex.get_state()[0]
    umax = torch.tensor([float(np.abs(Uh).max())], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(umax, op=dist.ReduceOp.MAX)
    ke = ex.kinetic_energy()
    nrows, knnz, nruns, idx_entries = ex.layout()
    info = ex.dist_info() if world > 1 else None
    # f64 value per entry + one i32 column index per entry of every distinct row pattern (runs of <= 6 rows share
    # one), ~10 vector passes.  SURVEY 8(d)'s figure for the plain Int32 CSR form is 12 B per entry.
    tot = torch.tensor([8.0 * knnz + 4.0 * idx_entries + 10 * 8.0 * nrows, 12.0 * knnz + 10 * 8.0 * nrows, float(knnz)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    bytes_step, bytes_step_csr, knnz_g = float(tot[0].item()), float(tot[1].item()), float(tot[2].item())
    expl = {"workload": f"explicit central differences, T3FF, ONE {nelem_g}-element RCM-numbered panel row-partitioned over {world} rank(s), SpMV form (K_ff CSR, lumped M)",
            "scaling": "strong", "steps_per_s": 1e3 / step_ms, "element_steps_per_s": nelem_g * 1e3 / step_ms, "ms_per_step": step_ms, "dt": dt,
            "omega_max": float(np.sqrt(lam)), "nsteps_timed": nsteps, "gpu_launches": int(launches), "max_abs_U": float(umax.item()),
            "kinetic_energy": ke, "nnz_global": int(knnz_g),
            "interface_exchange": ("halo displacements written by the fused step kernel into the neighbours' peer-mapped windows "
                                   "(cudaIpc over NVLink) + flag; boundary rows first, interior rows overlap the exchange; no NCCL on the step path")
            if world > 1 else "none",
            "partition": None if info is None else {"own_rows": info[0], "halo_entries": info[1], "pushed_entries_per_step": info[2],
                                                    "boundary_runs": info[3], "neighbours": info[4], "plan_host_s": plan_s},
            "roofline": {"bound": "hbm", "achieved": bytes_step / world / (step_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": bytes_step / world / (step_ms * 1e-3) / 1e9 / hbm_peak, "basis": "per GPU: the job's algorithmic bytes / ranks / step time",
                         "algorithmic_bytes_per_element_step": bytes_step / nelem_g,
                         "plain_csr_bytes_per_element_step": bytes_step_csr / nelem_g, "row_runs_this_rank": int(nruns),
                         "kernel": "k_spmv_step<dist>" if world > 1 else "k_spmv_step"},
            "mesh_and_rcm_host_s": mesh_s}
    # CPU arm for the explicit loop (rank 0, bounded sample: a 1/100-size strip, all host threads)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            cpu = cpu_explicit(local_rank)
        except Exception as ex_:  # never let the side measurement kill the bench line
            cpu = {"error": repr(ex_)}
    barrier()
    ex.close()
    femm.ctx.close()
    return {"t3ff_assembly_C4": asm, "explicit_C4": expl, "explicit_cpu_baseline": cpu}


def extras_c3_c5(args, rank, local_rank, world, stream, hbm_peak, fp64_peak):
    """BASELINE configs[2] (T3FFComp laminated cylinder, 2M triangles: stiffness + mass) and configs[4]
    (corotational beam lattice, 1.01M elements: restoringforce + stiffness + geostiffness = one Newton
    iteration's assembly).  Numeric phase, device-resident, max over ranks."""
    import torch
    import torch.distributed as dist

    import fsb200
    from fsb200 import workloads as wl

    f = fsb200.femm
    out = {}

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        tm = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm.item())

    def field(values=None, dofnums=None, nfree=0):
        x = f.NodalField.__new__(f.NodalField)
        x.values, x.dofnums, x._nfree = values, dofnums, nfree
        return x

    # ---- C3 ----
    w = wl.c3_t3ffcomp_cylinder(args.c3_n, args.c3_n)
    mat = f.lamina_material(*w["lamina"])
    t = w["thickness"]
    plies = [f.Ply(f"p{k}", mat, t / 4, a) for k, a in enumerate(w["angles"])]
    # the reference example's `cylindrical!` csys (clamp_cyl_expl_examples.jl:62-68: e3 radial, e2 = axis, e1 = e2 x e3) as a
    # built-in kind evaluated on the device; a Python callback over 6 M element nodes took 1.2 - 1.6 s here in round 1
    layup = f.CompositeLayup("C3", plies, f.CSysKind.cylindrical(axis=(0.0, 0.0, 1.0)))
    femm = f.FEMMShellT3FFComp(f.IntegDomain(w["conn"], None, t), layup, device=local_rank)
    femm.ctx.set_stream(stream.cuda_stream)
    geom0, dchi = field(w["xyz"]), field(None, w["dofnums"], w["nfree"])
    # associategeometry!: csys evaluated at every node of every element, accumulation, normalisation and validity
    # pass all on the device (fsgpu_associategeometry_csys)
    f.associategeometry(femm, geom0)  # (first call: mesh upload + device allocations)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f.associategeometry(femm, geom0)
    torch.cuda.synchronize()
    assoc_s = time.perf_counter() - t0
    assert np.abs(femm._normals - w["normals"]).max() < 1e-12 and femm._normal_valid.all()  # radial, all valid
    femm._startassembly(f.SysmatAssemblerFFBlock(), dchi)
    femm._sync_stab()
    p = femm._params()
    ne = w["conn"].shape[0]
    ms_k = timed(lambda: femm.ctx.shell_op("t3ffcomp_stiffness", p))
    kms = femm.ctx.last_kernel_ms
    ms_m = timed(lambda: femm.ctx.shell_op("t3ffcomp_mass", p))

    def roof(flops, in_bytes, nnz_, ne_, ms):
        t_f = flops * ne_ / (fp64_peak * 1e12) * 1e3
        t_h = (in_bytes * ne_ + 8.0 * nnz_) / (hbm_peak * 1e9) * 1e3
        return {"bound": "fp64" if t_f >= t_h else "hbm", "t_roof_ms": max(t_f, t_h), "frac": max(t_f, t_h) / ms, "t_fp64_ms": t_f, "t_hbm_ms": t_h,
                "flops_per_element": flops, "algorithmic_bytes_per_element": in_bytes + 8.0 * nnz_ / ne_, "basis": "step = value-array clear + kernel"}

    out["t3ffcomp_C3"] = {"roofline": roof(10.5e3, 73.0 + 72.0 + 8.0, femm.ctx.result_size()[2], ne, ms_k),"workload": f"T3FFComp 4-ply [0/90/90/0] cylinder, {ne} triangles per rank, per-element layup csys: stiffness and lumped mass -> CSC (FFBlock)",
                          "stiffness_elements_per_s": ne * world / (ms_k * 1e-3), "stiffness_ms": ms_k, "stiffness_kernel_ms": kms,
                          "mass_elements_per_s": ne * world / (ms_m * 1e-3), "mass_ms": ms_m, "nnz": int(femm.ctx.result_size()[2]),
                          "associategeometry_s_device_csys": assoc_s}
    femm.ctx.close()

    # ---- C5 ----
    w = wl.c5_beam_lattice(args.c5_n)
    sc = w["sections"]
    secs = f.FESetL2Beam(sc["A"], sc["I1"], sc["I2"], sc["I3"], sc["J"], sc["A2s"], sc["A3s"], sc["x1x2"])
    bf = f.FEMMCorotBeam(f.IntegDomain(w["conn"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]), secs, device=local_rank)
    bf.ctx.set_stream(stream.cuda_stream)
    geom0, dchi = field(w["xyz"]), field(None, w["dofnums"], w["nfree"])
    bf._sync_mesh(geom0)
    bf._startassembly(f.SysmatAssemblerFFBlock(), dchi)
    bf.ctx.set_state(w["u1"], w["Rfield1"])
    bp = bf._params()
    ne = w["conn"].shape[0]
    ms_r = timed(lambda: bf.ctx.beam_op("restoringforce", bp, 1))
    ms_s = timed(lambda: bf.ctx.beam_op("stiffness", bp))
    ms_g = timed(lambda: bf.ctx.beam_op("geostiffness", bp))
    tot = ms_r + ms_s + ms_g
    nnz5 = bf.ctx.result_size()[2]
    out["corotbeam_C5"] = {"roofline": {"stiffness": roof(1.5e3, 152.0, nnz5, ne, ms_s), "geostiffness": roof(1.5e3, 152.0, nnz5, ne, ms_g),
                                        "restoringforce": roof(450.0, 170.0, 0, ne, ms_r)},"workload": f"corotational beam lattice, {ne} elements per rank: restoringforce + stiffness + geostiffness (one Newton iteration's assembly, FFBlock)",
                           "newton_assemblies_per_s": 1e3 / tot, "elements_per_s": ne * world / (tot * 1e-3), "restoringforce_ms": ms_r,
                           "stiffness_ms": ms_s, "geostiffness_ms": ms_g, "nnz": int(bf.ctx.result_size()[2])}
    bf.ctx.close()

    # ---- Q4RSComp on the C2 mesh (4-ply layup, cartesian layup csys), RED path and the order-fixed gather path ----
    w = wl.c2_q4rs_plate(args.n)
    t = w["thickness"]
    mat = f.lamina_material(1500.0, 133860e6, 7706e6, 0.301, 4306e6, 4306e6, 2760e6)
    plies = [f.Ply(f"p{k}", mat, t / 4, a) for k, a in enumerate((0.0, 90.0, 90.0, 0.0))]
    fq = f.FEMMShellQ4RSComp(f.IntegDomain(w["conn"], f.GaussRule2x2(), t), f.CompositeLayup("Q4C", plies, np.eye(3)), device=local_rank)
    fq.ctx.set_stream(stream.cuda_stream)
    geom0, dchi = field(w["xyz"]), field(None, w["dofnums"], w["nfree"])
    f.associategeometry(fq, geom0)
    fq._startassembly(f.SysmatAssemblerFFBlock(), dchi)
    fq._sync_stab()
    p = fq._params()
    ne = w["conn"].shape[0]
    ms_qc = timed(lambda: fq.ctx.shell_op("q4rscomp_stiffness", p))
    kms_qc = fq.ctx.last_kernel_ms
    nnzq = fq.ctx.result_size()[2]
    out["q4rscomp_C2mesh"] = {"workload": f"Q4RSComp 4-ply [0/90/90/0], {ne} quads per rank (the C2 mesh), GaussRule(2,2): stiffness -> CSC (FFBlock)",
                              "elements_per_s": ne * world / (ms_qc * 1e-3), "stiffness_ms": ms_qc, "stiffness_kernel_ms": kms_qc,
                              "roofline": roof(48.0e3, 129.0 + 8.0, nnzq, ne, ms_qc)}
    fq.ctx.close()
    fh = f.FEMMShellQ4RS(f.IntegDomain(w["conn"], f.GaussRule2x2(), t), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]), device=local_rank)
    fh.ctx.set_stream(stream.cuda_stream)
    fh.ctx.set_deterministic(True)
    f.associategeometry(fh, geom0)
    fh._startassembly(f.SysmatAssemblerFFBlock(), dchi)
    fh._sync_stab()
    p = fh._params()
    ms_det = timed(lambda: fh.ctx.shell_op("q4rs_stiffness", p))
    out["q4rs_deterministic_C2"] = {"workload": "C2 stiffness with fsgpu_set_deterministic: element matrices -> dense buffer -> one owner per matrix block sums "
                                                "in ascending element order (the reference loop's order), no atomics, no clearing",
                                    "ms_per_step": ms_det, "elements_per_s": ne * world / (ms_det * 1e-3), "scatter_path": fh.ctx.scatter_path,
                                    "bitwise_reproducible": True}
    # ---- COO -> CSC (Julia sparse()) of the COO list the reference materialises for the first 64k C2 elements ----
    if rank == 0 and not args.no_cpu_baseline:
        try:
            out["coo_to_csc"] = coo_bench(w, fh, min(ne, 64000))
        except Exception as ex_:
            out["coo_to_csc"] = {"error": repr(ex_)}
    fh.ctx.close()
    return out


def coo_bench(w, femm, nsample):
    """fsgpu_coo_to_csc (host arrays in, host arrays out: H2D + sort over the significant key bits + in-order segment
    sums + D2H) against the C port's serial counting-sort `ref_coo_to_csc` on the same triples.  The full C2 list
    (576 M triples, 13.8 GB) fits the device but not a bounded CPU sample; the sample is stated."""
    from oracle import fe_external as fx
    from oracle import shells as osh

    lib = load_refport()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    conn = np.ascontiguousarray(w["conn"][:nsample])
    n = 24
    nt = nsample * n * n
    Dps, Dt = osh.shell_material_stiffness(fx.moduli_iso(w["E"], w["nu"]))
    Dps, Dt = np.ascontiguousarray(Dps), np.ascontiguousarray(Dt)
    pc, wt = fx.gauss_rule_2x2()
    pc = np.ascontiguousarray(pc)
    I, J, V = np.zeros(nt, np.int64), np.zeros(nt, np.int64), np.zeros(nt)
    v8 = np.ascontiguousarray(np.asarray(femm._normal_valid).astype(np.uint8))
    nF = np.asfortranarray(femm._normals)
    nnodes, nall = w["xyz"].shape[0], w["dofnums"].size
    lib.ref_shell_stiffness_coo(4, C.c_int64(nsample), P(conn), C.c_int64(nnodes), P(w["xyz"]), P(nF), P(v8), P(w["dofnums"]), P(Dps), P(Dt),
                                C.c_double(w["thickness"]), C.c_double(0.1), C.c_double(1.0), 4, P(pc), P(wt), os.cpu_count() or 1, P(I), P(J), P(V))
    cp, rv, nz = np.zeros(nall + 1, np.int64), np.zeros(nt, np.int64), np.zeros(nt)
    t0 = time.perf_counter()
    nnz = lib.ref_coo_to_csc(C.c_int64(nt), P(I), P(J), P(V), C.c_int64(nall), C.c_int64(nall), C.c_int64(nall), C.c_int64(nall), P(cp), P(rv), P(nz))
    cpu_s = time.perf_counter() - t0
    femm.ctx.coo_to_csc(I, J, V, nall, nall)  # warm-up (allocations)
    t0 = time.perf_counter()
    S = femm.ctx.coo_to_csc(I, J, V, nall, nall)
    gpu_s = time.perf_counter() - t0
    same = bool(np.array_equal(S.colptr, cp) and np.array_equal(S.rowval, rv[:nnz]) and np.array_equal(S.nzval, nz[:nnz]))
    return {"workload": f"COO -> CSC of the {nt} triples (24 B each) the reference materialises for the first {nsample} C2 elements, {nall} x {nall}",
            "triples": int(nt), "nnz": int(nnz), "gpu_e2e_s": gpu_s, "gpu_triples_per_s": nt / gpu_s, "cpu_serial_s": cpu_s,
            "cpu_triples_per_s": nt / cpu_s, "speedup": cpu_s / gpu_s, "bitwise_equal_to_cpu": same,
            "includes": "H2D of I, J, V (Int64/Int64/f64 host arrays) + conversion + D2H of colptr/rowval/nzval; CPU: serial counting sort (Julia sparse() is serial)"}


def gathered_c2(args, rank, local_rank, world, stream, w, normals, valid):
    """N > 1 only: ONE global C2 matrix assembled by all ranks (strong scaling of the 1M-element case): every rank
    assembles the elements touching the nodes whose columns it owns (partition.ColumnBlockPlan), then the column
    blocks are gathered over NCCL into the global CSC (Int64 / Float64, Julia layout) on every rank's device --
    SURVEY section 8(e) 'gathering assembled blocks'; the solver hand-off.  Timed with CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist

    import fsb200
    from fsb200 import partition as pt

    f = fsb200.femm
    dev = torch.device("cuda", local_rank)
    t0 = time.perf_counter()
    plan = pt.ColumnBlockPlan(w["conn"], w["dofnums"], w["nfree"], "ffblock", rank, world)
    plan_s = time.perf_counter() - t0
    mat = f.MatDeforElastIso(w["E"], w["nu"], w["rho"])
    femm = f.FEMMShellQ4RS(f.IntegDomain(plan.conn, f.GaussRule2x2(), w["thickness"]), mat, device=local_rank)
    femm.ctx.set_stream(stream.cuda_stream)
    femm._normals, femm._normal_valid = np.asfortranarray(plan.restrict_nodes(normals)), plan.restrict_nodes(valid)
    femm._associatedgeometry = True
    g = f.NodalField.__new__(f.NodalField)
    g.values = np.asfortranarray(plan.restrict_nodes(w["xyz"]))
    d = f.NodalField.__new__(f.NodalField)
    d.values, d.dofnums, d._nfree = None, plan.dofnums, plan.nfree
    with torch.cuda.stream(stream):
        femm._sync_mesh(g)
        femm._startassembly(f.SysmatAssemblerFFBlock(), d)
        femm._sync_stab()
        params = femm._params()
        res = None
        times = []
        for it in range(4):
            torch.cuda.synchronize()
            dist.barrier()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(stream)
            femm.ctx.shell_op("q4rs_stiffness", params)
            e[1].record(stream)
            res = None  # release the previous global arrays first
            res = pt.gather_matrix(femm.ctx, plan, dev, to_host=False)
            e[2].record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[0].elapsed_time(e[2])], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it > 0:
                times.append(t.cpu().numpy())
        n, colptr, rowval, nzval = res
        nnz = int(rowval.numel())
        chk = float(nzval.sum().item())
    tm = np.mean(times, axis=0)
    nelem = w["conn"].shape[0]
    out = {"workload": f"ONE global C2 matrix ({nelem} elements) assembled by {world} ranks (owned-column blocks, interface elements "
                       "computed by both neighbours) and gathered over NCCL into the global CSC on every rank's device",
           "elements_this_rank": int(len(plan.elems)), "assembly_ms": float(tm[0]), "gather_ms": float(tm[1]), "total_ms": float(tm[2]),
           "elements_per_s": nelem / (float(tm[2]) * 1e-3), "scaling": "strong", "ncols": int(n), "nnz": nnz,
           "gathered_bytes_per_rank": int(nnz * 16 + n * 8), "gather_GBps_received_per_rank": (nnz * 16 + n * 8) * (world - 1) / world / (float(tm[1]) * 1e-3) / 1e9,
           "plan_host_s": plan_s, "nzval_checksum": chk}
    del res, colptr, rowval, nzval
    femm.ctx.close()
    return out


def cpu_explicit(local_rank, nx=400):
    """Reference explicit loop (SpMV + vector updates) on the host cores: K of a nx x nx/2 x 2
    T3FF strip (assembled on the GPU, fetched), oracle C port `ref_explicit_steps`."""
    import scipy.sparse as sp

    import fsb200
    from fsb200 import workloads as wl

    f = fsb200.femm
    w = wl.c4_t3ff_panel(nx, nx // 2)
    femm = f.FEMMShellT3FF(f.IntegDomain(w["conn"], None, w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]), device=local_rank)
    geom0 = f.NodalField.__new__(f.NodalField)
    geom0.values = w["xyz"]
    dchi = f.NodalField.__new__(f.NodalField)
    dchi.values, dchi.dofnums, dchi._nfree = None, w["dofnums"], w["nfree"]
    f.associategeometry(femm, geom0)
    K = f.stiffness(femm, f.SysmatAssemblerFFBlock(), geom0, None, None, dchi)
    femm.ctx.shell_mass_diag(femm._params(), 3, nfree_only=True)
    M = femm.ctx.fetch_vector(w["nfree"])
    Kc = K.to_scipy().tocsr()
    Kc.sort_indices()
    rp, cv, nz = (Kc.indptr + 1).astype(np.int64), (Kc.indices + 1).astype(np.int64), Kc.data
    lib = load_refport()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    n = w["nfree"]
    U, V, A, F0 = np.zeros(n), np.zeros(n), np.zeros(n), np.ones(n)
    ncores = os.cpu_count() or 1
    nst = 50
    lib.ref_explicit_steps(C.c_int64(n), P(rp), P(cv), P(nz), P(M), C.c_double(100.0), C.c_double(1e-8), P(F0), None, C.c_int64(5), P(U), P(V), P(A), ncores)
    t0 = time.perf_counter()
    lib.ref_explicit_steps(C.c_int64(n), P(rp), P(cv), P(nz), P(M), C.c_double(100.0), C.c_double(1e-8), P(F0), None, C.c_int64(nst), P(U), P(V), P(A), ncores)
    dt = time.perf_counter() - t0
    ne = w["conn"].shape[0]
    return {"value": ne * nst / dt, "unit": "element-steps/s", "cores": ncores, "kind": "port",
            "sample": f"{ne}-element T3FF strip, {nst} steps, CSR SpMV (Int64 indices) + vector updates, OpenMP row-parallel"}


# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1000, help="quads per side (1000 -> 1M elements, the BASELINE config)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-sample", type=int, default=256000, help="elements per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads (T3FF assembly / explicit loop on C4)")
    ap.add_argument("--all-extras", action="store_true", help="N > 1: also run the C3 / C5 workloads (replicas) on every rank")
    ap.add_argument("--c4-nx", type=int, default=2000, help="C4 strip: nx x nx/2 cells x 2 triangles (2000 -> 4M elements per rank)")
    ap.add_argument("--expl-steps", type=int, default=1000)
    ap.add_argument("--power-its", type=int, default=30)
    ap.add_argument("--c3-n", type=int, default=1000, help="C3 cylinder: n x n x 2 triangles (1000 -> 2M)")
    ap.add_argument("--c5-n", type=int, default=69, help="C5 lattice cells per side (69 -> 1,014,300 beams)")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    metric, unit = "element matrices assembled/sec (Q4RS stiffness -> CSC)", "elements/s"
    wl = load_workloads()

    w = wl.c2_q4rs_plate(args.n)
    nelem = w["conn"].shape[0]
    config = {"workload": f"BASELINE configs[1]: Q4RS homogeneous square plate, synthetic {args.n}x{args.n} quad mesh "
                          f"({nelem} elements), stiffness assembly to CSC via SysmatAssemblerFFBlock, GaussRule(2,2)",
              "nelem": nelem, "nnodes": int(w["xyz"].shape[0]), "nfree": int(w["nfree"]),
              "l2": "inputs+outputs (slot map 2.3 GB + values 2.6 GB at 1M elements) are larger than the 126 MB L2",
              "pattern": "reused across steps for `value` (symbolic phase reported separately); rebuilt every step for `e2e`",
              "parallelism": f"element-partitioned, {world} rank(s), no data-path collective"}

    if args.impl == "reference":
        if rank != 0:
            return
        ncores = os.cpu_count() or 1
        normals, valid = oracle_normals(w)
        sample = min(nelem, args.ref_sample)
        vals = []
        for s in range(args.warmup + args.steps):
            r = cpu_q4rs_assembly(w, sample, ncores, normals, valid)
            if s >= args.warmup:
                vals.append(r)
        v = float(np.mean([x["value"] for x in vals]))
        ms = float(np.mean([x["seconds"] for x in vals])) * 1e3
        loop_s, sparse_s = float(np.mean([x["loop_s"] for x in vals])), float(np.mean([x["sparse_s"] for x in vals]))
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": unit, "cores": ncores, "kind": "port",
                                 "sample": f"first {sample} elements of the workload per step: element loop + COO append (OpenMP element-parallel, "
                                           f"{loop_s:.2f} s) + ONE serial COO->CSC pass as Julia's sparse() does ({sparse_s:.2f} s); buffers preallocated",
                                 "element_loop_only_value": float(np.mean([x["loop_value"] for x in vals]))},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # stdout carries exactly ONE JSON line: anything a library prints there during the run (NCCL's version banner
    # at communicator creation, for one) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    import fsb200

    f = fsb200.femm
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    hbm_peak, peak_kind = peaks()
    mat = f.MatDeforElastIso(w["E"], w["nu"], w["rho"])
    stream = torch.cuda.Stream()
    wg = w  # the global workload
    plan = None
    if world > 1:
        # STRONG scaling: the ONE mesh of the configuration is split over the ranks.  Rank r owns a contiguous block
        # of columns of the global matrix and assembles every element touching a node of that block
        # (partition.ColumnBlockPlan; interface elements are computed by both neighbours, no partial sums travel).
        # Nodal normals come from the global mesh (one device pass per rank, untimed setup).
        from fsb200 import partition as pt

        fg = f.FEMMShellQ4RS(f.IntegDomain(wg["conn"], f.GaussRule2x2(), wg["thickness"]), mat, device=local_rank)
        gg = f.NodalField.__new__(f.NodalField)
        gg.values = wg["xyz"]
        f.associategeometry(fg, gg)
        nrm_g, val_g = fg._normals, fg._normal_valid
        fg.ctx.close()
        plan = pt.ColumnBlockPlan(wg["conn"], wg["dofnums"], wg["nfree"], "ffblock", rank, world)
        w = dict(wg, conn=plan.conn, xyz=np.asfortranarray(plan.restrict_nodes(wg["xyz"])), dofnums=plan.dofnums, nfree=plan.nfree)
        config["parallelism"] = (f"STRONG scaling: the one {nelem}-element mesh split over {world} ranks by owned column blocks "
                                 f"(this rank: {len(plan.elems)} elements incl. interface elements); no data-path collective, "
                                 "the matrix stays distributed (column block per rank)")

    # host (pinned) inputs, as a Julia host would hold them
    xyz_p, k1 = pin_copy(w["xyz"])
    conn_p, k2 = pin_copy(np.ascontiguousarray(w["conn"]))
    dof_p, k3 = pin_copy(w["dofnums"])
    femm = f.FEMMShellQ4RS(f.IntegDomain(conn_p, f.GaussRule2x2(), w["thickness"]), mat, device=local_rank)
    femm.ctx.set_stream(stream.cuda_stream)
    geom0 = f.NodalField.__new__(f.NodalField)
    geom0.values = xyz_p
    dchi = f.NodalField.__new__(f.NodalField)
    dchi.values, dchi.dofnums, dchi._nfree = None, dof_p, w["nfree"]
    u0 = R0 = None

    # --- setup (untimed): nodal normals, first symbolic phase -------------------------------
    if plan is None:
        f.associategeometry(femm, geom0)
    else:
        femm._normals, femm._normal_valid = np.asfortranarray(plan.restrict_nodes(nrm_g)), plan.restrict_nodes(val_g)
        femm._associatedgeometry = True
        femm._sync_mesh(geom0)
    t0 = time.perf_counter()
    femm._startassembly(f.SysmatAssemblerFFBlock(), dchi)
    femm.ctx.sync()
    symbolic_ms = (time.perf_counter() - t0) * 1e3
    nr, nc, nnz = femm.ctx.symbolic(fsb200._lib.FFBLOCK)  # (second build, for a warm number)
    t0 = time.perf_counter()
    nr, nc, nnz = femm.ctx.symbolic(fsb200._lib.FFBLOCK)
    femm.ctx.sync()
    symbolic_warm_ms = (time.perf_counter() - t0) * 1e3
    params = femm._params()
    femm._sync_stab()
    fp64_peak, copy_bw = femm.ctx.measure_peaks()

    # --- device-resident numeric phase: `value` --------------------------------------------
    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler is started (and has delivered its first row) BEFORE the warm-up: nvidia-smi's start-up -- one process
    # per rank -- otherwise competes with the launching threads for host cores inside a timed region of 16 - 100 ms;
    # its rows cover the warm-up and the timed steps (the same load)
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.perf_counter()
    while not sampler.rows and time.perf_counter() - t_wait < 3.0:
        time.sleep(0.01)
    sampler.mark()
    for _ in range(max(3, args.warmup)):
        femm.ctx.shell_op("q4rs_stiffness", params)
    barrier()
    l0 = femm.ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    e0.record(stream)
    for _ in range(args.steps):
        femm.ctx.shell_op("q4rs_stiffness", params)
        kms.append(femm.ctx.last_kernel_ms)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = femm.ctx.launch_count - l0
    ms_total = e0.elapsed_time(e1)
    tmax = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = float(tmax.item()) / args.steps
    value = nelem / (ms_per_step * 1e-3)  # N > 1: the same global mesh, split (strong scaling)
    kernel_ms = float(np.mean(kms))
    nelem_rank = int(w["conn"].shape[0])

    # --- end to end through the reference-facing operator API: `e2e` -------------------------
    cp_p, k4 = pinned((nc + 1,), np.int64)
    rv_p, k5 = pinned((nnz,), np.int64)
    nz_p, k6 = pinned((nnz,), np.float64)
    nrm_host, val_host = femm._normals, femm._normal_valid
    h2d = conn_p.nbytes + xyz_p.nbytes + dof_p.nbytes + nrm_host.nbytes + val_host.size + 8
    d2h = cp_p.nbytes + rv_p.nbytes + nz_p.nbytes

    def e2e_step():
        g = f.NodalField.__new__(f.NodalField)  # fresh field objects -> mesh, dofs, normals re-uploaded
        g.values = xyz_p
        d = f.NodalField.__new__(f.NodalField)
        d.values, d.dofnums, d._nfree = None, dof_p, w["nfree"]
        femm.reset_uploads()
        return f.stiffness(femm, f.SysmatAssemblerFFBlock(), g, u0, R0, d, out=(cp_p, rv_p, nz_p))

    e2e_includes = ("H2D mesh+dofs+normals, symbolic phase, numeric phase, D2H of the CSC result into host colptr+rowval+nzval (Int64/f64; "
                    "the row indices cross PCIe run-length coded and are expanded by host threads inside the call)")
    if world > 1:
        e2e_includes += (f"; N > 1: every rank does this for its share of the ONE mesh (its elements, its local matrix in local numbering "
                         "with the local->global row map; the owned column block is a contiguous slice of it), all ranks concurrently")
    e2e_step()
    # PCIe probe (pinned, this box): explains the end-to-end number, which is dominated by the CSC D2H
    probe = torch.empty(1 << 27, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    tp = time.perf_counter()
    nz_p_t = k6[: 1 << 27] if k6.numel() >= (1 << 27) else k6
    nz_p_t.copy_(probe[: nz_p_t.numel()], non_blocking=False)
    torch.cuda.synchronize()
    d2h_gbs = nz_p_t.numel() * 8 / (time.perf_counter() - tp) / 1e9
    del probe
    barrier()
    b0 = femm.ctx.d2h_bytes
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        K = e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    d2h_link = (femm.ctx.d2h_bytes - b0) / args.e2e_steps  # bytes that crossed PCIe (row indices travel run-length coded)
    te = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = nelem / float(te.item())
    # secondary figure (not the headline): a re-assembly on the SAME mesh and numbering -- only the values
    # change (a Newton / time-stepping loop), so only nzval crosses PCIe
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        femm.ctx.shell_op("q4rs_stiffness", params)
        femm.ctx.fetch_values(nz_p)
    barrier()
    refresh_ms = (time.perf_counter() - t0) / args.e2e_steps * 1e3

    # checksum of the values of the owned column blocks: the gathered multi-GPU assembly of the same matrix must reproduce it
    nz_checksum_blocks = None
    if world > 1:
        c0, c1 = int(cp_p[plan.lcol_lo]) - 1, int(cp_p[plan.lcol_hi]) - 1
        t = torch.tensor([float(np.sum(nz_p[c0:c1]))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        nz_checksum_blocks = float(t.item())
    # secondary figure: the same end-to-end call for a consumer of ONE triangle (cholesky(Symmetric(K, :L))): half the
    # bytes cross PCIe (fsgpu_fetch_matrix_uplo); the headline e2e above is the full matrix
    asm_l = f.SysmatAssemblerFFBlock()
    asm_l.uplo = "L"

    def e2e_step_lower():
        g = f.NodalField.__new__(f.NodalField)
        g.values = xyz_p
        d = f.NodalField.__new__(f.NodalField)
        d.values, d.dofnums, d._nfree = None, dof_p, w["nfree"]
        femm.reset_uploads()
        return f.stiffness(femm, asm_l, g, u0, R0, d, out=(cp_p, rv_p, nz_p))

    e2e_step_lower()
    barrier()
    b0 = femm.ctx.d2h_bytes
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step_lower()
    barrier()
    tl = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
    e2e_lower = {"ms_per_step": float(tl.item()) * 1e3, "value": nelem / float(tl.item()), "unit": unit,
                 "d2h_bytes_per_step": int((femm.ctx.d2h_bytes - b0) / args.e2e_steps),
                 "what": "lower triangle incl. diagonal (uplo = L) of the same matrix, same H2D + symbolic + numeric phases"}

    # secondary figure: the same call with PAGEABLE destination arrays (what a plain Julia `Vector` is): the values and colptr
    # reach them through the library's pinned staging ring and host threads instead of a driver-staged DMA copy
    e2e_pageable = None
    if world == 1:
        outs = (np.empty(nc + 1, np.int64), np.empty(nnz, np.int64), np.empty(nnz, np.float64))
        for a in outs:
            a[...] = 0  # touch the pages outside the timed region (as the pinned arrays are)

        def e2e_step_pageable():
            g = f.NodalField.__new__(f.NodalField)
            g.values = xyz_p
            d = f.NodalField.__new__(f.NodalField)
            d.values, d.dofnums, d._nfree = None, dof_p, w["nfree"]
            femm.reset_uploads()
            return f.stiffness(femm, f.SysmatAssemblerFFBlock(), g, u0, R0, d, out=outs)

        e2e_step_pageable()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step_pageable()
        tp_s = (time.perf_counter() - t0) / args.e2e_steps
        e2e_pageable = {"ms_per_step": tp_s * 1e3, "value": nelem / tp_s, "unit": unit,
                        "what": "same call, colptr / rowval / nzval are ordinary (pageable) numpy arrays; inputs still pinned"}
        del outs

    # free the C2 buffers before the 4M-element workload
    del K, cp_p, rv_p, nz_p, k4, k5, k6
    femm.ctx.close()
    extras = None
    if not args.no_extras:
        extras = explicit_c4(args, rank, local_rank, world, stream, hbm_peak, fp64_peak)
        if world == 1 or args.all_extras:
            with torch.cuda.stream(stream):
                extras.update(extras_c3_c5(args, rank, local_rank, world, stream, hbm_peak, fp64_peak))

    if world > 1 and not args.no_extras:
        gathered = gathered_c2(args, rank, local_rank, world, stream, wg, nrm_g, val_g)
        gathered["nzval_checksum_of_the_distributed_blocks"] = nz_checksum_blocks
        gathered["checksum_rel_diff"] = abs(gathered["nzval_checksum"] - nz_checksum_blocks) / abs(nz_checksum_blocks)
        extras["gathered_assembly_C2"] = gathered

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # --- roofline of the dominant kernel: north star = "the slower of FP64 peak and HBM bytes at peak bandwidth",
    # taken on the STEP (value-array clear + element kernel), for the elements this rank processes ----------------
    out_bytes = 8.0 * nnz / nelem_rank  # values written once (pattern reused)
    alg_bytes = Q4_IN_BYTES + out_bytes
    t_hbm = alg_bytes * nelem_rank / (hbm_peak * 1e9) * 1e3
    t_fp64 = Q4_FLOPS * nelem_rank / (fp64_peak * 1e12) * 1e3
    # DRAM bytes of one launch of this kernel from the committed `ncu --set full` capture (same workload size)
    traffic, traffic_src = None, None
    for name in ("r02_q4_traffic.json", "r01_q4_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                if int(tj.get("nelem", -1)) == int(nelem_rank):
                    traffic, traffic_src = float(tj["dram_bytes_per_launch"]), tj.get("source")
                    break
            except (OSError, ValueError, KeyError):
                pass
    fp64_bound = t_fp64 >= t_hbm
    fl_step = Q4_FLOPS * nelem_rank / (ms_per_step * 1e-3) / 1e12
    gb_step = alg_bytes * nelem_rank / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "fp64" if fp64_bound else "hbm",
                "achieved": fl_step if fp64_bound else gb_step, "peak": fp64_peak if fp64_bound else hbm_peak,
                "unit": "TFLOP/s" if fp64_bound else "GB/s", "frac": max(t_fp64, t_hbm) / ms_per_step,
                "basis": "step = value-array clear + element kernel + status read-back, this rank's elements",
                "t_roof_ms": max(t_fp64, t_hbm), "frac_kernel_only": max(t_fp64, t_hbm) / kernel_ms,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "k_q4_stiffness<false,false,EmitRuns>",
                "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / ms_per_step, "flops_per_element": Q4_FLOPS,
                "peak_source": "FP64: fsgpu_measure_peaks DFMA micro-kernel on this device in this run (record with clocks: "
                               "profiles/r02_fp64_peak.json); HBM: MEASURED_PEAKS.json (" + peak_kind + ")",
                "hbm": {"achieved": gb_step, "peak": hbm_peak, "unit": "GB/s", "frac": t_hbm / ms_per_step,
                        "algorithmic_bytes_per_element": alg_bytes},
                "copy_gbs_this_device": copy_bw}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ncores = os.cpu_count() or 1
        s1 = min(nelem, 32000)
        r1 = cpu_q4rs_assembly(wg, s1, 1, nrm_host, val_host)
        sN = min(nelem, max(64000, 16000 * ncores))
        rN = cpu_q4rs_assembly(wg, sN, ncores, nrm_host, val_host)
        cpu = {"value": rN["value"], "unit": unit, "cores": ncores, "kind": "port",
               "sample": f"first {sN} elements of the workload (all {ncores} threads): element loop {rN['loop_s']:.2f} s + COO->CSC "
                         f"{rN['sparse_s']:.2f} s; single thread (= the reference's serial loop) on the first {s1} elements: "
                         f"{r1['value']:.0f} elements/s (loop {r1['loop_s']:.2f} s + COO->CSC {r1['sparse_s']:.2f} s)",
               "single_thread_value": r1["value"], "element_loop_only_value": rN["loop_value"],
               "single_thread_element_loop_only_value": r1["loop_value"]}

    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h_link),
                    "host_result_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3, "includes": e2e_includes,
                    "pinned_d2h_gbs_this_box": d2h_gbs,
                    "values_refresh_ms_same_pattern": refresh_ms, "lower_triangle": e2e_lower, "pageable_destination": e2e_pageable},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "symbolic_ms": {"first": symbolic_ms, "warm": symbolic_warm_ms}, "nnz": int(nnz), "other_workloads": extras}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
