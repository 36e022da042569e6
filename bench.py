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
        self.index, self.rows, self.proc = index, [], None

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
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


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


def cpu_q4rs_assembly(w, nelem_sample, nthreads, normals, valid):
    """Times the reference algorithm (element loop + COO append + sparse()) on the first
    `nelem_sample` elements of the workload.  Returns elements/s."""
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
    I, J, V = np.empty(nt, np.int64), np.empty(nt, np.int64), np.empty(nt)
    v8 = np.ascontiguousarray(valid.astype(np.uint8))
    nF = np.asfortranarray(normals)
    nnodes = w["xyz"].shape[0]
    nall, nfree = w["dofnums"].size, w["nfree"]
    alpha = 0.1 if nn == 4 else 5 / 12 / 1.5
    t0 = time.perf_counter()
    lib.ref_shell_stiffness_coo(nn, C.c_int64(nelem_sample), P(conn), C.c_int64(nnodes), P(w["xyz"]), P(nF), P(v8),
                                P(w["dofnums"]), P(Dps), P(Dt), C.c_double(w["thickness"]), C.c_double(alpha), C.c_double(1.0), 4,
                                P(pc), P(wt), nthreads, P(I), P(J), P(V))
    nnz = lib.ref_coo_to_csc(C.c_int64(nt), P(I), P(J), P(V), C.c_int64(nall), C.c_int64(nall), C.c_int64(nfree), C.c_int64(nfree), None, None, None)
    cp, rv, nz = np.empty(nfree + 1, np.int64), np.empty(nnz, np.int64), np.empty(nnz)
    lib.ref_coo_to_csc(C.c_int64(nt), P(I), P(J), P(V), C.c_int64(nall), C.c_int64(nall), C.c_int64(nfree), C.c_int64(nfree), P(cp), P(rv), P(nz))
    dt = time.perf_counter() - t0
    return nelem_sample / dt, dt


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
    import torch
    import torch.distributed as dist

    import fsb200
    from fsb200 import partition as pt
    from fsb200 import workloads as wl

    f = fsb200.femm
    nx, ny = args.c4_nx, args.c4_nx // 2
    w = wl.c4_strip(rank, world, nx, ny)
    nelem = w["conn"].shape[0]
    femm = f.FEMMShellT3FF(f.IntegDomain(w["conn"], None, w["thickness"]), f.MatDeforElastIso(w["E"], w["nu"], w["rho"]), device=local_rank)
    femm.ctx.set_stream(stream.cuda_stream)
    geom0 = f.NodalField.__new__(f.NodalField)
    geom0.values = w["xyz"]
    dchi = f.NodalField.__new__(f.NodalField)
    dchi.values, dchi.dofnums, dchi._nfree = None, w["dofnums"], w["nfree"]
    if world > 1:  # nodal normals of interface nodes see the elements of both partitions
        with torch.cuda.stream(stream):
            f.associategeometry(femm, geom0, interface=(pt.strip_links(rank, world, w["lo_nodes"], w["hi_nodes"]), torch.device("cuda", local_rank)))
    else:
        f.associategeometry(femm, geom0)
    t0 = time.perf_counter()
    femm._startassembly(f.SysmatAssemblerFFBlock(), dchi)
    femm.ctx.sync()
    sym_ms = (time.perf_counter() - t0) * 1e3
    femm._sync_stab()
    p = femm._params()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- T3FF stiffness assembly (numeric phase, device resident) ---
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
    tm = torch.tensor([e0.elapsed_time(e1) / nrep], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    asm_ms = float(tm.item())
    nnz = femm.ctx.result_size()[2]
    kernel_ms = float(np.mean(kms))
    alg = T3_IN_BYTES + 8.0 * nnz / nelem
    # associategeometry! on the device: nodal normals + crease detection, single rank only
    ag_ms = None
    if world == 1:
        femm.ctx.associategeometry(femm.threshold_angle, None, True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            femm.ctx.associategeometry(femm.threshold_angle, None, True)
        barrier()
        ag_ms = (time.perf_counter() - t0) / 5 * 1e3
    asm = {"workload": f"T3FF stiffness -> CSC (FFBlock), {nelem} elements per rank", "value": nelem * world / (asm_ms * 1e-3),
           "unit": "elements/s", "ms_per_step": asm_ms, "kernel_ms": kernel_ms, "symbolic_ms": sym_ms, "nnz": int(nnz),
           "associategeometry_ms": ag_ms,
           "roofline": {"bound": "hbm", "achieved": alg * nelem / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg * nelem / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_element": alg,
                        "kernel": "k_t3_stiffness<false,false,EmitRuns>"},
           "fp64": {"achieved_tflops": T3_FLOPS * nelem / (kernel_ms * 1e-3) / 1e12, "peak_tflops": fp64_peak,
                    "frac": T3_FLOPS * nelem / (kernel_ms * 1e-3) / 1e12 / fp64_peak, "flops_per_element": T3_FLOPS}}

    # --- explicit central differences: K = FF block (device resident), lumped M ---
    femm.ctx.shell_mass_diag(p, 3, nfree_only=True)
    nf = w["nfree"]
    links = pt.strip_links(rank, world, w["lo_dofs"], w["hi_dofs"])
    ex_if = pt.InterfaceExchange(links, torch.device("cuda", local_rank)) if world > 1 else None
    with torch.cuda.stream(stream):
        if world > 1:  # lumped mass of interface nodes: sum of both partitions' contributions
            import ctypes as C

            vp, vn = C.c_void_p(), C.c_int64()
            fsb200._lib.check(fsb200._lib.lib.fsgpu_vector_device(femm.ctx._h, C.byref(vp), C.byref(vn)))
            Mt = torch.as_tensor(pt.DevicePointer(vp.value, vn.value), device="cuda")
            ex_if.exchange_sum(Mt)
            torch.cuda.synchronize()
        ex = fsb200.Explicit(femm.ctx, c_scale=100.0, dt=0.0)
        lam = ex.omega_max_sq(args.power_its)
        if world > 1:
            lt = torch.tensor([lam], device="cuda", dtype=torch.float64)
            dist.all_reduce(lt, op=dist.ReduceOp.MAX)
            lam = float(lt.item())
        ex.close()
        dt = 0.9 * 2 / np.sqrt(lam)
        ex = fsb200.Explicit(femm.ctx, c_scale=100.0, dt=dt)
        F0 = np.zeros(nf)
        F0[2::6][: nf // 600] = 1.0  # a small pressure patch
        ex.set_load(F0)
        ex.start(1.0)
        U, V, A, E = ex.device_state()
        Et = torch.as_tensor(pt.DevicePointer(E, nf), device="cuda")
        nsteps = args.expl_steps

        def run(n):
            if world == 1:
                ex.step(n)
            else:
                for _ in range(n):
                    ex.step_begin()
                    ex_if.exchange_sum(Et)
                    ex.step_end(1.0)

        run(20)
        barrier()
        l0 = femm.ctx.launch_count
        e0.record(stream)
        run(nsteps)
        e1.record(stream)
        barrier()
        launches = femm.ctx.launch_count - l0
        tm = torch.tensor([e0.elapsed_time(e1) / nsteps], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        step_ms = float(tm.item())
        Uh = ex.get_state()[0]
        ke = ex.kinetic_energy()
    knnz = nnz
    _, _, nruns, idx_entries = ex.layout()
    # f64 value per entry + one i32 column index per entry of every distinct row pattern (runs of <= 6 rows share
    # one), ~10 vector passes.  SURVEY 8(d)'s figure for the plain Int32 CSR form is 12 B per entry.
    alg_step = (8.0 * knnz + 4.0 * idx_entries + 10 * 8.0 * nf) / nelem
    alg_step_csr = (12.0 * knnz + 10 * 8.0 * nf) / nelem
    expl = {"workload": f"explicit central differences, T3FF, {nelem} elements per rank x {world} rank(s), SpMV form (K_ff CSR, lumped M)",
            "steps_per_s": 1e3 / step_ms, "element_steps_per_s": nelem * world * 1e3 / step_ms, "ms_per_step": step_ms, "dt": dt,
            "omega_max": float(np.sqrt(lam)), "nsteps_timed": nsteps, "gpu_launches": int(launches), "max_abs_U": float(np.abs(Uh).max()),
            "kinetic_energy": ke, "interface_exchange": "pairwise NCCL isend/irecv of packed interface E entries" if world > 1 else "none",
            "roofline": {"bound": "hbm", "achieved": alg_step * nelem / (step_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": alg_step * nelem / (step_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_element_step": alg_step,
                         "plain_csr_bytes_per_element_step": alg_step_csr, "row_runs": int(nruns), "index_entries_read": int(idx_entries),
                         "kernel": "k_spmv_step + k_update_u"}}
    # CPU arm for the explicit loop (rank 0, bounded sample: a 1/100-size strip, all host threads)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            cpu = cpu_explicit(local_rank)
        except Exception as ex_:  # never let the side measurement kill the bench line
            cpu = {"error": repr(ex_)}
    ex.close()
    return {"t3ff_assembly_C4": asm, "explicit_C4": expl, "explicit_cpu_baseline": cpu}


def extras_c3_c5(args, rank, local_rank, world, stream):
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
    layup = f.CompositeLayup("C3", plies, wl.cylindrical_csys)  # a csys CALLBACK, as in the reference example
    femm = f.FEMMShellT3FFComp(f.IntegDomain(w["conn"], None, t), layup, device=local_rank)
    femm.ctx.set_stream(stream.cuda_stream)
    geom0, dchi = field(w["xyz"]), field(None, w["dofnums"], w["nfree"])
    # associategeometry!: the host evaluates the callback at every node of every element, the device accumulates,
    # normalises and validates (fsgpu_associategeometry_dirs)
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
    out["t3ffcomp_C3"] = {"workload": f"T3FFComp 4-ply [0/90/90/0] cylinder, {ne} triangles per rank, per-element layup csys: stiffness and lumped mass -> CSC (FFBlock)",
                          "stiffness_elements_per_s": ne * world / (ms_k * 1e-3), "stiffness_ms": ms_k, "stiffness_kernel_ms": kms,
                          "mass_elements_per_s": ne * world / (ms_m * 1e-3), "mass_ms": ms_m, "nnz": int(femm.ctx.result_size()[2]),
                          "associategeometry_s_incl_host_csys_callback": assoc_s}
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
    out["corotbeam_C5"] = {"workload": f"corotational beam lattice, {ne} elements per rank: restoringforce + stiffness + geostiffness (one Newton iteration's assembly, FFBlock)",
                           "newton_assemblies_per_s": 1e3 / tot, "elements_per_s": ne * world / (tot * 1e-3), "restoringforce_ms": ms_r,
                           "stiffness_ms": ms_s, "geostiffness_ms": ms_g, "nnz": int(bf.ctx.result_size()[2])}
    bf.ctx.close()
    return out


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads (T3FF assembly / explicit loop on C4)")
    ap.add_argument("--c4-nx", type=int, default=2000, help="C4 strip: nx x nx/2 cells x 2 triangles (2000 -> 4M elements per rank)")
    ap.add_argument("--expl-steps", type=int, default=200)
    ap.add_argument("--power-its", type=int, default=30)
    ap.add_argument("--c3-n", type=int, default=1000, help="C3 cylinder: n x n x 2 triangles (1000 -> 2M)")
    ap.add_argument("--c5-n", type=int, default=69, help="C5 lattice cells per side (69 -> 1,014,300 beams)")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    metric, unit = "element matrices assembled/sec (Q4RS stiffness -> CSC)", "elements/s"
    from fsb200 import workloads as wl  # noqa: E402  (pure numpy part of the package)

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
        sample = min(nelem, max(20000, 4000 * ncores))
        vals = []
        for s in range(args.warmup + args.steps):
            v, dt = cpu_q4rs_assembly(w, sample, ncores, normals, valid)
            if s >= args.warmup:
                vals.append((v, dt))
        v = float(np.mean([x[0] for x in vals]))
        ms = float(np.mean([x[1] for x in vals])) * 1e3
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": unit, "cores": ncores, "kind": "port",
                                 "sample": f"first {sample} elements of the workload per step: element loop + COO append + COO->CSC, OpenMP element-parallel"},
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

    # host (pinned) inputs, as a Julia host would hold them
    xyz_p, k1 = pin_copy(w["xyz"])
    conn_p, k2 = pin_copy(np.ascontiguousarray(w["conn"]))
    dof_p, k3 = pin_copy(w["dofnums"])
    mat = f.MatDeforElastIso(w["E"], w["nu"], w["rho"])
    femm = f.FEMMShellQ4RS(f.IntegDomain(conn_p, f.GaussRule2x2(), w["thickness"]), mat, device=local_rank)
    stream = torch.cuda.Stream()
    femm.ctx.set_stream(stream.cuda_stream)
    geom0 = f.NodalField.__new__(f.NodalField)
    geom0.values = xyz_p
    dchi = f.NodalField.__new__(f.NodalField)
    dchi.values, dchi.dofnums, dchi._nfree = None, dof_p, w["nfree"]
    u0 = R0 = None

    # --- setup (untimed): nodal normals, first symbolic phase -------------------------------
    f.associategeometry(femm, geom0)
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

    for _ in range(max(3, args.warmup)):
        femm.ctx.shell_op("q4rs_stiffness", params)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
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
    value = nelem * world / (ms_per_step * 1e-3)
    kernel_ms = float(np.mean(kms))

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
    e2e_value = nelem * world / float(te.item())
    # secondary figure (not the headline): a re-assembly on the SAME mesh and numbering -- only the values
    # change (a Newton / time-stepping loop), so only nzval crosses PCIe
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        femm.ctx.shell_op("q4rs_stiffness", params)
        femm.ctx.fetch_values(nz_p)
    barrier()
    refresh_ms = (time.perf_counter() - t0) / args.e2e_steps * 1e3

    # checksum of the single-GPU values: the gathered multi-GPU assembly of the same matrix must reproduce it
    nz_checksum_single = float(np.sum(nz_p)) if world > 1 else None
    # free the C2 buffers before the 4M-element workload
    del K, cp_p, rv_p, nz_p, k4, k5, k6
    femm.ctx.close()
    extras = None
    if not args.no_extras:
        extras = explicit_c4(args, rank, local_rank, world, stream, hbm_peak, fp64_peak)
        with torch.cuda.stream(stream):
            extras.update(extras_c3_c5(args, rank, local_rank, world, stream))

    if world > 1 and not args.no_extras:
        gathered = gathered_c2(args, rank, local_rank, world, stream, w, nrm_host, val_host)
        if extras is None:
            extras = {}
        gathered["nzval_checksum_single_gpu"] = nz_checksum_single
        gathered["checksum_rel_diff"] = abs(gathered["nzval_checksum"] - nz_checksum_single) / abs(nz_checksum_single)
        extras["gathered_assembly_C2"] = gathered

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # --- roofline of the dominant kernel ------------------------------------------------------
    out_bytes = 8.0 * nnz / nelem  # values written once (pattern reused)
    alg_bytes = Q4_IN_BYTES + out_bytes
    achieved = alg_bytes * nelem / (kernel_ms * 1e-3) / 1e9
    # DRAM bytes of one launch of this kernel from the committed `ncu --set full` capture (same workload size)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_q4_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if int(tj.get("nelem", -1)) == int(nelem):
                traffic, traffic_src = float(tj["dram_bytes_per_launch"]), tj.get("source")
        except (OSError, ValueError, KeyError):
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": f"MEASURED_PEAKS.json ({peak_kind})", "kernel": "k_q4_stiffness<false,false,EmitRuns>",
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_element": alg_bytes,
                "kernel_share_of_step": kernel_ms / ms_per_step}
    # north_star: "the slower of FP64 peak and HBM bytes at peak bandwidth"
    t_hbm = alg_bytes * nelem / (hbm_peak * 1e9)
    t_fp64 = Q4_FLOPS * nelem / (fp64_peak * 1e12)
    roofline["north_star"] = {"binding": "fp64" if t_fp64 >= t_hbm else "hbm", "t_roof_ms": max(t_hbm, t_fp64) * 1e3,
                              "frac": max(t_hbm, t_fp64) * 1e3 / kernel_ms}
    fl = Q4_FLOPS * nelem / (kernel_ms * 1e-3) / 1e12
    fp64 = {"achieved_tflops": fl, "peak_tflops": fp64_peak, "frac": fl / fp64_peak, "flops_per_element": Q4_FLOPS,
            "peak_source": "fsgpu_measure_peaks DFMA micro-kernel on this device", "copy_gbs_this_device": copy_bw}

    cpu = None
    if not args.no_cpu_baseline:
        ncores = os.cpu_count() or 1
        normals, valid = nrm_host, val_host
        s1 = min(nelem, 20000)
        v1, t1 = cpu_q4rs_assembly(w, s1, 1, normals, valid)
        sN = min(nelem, max(40000, 8000 * ncores))
        vN, tN = cpu_q4rs_assembly(w, sN, ncores, normals, valid)
        cpu = {"value": vN, "unit": unit, "cores": ncores, "kind": "port",
               "sample": f"first {sN} elements (all {ncores} threads, {tN:.1f} s); single-thread = reference behaviour: "
                         f"{v1:.0f} elements/s on the first {s1} elements ({t1:.1f} s)",
               "single_thread_value": v1}

    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h_link),
                    "host_result_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3, "includes": "H2D mesh+dofs+normals, symbolic phase, numeric phase, D2H of the CSC result into host colptr+rowval+nzval (Int64/f64; the row indices cross PCIe run-length coded and are expanded by 8 host threads inside the call)",
                    "pinned_d2h_gbs_this_box": d2h_gbs,
                    "values_refresh_ms_same_pattern": refresh_ms},
            "gpu_launches": int(launches), "roofline": roofline, "fp64": fp64, "cpu_baseline": cpu,
            "symbolic_ms": {"first": symbolic_ms, "warm": symbolic_warm_ms}, "nnz": int(nnz), "other_workloads": extras}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
