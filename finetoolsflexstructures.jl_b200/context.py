"""Low-level object wrapper of the fsgpu C ABI: one `Context` per (mesh, device)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from ._lib import BeamParams, ShellParams, check, f64, i64, lib, ptr


@dataclass
class SparseMatrixCSC:
    """What Julia's `SparseMatrixCSC{Float64,Int64}` holds: 1-based colptr/rowval."""

    m: int
    n: int
    colptr: np.ndarray
    rowval: np.ndarray
    nzval: np.ndarray
    csr: bool = False  # CSR_SYMM target: colptr is rowptr, rowval is colval

    def to_scipy(self):
        import scipy.sparse as sp

        cls = sp.csr_matrix if self.csr else sp.csc_matrix
        return cls((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.m, self.n))


class Context:
    def __init__(self, device=0):
        h = C.c_void_p()
        check(lib.fsgpu_create(C.byref(h), device))
        self._h = h
        self.target = None

    def close(self):
        if self._h:
            lib.fsgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data ------------------------------------------------------------------------
    def set_mesh(self, conn, xyz):
        """conn: (nelem, nnpe) 1-based ints (row e = fes.conn[e]); xyz: (nnodes, 3)."""
        conn = np.ascontiguousarray(conn, dtype=np.int64)  # C order == nnpe x nelem column-major
        xyz = f64(xyz, "F")
        self.nelem, self.nnpe = conn.shape
        self.nnodes = xyz.shape[0]
        check(lib.fsgpu_set_mesh(self._h, self.nnpe, self.nelem, ptr(conn), self.nnodes, ptr(xyz)))

    def set_dofnums(self, dofnums, nfree, nall=None):
        d = i64(dofnums, "F")
        self.nfree = int(nfree)
        self.nall = int(d.size if nall is None else nall)
        check(lib.fsgpu_set_dofnums(self._h, ptr(d), self.nfree, self.nall))

    def set_normals(self, normals, valid):
        n = f64(normals, "F")
        v = np.ascontiguousarray(valid, dtype=np.uint8)
        check(lib.fsgpu_set_normals(self._h, ptr(n), ptr(v)))

    def associategeometry(self, threshold_angle=30.0, fixed_dir=None, accumulate=False):
        fd = None if fixed_dir is None else f64(fixed_dir)
        check(lib.fsgpu_associategeometry(self._h, float(threshold_angle), ptr(fd), 1 if accumulate else 0))

    def associategeometry_dirs(self, dirs, threshold_angle=30.0, accumulate=False):
        """dirs: (nelem, nnpe, 3) csys normal direction per element and node."""
        d = np.ascontiguousarray(np.asarray(dirs, dtype=np.float64).reshape(self.nelem, self.nnpe, 3))
        check(lib.fsgpu_associategeometry_dirs(self._h, float(threshold_angle), ptr(d), 1 if accumulate else 0))

    def associategeometry_csys(self, kind, origin, axis, threshold_angle=30.0, accumulate=False):
        """Nodal normals from a built-in csys kind (L.CSYS_*), evaluated on the device."""
        o = None if origin is None else f64(origin)
        check(lib.fsgpu_associategeometry_csys(self._h, float(threshold_angle), int(kind), ptr(o), ptr(f64(axis)), 1 if accumulate else 0))

    def set_layup_csys(self, kind, origin, axis):
        """Layup csys matrices from a built-in csys kind, evaluated on the device (after set_layup)."""
        o = None if origin is None else f64(origin)
        check(lib.fsgpu_set_layup_csys(self._h, int(kind), ptr(o), ptr(f64(axis))))

    def normals_accumulate(self, fixed_dir=None, accumulate=False):
        fd = None if fixed_dir is None else f64(fixed_dir)
        p = C.c_void_p()
        check(lib.fsgpu_normals_accumulate(self._h, ptr(fd), 1 if accumulate else 0, C.byref(p)))
        return p.value  # device address of the [nnodes][3] sums

    def normals_finish(self, threshold_angle=30.0, fixed_dir=None):
        fd = None if fixed_dir is None else f64(fixed_dir)
        p = C.c_void_p()
        check(lib.fsgpu_normals_finish(self._h, float(threshold_angle), ptr(fd), C.byref(p)))
        return p.value  # device address of the packed [nnodes][4] normals (4th = valid flag)

    def get_normals(self):
        n = np.zeros((self.nnodes, 3), order="F")
        v = np.zeros(self.nnodes, dtype=np.uint8)
        check(lib.fsgpu_get_normals(self._h, ptr(n), ptr(v)))
        return n, v.astype(bool)

    def set_thickness(self, t):
        t = f64(np.atleast_1d(t), "C").ravel()
        check(lib.fsgpu_set_thickness(self._h, ptr(t), t.size))

    def set_stab_factor(self, f):
        if f is None:
            check(lib.fsgpu_set_stab_factor(self._h, None, 0))
        else:
            f = f64(f, "C").ravel()
            check(lib.fsgpu_set_stab_factor(self._h, ptr(f), f.size))

    def element_sizes(self):
        h = np.zeros(self.nelem)
        check(lib.fsgpu_element_sizes(self._h, ptr(h)))
        return h

    def set_rule(self, pc, w):
        pc = np.asarray(pc, dtype=np.float64)
        xi = np.ascontiguousarray(pc[:, 0])
        eta = np.ascontiguousarray(pc[:, 1])
        w = np.ascontiguousarray(w, dtype=np.float64)
        self.npts = len(w)
        check(lib.fsgpu_set_rule(self._h, len(w), ptr(xi), ptr(eta), ptr(w)))

    def set_layup(self, group_data, group_of_elem, csmat):
        """group_data (ngroups, 34); group_of_elem (nelem,) 1-based or None; csmat (..., 3, 3)."""
        g = f64(group_data, "C").reshape(-1, 34)
        go = None if group_of_elem is None else i64(group_of_elem, "C")
        cs = np.asarray(csmat, dtype=np.float64).reshape(-1, 3, 3)
        cs_cm = np.ascontiguousarray(np.transpose(cs, (0, 2, 1)))  # each 3x3 column-major
        check(lib.fsgpu_set_layup(self._h, g.shape[0], ptr(g), ptr(go), ptr(cs_cm), cs.shape[0]))

    def set_beam_sections(self, A, I1, I2, I3, J, A2s, A3s, x1x2):
        arrs = [f64(a, "C").ravel() for a in (A, I1, I2, I3, J, A2s, A3s)]
        xx = np.ascontiguousarray(np.asarray(x1x2, dtype=np.float64).reshape(-1, 3))  # 3 x nelem column-major
        check(lib.fsgpu_set_beam_sections(self._h, *[ptr(a) for a in arrs], ptr(xx)))

    def set_state(self, u1, Rfield1):
        u = f64(u1, "F")
        R = f64(Rfield1, "F")
        check(lib.fsgpu_set_state(self._h, ptr(u), ptr(R)))

    def set_velocity(self, v1):
        v = f64(v1, "F")
        check(lib.fsgpu_set_velocity(self._h, ptr(v)))

    def beam_distribloads(self, params, force, nfree_only=False):
        fo = np.ascontiguousarray(np.asarray(force, dtype=np.float64).reshape(-1, 3))
        check(lib.fsgpu_corotbeam_distribloads(self._h, C.byref(params), ptr(fo), fo.shape[0], 1 if nfree_only else 0))

    def set_stream(self, cuda_stream):
        check(lib.fsgpu_set_stream(self._h, C.c_void_p(cuda_stream)))

    def sync(self):
        check(lib.fsgpu_sync(self._h))

    @property
    def launch_count(self):
        return int(lib.fsgpu_launch_count(self._h))

    @property
    def last_kernel_ms(self):
        ms = C.c_double()
        check(lib.fsgpu_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    @property
    def d2h_bytes(self):
        v = C.c_int64()
        check(lib.fsgpu_d2h_bytes(self._h, C.byref(v)))
        return v.value

    def set_deterministic(self, on=True):
        """Prefer the atomics-free T3 tile kernel (bitwise reproducible); effective at the next symbolic phase."""
        check(lib.fsgpu_set_deterministic(self._h, 1 if on else 0))

    @property
    def scatter_path(self):
        """0 slot map + RED, 1 run-structured + RED, 2 owner-computes tile, -1 none yet."""
        v = C.c_int(-1)
        check(lib.fsgpu_scatter_path(self._h, C.byref(v)))
        return v.value

    def measure_peaks(self):
        f, b = C.c_double(), C.c_double()
        check(lib.fsgpu_measure_peaks(self._h, C.byref(f), C.byref(b)))
        return f.value, b.value

    # ---- symbolic / numeric -------------------------------------------------------------
    def symbolic(self, target):
        nr, nc, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.fsgpu_symbolic(self._h, int(target), C.byref(nr), C.byref(nc), C.byref(nnz)))
        self.target = int(target)
        return nr.value, nc.value, nnz.value

    def shell_op(self, name, params: ShellParams):
        check(getattr(lib, f"fsgpu_{name}")(self._h, C.byref(params)))

    def beam_op(self, name, params: BeamParams, *extra):
        check(getattr(lib, f"fsgpu_corotbeam_{name}")(self._h, C.byref(params), *extra))

    def shell_mass_diag(self, params, kind, nfree_only=False):
        check(lib.fsgpu_shell_mass_diag(self._h, C.byref(params), kind, 1 if nfree_only else 0))

    def result_size(self):
        nr, nc, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.fsgpu_result_size(self._h, C.byref(nr), C.byref(nc), C.byref(nnz)))
        return nr.value, nc.value, nnz.value

    def fetch_matrix(self, out=None):
        """makematrix!: returns SparseMatrixCSC (Julia-layout arrays).  `out`: optional
        preallocated (colptr, rowval, nzval) numpy arrays (e.g. views of pinned memory)."""
        m, n, nnz = self.result_size()
        if out is None:
            colptr = np.empty(n + 1, dtype=np.int64)
            rowval = np.empty(nnz, dtype=np.int64)
            nzval = np.empty(nnz, dtype=np.float64)
        else:
            colptr, rowval, nzval = out
        check(lib.fsgpu_fetch_matrix(self._h, ptr(colptr), ptr(rowval), ptr(nzval)))
        return SparseMatrixCSC(m, n, colptr, rowval, nzval, csr=(self.target == L.CSR_SYMM))

    def result_size_uplo(self, uplo):
        nnz = C.c_int64()
        check(lib.fsgpu_result_size_uplo(self._h, ord(uplo), C.byref(nnz)))
        return nnz.value

    def fetch_matrix_uplo(self, uplo="L", out=None):
        """One triangle (diagonal included) of a square result, e.g. for `cholesky(Symmetric(K, :L))`:
        half the bytes of fetch_matrix cross PCIe.  `out`: optional (colptr, rowval, nzval) arrays,
        rowval / nzval at least result_size_uplo(uplo) long."""
        m, n, _ = self.result_size()
        if out is None:
            nnz = self.result_size_uplo(uplo)
            colptr = np.empty(n + 1, dtype=np.int64)
            rowval = np.empty(nnz, dtype=np.int64)
            nzval = np.empty(nnz, dtype=np.float64)
        else:
            colptr, rowval, nzval = out
        check(lib.fsgpu_fetch_matrix_uplo(self._h, ord(uplo), ptr(colptr), ptr(rowval), ptr(nzval)))
        nnz = int(colptr[n]) - 1
        return SparseMatrixCSC(m, n, colptr, rowval[:nnz], nzval[:nnz])

    def result_block(self, col_lo, col_hi, row_map=None, colcount=None, rowval=None, nzval=None):
        """Columns [col_lo, col_hi) of the device-resident result as pieces of the global CSC, written into
        DEVICE buffers (torch tensors or raw addresses; None skips); returns the number of stored entries."""
        nb = C.c_int64()
        check(lib.fsgpu_result_block(self._h, int(col_lo), int(col_hi), ptr(row_map), C.byref(nb), ptr(colcount), ptr(rowval), ptr(nzval)))
        return nb.value

    def vector_device(self):
        """(device address, length) of the last vector result."""
        vp, vn = C.c_void_p(), C.c_int64()
        check(lib.fsgpu_vector_device(self._h, C.byref(vp), C.byref(vn)))
        return vp.value, vn.value

    def fetch_values(self, nzval):
        check(lib.fsgpu_fetch_matrix(self._h, None, None, ptr(nzval)))
        return nzval

    def fetch_vector(self, n):
        out = np.empty(n)
        check(lib.fsgpu_fetch_vector(self._h, ptr(out), n))
        return out

    def element_matrices(self, kind, op, params):
        n = 6 * self.nnpe
        out = np.zeros((self.nelem, n, n))  # [e][col][row]
        check(lib.fsgpu_element_matrices(self._h, kind, op, C.byref(params), ptr(out)))
        return np.transpose(out, (0, 2, 1))  # -> [e][row][col]

    def element_vectors(self, params):
        out = np.zeros((self.nelem, 12))
        check(lib.fsgpu_element_vectors(self._h, C.byref(params), ptr(out)))
        return out

    def shell_resultants(self, params, kind, quantity, u, outputcsys=None, npts=1):
        uu = f64(u, "F")
        out = np.zeros((self.nelem, npts, 3))
        if outputcsys is None:
            cs, ncs = None, 0
        else:
            cs = np.asarray(outputcsys, dtype=np.float64).reshape(-1, 3, 3)
            ncs = cs.shape[0]
            cs = np.ascontiguousarray(np.transpose(cs, (0, 2, 1)))
        check(lib.fsgpu_shell_resultants(self._h, C.byref(params), kind, quantity, ptr(uu), ptr(cs), ncs, ptr(out)))
        return out

    def shell_nodal_field(self, params, kind, quantity, u, outputcsys=None):
        """fieldfromintegpoints on the device: (nnodes, 3) nodal means of the resultants (inverse squared distance)."""
        uu = f64(u, "F")
        out = np.zeros((self.nnodes, 3), order="F")
        if outputcsys is None:
            cs, ncs = None, 0
        else:
            cs = np.asarray(outputcsys, dtype=np.float64).reshape(-1, 3, 3)
            ncs = cs.shape[0]
            cs = np.ascontiguousarray(np.transpose(cs, (0, 2, 1)))
        check(lib.fsgpu_shell_nodal_field(self._h, C.byref(params), kind, quantity, ptr(uu), ptr(cs), ncs, ptr(out)))
        return out

    def update_rotation_field(self, dchi_values):
        d = f64(dchi_values, "F")
        out = np.zeros((self.nnodes, 9), order="F")
        check(lib.fsgpu_update_rotation_field(self._h, ptr(d), ptr(out)))
        return out

    def coo_to_csc(self, I, J, V, m, n):
        I, J, V = i64(I, "C"), i64(J, "C"), f64(V, "C")
        nnz = C.c_int64()
        check(lib.fsgpu_coo_to_csc(self._h, m, n, I.size, ptr(I), ptr(J), ptr(V), C.byref(nnz), None, None, None))
        colptr = np.empty(n + 1, dtype=np.int64)
        rowval = np.empty(nnz.value, dtype=np.int64)
        nzval = np.empty(nnz.value)
        check(lib.fsgpu_coo_to_csc(self._h, m, n, I.size, ptr(I), ptr(J), ptr(V), C.byref(nnz), ptr(colptr), ptr(rowval), ptr(nzval)))
        return SparseMatrixCSC(m, n, colptr, rowval, nzval)


class Explicit:
    """Device-resident central-difference integrator (fsgpu_explicit_*)."""

    def __init__(self, ctx: Context, K=None, mdiag=None, c_scale=0.0, dt=0.0):
        self.ctx = ctx
        h = C.c_void_p()
        if K is None:
            check(lib.fsgpu_explicit_create_from_ctx(C.byref(h), ctx._h, float(c_scale), float(dt)))
            self.n = ctx.result_size()[0]
        else:
            rowptr, colval, nzval = K  # Int64 1-based CSR
            rowptr, colval, nzval = i64(rowptr, "C"), i64(colval, "C"), f64(nzval, "C")
            md = f64(mdiag, "C")
            self.n = md.size
            check(lib.fsgpu_explicit_create(C.byref(h), ctx._h, self.n, ptr(rowptr), ptr(colval), ptr(nzval), ptr(md), float(c_scale), float(dt)))
        self._h = h

    @classmethod
    def create_dist(cls, ctx: Context, rank, world, row_lo, row_hi, loc2glob, bounds, c_scale=0.0, dt=0.0):
        """Row-partitioned run (fsgpu_explicit_create_dist): this rank owns local rows [row_lo, row_hi) of the
        FFBLOCK result of `ctx` = global rows [bounds[rank], bounds[rank+1]).  Follow with `export()` on every
        rank, an all-gather of the blobs by the host, and `connect(blobs)`; after that every call is collective."""
        self = cls.__new__(cls)
        self.ctx = ctx
        l2g = None if loc2glob is None else i64(loc2glob, "C")
        b = i64(bounds, "C")
        h = C.c_void_p()
        check(lib.fsgpu_explicit_create_dist(C.byref(h), ctx._h, int(rank), int(world), int(row_lo), int(row_hi), ptr(l2g), ptr(b),
                                             float(c_scale), float(dt)))
        self._h = h
        self.n = int(row_hi) - int(row_lo)
        self.rank, self.world = int(rank), int(world)
        return self

    def export(self):
        buf = np.zeros(L.EXPLICIT_BLOB_BYTES, dtype=np.uint8)
        check(lib.fsgpu_explicit_export(self._h, ptr(buf)))
        return buf

    def connect(self, blobs):
        """blobs: the exports of all ranks, in rank order (sequence of uint8 arrays, or one (world, BLOB) array)."""
        b = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.uint8).ravel() for x in blobs]))
        assert b.size == self.world * L.EXPLICIT_BLOB_BYTES
        check(lib.fsgpu_explicit_connect(self._h, ptr(b)))

    def dist_info(self):
        """(own rows, halo entries, entries pushed per step, boundary runs, neighbouring ranks)."""
        v = [C.c_int64() for _ in range(4)]
        k = C.c_int32()
        check(lib.fsgpu_explicit_dist_info(self._h, *[C.byref(x) for x in v], C.byref(k)))
        return tuple(x.value for x in v) + (k.value,)

    def close(self):
        if self._h:
            lib.fsgpu_explicit_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, U0=None, V0=None):
        U0 = None if U0 is None else f64(U0, "C")
        V0 = None if V0 is None else f64(V0, "C")
        check(lib.fsgpu_explicit_set_state(self._h, ptr(U0), ptr(V0)))

    def set_timestep(self, c_scale, dt):
        check(lib.fsgpu_explicit_set_timestep(self._h, float(c_scale), float(dt)))

    def set_load(self, F0):
        F0 = None if F0 is None else f64(F0, "C")
        check(lib.fsgpu_explicit_set_load(self._h, ptr(F0)))

    def start(self, fscale0=1.0):
        check(lib.fsgpu_explicit_start(self._h, float(fscale0)))

    def step(self, nsteps, fscale=None):
        fs = None if fscale is None else f64(fscale, "C")
        check(lib.fsgpu_explicit_step(self._h, int(nsteps), ptr(fs)))

    def run(self, nsteps, dt, force=None, F0=None, fscale=None, peek=None, nbtw=0):
        """The loop of plate_expl_examples.jl:61-94 with its two closures.
        `force(t) -> F` general (`force!(F, t)`, :86: any spatial distribution per step): the load vector is re-sent
        every step.  The separable form F(t) = fscale(t) * F0 stays on the device: the factors of a whole stretch
        are sampled into a table and the stretch runs without host interaction.  `peek(step, U, V, t)` (:92) is
        called with the fetched state at step 0 and every `nbtw` steps.  Returns (U, V, A)."""
        if force is not None:
            self.set_load(force(0.0))
            self.start(1.0)
        else:
            self.set_load(F0)
            self.start(1.0 if fscale is None else float(fscale(0.0)))
        if peek is not None:
            U, V, _ = self.get_state()
            peek(0, U, V, 0.0)
        step = 0
        while step < nsteps:
            if force is not None:
                m = 1
                self.set_load(force((step + 1) * dt))
                self.step(1, [1.0])
            else:
                m = min(nbtw - step % nbtw, nsteps - step) if nbtw > 0 else nsteps - step
                self.step(m, None if fscale is None else [float(fscale((step + k) * dt)) for k in range(1, m + 1)])
            step += m
            if peek is not None and nbtw > 0 and step % nbtw == 0:
                U, V, _ = self.get_state()
                peek(step, U, V, step * dt)
        return self.get_state()

    def step_begin(self):
        check(lib.fsgpu_explicit_step_begin(self._h))

    def step_end(self, fscale=1.0):
        check(lib.fsgpu_explicit_step_end(self._h, float(fscale)))

    def get_state(self):
        U, V, A = np.empty(self.n), np.empty(self.n), np.empty(self.n)
        check(lib.fsgpu_explicit_get_state(self._h, ptr(U), ptr(V), ptr(A)))
        return U, V, A

    def spmv(self, x):
        x = f64(x, "C")
        y = np.empty(self.n)
        check(lib.fsgpu_explicit_spmv(self._h, ptr(x), ptr(y)))
        return y

    def omega_max_sq(self, maxit=30):
        lam = C.c_double()
        check(lib.fsgpu_explicit_omega_max(self._h, int(maxit), C.byref(lam)))
        return lam.value

    def layout(self):
        """(rows, stored entries, runs of rows sharing one column pattern, column indices read per SpMV)."""
        v = [C.c_int64() for _ in range(4)]
        check(lib.fsgpu_explicit_layout(self._h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def kinetic_energy(self):
        ke = C.c_double()
        check(lib.fsgpu_explicit_kinetic_energy(self._h, C.byref(ke)))
        return ke.value

    def device_state(self):
        p = [C.c_void_p() for _ in range(4)]
        check(lib.fsgpu_explicit_device_state(self._h, *[C.byref(x) for x in p]))
        return tuple(x.value for x in p)  # U, V, A, E device addresses
