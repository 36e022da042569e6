"""Host-side mirror of the reference's operator API for the GPU path.

The reference host stays in Julia (julia/FlexStructuresGPU.jl is the `ccall` glue); this
module is the same thin layer in Python, because the build image has no Julia.  Names,
argument order and error behaviour follow the reference:

    femm = FEMMShellT3FF(integdomain, material)            # make(...)
    associategeometry(femm, geom0)                          # associategeometry!
    K = stiffness(femm, assembler, geom0, u1, Rfield1, dchi)
    M = mass(femm, assembler, geom0, dchi)
    F = restoringforce(femm, vassembler, geom0, u1, Rfield1, dchi)

(src/FEMMShellT3FFModule.jl:255,570,635,757; src/FEMMShellQ4RSModule.jl:234,472,877,968;
src/FEMMShellT3FFCompModule.jl:199,489,561,710; src/FEMMShellQ4RSCompModule.jl:240,447,861,979;
src/FEMMCorotBeamModule.jl:810,972,1042,1112.)  Each operator is ONE call into libfsgpu
per reference operator call; there is no per-element host loop and no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib as L
from ._lib import BeamParams, FsgpuError, ShellParams
from .context import Context, SparseMatrixCSC

# ---- FinEtools-like host objects (plain arrays, as in the reference) ----------------------


class NodalField:
    """`NodalField`: values (nnodes x ndof), dofnums, is_fixed; `numberdofs!` numbers free
    dofs first, node-major, prescribed dofs after (FinEtools >= 7 convention)."""

    def __init__(self, values):
        self.values = np.array(values, dtype=np.float64, order="F")
        self.is_fixed = np.zeros(self.values.shape, dtype=bool)
        self.dofnums = np.zeros(self.values.shape, dtype=np.int64, order="F")
        self._nfree = 0
        self._version = 0  # bumped by every call that edits the field in place: the FEMMs' upload caches key on it

    def touch(self):
        """Call after editing `values` / `dofnums` in place by hand: operators then re-upload this field."""
        self._version = getattr(self, "_version", 0) + 1

    def setebc(self, nodes, comp):
        self.is_fixed[np.asarray(nodes, dtype=np.int64), comp - 1] = True
        self.touch()

    def numberdofs(self, perm=None):
        nn, nd = self.is_fixed.shape
        order = np.arange(nn) if perm is None else np.asarray(perm, dtype=np.int64)
        free = ~self.is_fixed[order].ravel()
        nums = np.empty(nn * nd, dtype=np.int64)
        self._nfree = int(free.sum())
        nums[free] = np.arange(1, self._nfree + 1)
        nums[~free] = np.arange(self._nfree + 1, nn * nd + 1)
        self.dofnums[order] = nums.reshape(nn, nd)
        self.touch()
        return self


def nfreedofs(f):
    return f._nfree


def nalldofs(f):
    return f.dofnums.size


@dataclass
class MatDeforElastIso:
    E: float
    nu: float
    rho: float = 0.0

    def moduli(self):
        lam = self.E * self.nu / (1 + self.nu) / (1 - 2 * self.nu)
        mu = self.E / 2.0 / (1 + self.nu)
        D = np.zeros((6, 6))
        D[:3, :3] = lam
        D[np.arange(3), np.arange(3)] += 2 * mu
        D[3, 3] = D[4, 4] = D[5, 5] = mu
        return D


@dataclass
class MatDeforElastOrtho:
    E1: float
    E2: float
    E3: float
    nu12: float
    nu13: float
    nu23: float
    G12: float
    G13: float
    G23: float
    rho: float = 0.0

    def moduli(self):
        S = np.zeros((6, 6))
        S[0, 0], S[1, 1], S[2, 2] = 1 / self.E1, 1 / self.E2, 1 / self.E3
        S[0, 1] = S[1, 0] = -self.nu12 / self.E1
        S[0, 2] = S[2, 0] = -self.nu13 / self.E1
        S[1, 2] = S[2, 1] = -self.nu23 / self.E2
        S[3, 3], S[4, 4], S[5, 5] = 1 / self.G12, 1 / self.G13, 1 / self.G23
        return np.linalg.inv(S)


def _shell_material_stiffness(D6):
    """Plane-stress / transverse-shear reduction (src/FEMMShellT3FFModule.jl:323-342)."""
    Dps = np.zeros((3, 3))
    Dps[:2, :2] = D6[:2, :2] - np.outer(D6[:2, 2], D6[2, :2]) / D6[2, 2]
    for i, k in enumerate((0, 1, 3)):
        Dps[2, i] = Dps[i, 2] = D6[3, k]
    Dt = np.diag([D6[4, 4], D6[5, 5]])
    return Dps, Dt


def lamina_material(*a):
    """`lamina_material` overloads (src/CompositeLayupModule.jl:105-167)."""
    if len(a) == 6:
        E1, E2, nu12, G12, G13, G23 = a
        return MatDeforElastOrtho(E1, E2, E2, nu12, nu12, 0.0, G12, G13, G23, 0.0)
    if len(a) == 7:
        rho, E1, E2, nu12, G12, G13, G23 = a
        return MatDeforElastOrtho(E1, E2, E2, nu12, nu12, 0.0, G12, G13, G23, rho)
    if len(a) == 2:
        return MatDeforElastIso(a[0], a[1], 0.0)
    rho, E, nu = a
    return MatDeforElastIso(E, nu, rho)


@dataclass
class Ply:
    name: str
    material: object
    thickness: float
    angle: float  # degrees


def cartesian_csys(axes):
    """Constant layup csys matrix (src/CompositeLayupModule.jl:26-44)."""
    M = np.zeros((3, 3))
    for j in range(3):
        M[abs(axes[j]) - 1, j] = 1.0 if axes[j] > 0 else -1.0
    return M


@dataclass
class CompositeLayup:
    """Layup + the O(nplies) through-thickness integration that the reference runs once per
    operator call per layup group (src/CompositeLayupModule.jl:232-330)."""

    name: str
    plies: list
    csys: object  # 3x3 matrix, or callable(centroids (n,3)) -> (n,3,3)
    offset: float = 0.0
    transverse_shear_constant: float | None = None  # None -> Vinson-Sierakowski

    def thickness(self):
        return sum(p.thickness for p in self.plies)

    def group_record(self):
        t = self.thickness()
        A, B, D, H = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((2, 2))
        md = mi = 0.0
        zs = -t / 2 - self.offset
        for p in self.plies:
            ze = zs + p.thickness
            Dps, Dts = _shell_material_stiffness(p.material.moduli())
            a = p.angle / 180 * np.pi
            m, n = np.cos(-a), np.sin(-a)
            T = np.array([[m * m, n * n, 2 * m * n], [n * n, m * m, -2 * m * n], [-m * n, m * n, m * m - n * n]])
            Q = T @ (Dps @ T.T)
            A += (ze - zs) * Q
            B += (ze**2 - zs**2) / 2 * Q
            D += (ze**3 - zs**3) / 3 * Q
            m, n = np.cos(a), np.sin(a)
            Tt = np.array([[m, -n], [n, m]])
            Qt = Tt.T @ (Dts @ Tt)
            if self.transverse_shear_constant is None:
                H += 5 / 4 * (ze - zs - 4 / 3 * (ze**3 - zs**3) / t**2) * Qt
            else:
                H += self.transverse_shear_constant * (ze - zs) * Qt
            md += (ze - zs) * p.material.rho
            mi += (ze**3 - zs**3) * p.material.rho / 3
            zs += p.thickness
        return np.concatenate([A.ravel(), B.ravel(), D.ravel(), H.ravel(), [t, md, mi]])


@dataclass
class IntegDomain:
    """fes.conn + integration rule + `otherdimension` (thickness: number, per-element array,
    or callable(loc (n,3)) -> (n,))."""

    conn: np.ndarray  # (nelem, nnpe) 1-based
    rule: tuple | None = None  # (pc (npts,2), w (npts,)) for Q4
    otherdimension: object = 1.0


def GaussRule2x2():
    g = 0.577350269189626
    return np.array([[-g, -g], [-g, g], [g, -g], [g, g]]), np.ones(4)


def Simpson13Rule2():
    """Simpson13Rule(2): 3x3 tensor product, first coordinate outer loop, 1-D weights (1/3, 4/3, 1/3)
    (used by examples/shells/statics/homogeneous/plates/clamped_square_plate_udl_examples.jl)."""
    p1 = np.array([-1.0, 0.0, 1.0])
    w1 = np.array([1.0, 4.0, 1.0]) / 3.0
    pc = np.array([[p1[i], p1[j]] for i in range(3) for j in range(3)])
    w = np.array([w1[i] * w1[j] for i in range(3) for j in range(3)])
    return pc, w


def GaussRule1x1():
    return np.array([[0.0, 0.0]]), np.array([4.0])


# ---- assemblers: the plugin point (a new assembler type selects the GPU path) ----------------


@dataclass
class SysmatAssemblerGPU:
    """GPU assembler carrying the semantic target of a FinEtools assembler
    (SURVEY section 8(b)): kind in 'sparse' | 'symm' | 'diag' | 'ffblock' | 'ffblock_diag' | 'csrsymm'."""

    kind: str = "symm"
    uplo: str = ""  # "L" / "U": makematrix! returns that triangle only (for Symmetric(K, :L) consumers; half the PCIe bytes)
    TARGETS = {
        "sparse": L.SPARSE,
        "symm": L.SPARSE_SYMM,
        "diag": L.SPARSE_DIAG,
        "ffblock": L.FFBLOCK,
        "ffblock_diag": L.FFBLOCK_DIAG,
        "csrsymm": L.CSR_SYMM,
    }

    @property
    def target(self):
        return self.TARGETS[self.kind]


def SysmatAssemblerSparse():
    return SysmatAssemblerGPU("sparse")


def SysmatAssemblerSparseSymm():
    return SysmatAssemblerGPU("symm")


def SysmatAssemblerSparseDiag():
    return SysmatAssemblerGPU("diag")


def SysmatAssemblerFFBlock(inner=None):
    return SysmatAssemblerGPU("ffblock_diag" if (inner is not None and inner.kind == "diag") else "ffblock")


def SysmatAssemblerSparseCSRSymm():
    return SysmatAssemblerGPU("csrsymm")


@dataclass
class SysvecAssemblerGPU:
    nfree_only: bool = False


def SysvecAssembler():
    return SysvecAssemblerGPU(False)


def SysvecAssemblerFBlock():
    return SysvecAssemblerGPU(True)


# ---- FEMMs -------------------------------------------------------------------------------------


class CSysKind:
    """A built-in coordinate system the library evaluates ON THE DEVICE (`fsgpu_associategeometry_csys`,
    `fsgpu_set_layup_csys`): the CSys callbacks of the reference's examples -- `cylindrical!`
    (clamp_cyl_expl_examples.jl:62-68), `spherical!` (hemisphere_examples.jl:31-39), the surface-normal variant of
    pressurized_cylinder_free_examples.jl:16-23.  Calling the object evaluates the same formulas with NumPy
    (host-side consumers, tests)."""

    def __init__(self, kind, axis, origin=None):
        self.kind, self.axis = int(kind), np.asarray(axis, dtype=np.float64) / np.linalg.norm(axis)
        self.origin = np.zeros(3) if origin is None else np.asarray(origin, dtype=np.float64)

    @classmethod
    def cylindrical(cls, axis=(0.0, 1.0, 0.0), origin=None):
        return cls(L.CSYS_CYLINDRICAL, axis, origin)

    @classmethod
    def spherical(cls, axis=(0.0, 0.0, 1.0), origin=None):
        return cls(L.CSYS_SPHERICAL, axis, origin)

    @classmethod
    def normal_axis(cls, axis=(0.0, 1.0, 0.0)):
        return cls(L.CSYS_NORMAL_AXIS, axis, None)

    def __call__(self, XYZ, tangents=None, feid=None, qpid=None):
        X = np.asarray(XYZ, dtype=np.float64)[:, :3]
        a = np.broadcast_to(self.axis, X.shape)
        r = X - self.origin
        unit = lambda v: v / np.linalg.norm(v, axis=1, keepdims=True)
        if self.kind == L.CSYS_CYLINDRICAL:
            e3 = unit(r - np.sum(r * a, axis=1, keepdims=True) * a)
            e2 = a
            e1 = np.cross(e2, e3)
        elif self.kind == L.CSYS_SPHERICAL:
            e3 = unit(r)
            e1 = unit(np.cross(a, e3))
            e2 = np.cross(e3, e1)
        else:
            e3 = unit(np.cross(tangents[:, :, 0], tangents[:, :, 1]))
            e2 = a
            e1 = np.cross(e2, e3)
        return np.stack([e1, e2, e3], axis=-1)


_Q4_NODE_PC = ((-1.0, -1.0), (1.0, -1.0), (1.0, 1.0), (-1.0, 1.0))  # NodalTensorProductRule(2): the nodes, in node order


def _q4_shape(xi, eta):
    """Q4 shape functions and parametric derivatives (FinEtools FESetQ4: nodes (-1,-1), (1,-1), (1,1), (-1,1))."""
    N = 0.25 * np.array([(1 - xi) * (1 - eta), (1 + xi) * (1 - eta), (1 + xi) * (1 + eta), (1 - xi) * (1 + eta)])
    dN = 0.25 * np.array([[-(1 - eta), -(1 - xi)], [(1 - eta), -(1 + xi)], [(1 + eta), (1 + xi)], [-(1 + eta), (1 - xi)]])
    return N, dN


def _eval_csys(cs, XYZ, tangents, feid, qpid, elems=None):
    """FinEtools `updatecsmat!(csys, XYZ, tangents, feid, qpid)`, batched: a constant matrix is broadcast, a callable
    gets the locations (n, k) and -- if it accepts them -- tangents (n, 3, 2), feid, qpid; returns (n, 3, 3)."""
    n = XYZ.shape[0]
    if not callable(cs):
        c = np.asarray(cs, dtype=np.float64)
        return np.broadcast_to(c, (n, 3, 3)) if c.ndim == 2 else c.reshape(-1, 3, 3)[elems]  # per-element matrices
    import inspect

    try:
        npar = len(inspect.signature(cs).parameters)
    except (TypeError, ValueError):
        npar = 1
    out = cs(XYZ) if npar < 2 else cs(XYZ, tangents, feid, qpid)
    return np.asarray(out, dtype=np.float64).reshape(n, 3, 3)


class _FEMMBase:
    _nnpe = 0

    def __init__(self, integdomain, device=0):
        self.integdomain = integdomain
        self.ctx = Context(device)
        self._associatedgeometry = False
        self._mesh_key = None
        self._conn_key = None
        self._dof_key = None
        self._sym_key = None

    # upload mesh / dofs once per distinct array (the Julia glue keys on objectid)
    @staticmethod
    def _field_key(f, arr):
        # identity of the object (held strongly, so the id cannot be recycled), its edit counter, and the buffer the
        # device copy was made from: numberdofs!/setebc! on the same field, or a swapped array, invalidate the copy
        return (f, getattr(f, "_version", 0), None if arr is None else (arr.__array_interface__["data"][0], arr.shape))

    @staticmethod
    def _same_key(a, b):
        return a is not None and b is not None and a[0] is b[0] and a[1:] == b[1:]

    def _sync_mesh(self, geom0):
        conn_obj = self.integdomain.conn
        key = self._field_key(geom0, np.asarray(geom0.values))
        if not (self._same_key(self._mesh_key, key) and self._conn_key is conn_obj):
            self._conn_key = conn_obj
            conn = np.asarray(self.integdomain.conn)
            if conn.shape[1] != self._nnpe:
                raise FsgpuError(L.ERR_ARG, f"element set has {conn.shape[1]} nodes per element, expected {self._nnpe}")
            self._xyz = np.asarray(geom0.values)
            self.ctx.set_mesh(conn, geom0.values)
            self._mesh_key = key
            self._dof_key = self._sym_key = None
            self._after_mesh()

    def _after_mesh(self):
        pass

    def reset_uploads(self):
        """Forget what is resident on the device: the next operator re-uploads everything."""
        self._mesh_key = self._dof_key = self._sym_key = None

    def _sync_dofs(self, dchi):
        key = self._field_key(dchi, np.asarray(dchi.dofnums)) + (nfreedofs(dchi),)
        if not self._same_key(self._dof_key, key):
            self.ctx.set_dofnums(dchi.dofnums, nfreedofs(dchi), nalldofs(dchi))
            self._dof_key = key
            self._sym_key = None

    def _startassembly(self, assembler, dchi):
        self._sync_dofs(dchi)
        if self._sym_key != assembler.target:  # reset to None whenever the mesh or the dofs are re-uploaded
            self.ctx.symbolic(assembler.target)
            self._sym_key = assembler.target


class _FEMMShell(_FEMMBase):
    _comp = False
    _default_alpha = 0.0

    def __init__(self, integdomain, material_or_layup, stab_alpha=None, device=0):
        super().__init__(integdomain, device)
        if self._comp:
            layup = material_or_layup
            self.layup_groups = [(layup, None)] if isinstance(layup, CompositeLayup) else list(layup)
            self.material = None
        else:
            self.material = material_or_layup
        self.drilling_stiffness_scale = 1.0
        self.threshold_angle = 30.0
        self.transv_shear_formulation = 0
        self.stab_alpha = self._default_alpha if stab_alpha is None else stab_alpha
        self.stab_fun = None  # optional python callable (t, h) -> factor, evaluated on the host per element
        self._normals = None
        self._normal_valid = None

    def _kind(self):
        return self._nnpe + (10 if self._comp else 0)

    def _after_mesh(self):
        ctx, idom = self.ctx, self.integdomain
        if self._nnpe == 4:
            pc, w = idom.rule if idom.rule is not None else GaussRule2x2()
            ctx.set_rule(pc, w)
        if self._comp:
            recs = np.stack([lg[0].group_record() for lg in self.layup_groups])
            gof = None
            if len(self.layup_groups) > 1:
                gof = np.zeros(ctx.nelem, dtype=np.int64)
                for gi, (_, eset) in enumerate(self.layup_groups):
                    gof[np.asarray(eset) - 1] = gi + 1
            css = [lg[0].csys for lg in self.layup_groups]
            if isinstance(css[0], CSysKind) and all(c is css[0] for c in css):
                ctx.set_layup(recs, gof, np.eye(3))
                ctx.set_layup_csys(css[0].kind, css[0].origin, css[0].axis)  # evaluated on the device
            else:
                ctx.set_layup(recs, gof, self._layup_csmats())
        else:
            t = idom.otherdimension
            ctx.set_thickness(t)
        # the nodal normals belong to the FEMM (femm._normals) and survive a re-upload
        if self._associatedgeometry and self._normals is not None:
            ctx.set_normals(self._normals, self._normal_valid)

    def _group_elements(self, gi):
        eset = self.layup_groups[gi][1]
        return np.arange(self.ctx.nelem) if eset is None else np.asarray(eset) - 1

    def _layup_csmats(self):
        """Layup csys matrices for `fsgpu_set_layup`: one (3, 3) matrix when every group has the same constant csys,
        else one per element (T3FFComp: `updatecsmat!(layup.csys, centroid, J0, i, 0)`,
        src/FEMMShellT3FFCompModule.jl:617) or one per element and integration point (Q4RSComp:
        `updatecsmat!(layup.csys, Ns[j], J, -1, 0)`, src/FEMMShellQ4RSCompModule.jl:929 -- the SHAPE-FUNCTION VALUES
        are passed as the location; reproduced here).  Each group uses its own csys."""
        css = [lg[0].csys for lg in self.layup_groups]
        if not any(callable(c) for c in css) and all(np.ndim(c) == 2 for c in css) and all(np.array_equal(css[0], c) for c in css[1:]):
            return np.asarray(css[0], dtype=np.float64)
        conn = np.asarray(self.integdomain.conn)
        ne = conn.shape[0]
        X = self._xyz[conn - 1]  # (ne, nnpe, 3)
        if self._nnpe == 3:
            out = np.zeros((ne, 3, 3))
            J0 = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=-1)
            for gi, c in enumerate(css):
                el = self._group_elements(gi)
                out[el] = _eval_csys(c, X[el].mean(axis=1), J0[el], el + 1, 0, el)
            return out
        pc, _ = self.integdomain.rule if self.integdomain.rule is not None else GaussRule2x2()
        out = np.zeros((ne, len(pc), 3, 3))
        for j, (xi, eta) in enumerate(pc):
            N, dN = _q4_shape(xi, eta)
            J = np.einsum("eai,ak->eik", X, dN)
            for gi, c in enumerate(css):
                el = self._group_elements(gi)
                if not callable(c) and np.ndim(c) == 4:  # matrices given per element and integration point
                    out[el, j] = np.asarray(c, dtype=np.float64)[el, j]
                    continue
                out[el, j] = _eval_csys(c, np.broadcast_to(N, (len(el), 4)), J[el], -1, 0, el)
        return out

    def _params(self):
        p = ShellParams()
        if not self._comp:
            Dps, Dt = _shell_material_stiffness(self.material.moduli())
            p.Dps[:] = Dps.ravel().tolist()
            p.Dt[:] = Dt.ravel().tolist()
            p.rho = float(self.material.rho)
        p.stab_alpha = float(self.stab_alpha)
        p.drilling_stiffness_scale = float(self.drilling_stiffness_scale)
        p.transv_shear_formulation = int(self.transv_shear_formulation)
        return p

    def _sync_stab(self):
        if self.stab_fun is not None:
            h = self.ctx.element_sizes()
            if self._comp:
                t = np.zeros(self.ctx.nelem)
                for gi, (lay, _) in enumerate(self.layup_groups):  # `thickness(layup)` of the element's group
                    t[self._group_elements(gi)] = lay.thickness()
            else:
                t = np.broadcast_to(np.asarray(self.integdomain.otherdimension, dtype=np.float64), (self.ctx.nelem,))
            self.ctx.set_stab_factor(self.stab_fun(t, h))
        else:
            self.ctx.set_stab_factor(None)


def associategeometry(femm, geom0, interface=None):
    """`associategeometry!(femm, geom0)`: nodal normals + validity, computed on the device
    with the default csys (isoparametric) or the layup's cartesian csys for composites.
    `interface`: (InterfaceExchange over node indices, device) for element-partitioned runs -- the
    normal sums and the validity flags of interface nodes are combined across ranks."""
    femm._sync_mesh(geom0)
    fixed = None
    if femm._comp:
        css = [lg[0].csys for lg in femm.layup_groups]
        if isinstance(css[0], CSysKind) and all(c is css[0] for c in css) and interface is None:
            femm.ctx.associategeometry_csys(css[0].kind, css[0].origin, css[0].axis, femm.threshold_angle, False)
            femm._normals, femm._normal_valid = femm.ctx.get_normals()
            femm._associatedgeometry = True
            return femm
        same_const = not any(callable(c) for c in css) and all(np.ndim(c) == 2 for c in css) and all(np.array_equal(css[0], c) for c in css[1:])
        if not same_const:
            # `_compute_nodal_normal!(nnormal, layup.csys, geom.values[n, :], J0, el, 0)`: each group's csys evaluated AT
            # THE NODE (src/FEMMShellT3FFCompModule.jl:203-207,509; src/FEMMShellQ4RSCompModule.jl:228-232,471)
            if interface is not None:
                raise FsgpuError(L.ERR_ARG, "partitioned associategeometry supports the default and the cartesian csys")
            conn = np.asarray(femm.integdomain.conn)
            xyz = np.asarray(geom0.values)
            X = xyz[conn - 1]
            nn = conn.shape[1]
            dirs = np.zeros((conn.shape[0], nn, 3))
            for gi, c in enumerate(css):
                el = femm._group_elements(gi)
                for k in range(nn):
                    if nn == 3:
                        J = np.stack([X[el, 1] - X[el, 0], X[el, 2] - X[el, 0]], axis=-1)
                    else:
                        J = np.einsum("eai,ak->eik", X[el], _q4_shape(*_Q4_NODE_PC[k])[1])
                    dirs[el, k] = _eval_csys(c, X[el, k], J, el + 1, k + 1, el)[:, :, 2]
            femm.ctx.associategeometry_dirs(dirs, femm.threshold_angle, False)
            femm._normals, femm._normal_valid = femm.ctx.get_normals()
            femm._associatedgeometry = True
            return femm
        fixed = np.asarray(css[0], dtype=np.float64)[:, 2].copy()
    # homogeneous T3FF never resets its arrays
    accumulate = (femm._nnpe == 3) and (not femm._comp)
    if interface is None:
        femm.ctx.associategeometry(femm.threshold_angle, fixed, accumulate)
    else:
        import torch

        from .partition import DevicePointer, InterfaceExchange

        nn = femm.ctx.nnodes
        node_links, device = interface  # [(peer, local node indices)], torch device
        ex3 = InterfaceExchange([(p, (np.asarray(ix)[:, None] * 3 + np.arange(3)[None, :]).ravel()) for p, ix in node_links], device)
        ex4 = InterfaceExchange([(p, np.asarray(ix) * 4 + 3) for p, ix in node_links], device)
        sums = torch.as_tensor(DevicePointer(femm.ctx.normals_accumulate(fixed, accumulate), nn * 3), device=device)
        ex3.exchange_sum(sums)
        torch.cuda.synchronize()
        n4 = torch.as_tensor(DevicePointer(femm.ctx.normals_finish(femm.threshold_angle, fixed), nn * 4), device=device)
        inval = 1.0 - n4  # only the 4th components are exchanged
        ex4.exchange_sum(inval)
        n4[3::4] = (inval[3::4] == 0.0).to(torch.float64)
        torch.cuda.synchronize()
    femm._normals, femm._normal_valid = femm.ctx.get_normals()
    femm._associatedgeometry = True
    return femm


def _require_associated(femm):
    if not femm._associatedgeometry:
        # `@assert self._associatedgeometry == true` (src/FEMMShellT3FFModule.jl:643)
        raise FsgpuError(L.ERR_STATE, "geometry not associated: call associategeometry(femm, geom0) first")


def stiffness(femm, *args, out=None):
    """stiffness(femm, [assembler,] geom0, u1, Rfield1, dchi); `out`: optional caller-owned
    (colptr, rowval, nzval) arrays (e.g. pinned) that receive the result."""
    if len(args) == 4:
        args = (SysmatAssemblerSparseSymm(),) + args
    assembler, geom0, u1, Rfield1, dchi = args
    if isinstance(femm, FEMMCorotBeam):
        return femm._matrix_op("stiffness", assembler, geom0, u1, Rfield1, dchi)
    _require_associated(femm)
    femm._sync_mesh(geom0)
    femm._startassembly(assembler, dchi)
    femm._sync_stab()
    femm.ctx.shell_op(femm._opname + "_stiffness", femm._params())
    if assembler.uplo:
        return femm.ctx.fetch_matrix_uplo(assembler.uplo, out)
    return femm.ctx.fetch_matrix(out)


def mass(femm, *args, mass_type=1):
    """shells: mass(femm, [assembler,] geom0, dchi); beam: mass(femm, [assembler,] geom0, u1, Rfield1, dchi; mass_type)"""
    if isinstance(femm, FEMMCorotBeam):
        if len(args) == 4:
            args = (SysmatAssemblerSparseSymm(),) + args
        return femm._matrix_op("mass", *args, mass_type=mass_type)
    if len(args) == 2:
        args = (SysmatAssemblerSparseSymm(),) + args
    assembler, geom0, dchi = args
    _require_associated(femm)
    femm._sync_mesh(geom0)
    femm._startassembly(assembler, dchi)
    femm.ctx.shell_op(femm._opname + "_mass", femm._params())
    if assembler.uplo:
        return femm.ctx.fetch_matrix_uplo(assembler.uplo)
    return femm.ctx.fetch_matrix()


def geostiffness(femm, *args):
    if len(args) == 4:
        args = (SysmatAssemblerSparseSymm(),) + args
    return femm._matrix_op("geostiffness", *args)


_QUANTITY = {"bending": 1, "moment": 1, "bending_moment": 1, "transverse_shear": 2, "transverse": 2, "shear": 2,
             "membrane_force": 3, "membrane": 3}


def inspectintegpoints(femm, geom0, u, felist=None, quantity="moment", outputcsys=None):
    """Batched `inspectintegpoints` (src/FEMMShellT3FFModule.jl:850-962): instead of calling a host
    `inspector` closure per point, returns the array of resultants, shape (len(felist), npts, 3),
    in the output csys.  Default output csys: the element triad for the homogeneous shells, the layup csys for
    the laminated ones (src/FEMMShellT3FFCompModule.jl:846, src/FEMMShellQ4RSCompModule.jl:1119)."""
    _require_associated(femm)
    femm._sync_mesh(geom0)
    femm._sync_stab()
    npts = 1 if femm._nnpe == 3 else femm.ctx.npts
    out = femm.ctx.shell_resultants(femm._params(), femm._kind(), _QUANTITY[quantity], u.values, outputcsys, npts)
    return out if felist is None else out[np.asarray(felist) - 1]


def fieldfromintegpoints(femm, geom0, u, quantity, component, outputcsys=None):
    """FinEtools `fieldfromintegpoints(femm, geom, u, quantity, component; outputcsys)` with its default
    `nodevalmethod = :invdistance` (as called by test/test_shell_resultants.jl:123 and the shell examples): the batched
    resultants averaged to the nodes on the device (`fsgpu_shell_nodal_field`: weights 1 / squared distance between the
    node and the integration point -- the centroid for the T3 shells).  `component`: 1-based index or indices; returns a
    NodalField with one column per component."""
    _require_associated(femm)
    femm._sync_mesh(geom0)
    femm._sync_stab()
    comp = np.atleast_1d(np.asarray(component, dtype=np.int64)) - 1
    fld = femm.ctx.shell_nodal_field(femm._params(), femm._kind(), _QUANTITY[quantity], u.values, outputcsys)
    return NodalField(fld[:, comp])


def elemfieldfromintegpoints(femm, geom0, u, quantity, component, outputcsys=None):
    """FinEtools `elemfieldfromintegpoints`: per element the mean of the integration-point values, (nelem, ncomp)."""
    res = inspectintegpoints(femm, geom0, u, None, quantity, outputcsys)
    comp = np.atleast_1d(np.asarray(component, dtype=np.int64)) - 1
    return res[:, :, comp].mean(axis=1)


def gyroscopic(femm, *args, mass_type=1):
    """gyroscopic(femm, [assembler,] geom0, u1, Rfield1, v1, dchi; mass_type)
    (src/FEMMCorotBeamModule.jl:883-952)"""
    if len(args) == 5:
        args = (SysmatAssemblerSparseSymm(),) + args
    assembler, geom0, u1, Rfield1, v1, dchi = args
    femm._sync_mesh(geom0)
    femm._startassembly(assembler, dchi)
    femm.ctx.set_state(u1.values, Rfield1.values)
    femm.ctx.set_velocity(v1.values)
    femm.ctx.beam_op("gyroscopic", femm._params(mass_type))
    return femm.ctx.fetch_matrix()


def distribloads_global(femm, *args):
    """distribloads_global(femm, [assembler,] geom0, u1, Rfield1, dchi, fi); `fi`: force per unit
    length in global components, (3,) or (nelem, 3) (src/FEMMCorotBeamModule.jl:1186-1247)"""
    if len(args) == 5:
        args = (SysvecAssembler(),) + args
    assembler, geom0, u1, Rfield1, dchi, fi = args
    femm._sync_mesh(geom0)
    femm._sync_dofs(dchi)
    femm.ctx.set_state(u1.values, Rfield1.values)
    femm.ctx.beam_distribloads(femm._params(), fi, assembler.nfree_only)
    return femm.ctx.fetch_vector(nfreedofs(dchi) if assembler.nfree_only else nalldofs(dchi))


def restoringforce(femm, *args):
    if len(args) == 4:
        args = (SysvecAssembler(),) + args
    assembler, geom0, u1, Rfield1, dchi = args
    femm._sync_mesh(geom0)
    femm._sync_dofs(dchi)
    femm.ctx.set_state(u1.values, Rfield1.values)
    femm.ctx.beam_op("restoringforce", femm._params(), 1 if assembler.nfree_only else 0)
    return femm.ctx.fetch_vector(nfreedofs(dchi) if assembler.nfree_only else nalldofs(dchi))


class FEMMShellT3FF(_FEMMShell):
    _nnpe, _comp, _opname, _default_alpha = 3, False, "t3ff", 5 / 12 / 1.5


class FEMMShellQ4RS(_FEMMShell):
    _nnpe, _comp, _opname, _default_alpha = 4, False, "q4rs", 0.1


class FEMMShellT3FFComp(_FEMMShell):
    _nnpe, _comp, _opname, _default_alpha = 3, True, "t3ffcomp", 5 / 12 / 1.5


class FEMMShellQ4RSComp(_FEMMShell):
    _nnpe, _comp, _opname, _default_alpha = 4, True, "q4rscomp", 0.1


@dataclass
class FESetL2Beam:
    """Per-element section SoA (src/FESetL2BeamModule.jl:20-31)."""

    A: np.ndarray
    I1: np.ndarray
    I2: np.ndarray
    I3: np.ndarray
    J: np.ndarray
    A2s: np.ndarray
    A3s: np.ndarray
    x1x2_vector: np.ndarray  # (nelem, 3)


class FEMMCorotBeam(_FEMMBase):
    _nnpe = 2
    _comp = False

    def __init__(self, integdomain, material, sections: FESetL2Beam, device=0):
        super().__init__(integdomain, device)
        self.material = material
        self.sections = sections

    def _after_mesh(self):
        s = self.sections
        self.ctx.set_beam_sections(s.A, s.I1, s.I2, s.I3, s.J, s.A2s, s.A3s, s.x1x2_vector)

    def _params(self, mass_type=1):
        p = BeamParams()
        p.E, p.nu, p.rho, p.mass_type = float(self.material.E), float(self.material.nu), float(self.material.rho), int(mass_type)
        return p

    def _matrix_op(self, name, assembler, geom0, u1, Rfield1, dchi, mass_type=1):
        self._sync_mesh(geom0)
        self._startassembly(assembler, dchi)
        self.ctx.set_state(u1.values, Rfield1.values)
        self.ctx.beam_op(name, self._params(mass_type))
        if assembler.uplo:
            return self.ctx.fetch_matrix_uplo(assembler.uplo)
        return self.ctx.fetch_matrix()


def initial_Rfield(nnodes):
    """src/RotUtilModule.jl:16-22"""
    R = np.zeros((nnodes, 9), order="F")
    R[:, 0] = R[:, 4] = R[:, 8] = 1.0
    return NodalField(R)


def update_rotation_field(femm, Rfield, dchi):
    """`update_rotation_field!(Rfield, dchi)` on the device copy held by `femm`
    (src/RotUtilModule.jl:29-42)."""
    femm.ctx.set_state(np.zeros((femm.ctx.nnodes, 3)), Rfield.values)
    Rfield.values[:] = femm.ctx.update_rotation_field(dchi.values)
    return Rfield
