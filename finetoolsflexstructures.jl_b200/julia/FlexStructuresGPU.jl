"""
    FlexStructuresGPU

`ccall` glue that makes libfsgpu.so a drop-in for the element-level hot path of
FinEtoolsFlexStructures.jl.  The plugin point is the assembler type: passing a
`SysmatAssemblerGPU` / `SysvecAssemblerGPU` to the EXISTING operators

    stiffness(femm, assembler, geom0, u1, Rfield1, dchi)
    mass(femm, assembler, geom0, dchi)                                  # shells
    mass(femm, assembler, geom0, u1, Rfield1, dchi; mass_type)          # beam
    geostiffness(femm, assembler, geom0, u1, Rfield1, dchi)
    gyroscopic(femm, assembler, geom0, u1, Rfield1, v1, dchi; mass_type)
    restoringforce(femm, assembler, geom0, u1, Rfield1, dchi)
    distribloads_global(femm, assembler, geom0, u1, Rfield1, dchi, fi)

dispatches to the methods below, which replace the per-element Julia loop
(e.g. src/FEMMShellT3FFModule.jl:670-734) by ONE library call per operator.  Operators without an
assembler argument get a method with a trailing / leading `Context`:

    associategeometry!(femm, geom0, ctx)                                # nodal normals on the device
    inspectintegpoints(ctx, femm, geom0, u, dT, felist, inspector, idat, quantity; context...)

Every other assembler keeps using the original Julia methods, as does any FEMM whose number type is not
Float64 (the ForwardDiff use in examples/shells/statics/homogeneous/plates/
ss_circular_plate_udl_examples.jl:129-187).

State on the device is cached per `Context`: the mesh, the dof numbers, the nodal normals, the
thickness / layup data and the symbolic phase are uploaded / rebuilt only when the arrays they came
from change (identity, length and a strided checksum; `invalidate!(ctx)` forces a refresh after an
in-place edit the checksum cannot see).  A K + M pair on the same mesh therefore pays for one upload
and one symbolic phase.  Assemblers share one context per device (`default_context`).

NOTE: neither the build image of the B200 port nor its GPU boxes have Julia (`command -v julia`
fails on both, recorded in profiles/), so this file has been written against include/fsgpu.h but
never executed.  The same call sequences are exercised, through the same C ABI, by the Python
mirror (finetoolsflexstructures.jl_b200/femm.py) and its tests.  `julia/smoke.jl` runs a small
end-to-end check wherever Julia and the reference packages are available.
"""
module FlexStructuresGPU

using LinearAlgebra
using SparseArrays
using FinEtools
using FinEtools.AssemblyModule: AbstractSysmatAssembler, AbstractSysvecAssembler
using FinEtoolsDeforLinear
using FinEtoolsFlexStructures
using FinEtoolsFlexStructures.FEMMShellT3FFModule: FEMMShellT3FF
using FinEtoolsFlexStructures.FEMMShellQ4RSModule: FEMMShellQ4RS
using FinEtoolsFlexStructures.FEMMShellT3FFCompModule: FEMMShellT3FFComp
using FinEtoolsFlexStructures.FEMMShellQ4RSCompModule: FEMMShellQ4RSComp
using FinEtoolsFlexStructures.FEMMCorotBeamModule: FEMMCorotBeam, properties
using FinEtoolsFlexStructures.CompositeLayupModule: thickness, laminate_stiffnesses!, laminate_transverse_stiffness!, laminate_inertia!
import FinEtoolsFlexStructures.FEMMShellT3FFModule
import FinEtoolsFlexStructures.FEMMShellQ4RSModule
import FinEtoolsFlexStructures.FEMMShellT3FFCompModule
import FinEtoolsFlexStructures.FEMMShellQ4RSCompModule
import FinEtoolsFlexStructures.FEMMCorotBeamModule

const libfsgpu = get(ENV, "LIBFSGPU", joinpath(@__DIR__, "..", "libfsgpu.so"))

# enum fsgpu_target, enum fsgpu_csys_kind (include/fsgpu.h)
const SPARSE, SPARSE_SYMM, SPARSE_DIAG, FFBLOCK, FFBLOCK_DIAG, CSR_SYMM = Int32.(0:5)
const CSYS_CYLINDRICAL, CSYS_SPHERICAL, CSYS_NORMAL_AXIS = Int32.(1:3)

struct FsgpuError <: Exception
    code::Int
    msg::String
end

function _check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:fsgpu_last_error, libfsgpu), Cstring, ()))
    # reproduce the reference's failure modes
    rc == 3 && throw(AssertionError(msg))              # @assert self._associatedgeometry
    rc == 5 && error("Singular metric matrix in _gradN_e!")
    throw(FsgpuError(Int(rc), msg))
end

# ---- context with upload caches ------------------------------------------------------------------
const Key = Tuple{UInt,Int,Float64}
const NOKEY = (UInt(0), -1, 0.0)

mutable struct Context
    h::Ptr{Cvoid}
    device::Int
    conn_key::Key
    xyz_key::Key
    dof_key::Tuple{Key,Int}
    normals_key::Key
    aux_key::Any          # thickness / rule / layup / sections of the FEMM last uploaded
    sym_target::Int32     # target of the symbolic phase on the device (-1: none)
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        _check(ccall((:fsgpu_create, libfsgpu), Cint, (Ref{Ptr{Cvoid}}, Cint), r, device))
        c = new(r[], Int(device), NOKEY, NOKEY, (NOKEY, -1), NOKEY, nothing, Int32(-1))
        finalizer(x -> ccall((:fsgpu_destroy, libfsgpu), Cint, (Ptr{Cvoid},), x.h), c)
        return c
    end
end

"Forget what is resident on the device: the next operator re-uploads everything and rebuilds the pattern."
function invalidate!(c::Context)
    c.conn_key = c.xyz_key = c.normals_key = NOKEY
    c.dof_key = (NOKEY, -1)
    c.aux_key = nothing
    c.sym_target = Int32(-1)
    return c
end

const _contexts = Dict{Int,Context}()
"One shared context per device: every `SysmatAssemblerGPU(...)` reuses it (a CUDA context per assembler cost ~0.2 s)."
default_context(device::Integer = 0) = get!(() -> Context(device), _contexts, Int(device))

# identity + length + a strided checksum (<= 257 samples): cheap, and catches renumbering / moved nodes
function _key(a::AbstractArray)
    n = length(a)
    n == 0 && return (objectid(a), 0, 0.0)
    st = max(1, n ÷ 256)
    s = 0.0
    @inbounds for i in 1:st:n
        s += Float64(a[i]) * (1 + (i % 7))
    end
    return (objectid(a), n, s + Float64(a[n]))
end
# Vector{NTuple{N,Int}} (fes.conn) as a flat Int64 view, no copy
_flat(conn::Vector{NTuple{N,Int}}) where {N} = reinterpret(Int, conn)

"""
    SysmatAssemblerGPU(target = SPARSE_SYMM; uplo = :full, ctx = default_context(0))

GPU assembler carrying the semantic target of a FinEtools assembler:
`SPARSE` (SysmatAssemblerSparse), `SPARSE_SYMM` (SysmatAssemblerSparseSymm, the default of
the convenience methods), `SPARSE_DIAG`, `FFBLOCK` (SysmatAssemblerFFBlock(nfreedofs)),
`FFBLOCK_DIAG`, `CSR_SYMM` (SysmatAssemblerSparseCSRSymm, src/AssemblyModule.jl:20-53).
`uplo = :L` / `:U`: `makematrix!` returns that triangle only (diagonal included), for
`cholesky(Symmetric(K, :L))`-style consumers -- half the bytes cross PCIe.
"""
struct SysmatAssemblerGPU <: AbstractSysmatAssembler
    target::Int32
    uplo::Symbol
    ctx::Context
end
SysmatAssemblerGPU(target = SPARSE_SYMM; uplo::Symbol = :full, ctx::Context = default_context(0)) =
    SysmatAssemblerGPU(Int32(target), uplo, ctx)

struct SysvecAssemblerGPU <: AbstractSysvecAssembler
    nfree_only::Bool
    ctx::Context
end
SysvecAssemblerGPU(nfree_only::Bool = false; ctx::Context = default_context(0)) = SysvecAssemblerGPU(nfree_only, ctx)

# struct fsgpu_shell_params / fsgpu_beam_params (include/fsgpu.h)
struct ShellParams
    Dps::NTuple{9,Float64}
    Dt::NTuple{4,Float64}
    rho::Float64
    stab_alpha::Float64
    drilling_stiffness_scale::Float64
    transv_shear_formulation::Int32
    reserved::Int32
end
struct BeamParams
    E::Float64
    nu::Float64
    rho::Float64
    mass_type::Int32
    reserved::Int32
end

# ---- data hand-over (cached) -----------------------------------------------------------------------
function _set_mesh!(c::Context, fes, geom0)
    conn = fes.conn  # Vector{NTuple{nnpe,Int}}: nnpe x nelem Int64, inline
    kc, kx = _key(_flat(conn)), _key(geom0.values)
    (kc == c.conn_key && kx == c.xyz_key) && return false
    GC.@preserve conn geom0 _check(ccall((:fsgpu_set_mesh, libfsgpu), Cint,
        (Ptr{Cvoid}, Int32, Int64, Ptr{Int64}, Int64, Ptr{Float64}),
        c.h, nodesperelem(fes), count(fes), pointer(conn), size(geom0.values, 1), pointer(geom0.values)))
    c.conn_key, c.xyz_key = kc, kx
    c.dof_key = (NOKEY, -1); c.normals_key = NOKEY; c.aux_key = nothing; c.sym_target = Int32(-1)
    return true
end
function _set_dofs!(c::Context, dchi)
    k = (_key(dchi.dofnums), Int(nfreedofs(dchi)))
    k == c.dof_key && return false
    GC.@preserve dchi _check(ccall((:fsgpu_set_dofnums, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Int64),
        c.h, pointer(dchi.dofnums), nfreedofs(dchi), nalldofs(dchi)))
    c.dof_key = k
    c.sym_target = Int32(-1)
    return true
end
function _set_normals!(c::Context, femm)
    n, v = femm._normals, femm._normal_valid   # Matrix{Float64} nnodes x 3, Vector{Bool}
    k = _key(n)
    k = (k[1], k[2], k[3] + count(v))
    k == c.normals_key && return false
    GC.@preserve n v _check(ccall((:fsgpu_set_normals, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}), c.h, pointer(n), pointer(v)))
    c.normals_key = k
    return true
end
function _element_sizes(c::Context, nelem::Integer)
    h = Vector{Float64}(undef, nelem)
    GC.@preserve h _check(ccall((:fsgpu_element_sizes, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.h, pointer(h)))
    return h
end
function _set_rule!(c::Context, femm)
    npts, Ns, gradNparams, w, pc = integrationdata(femm.integdomain, femm.integdomain.integration_rule)
    xi, eta, ww = pc[:, 1], pc[:, 2], vec(w)
    GC.@preserve xi eta ww _check(ccall((:fsgpu_set_rule, libfsgpu), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        c.h, npts, pointer(xi), pointer(eta), pointer(ww)))
    return npts, Ns, gradNparams
end
# thickness callback evaluated on the host, once per element (T3: at the centroid, src/FEMMShellT3FFModule.jl:676)
# or per element and integration point (Q4, src/FEMMShellQ4RSModule.jl:921); returns the per-element thickness
function _set_thickness!(c::Context, femm, geom0, isq4::Bool)
    fes = femm.integdomain.fes
    ne = count(fes)
    if !isq4
        ipc = [(1.0 / 3) (1.0 / 3)]
        t = Vector{Float64}(undef, ne)
        centroid = fill(0.0, 1, 3)
        for i in eachindex(fes)
            cn = fes.conn[i]
            centroid .= (geom0.values[cn[1]:cn[1], :] .+ geom0.values[cn[2]:cn[2], :] .+ geom0.values[cn[3]:cn[3], :]) ./ 3
            t[i] = femm.integdomain.otherdimension(centroid, cn, ipc)
        end
        tv = all(==(t[1]), t) ? t[1:1] : t
        GC.@preserve tv _check(ccall((:fsgpu_set_thickness, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), c.h, pointer(tv), length(tv)))
        return t
    end
    npts, Ns, _ = _set_rule!(c, femm)
    t = Matrix{Float64}(undef, npts, ne)
    loc = fill(0.0, 1, 3)
    for i in eachindex(fes), j in 1:npts
        loc .= Ns[j]' * geom0.values[collect(fes.conn[i]), :]
        t[j, i] = femm.integdomain.otherdimension(loc, fes.conn[i], Ns[j])
    end
    tv = all(==(t[1]), t) ? t[1:1] : vec(t)
    GC.@preserve tv _check(ccall((:fsgpu_set_thickness, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), c.h, pointer(tv), length(tv)))
    return vec(t[1, :])
end
# stab_fun: the library evaluates t^2/(t^2 + alpha h^2) itself; any other closure is SAMPLED per element on the host
# (t = the element's thickness, h = the element size the reference uses: T3 sqrt(2 Ae), Q4 the "diameter" from node 1)
# and handed over with fsgpu_set_stab_factor
function _stab!(c::Context, femm, tper::AbstractVector{Float64}, nelem::Integer)
    f = femm.stab_fun
    t, h = 0.37, 1.91
    alpha = (t^2 / f(t, h) - t^2) / h^2
    if isapprox(f(0.11, 0.7), 0.11^2 / (0.11^2 + alpha * 0.7^2); rtol = 1e-13) && isapprox(f(2.3, 0.05), 2.3^2 / (2.3^2 + alpha * 0.05^2); rtol = 1e-13)
        _check(ccall((:fsgpu_set_stab_factor, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), c.h, C_NULL, 0))
        return alpha
    end
    hs = _element_sizes(c, nelem)
    fac = [f(length(tper) == 1 ? tper[1] : tper[i], hs[i]) for i in 1:nelem]
    GC.@preserve fac _check(ccall((:fsgpu_set_stab_factor, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), c.h, pointer(fac), nelem))
    return 0.0
end

function _shell_params(femm, alpha::Float64; comp = false)
    Dps, Dt = comp ? (zeros(3, 3), zeros(2, 2)) : FEMMShellT3FFModule._shell_material_stiffness(femm.material)
    rho = comp ? 0.0 : massdensity(femm.material)
    tsf = hasproperty(femm, :transv_shear_formulation) ? femm.transv_shear_formulation : 0
    ShellParams(Tuple(permutedims(Dps)), Tuple(permutedims(Dt)), rho, alpha, femm.drilling_stiffness_scale, tsf, 0)
end

function _fetch(a::SysmatAssemblerGPU)
    m, n, nnz = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    _check(ccall((:fsgpu_result_size, libfsgpu), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), a.ctx.h, m, n, nnz))
    if a.uplo != :full
        ul = Int32(a.uplo == :L ? 'L' : 'U')
        _check(ccall((:fsgpu_result_size_uplo, libfsgpu), Cint, (Ptr{Cvoid}, Int32, Ref{Int64}), a.ctx.h, ul, nnz))
        colptr, rowval, nzval = Vector{Int64}(undef, n[] + 1), Vector{Int64}(undef, nnz[]), Vector{Float64}(undef, nnz[])
        GC.@preserve colptr rowval nzval _check(ccall((:fsgpu_fetch_matrix_uplo, libfsgpu), Cint, (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
            a.ctx.h, ul, pointer(colptr), pointer(rowval), pointer(nzval)))
        return SparseMatrixCSC(m[], n[], colptr, rowval, nzval)
    end
    colptr, rowval, nzval = Vector{Int64}(undef, n[] + 1), Vector{Int64}(undef, nnz[]), Vector{Float64}(undef, nnz[])
    GC.@preserve colptr rowval nzval _check(ccall((:fsgpu_fetch_matrix, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
        a.ctx.h, pointer(colptr), pointer(rowval), pointer(nzval)))
    if a.target == CSR_SYMM   # SparseMatricesCSR.SparseMatrixCSR{1}(m, n, rowptr, colval, nzval)
        return FinEtoolsFlexStructures.AssemblyModule.SparseMatricesCSR.SparseMatrixCSR{1}(m[], n[], colptr, rowval, nzval)
    end
    return SparseMatrixCSC(m[], n[], colptr, rowval, nzval)
end
function _fetch_vector(c::Context, n::Integer)
    F = Vector{Float64}(undef, n)
    GC.@preserve F _check(ccall((:fsgpu_fetch_vector, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), c.h, pointer(F), n))
    return F
end

"""
    pinned_vector(T, n) -> Vector{T}

A `Vector{T}` over page-locked host memory (`fsgpu_host_alloc`), released by its finalizer.  Results fetched into such
arrays cross PCIe by direct DMA (C2: 48 ms for the values); an ordinary `Vector` is pageable and is filled through the
library's staging ring by host threads (73 ms).  Allocation is slow (page locking): allocate once, reuse.
"""
function pinned_vector(::Type{T}, n::Integer) where {T}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:fsgpu_host_alloc, libfsgpu), Cint, (Ref{Ptr{Cvoid}}, Int64), p, Int64(n) * sizeof(T)))
    ptr = p[]
    v = unsafe_wrap(Array, Ptr{T}(ptr), Int(n); own = false)
    finalizer(_ -> ccall((:fsgpu_host_free, libfsgpu), Cint, (Ptr{Cvoid},), ptr), v)
    return v
end

"""
    refresh_values!(K, assembler) -> K

Re-assembly on the same mesh, numbering and target (a Newton or time-stepping loop): after the operator has run again,
only the values cross PCIe, into `K.nzval` (`fsgpu_fetch_matrix(ctx, NULL, NULL, nzval)`); `colptr` / `rowval` of `K` are
the pattern of the first fetch.  Not for `SysmatAssemblerSparseSymm` targets (value-dependent pattern).
"""
function refresh_values!(K::SparseMatrixCSC{Float64,Int64}, a::SysmatAssemblerGPU)
    a.target == SPARSE_SYMM && error("refresh_values!: the SparseSymm pattern depends on the values; fetch the matrix")
    m, n, nnz = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    _check(ccall((:fsgpu_result_size, libfsgpu), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), a.ctx.h, m, n, nnz))
    (size(K, 1) == m[] && size(K, 2) == n[] && length(K.nzval) == nnz[]) || error("refresh_values!: the pattern changed")
    nz = K.nzval
    GC.@preserve nz _check(ccall((:fsgpu_fetch_matrix, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
        a.ctx.h, C_NULL, C_NULL, pointer(nz)))
    return K
end

# startassembly!: pattern + addressing, once per (mesh, dof numbering, target)
function _symbolic!(a::SysmatAssemblerGPU)
    a.ctx.sym_target == a.target && return nothing
    nr, nc, nnz = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    _check(ccall((:fsgpu_symbolic, libfsgpu), Cint, (Ptr{Cvoid}, Int32, Ref{Int64}, Ref{Int64}, Ref{Int64}), a.ctx.h, a.target, nr, nc, nnz))
    a.ctx.sym_target = a.target
    return nothing
end

# ---- homogeneous shells ---------------------------------------------------------------------------
function _shell_setup!(c::Context, self, geom0, dchi, isq4::Bool)
    @assert self._associatedgeometry == true
    fes = self.integdomain.fes
    changed = _set_mesh!(c, fes, geom0)
    _set_dofs!(c, dchi)
    _set_normals!(c, self)
    key = (:homog, objectid(self), objectid(self.integdomain.otherdimension))
    if changed || c.aux_key === nothing || c.aux_key[1] != key
        tper = _set_thickness!(c, self, geom0, isq4)
        c.aux_key = (key, tper)
    end
    return _stab!(c, self, c.aux_key[2], count(fes))
end
for (FEMM, mod, op, isq4) in ((:FEMMShellT3FF, :FEMMShellT3FFModule, :t3ff, false), (:FEMMShellQ4RS, :FEMMShellQ4RSModule, :q4rs, true))
    kstiff, kmass = QuoteNode(Symbol(:fsgpu_, op, :_stiffness)), QuoteNode(Symbol(:fsgpu_, op, :_mass))
    @eval function $mod.stiffness(self::$FEMM{ID,Float64}, assembler::SysmatAssemblerGPU, geom0::NodalField{Float64},
            u1::NodalField{TI}, Rfield1::NodalField{TI}, dchi::NodalField{TI}) where {ID,TI<:Number}
        c = assembler.ctx
        alpha = _shell_setup!(c, self, geom0, dchi, $isq4)
        _symbolic!(assembler)
        p = Ref(_shell_params(self, alpha))
        _check(ccall(($kstiff, libfsgpu), Cint, (Ptr{Cvoid}, Ref{ShellParams}), c.h, p))
        return _fetch(assembler)
    end
    @eval function $mod.mass(self::$FEMM{ID,Float64}, assembler::SysmatAssemblerGPU, geom0::NodalField{Float64}, dchi::NodalField{TI}) where {ID,TI<:Number}
        c = assembler.ctx
        alpha = _shell_setup!(c, self, geom0, dchi, $isq4)
        _symbolic!(assembler)
        p = Ref(_shell_params(self, alpha))
        _check(ccall(($kmass, libfsgpu), Cint, (Ptr{Cvoid}, Ref{ShellParams}), c.h, p))
        return _fetch(assembler)
    end
end

# ---- layered shells: the O(nplies) through-thickness integration stays on the host (reference functions) ----------
"""
    CSysKind(kind, axis; origin = (0,0,0))

Marker for a layup / normal csys the library evaluates ON THE DEVICE (`fsgpu_associategeometry_csys`,
`fsgpu_set_layup_csys`): `CSYS_CYLINDRICAL` (clamp_cyl_expl_examples.jl:62-68), `CSYS_SPHERICAL`
(hemisphere_examples.jl:31-39), `CSYS_NORMAL_AXIS` (pressurized_cylinder_free_examples.jl:16-23).  Register it for a
`CSys` object with `register_csys_kind!(csys, kind)`: FEMMs whose layup csys is registered skip the host callback loop.
"""
struct CSysKind
    kind::Int32
    axis::NTuple{3,Float64}
    origin::NTuple{3,Float64}
end
CSysKind(kind, axis; origin = (0.0, 0.0, 0.0)) = CSysKind(Int32(kind), Tuple(Float64.(axis)), Tuple(Float64.(origin)))
const _csys_kinds = IdDict{Any,CSysKind}()
register_csys_kind!(csys, k::CSysKind) = (_csys_kinds[csys] = k; csys)

function _set_layup!(c::Context, femm, geom0, nnpe)
    groups = femm.layup_groups
    rec = Matrix{Float64}(undef, 34, length(groups))
    for (g, (layup, _)) in enumerate(groups)
        A, B, D, H = zeros(3, 3), zeros(3, 3), zeros(3, 3), zeros(2, 2)
        laminate_stiffnesses!(layup, A, B, D); laminate_transverse_stiffness!(layup, H)
        md, mi = laminate_inertia!(layup)
        rec[:, g] = vcat(vec(permutedims(A)), vec(permutedims(B)), vec(permutedims(D)), vec(permutedims(H)), thickness(layup), md, mi)
    end
    gof = femm._layup_group_lookup
    fes = femm.integdomain.fes
    ne = count(fes)
    tper = [thickness(groups[gof[i]][1]) for i in 1:ne]
    kinds = [get(_csys_kinds, g[1].csys, nothing) for g in groups]
    if all(k -> k !== nothing && k == kinds[1], kinds)
        # every group's csys is the same built-in kind: evaluated on the device
        eye = Matrix{Float64}(I, 3, 3)
        GC.@preserve rec gof eye _check(ccall((:fsgpu_set_layup, libfsgpu), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Int64),
            c.h, length(groups), pointer(rec), pointer(gof), pointer(eye), 1))
        k = kinds[1]
        o, a = collect(k.origin), collect(k.axis)
        GC.@preserve o a _check(ccall((:fsgpu_set_layup_csys, libfsgpu), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), c.h, k.kind, pointer(o), pointer(a)))
        return tper
    end
    if nnpe == 3
        # `updatecsmat!(layup.csys, centroid, J0, i, 0)` per element (src/FEMMShellT3FFCompModule.jl:617)
        cs = Array{Float64}(undef, 3, 3, ne)
        centroid, J0 = fill(0.0, 1, 3), fill(0.0, 3, 2)
        for i in eachindex(fes)
            layup = groups[gof[i]][1]
            cn = fes.conn[i]
            centroid .= sum(geom0.values[collect(cn), :]; dims = 1) ./ 3
            J0[:, 1] .= geom0.values[cn[2], :] .- geom0.values[cn[1], :]
            J0[:, 2] .= geom0.values[cn[3], :] .- geom0.values[cn[1], :]
            updatecsmat!(layup.csys, centroid, J0, i, 0)
            cs[:, :, i] .= csmat(layup.csys)
        end
    else
        # Q4RSComp: per element AND integration point, `updatecsmat!(layup.csys, Ns[j], J, -1, 0)`
        # (src/FEMMShellQ4RSCompModule.jl:929): the reference passes the SHAPE-FUNCTION VALUES as the location
        # (SURVEY App. B.9); reproduced.  Layout: 3 x 3 x npts x nelem  (ncs = nelem * npts, point index fastest)
        npts, Ns, gradNparams, w, pc = integrationdata(femm.integdomain, femm.integdomain.integration_rule)
        cs = Array{Float64}(undef, 3, 3, npts * ne)
        J = fill(0.0, 3, 2)
        for i in eachindex(fes)
            layup = groups[gof[i]][1]
            ecoords = geom0.values[collect(fes.conn[i]), :]
            for j in 1:npts
                J .= transpose(ecoords) * gradNparams[j]
                updatecsmat!(layup.csys, Ns[j], J, -1, 0)
                cs[:, :, (i - 1) * npts + j] .= csmat(layup.csys)
            end
        end
    end
    ncs = all(cs[:, :, i] == cs[:, :, 1] for i in axes(cs, 3)) ? 1 : size(cs, 3)
    GC.@preserve rec gof cs _check(ccall((:fsgpu_set_layup, libfsgpu), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Int64),
        c.h, length(groups), pointer(rec), pointer(gof), pointer(cs), ncs))
    return tper
end
function _comp_setup!(c::Context, self, geom0, dchi, nnpe)
    @assert self._associatedgeometry == true
    fes = self.integdomain.fes
    changed = _set_mesh!(c, fes, geom0)
    _set_dofs!(c, dchi)
    _set_normals!(c, self)
    key = (:comp, objectid(self))
    if changed || c.aux_key === nothing || c.aux_key[1] != key
        nnpe == 4 && _set_rule!(c, self)
        tper = _set_layup!(c, self, geom0, nnpe)
        c.aux_key = (key, tper)
    end
    return _stab!(c, self, c.aux_key[2], count(fes))
end
for (FEMM, mod, op, nn) in ((:FEMMShellT3FFComp, :FEMMShellT3FFCompModule, :t3ffcomp, 3), (:FEMMShellQ4RSComp, :FEMMShellQ4RSCompModule, :q4rscomp, 4))
    for (fname, sym, args) in ((:stiffness, Symbol(:fsgpu_, op, :_stiffness), :(geom0::NodalField{Float64}, u1::NodalField{TI}, Rfield1::NodalField{TI}, dchi::NodalField{TI})),
                               (:mass, Symbol(:fsgpu_, op, :_mass), :(geom0::NodalField{Float64}, dchi::NodalField{TI})))
        q = QuoteNode(sym)
        @eval function $mod.$fname(self::$FEMM{ID,Float64}, assembler::SysmatAssemblerGPU, $(args.args...)) where {ID,TI<:Number}
            c = assembler.ctx
            alpha = _comp_setup!(c, self, geom0, dchi, $nn)
            _symbolic!(assembler)
            p = Ref(_shell_params(self, alpha; comp = true))
            _check(ccall(($q, libfsgpu), Cint, (Ptr{Cvoid}, Ref{ShellParams}), c.h, p))
            return _fetch(assembler)
        end
    end
end

# ---- associategeometry! on the device (src/FEMMShellT3FFModule.jl:570-616 and the three sibling methods) --------------
# No assembler takes part in the reference signature `associategeometry!(femm, geom0)`; the GPU method has a trailing Context.
function _store_normals!(c::Context, self)
    n, v = self._normals, self._normal_valid
    GC.@preserve n v _check(ccall((:fsgpu_get_normals, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}), c.h, pointer(n), pointer(v)))
    self._associatedgeometry = true
    c.normals_key = NOKEY
    return self
end
# the csys evaluated AT EVERY NODE OF EVERY ELEMENT (`_compute_nodal_normal!`, src/FEMMShellT3FFCompModule.jl:203-207,509;
# src/FEMMShellQ4RSCompModule.jl:228-232,471): built-in kinds on the device, any other CSys callback on the host
function _csys_dirs(fes, geom0, csys_of_element)
    nnpe = nodesperelem(fes)
    dirs = Array{Float64,3}(undef, 3, nnpe, count(fes))
    J = zeros(3, 2)
    Ns, gradNparams = nothing, nothing
    if nnpe == 4
        _, Ns, gradNparams, _, _ = integrationdata(IntegDomain(fes, NodalTensorProductRule(2)))
    end
    for (el, conn) in enumerate(fes.conn)
        csys = csys_of_element(el)
        for (k, n) in enumerate(conn)
            if nnpe == 3
                J[:, 1] .= geom0.values[conn[2], :] .- geom0.values[conn[1], :]
                J[:, 2] .= geom0.values[conn[3], :] .- geom0.values[conn[1], :]
            else
                J .= transpose(geom0.values[collect(conn), :]) * gradNparams[k]
            end
            updatecsmat!(csys, reshape(geom0.values[n, :], 1, 3), J, el, nnpe == 3 ? 0 : k)
            dirs[:, k, el] .= view(csmat(csys), :, 3)
        end
    end
    return dirs
end
function _associate!(c::Context, self, geom0, csys_of_element, kinds, accumulate::Bool)
    fes = self.integdomain.fes
    _set_mesh!(c, fes, geom0)
    ta = Float64(self.threshold_angle)
    if kinds === nothing      # default isoparametric csys: element normals
        _check(ccall((:fsgpu_associategeometry, libfsgpu), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}, Int32), c.h, ta, C_NULL, accumulate ? 1 : 0))
    elseif all(k -> k !== nothing && k == kinds[1], kinds)
        k = kinds[1]
        o, a = collect(k.origin), collect(k.axis)
        GC.@preserve o a _check(ccall((:fsgpu_associategeometry_csys, libfsgpu), Cint, (Ptr{Cvoid}, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Int32),
            c.h, ta, k.kind, pointer(o), pointer(a), accumulate ? 1 : 0))
    else
        dirs = _csys_dirs(fes, geom0, csys_of_element)
        GC.@preserve dirs _check(ccall((:fsgpu_associategeometry_dirs, libfsgpu), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}, Int32),
            c.h, ta, pointer(dirs), accumulate ? 1 : 0))
    end
    return _store_normals!(c, self)
end
# homogeneous shells: `mcsys` of the FEMM; isoparametric (the default) -> element normals.  T3FF never resets its arrays
# (src/FEMMShellT3FFModule.jl:575-587, SURVEY App. B.6): accumulate = true reproduces a repeated call.
# is `csys` the FEMM's default isoparametric csys (the element normal)?  FinEtools keeps the callback in a field of the CSys;
# if the field cannot be read the host-callback path below is used instead -- slower, same result
_is_iso(csys, f) = try
    getfield(csys, Symbol("__updatebuffer!")) === f
catch
    false
end
function FEMMShellT3FFModule.associategeometry!(self::FEMMShellT3FF{ID,Float64}, geom::NodalField{Float64}, c::Context) where {ID}
    kinds = _is_iso(self.mcsys, FEMMShellT3FFModule.isoparametric!) ? nothing : Union{Nothing,CSysKind}[get(_csys_kinds, self.mcsys, nothing)]
    again = self._associatedgeometry
    if again   # a repeated call accumulates onto the FEMM's current normals and keeps its invalid flags
        _set_mesh!(c, self.integdomain.fes, geom)
        _set_normals!(c, self)
    end
    return _associate!(c, self, geom, el -> self.mcsys, kinds, again)
end
function FEMMShellQ4RSModule.associategeometry!(self::FEMMShellQ4RS{ID,Float64}, geom::NodalField{Float64}, c::Context) where {ID}
    kinds = _is_iso(self.mcsys, FEMMShellQ4RSModule._isoparametric!) ? nothing : Union{Nothing,CSysKind}[get(_csys_kinds, self.mcsys, nothing)]
    return _associate!(c, self, geom, el -> self.mcsys, kinds, false)
end
for (FEMM, mod) in ((:FEMMShellT3FFComp, :FEMMShellT3FFCompModule), (:FEMMShellQ4RSComp, :FEMMShellQ4RSCompModule))
    @eval function $mod.associategeometry!(self::$FEMM{ID,Float64}, geom::NodalField{Float64}, c::Context) where {ID}
        gof = self._layup_group_lookup
        groups = self.layup_groups
        kinds = Union{Nothing,CSysKind}[get(_csys_kinds, g[1].csys, nothing) for g in groups]
        return _associate!(c, self, geom, el -> groups[gof[el]][1].csys, kinds, false)
    end
end

# ---- batched inspectintegpoints (src/FEMMShellT3FFModule.jl:850-962 and the three sibling methods) --------------------
# Same arguments as the reference method after the leading Context.  The resultants of all points come back as one
# array; the caller's `inspector(idat, i, conn, ecoords, out, loc)` is folded over it in the reference's order.
const _QUANTITY = Dict(:bending => 1, :moment => 1, :bending_moment => 1, :transverse_shear => 2, :transverse => 2,
                       :shear => 2, :membrane_force => 3, :membrane => 3)
"kind: 3 T3FF, 4 Q4RS, 13 T3FFComp, 14 Q4RSComp; outputcsys: nothing (element triad / layup csys) or 3x3xN matrices"
function shell_resultants(c::Context, params::ShellParams, kind::Integer, quantity::Symbol, u::NodalField{Float64}, npts::Integer,
        nelem::Integer; outputcsys::Union{Nothing,Array{Float64,3}} = nothing)
    out = Array{Float64,3}(undef, 3, npts, nelem)
    cs, ncs = outputcsys === nothing ? (C_NULL, 0) : (pointer(outputcsys), size(outputcsys, 3))
    uv = u.values
    GC.@preserve uv outputcsys out _check(ccall((:fsgpu_shell_resultants, libfsgpu), Cint,
        (Ptr{Cvoid}, Ref{ShellParams}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
        c.h, Ref(params), kind, _QUANTITY[quantity], pointer(uv), cs, ncs, pointer(out)))
    return out
end
"""
    shell_nodal_field(c, params, kind, quantity, u, nnodes; outputcsys) -> Matrix{Float64}(nnodes, 3)

`fieldfromintegpoints` with the default `nodevalmethod = :invdistance`, evaluated on the device (`fsgpu_shell_nodal_field`):
the three components of `quantity` as nodal fields.  FinEtools' own `fieldfromintegpoints` also works unchanged on top of
the `inspectintegpoints` methods below; this call avoids the per-point inspector callback.
"""
function shell_nodal_field(c::Context, params::ShellParams, kind::Integer, quantity::Symbol, u::NodalField{Float64}, nnodes::Integer;
        outputcsys::Union{Nothing,Array{Float64,3}} = nothing)
    out = Matrix{Float64}(undef, nnodes, 3)
    cs, ncs = outputcsys === nothing ? (C_NULL, 0) : (pointer(outputcsys), size(outputcsys, 3))
    uv = u.values
    GC.@preserve uv outputcsys out _check(ccall((:fsgpu_shell_nodal_field, libfsgpu), Cint,
        (Ptr{Cvoid}, Ref{ShellParams}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
        c.h, Ref(params), kind, _QUANTITY[quantity], pointer(uv), cs, ncs, pointer(out)))
    return out
end
function _inspect(c::Context, self, kind::Integer, comp::Bool, geom0, u, felist, inspector, idat, quantity; context...)
    fes = self.integdomain.fes
    nnpe = nodesperelem(fes)
    dchi = NodalField(zeros(size(geom0.values, 1), 6)); numberdofs!(dchi)   # the operator needs no numbering; any valid one
    alpha = comp ? _comp_setup!(c, self, geom0, dchi, nnpe) : _shell_setup!(c, self, geom0, dchi, nnpe == 4)
    npts = nnpe == 3 ? 1 : integrationdata(self.integdomain, self.integdomain.integration_rule)[1]
    ocs = nothing
    for (k, v) in context
        if k == :outputcsys      # evaluated per element centroid, like `updatecsmat!(outputcsys, centroid, J0, i, 0)`
            ocs = Array{Float64,3}(undef, 3, 3, count(fes))
            loc, J0 = fill(0.0, 1, 3), fill(0.0, 3, 2)
            for i in eachindex(fes)
                loc .= sum(geom0.values[collect(fes.conn[i]), :]; dims = 1) ./ nnpe
                updatecsmat!(v, loc, J0, i, 0)
                ocs[:, :, i] .= csmat(v)
            end
        end
    end
    res = shell_resultants(c, _shell_params(self, alpha; comp = comp), kind, quantity, u, npts, count(fes); outputcsys = ocs)
    Ns = nnpe == 4 ? integrationdata(self.integdomain, self.integdomain.integration_rule)[2] : nothing
    out = fill(0.0, 3)
    for i in felist
        ecoords = geom0.values[collect(fes.conn[i]), :]
        for j in 1:npts
            loc = nnpe == 3 ? sum(ecoords; dims = 1) ./ 3 : Ns[j]' * ecoords
            out .= view(res, :, j, i)
            idat = inspector(idat, i, fes.conn[i], ecoords, quantity in (:transverse_shear, :transverse, :shear) ? out[1:2] : out, loc)
        end
    end
    return idat
end
for (FEMM, mod, kind, comp) in ((:FEMMShellT3FF, :FEMMShellT3FFModule, 3, false), (:FEMMShellQ4RS, :FEMMShellQ4RSModule, 4, false),
                                (:FEMMShellT3FFComp, :FEMMShellT3FFCompModule, 13, true), (:FEMMShellQ4RSComp, :FEMMShellQ4RSCompModule, 14, true))
    @eval function $mod.inspectintegpoints(c::Context, self::$FEMM{ID,Float64}, geom0::NodalField{Float64}, u::NodalField{Float64}, dT::NodalField{Float64},
            felist::Vector{Int}, inspector::F, idat, quantity = :moment; context...) where {ID,F<:Function}
        return _inspect(c, self, $kind, $comp, geom0, u, felist, inspector, idat, quantity; context...)
    end
    @eval function $mod.inspectintegpoints(c::Context, self::$FEMM{ID,Float64}, geom0::NodalField{Float64}, u::NodalField{Float64},
            felist::Vector{Int}, inspector::F, idat, quantity = :moment; context...) where {ID,F<:Function}
        return _inspect(c, self, $kind, $comp, geom0, u, felist, inspector, idat, quantity; context...)
    end
end

# ---- corotational beam --------------------------------------------------------------------------
function _beam_setup!(c::Context, self::FEMMCorotBeam, geom0, u1, Rfield1, dchi)
    fes = self.integdomain.fes
    changed = _set_mesh!(c, fes, geom0)
    _set_dofs!(c, dchi)
    key = (:beam, objectid(self))
    if changed || c.aux_key === nothing || c.aux_key[1] != key
        A, I1, I2, I3, J, A2s, A3s, x1x2_vector, dimensions = properties(fes)
        xx = reduce(hcat, x1x2_vector)   # 3 x nelem
        GC.@preserve A I1 I2 I3 J A2s A3s xx _check(ccall((:fsgpu_set_beam_sections, libfsgpu), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            c.h, pointer(A), pointer(I1), pointer(I2), pointer(I3), pointer(J), pointer(A2s), pointer(A3s), pointer(xx)))
        c.aux_key = (key, Float64[])
    end
    # the displaced state changes every Newton iteration: always sent (nnodes x 12 doubles)
    GC.@preserve u1 Rfield1 _check(ccall((:fsgpu_set_state, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), c.h, pointer(u1.values), pointer(Rfield1.values)))
    return BeamParams(self.material.E, self.material.nu, massdensity(self.material), 1, 0)
end
for (fname, sym) in ((:stiffness, :fsgpu_corotbeam_stiffness), (:geostiffness, :fsgpu_corotbeam_geostiffness))
    q = QuoteNode(sym)
    @eval function FEMMCorotBeamModule.$fname(self::FEMMCorotBeam, assembler::SysmatAssemblerGPU, geom0::NodalField{Float64},
            u1::NodalField{T}, Rfield1::NodalField{T}, dchi::NodalField{TI}) where {T<:Number,TI<:Number}
        p = Ref(_beam_setup!(assembler.ctx, self, geom0, u1, Rfield1, dchi))
        _symbolic!(assembler)
        _check(ccall(($q, libfsgpu), Cint, (Ptr{Cvoid}, Ref{BeamParams}), assembler.ctx.h, p))
        return _fetch(assembler)
    end
end
function FEMMCorotBeamModule.mass(self::FEMMCorotBeam, assembler::SysmatAssemblerGPU, geom0::NodalField{Float64}, u1::NodalField{T},
        Rfield1::NodalField{T}, dchi::NodalField{TI}; mass_type = FEMMCorotBeamModule.MASS_TYPE_CONSISTENT_WITH_ROTATION_INERTIA) where {T<:Number,TI<:Number}
    p0 = _beam_setup!(assembler.ctx, self, geom0, u1, Rfield1, dchi)
    p = Ref(BeamParams(p0.E, p0.nu, p0.rho, mass_type, 0))
    _symbolic!(assembler)
    _check(ccall((:fsgpu_corotbeam_mass, libfsgpu), Cint, (Ptr{Cvoid}, Ref{BeamParams}), assembler.ctx.h, p))
    return _fetch(assembler)
end
# gyroscopic(self, assembler, geom0, u1, Rfield1, v1, dchi; mass_type)   (src/FEMMCorotBeamModule.jl:883-952)
function FEMMCorotBeamModule.gyroscopic(self::FEMMCorotBeam, assembler::SysmatAssemblerGPU, geom0::NodalField{Float64}, u1::NodalField{T},
        Rfield1::NodalField{T}, v1::NodalField{T}, dchi::NodalField{TI};
        mass_type = FEMMCorotBeamModule.MASS_TYPE_CONSISTENT_WITH_ROTATION_INERTIA) where {T<:Number,TI<:Number}
    p0 = _beam_setup!(assembler.ctx, self, geom0, u1, Rfield1, dchi)
    vv = v1.values      # nnodes x 6, column-major
    GC.@preserve vv _check(ccall((:fsgpu_set_velocity, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}), assembler.ctx.h, pointer(vv)))
    p = Ref(BeamParams(p0.E, p0.nu, p0.rho, mass_type, 0))
    _symbolic!(assembler)
    _check(ccall((:fsgpu_corotbeam_gyroscopic, libfsgpu), Cint, (Ptr{Cvoid}, Ref{BeamParams}), assembler.ctx.h, p))
    return _fetch(assembler)
end
function FEMMCorotBeamModule.restoringforce(self::FEMMCorotBeam, assembler::SysvecAssemblerGPU, geom0::NodalField{Float64},
        u1::NodalField{T}, Rfield1::NodalField{T}, dchi::NodalField{TI}) where {T<:Number,TI<:Number}
    p = Ref(_beam_setup!(assembler.ctx, self, geom0, u1, Rfield1, dchi))
    _check(ccall((:fsgpu_corotbeam_restoringforce, libfsgpu), Cint, (Ptr{Cvoid}, Ref{BeamParams}, Int32), assembler.ctx.h, p, assembler.nfree_only))
    return _fetch_vector(assembler.ctx, assembler.nfree_only ? nfreedofs(dchi) : nalldofs(dchi))
end
# distribloads_global(self, assembler, geom0, u1, Rfield1, dchi, fi)   (src/FEMMCorotBeamModule.jl:1186-1247):
# `fi` is a ForceIntensity; `updateforce!(fi, ignore, ignore, i, 0)` is evaluated per element on the host (it may depend
# on the element number), one 3-vector per element -- or a single one when all are equal
function FEMMCorotBeamModule.distribloads_global(self::FEMMCorotBeam, assembler::SysvecAssemblerGPU, geom0::NodalField{Float64},
        u1::NodalField{T}, Rfield1::NodalField{T}, dchi::NodalField{TI}, fi) where {T<:Number,TI<:Number}
    p = Ref(_beam_setup!(assembler.ctx, self, geom0, u1, Rfield1, dchi))
    fes = self.integdomain.fes
    ignore = fill(0.0, 0, 0)
    force = Matrix{Float64}(undef, 3, count(fes))
    for i in eachindex(fes)
        force[:, i] .= vec(updateforce!(fi, ignore, ignore, i, 0))
    end
    nforce = all(force[:, i] == force[:, 1] for i in axes(force, 2)) ? 1 : size(force, 2)
    GC.@preserve force _check(ccall((:fsgpu_corotbeam_distribloads, libfsgpu), Cint, (Ptr{Cvoid}, Ref{BeamParams}, Ptr{Float64}, Int64, Int32),
        assembler.ctx.h, p, pointer(force), nforce, assembler.nfree_only))
    return _fetch_vector(assembler.ctx, assembler.nfree_only ? nfreedofs(dchi) : nalldofs(dchi))
end

# ---- COO -> CSC (makematrix! of any FinEtools assembler's buffers) ---------------------------------
# Julia `sparse(I, J, V, m, n)`: duplicates combined left to right in input order (bitwise), zeros kept.  The size
# query does the conversion and keeps the result on the device; the second call only downloads it.
function sparse_gpu(c::Context, I::Vector{Int64}, J::Vector{Int64}, V::Vector{Float64}, m::Integer, n::Integer)
    nnz = Ref{Int64}(0)
    GC.@preserve I J V _check(ccall((:fsgpu_coo_to_csc, libfsgpu), Cint,
        (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ref{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
        c.h, m, n, length(I), pointer(I), pointer(J), pointer(V), nnz, C_NULL, C_NULL, C_NULL))
    colptr, rowval, nzval = Vector{Int64}(undef, n + 1), Vector{Int64}(undef, nnz[]), Vector{Float64}(undef, nnz[])
    GC.@preserve I J V colptr rowval nzval _check(ccall((:fsgpu_coo_to_csc, libfsgpu), Cint,
        (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ref{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
        c.h, m, n, length(I), pointer(I), pointer(J), pointer(V), nnz, pointer(colptr), pointer(rowval), pointer(nzval)))
    return SparseMatrixCSC(m, n, colptr, rowval, nzval)
end

# ---- explicit central differences (examples/.../plate_expl_examples.jl:61-94) ----------------------
mutable struct ExplicitGPU
    h::Ptr{Cvoid}
    n::Int
end
"K: SparseMatricesCSR.SparseMatrixCSR{1} (rowptr, colval, nzval), M: diagonal of the lumped mass"
function ExplicitGPU(c::Context, K, Mdiag::Vector{Float64}, c_scale::Float64, dt::Float64)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    rp, cv, nz = K.rowptr, K.colval, K.nzval
    GC.@preserve rp cv nz Mdiag _check(ccall((:fsgpu_explicit_create, libfsgpu), Cint,
        (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64),
        r, c.h, length(Mdiag), pointer(rp), pointer(cv), pointer(nz), pointer(Mdiag), c_scale, dt))
    e = ExplicitGPU(r[], length(Mdiag))
    finalizer(x -> ccall((:fsgpu_explicit_destroy, libfsgpu), Cint, (Ptr{Cvoid},), x.h), e)
    return e
end
"K_ff and the lumped M_ff never leave the device: the context's last matrix result (FFBLOCK stiffness) and vector result (shell_mass_diag)"
function ExplicitGPU(c::Context, n::Integer, c_scale::Float64, dt::Float64)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:fsgpu_explicit_create_from_ctx, libfsgpu), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Float64, Float64), r, c.h, c_scale, dt))
    e = ExplicitGPU(r[], Int(n))
    finalizer(x -> ccall((:fsgpu_explicit_destroy, libfsgpu), Cint, (Ptr{Cvoid},), x.h), e)
    return e
end
set_load!(e::ExplicitGPU, F0::Vector{Float64}) = GC.@preserve F0 _check(ccall((:fsgpu_explicit_set_load, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.h, pointer(F0)))
set_timestep!(e::ExplicitGPU, c_scale::Float64, dt::Float64) = _check(ccall((:fsgpu_explicit_set_timestep, libfsgpu), Cint, (Ptr{Cvoid}, Float64, Float64), e.h, c_scale, dt))
start!(e::ExplicitGPU, fscale0 = 1.0) = _check(ccall((:fsgpu_explicit_start, libfsgpu), Cint, (Ptr{Cvoid}, Float64), e.h, fscale0))
step!(e::ExplicitGPU, nsteps::Integer, fscale::Vector{Float64}) = GC.@preserve fscale _check(ccall((:fsgpu_explicit_step, libfsgpu), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), e.h, nsteps, pointer(fscale)))
function state(e::ExplicitGPU)
    U, V, A = zeros(e.n), zeros(e.n), zeros(e.n)
    GC.@preserve U V A _check(ccall((:fsgpu_explicit_get_state, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), e.h, pointer(U), pointer(V), pointer(A)))
    return U, V, A
end
"largest eigenvalue of M^-1 K by power iteration (the `pwr` helper of spherical_cap_expl_examples.jl:160-166)"
function omega_max_sq(e::ExplicitGPU, maxit::Integer = 30)
    lam = Ref{Float64}(0.0)
    _check(ccall((:fsgpu_explicit_omega_max, libfsgpu), Cint, (Ptr{Cvoid}, Int32, Ref{Float64}), e.h, maxit, lam))
    return lam[]
end
function kinetic_energy(e::ExplicitGPU)
    ke = Ref{Float64}(0.0)
    _check(ccall((:fsgpu_explicit_kinetic_energy, libfsgpu), Cint, (Ptr{Cvoid}, Ref{Float64}), e.h, ke))
    return ke[]
end
"""
    run!(e, nsteps, dt; force! = nothing, F0 = nothing, fscale = t -> 1.0, peek = nothing, nbtw = 0)

The loop of plate_expl_examples.jl:61-94 with its two closures.  `force!(F, t)` general (any spatial distribution
per step): the load vector is re-sent every step.  The common separable form `F(t) = fscale(t) F0` stays on the
device: the factors of a whole stretch are sampled into a table and the stretch runs without host interaction.
`peek(step, U, V, t)` is called with the state fetched every `nbtw` steps (and at step 0), as the example does.
"""
function run!(e::ExplicitGPU, nsteps::Integer, dt::Float64; force! = nothing, F0 = nothing, fscale = t -> 1.0, peek = nothing, nbtw::Integer = 0)
    F = zeros(e.n)
    if force! !== nothing
        force!(F, 0.0); set_load!(e, F); start!(e, 1.0)
    else
        set_load!(e, F0); start!(e, fscale(0.0))
    end
    if peek !== nothing
        U, V, _ = state(e); peek(0, U, V, 0.0)
    end
    step = 0
    while step < nsteps
        m = force! !== nothing ? 1 : (nbtw > 0 ? min(nbtw - step % nbtw, nsteps - step) : nsteps - step)
        if force! !== nothing
            force!(F, (step + 1) * dt); set_load!(e, F)
            step!(e, 1, [1.0])
        else
            step!(e, m, [fscale((step + k) * dt) for k in 1:m])
        end
        step += m
        if peek !== nothing && nbtw > 0 && step % nbtw == 0
            U, V, _ = state(e); peek(step, U, V, step * dt)
        end
    end
    return state(e)
end

# ---- update_rotation_field! (src/RotUtilModule.jl:29-42): R <- exp(dtheta) R per node, on the device ---------------
# updates the context's copy of Rfield1 (uploaded by the last beam operator, fsgpu_set_state) and writes it back
function update_rotation_field_gpu!(c::Context, Rfield::NodalField{Float64}, dchi::NodalField{Float64})
    dv, Rv = dchi.values, Rfield.values          # nnodes x 6, nnodes x 9 column-major
    GC.@preserve dv Rv _check(ccall((:fsgpu_update_rotation_field, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
        c.h, pointer(dv), pointer(Rv)))
    return Rfield
end

# ---- multi-GPU: this rank's column block of the global matrix (SURVEY section 8(e)) --------------------
# One Julia process per GPU (MPI.jl / NCCL.jl for the exchange).  `row_map`, `colcount`, `rowval`, `nzval` are DEVICE
# pointers (CuPtr of CUDA.jl arrays): the block is written in Julia's CSC layout straight into this rank's slice of
# the gathered arrays; see finetoolsflexstructures.jl_b200/partition.py for the plan (owned columns, local mesh).
function result_block!(c::Context, col_lo::Integer, col_hi::Integer, row_map, colcount, rowval, nzval)
    nb = Ref{Int64}(0)
    _check(ccall((:fsgpu_result_block, libfsgpu), Cint,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ref{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
        c.h, col_lo, col_hi, row_map, nb, colcount, rowval, nzval))
    return nb[]
end
result_block_size(c::Context, col_lo::Integer, col_hi::Integer) = result_block!(c, col_lo, col_hi, C_NULL, C_NULL, C_NULL, C_NULL)
# row-partitioned explicit loop: fsgpu_explicit_create_dist / _export / _connect exchange 1 KB blobs between the ranks
# (MPI.Allgather of the blobs), after which the step kernel writes halo displacements straight into the neighbours'
# peer-mapped windows; see include/fsgpu.h:298-330 and partition.py `connect_ranks`.

"""
order-fixed assembly: T3FF / T3FFComp use the atomics-free owner-computes tile kernel, every other element kind the gather
path (dense element matrices, one owner per matrix block summing in ascending element order = the reference loop's
order); values are bitwise reproducible.  Takes effect at the next symbolic phase.
"""
function set_deterministic!(c::Context, on::Bool = true)
    _check(ccall((:fsgpu_set_deterministic, libfsgpu), Cint, (Ptr{Cvoid}, Cint), c.h, on ? 1 : 0))
    c.sym_target = Int32(-1)
    return c
end

export SysmatAssemblerGPU, SysvecAssemblerGPU, Context, default_context, invalidate!, ExplicitGPU, sparse_gpu
export set_load!, set_timestep!, start!, step!, state, run!, omega_max_sq, kinetic_energy, set_deterministic!
export shell_resultants, shell_nodal_field, result_block!, result_block_size, update_rotation_field_gpu!, CSysKind, register_csys_kind!
export pinned_vector, refresh_values!
export SPARSE, SPARSE_SYMM, SPARSE_DIAG, FFBLOCK, FFBLOCK_DIAG, CSR_SYMM, CSYS_CYLINDRICAL, CSYS_SPHERICAL, CSYS_NORMAL_AXIS

end # module
