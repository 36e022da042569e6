"""
    FlexStructuresGPU

`ccall` glue that makes libfsgpu.so a drop-in for the element-level hot path of
FinEtoolsFlexStructures.jl.  The plugin point is the assembler type: passing a
`SysmatAssemblerGPU` / `SysvecAssemblerGPU` to the EXISTING operators

    stiffness(femm, assembler, geom0, u1, Rfield1, dchi)
    mass(femm, assembler, geom0, dchi)                       # shells
    mass(femm, assembler, geom0, u1, Rfield1, dchi; mass_type)   # beam
    geostiffness(femm, assembler, geom0, u1, Rfield1, dchi)
    restoringforce(femm, assembler, geom0, u1, Rfield1, dchi)

dispatches to the methods below, which replace the per-element Julia loop
(e.g. src/FEMMShellT3FFModule.jl:670-734) by ONE library call per operator.  Every other
assembler keeps using the original Julia methods, as does any FEMM whose number type is not
Float64 (the ForwardDiff use in examples/shells/statics/homogeneous/plates/
ss_circular_plate_udl_examples.jl:129-187).

NOTE: the build image of the B200 port has no Julia, so this file is the mechanical binding a
maintainer adds; it has been written against include/fsgpu.h but not executed there.  The same
call sequence is exercised by the Python mirror (finetoolsflexstructures.jl_b200/femm.py).
"""
module FlexStructuresGPU

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

# enum fsgpu_target
const SPARSE, SPARSE_SYMM, SPARSE_DIAG, FFBLOCK, FFBLOCK_DIAG, CSR_SYMM = Int32.(0:5)

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

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        _check(ccall((:fsgpu_create, libfsgpu), Cint, (Ref{Ptr{Cvoid}}, Cint), r, device))
        c = new(r[])
        finalizer(x -> ccall((:fsgpu_destroy, libfsgpu), Cint, (Ptr{Cvoid},), x.h), c)
        return c
    end
end

"""
    SysmatAssemblerGPU(target; device = 0)

GPU assembler carrying the semantic target of a FinEtools assembler:
`SPARSE` (SysmatAssemblerSparse), `SPARSE_SYMM` (SysmatAssemblerSparseSymm, the default of
the convenience methods), `SPARSE_DIAG`, `FFBLOCK` (SysmatAssemblerFFBlock(nfreedofs)),
`FFBLOCK_DIAG`, `CSR_SYMM` (SysmatAssemblerSparseCSRSymm, src/AssemblyModule.jl:20-53).
"""
struct SysmatAssemblerGPU <: AbstractSysmatAssembler
    target::Int32
    ctx::Context
end
SysmatAssemblerGPU(target = SPARSE_SYMM; device = 0) = SysmatAssemblerGPU(Int32(target), Context(device))

struct SysvecAssemblerGPU <: AbstractSysvecAssembler
    nfree_only::Bool
    ctx::Context
end
SysvecAssemblerGPU(nfree_only = false; device = 0) = SysvecAssemblerGPU(nfree_only, Context(device))

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

# ---- data hand-over ---------------------------------------------------------------------------
function _set_mesh!(c::Context, fes, geom0)
    conn = fes.conn  # Vector{NTuple{nnpe,Int}}: nnpe x nelem Int64, inline
    GC.@preserve conn geom0 _check(ccall((:fsgpu_set_mesh, libfsgpu), Cint,
        (Ptr{Cvoid}, Int32, Int64, Ptr{Int64}, Int64, Ptr{Float64}),
        c.h, nodesperelem(fes), count(fes), pointer(conn), size(geom0.values, 1), pointer(geom0.values)))
end
function _set_dofs!(c::Context, dchi)
    GC.@preserve dchi _check(ccall((:fsgpu_set_dofnums, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Int64),
        c.h, pointer(dchi.dofnums), nfreedofs(dchi), nalldofs(dchi)))
end
function _set_normals!(c::Context, femm)
    n, v = femm._normals, femm._normal_valid   # Matrix{Float64} nnodes x 3, Vector{Bool}
    GC.@preserve n v _check(ccall((:fsgpu_set_normals, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}), c.h, pointer(n), pointer(v)))
end
# thickness callback evaluated on the host, once per element (T3: at the centroid)
function _set_thickness_t3!(c::Context, femm, geom0)
    fes = femm.integdomain.fes
    ipc = [(1.0 / 3) (1.0 / 3)]
    t = Vector{Float64}(undef, count(fes))
    centroid = fill(0.0, 1, 3)
    for i in eachindex(fes)
        cn = fes.conn[i]
        centroid .= (geom0.values[cn[1]:cn[1], :] .+ geom0.values[cn[2]:cn[2], :] .+ geom0.values[cn[3]:cn[3], :]) ./ 3
        t[i] = femm.integdomain.otherdimension(centroid, cn, ipc)
    end
    all(==(t[1]), t) && (t = t[1:1])
    GC.@preserve t _check(ccall((:fsgpu_set_thickness, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), c.h, pointer(t), length(t)))
end
function _set_rule_and_thickness_q4!(c::Context, femm, geom0)
    npts, Ns, gradNparams, w, pc = integrationdata(femm.integdomain, femm.integdomain.integration_rule)
    xi, eta, ww = pc[:, 1], pc[:, 2], vec(w)
    GC.@preserve xi eta ww _check(ccall((:fsgpu_set_rule, libfsgpu), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        c.h, npts, pointer(xi), pointer(eta), pointer(ww)))
    fes = femm.integdomain.fes
    t = Matrix{Float64}(undef, npts, count(fes))
    loc = fill(0.0, 1, 3)
    for i in eachindex(fes), j in 1:npts
        loc .= Ns[j]' * geom0.values[collect(fes.conn[i]), :]
        t[j, i] = femm.integdomain.otherdimension(loc, fes.conn[i], Ns[j])
    end
    tv = all(==(t[1]), t) ? t[1:1] : vec(t)
    GC.@preserve tv _check(ccall((:fsgpu_set_thickness, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), c.h, pointer(tv), length(tv)))
end
# stab_fun: the library evaluates t^2/(t^2 + alpha h^2); any other closure is sampled on the host
function _stab_alpha(femm, default)
    f = femm.stab_fun
    t, h = 0.37, 1.91
    v = f(t, h)
    alpha = (t^2 / v - t^2) / h^2
    ok = isapprox(f(0.11, 0.7), 0.11^2 / (0.11^2 + alpha * 0.7^2); rtol = 1e-13)
    return ok ? alpha : NaN
end

function _shell_params(femm; comp = false, default_alpha)
    Dps, Dt = comp ? (zeros(3, 3), zeros(2, 2)) : FEMMShellT3FFModule._shell_material_stiffness(femm.material)
    rho = comp ? 0.0 : massdensity(femm.material)
    tsf = hasproperty(femm, :transv_shear_formulation) ? femm.transv_shear_formulation : 0
    ShellParams(Tuple(permutedims(Dps)), Tuple(permutedims(Dt)), rho, _stab_alpha(femm, default_alpha), femm.drilling_stiffness_scale, tsf, 0)
end

function _fetch(a::SysmatAssemblerGPU)
    m, n, nnz = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    _check(ccall((:fsgpu_result_size, libfsgpu), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), a.ctx.h, m, n, nnz))
    colptr, rowval, nzval = Vector{Int64}(undef, n[] + 1), Vector{Int64}(undef, nnz[]), Vector{Float64}(undef, nnz[])
    GC.@preserve colptr rowval nzval _check(ccall((:fsgpu_fetch_matrix, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
        a.ctx.h, pointer(colptr), pointer(rowval), pointer(nzval)))
    if a.target == CSR_SYMM   # SparseMatricesCSR.SparseMatrixCSR{1}(m, n, rowptr, colval, nzval)
        return FinEtoolsFlexStructures.AssemblyModule.SparseMatricesCSR.SparseMatrixCSR{1}(m[], n[], colptr, rowval, nzval)
    end
    return SparseMatrixCSC(m[], n[], colptr, rowval, nzval)
end

function _symbolic!(a::SysmatAssemblerGPU)
    nr, nc, nnz = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    _check(ccall((:fsgpu_symbolic, libfsgpu), Cint, (Ptr{Cvoid}, Int32, Ref{Int64}, Ref{Int64}, Ref{Int64}), a.ctx.h, a.target, nr, nc, nnz))
end

# ---- shells -----------------------------------------------------------------------------------
for (FEMM, mod, op, alpha, isq4) in ((:FEMMShellT3FF, :FEMMShellT3FFModule, :t3ff, 5 / 12 / 1.5, false), (:FEMMShellQ4RS, :FEMMShellQ4RSModule, :q4rs, 0.1, true))
    kstiff, kmass = QuoteNode(Symbol(:fsgpu_, op, :_stiffness)), QuoteNode(Symbol(:fsgpu_, op, :_mass))
    @eval function $mod.stiffness(self::$FEMM{ID,Float64}, assembler::SysmatAssemblerGPU, geom0::NodalField{Float64},
            u1::NodalField{TI}, Rfield1::NodalField{TI}, dchi::NodalField{TI}) where {ID,TI<:Number}
        @assert self._associatedgeometry == true
        c = assembler.ctx
        _set_mesh!(c, self.integdomain.fes, geom0); _set_dofs!(c, dchi); _set_normals!(c, self)
        $(isq4 ? :(_set_rule_and_thickness_q4!(c, self, geom0)) : :(_set_thickness_t3!(c, self, geom0)))
        _symbolic!(assembler)
        p = Ref(_shell_params(self; default_alpha = $alpha))
        _check(ccall(($kstiff, libfsgpu), Cint, (Ptr{Cvoid}, Ref{ShellParams}), c.h, p))
        return _fetch(assembler)
    end
    @eval function $mod.mass(self::$FEMM{ID,Float64}, assembler::SysmatAssemblerGPU, geom0::NodalField{Float64}, dchi::NodalField{TI}) where {ID,TI<:Number}
        @assert self._associatedgeometry == true
        c = assembler.ctx
        _set_mesh!(c, self.integdomain.fes, geom0); _set_dofs!(c, dchi); _set_normals!(c, self)
        $(isq4 ? :(_set_rule_and_thickness_q4!(c, self, geom0)) : :(_set_thickness_t3!(c, self, geom0)))
        _symbolic!(assembler)
        p = Ref(_shell_params(self; default_alpha = $alpha))
        _check(ccall(($kmass, libfsgpu), Cint, (Ptr{Cvoid}, Ref{ShellParams}), c.h, p))
        return _fetch(assembler)
    end
end

# layered shells: the O(nplies) through-thickness integration stays on the host (reference functions)
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
    # layup csys evaluated per element centroid (T3) -- a constant csys collapses to one matrix
    cs = Array{Float64}(undef, 3, 3, count(fes))
    centroid, J0 = fill(0.0, 1, 3), fill(0.0, 3, 2)
    for i in eachindex(fes)
        layup = groups[gof[i]][1]
        centroid .= sum(geom0.values[collect(fes.conn[i]), :]; dims = 1) ./ nnpe
        updatecsmat!(layup.csys, centroid, J0, -1, 0)
        cs[:, :, i] .= csmat(layup.csys)
    end
    ncs = all(cs[:, :, i] == cs[:, :, 1] for i in axes(cs, 3)) ? 1 : size(cs, 3)
    GC.@preserve rec gof cs _check(ccall((:fsgpu_set_layup, libfsgpu), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Int64),
        c.h, length(groups), pointer(rec), pointer(gof), pointer(cs), ncs))
end
for (FEMM, mod, op, alpha, nn) in ((:FEMMShellT3FFComp, :FEMMShellT3FFCompModule, :t3ffcomp, 5 / 12 / 1.5, 3), (:FEMMShellQ4RSComp, :FEMMShellQ4RSCompModule, :q4rscomp, 0.1, 4))
    for (fname, sym, args) in ((:stiffness, Symbol(:fsgpu_, op, :_stiffness), :(geom0::NodalField{Float64}, u1::NodalField{TI}, Rfield1::NodalField{TI}, dchi::NodalField{TI})),
                               (:mass, Symbol(:fsgpu_, op, :_mass), :(geom0::NodalField{Float64}, dchi::NodalField{TI})))
        q = QuoteNode(sym)
        @eval function $mod.$fname(self::$FEMM{ID,Float64}, assembler::SysmatAssemblerGPU, $(args.args...)) where {ID,TI<:Number}
            @assert self._associatedgeometry == true
            c = assembler.ctx
            _set_mesh!(c, self.integdomain.fes, geom0); _set_dofs!(c, dchi); _set_normals!(c, self)
            if $nn == 4
                npts, Ns, gradNparams, w, pc = integrationdata(self.integdomain, self.integdomain.integration_rule)
                xi, eta, ww = pc[:, 1], pc[:, 2], vec(w)
                GC.@preserve xi eta ww _check(ccall((:fsgpu_set_rule, libfsgpu), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), c.h, npts, pointer(xi), pointer(eta), pointer(ww)))
            end
            _set_layup!(c, self, geom0, $nn)
            _symbolic!(assembler)
            p = Ref(_shell_params(self; comp = true, default_alpha = $alpha))
            _check(ccall(($q, libfsgpu), Cint, (Ptr{Cvoid}, Ref{ShellParams}), c.h, p))
            return _fetch(assembler)
        end
    end
end

# ---- corotational beam --------------------------------------------------------------------------
function _beam_setup!(c::Context, self::FEMMCorotBeam, geom0, u1, Rfield1, dchi)
    fes = self.integdomain.fes
    _set_mesh!(c, fes, geom0); _set_dofs!(c, dchi)
    A, I1, I2, I3, J, A2s, A3s, x1x2_vector, dimensions = properties(fes)
    xx = reduce(hcat, x1x2_vector)   # 3 x nelem
    GC.@preserve A I1 I2 I3 J A2s A3s xx _check(ccall((:fsgpu_set_beam_sections, libfsgpu), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        c.h, pointer(A), pointer(I1), pointer(I2), pointer(I3), pointer(J), pointer(A2s), pointer(A3s), pointer(xx)))
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
function FEMMCorotBeamModule.restoringforce(self::FEMMCorotBeam, assembler::SysvecAssemblerGPU, geom0::NodalField{Float64},
        u1::NodalField{T}, Rfield1::NodalField{T}, dchi::NodalField{TI}) where {T<:Number,TI<:Number}
    p = Ref(_beam_setup!(assembler.ctx, self, geom0, u1, Rfield1, dchi))
    _check(ccall((:fsgpu_corotbeam_restoringforce, libfsgpu), Cint, (Ptr{Cvoid}, Ref{BeamParams}, Int32), assembler.ctx.h, p, assembler.nfree_only))
    n = assembler.nfree_only ? nfreedofs(dchi) : nalldofs(dchi)
    F = Vector{Float64}(undef, n)
    GC.@preserve F _check(ccall((:fsgpu_fetch_vector, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), assembler.ctx.h, pointer(F), n))
    return F
end

# ---- COO -> CSC (makematrix! of any FinEtools assembler's buffers) ---------------------------------
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
set_load!(e::ExplicitGPU, F0::Vector{Float64}) = GC.@preserve F0 _check(ccall((:fsgpu_explicit_set_load, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.h, pointer(F0)))
start!(e::ExplicitGPU, fscale0 = 1.0) = _check(ccall((:fsgpu_explicit_start, libfsgpu), Cint, (Ptr{Cvoid}, Float64), e.h, fscale0))
step!(e::ExplicitGPU, nsteps::Integer, fscale::Vector{Float64}) = GC.@preserve fscale _check(ccall((:fsgpu_explicit_step, libfsgpu), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), e.h, nsteps, pointer(fscale)))
function state(e::ExplicitGPU)
    U, V, A = zeros(e.n), zeros(e.n), zeros(e.n)
    GC.@preserve U V A _check(ccall((:fsgpu_explicit_get_state, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), e.h, pointer(U), pointer(V), pointer(A)))
    return U, V, A
end

# ---- associategeometry! with a non-default csys (cylindrical, spherical, any CSys callback) ---------------------------
# The reference evaluates the csys AT EVERY NODE OF EVERY ELEMENT (`_compute_nodal_normal!`,
# src/FEMMShellT3FFCompModule.jl:203-207,509); the closure stays in Julia, the device accumulates, normalises, validates.
function associategeometry_dirs!(c::Context, fes, geom0::NodalField{Float64}, csys::CSys, threshold_angle::Float64; accumulate::Bool = false)
    nnpe = nodesperelem(fes)
    dirs = Array{Float64,3}(undef, 3, nnpe, count(fes))
    J0 = zeros(3, 2)
    for (el, conn) in enumerate(fes.conn)
        J0[:, 1] .= geom0.values[conn[2], :] .- geom0.values[conn[1], :]
        J0[:, 2] .= geom0.values[conn[end], :] .- geom0.values[conn[1], :]
        for (k, n) in enumerate(conn)
            updatecsmat!(csys, reshape(geom0.values[n, :], 1, 3), J0, el, 0)
            dirs[:, k, el] .= view(csmat(csys), :, 3)
        end
    end
    GC.@preserve dirs _check(ccall((:fsgpu_associategeometry_dirs, libfsgpu), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}, Int32),
        c.h, threshold_angle, pointer(dirs), accumulate ? 1 : 0))
    return c
end

# ---- update_rotation_field! (src/RotUtilModule.jl:29-42): R <- exp(dtheta) R per node, on the device ---------------
# updates the context's copy of Rfield1 (uploaded by the last beam operator, fsgpu_set_state) and writes it back
function update_rotation_field_gpu!(c::Context, Rfield::NodalField{Float64}, dchi::NodalField{Float64})
    dv, Rv = dchi.values, Rfield.values          # nnodes x 6, nnodes x 9 column-major
    GC.@preserve dv Rv _check(ccall((:fsgpu_update_rotation_field, libfsgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
        c.h, pointer(dv), pointer(Rv)))
    return Rfield
end

# ---- batched inspectintegpoints (src/FEMMShellT3FFModule.jl:850-962 and the three sibling methods) ----
# The reference calls `inspector(idat, i, conn, ecoords, out, loc)` per point; a Julia closure cannot cross the C
# boundary, so the GPU method returns the 3 x npts x nelem array and the caller folds its inspector over it.
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

"bitwise reproducible T3FF/T3FFComp stiffness (atomics-free owner-computes kernel); takes effect at the next symbolic phase"
set_deterministic!(c::Context, on::Bool = true) = _check(ccall((:fsgpu_set_deterministic, libfsgpu), Cint, (Ptr{Cvoid}, Cint), c.h, on ? 1 : 0))

export SysmatAssemblerGPU, SysvecAssemblerGPU, Context, ExplicitGPU, sparse_gpu, set_load!, start!, step!, state, set_deterministic!
export shell_resultants, result_block!, result_block_size, associategeometry_dirs!, update_rotation_field_gpu!
export SPARSE, SPARSE_SYMM, SPARSE_DIAG, FFBLOCK, FFBLOCK_DIAG, CSR_SYMM

end # module
