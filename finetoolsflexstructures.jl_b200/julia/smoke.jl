# Smoke test of the Julia glue: runs wherever `julia`, FinEtools, FinEtoolsDeforLinear and FinEtoolsFlexStructures
# are available next to a built libfsgpu.so and a CUDA device:
#
#     LIBFSGPU=/path/to/libfsgpu.so julia --project=<env with the packages> finetoolsflexstructures.jl_b200/julia/smoke.jl
#
# It assembles K and M of a small T3FF and a small Q4RS plate with the reference's own Julia loops and with the GPU
# assembler through the SAME operator calls, and compares colptr / rowval bit for bit and nzval to 1e-12.
# (scripts/gpu_final.sh calls it when `command -v julia` succeeds; neither the build image nor the GPU boxes of this
# project have Julia -- profiles/r02_julia_probe.txt records the probe.)
include(joinpath(@__DIR__, "FlexStructuresGPU.jl"))
using .FlexStructuresGPU
using FinEtools, FinEtoolsDeforLinear, FinEtoolsFlexStructures, LinearAlgebra, SparseArrays
using FinEtoolsFlexStructures.FESetShellT3Module: FESetShellT3
using FinEtoolsFlexStructures.FESetShellQ4Module: FESetShellQ4
using FinEtoolsFlexStructures.FEMMShellT3FFModule
using FinEtoolsFlexStructures.FEMMShellQ4RSModule
using FinEtoolsFlexStructures.RotUtilModule: initial_Rfield

function run(kind)
    E, nu, rho, t = 200e9, 0.3, 7850.0, 0.01
    fens, fes = kind == :t3 ? T3block(1.0, 1.0, 12, 12) : Q4block(1.0, 1.0, 12, 12)
    fens.xyz = xyz3(fens)
    for i in 1:count(fens)
        x, y = fens.xyz[i, 1], fens.xyz[i, 2]
        fens.xyz[i, 3] = 0.2 * sin(2x) + 0.1 * y^2
    end
    mater = MatDeforElastIso(DeforModelRed3D, rho, E, nu, 0.0)
    sfes = kind == :t3 ? FESetShellT3() : FESetShellQ4()
    accepttodelegate(fes, sfes)
    mod = kind == :t3 ? FEMMShellT3FFModule : FEMMShellQ4RSModule
    idom = kind == :t3 ? IntegDomain(fes, TriRule(1), t) : IntegDomain(fes, GaussRule(2, 2), t)
    femm = mod.make(idom, mater)
    geom0 = NodalField(fens.xyz)
    u0 = NodalField(zeros(size(fens.xyz, 1), 3))
    Rfield0 = initial_Rfield(fens)
    dchi = NodalField(zeros(size(fens.xyz, 1), 6))
    for i in selectnode(fens; box = [0.0 0.0 -Inf Inf -Inf Inf], inflate = 1e-6), d in 1:3
        setebc!(dchi, [i], true, d)
    end
    applyebc!(dchi); numberdofs!(dchi)
    mod.associategeometry!(femm, geom0)
    Kref = mod.stiffness(femm, SysmatAssemblerFFBlock(nfreedofs(dchi)), geom0, u0, Rfield0, dchi)
    Mref = mod.mass(femm, SysmatAssemblerFFBlock(nfreedofs(dchi)), geom0, dchi)
    a = SysmatAssemblerGPU(FFBLOCK)
    K = mod.stiffness(femm, a, geom0, u0, Rfield0, dchi)
    M = mod.mass(femm, a, geom0, dchi)          # shares the upload and the symbolic phase with K
    @assert K.colptr == Kref.colptr && K.rowval == Kref.rowval "pattern mismatch ($kind)"
    @assert norm(K.nzval - Kref.nzval) <= 1e-12 * norm(Kref.nzval) "stiffness values ($kind)"
    @assert M.colptr == Mref.colptr && M.rowval == Mref.rowval && norm(M.nzval - Mref.nzval) <= 1e-12 * norm(Mref.nzval) "mass ($kind)"
    # nodal normals on the device against the reference's
    n_ref, v_ref = copy(femm._normals), copy(femm._normal_valid)
    femm2 = mod.make(idom, mater)
    mod.associategeometry!(femm2, geom0, a.ctx)
    @assert maximum(abs.(femm2._normals .- n_ref)) < 1e-13 && femm2._normal_valid == v_ref "normals ($kind)"
    L = mod.stiffness(femm, SysmatAssemblerGPU(FFBLOCK; uplo = :L), geom0, u0, Rfield0, dchi)
    @assert norm(L - tril(Kref)) <= 1e-12 * norm(Kref) "lower triangle ($kind)"
    println("smoke $kind: nnz = ", nnz(K), "  rel. err K = ", norm(K.nzval - Kref.nzval) / norm(Kref.nzval))
end

run(:t3)
run(:q4)
println("FlexStructuresGPU smoke: OK")
