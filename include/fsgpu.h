/* fsgpu.h -- C ABI of libfsgpu.so: B200 (sm_100a) element-level hot path of
 * FinEtoolsFlexStructures.jl (reference v3.6.4, citations relative to /root/reference).
 *
 * What a host (Julia `ccall`, Python ctypes, C) binds.  Plain pointers and sizes only.
 * All host arrays use the JULIA layout of the reference objects, so they can be passed
 * zero-copy under `GC.@preserve`:
 *   conn      Int64, nnpe x nelem, 1-based  (Vector{NTuple{nnpe,Int}} of fes.conn)
 *   xyz       Float64, nnodes x 3 column-major      (geom0.values)
 *   dofnums   Int64,  nnodes x 6 column-major, 1-based (dchi.dofnums)
 *   normals   Float64, nnodes x 3 column-major      (femm._normals)
 *   valid     UInt8 (Bool), nnodes                   (femm._normal_valid)
 *   u1        Float64, nnodes x 3 column-major      (u1.values)
 *   Rfield1   Float64, nnodes x 9 column-major, each ROW a column-major 3x3 (Rfield1.values)
 * Returned sparse matrices are SparseMatrixCSC{Float64,Int64} pieces: colptr (ncols+1),
 * rowval (nnz), nzval (nnz), 1-based, rows ascending within a column, duplicates summed,
 * explicit zeros kept (except target SPARSE_SYMM) -- the semantics of Julia `sparse`.
 *
 * Every function returns 0 on success or an fsgpu_status code; the message is available
 * from fsgpu_last_error() (thread-local).  Nothing throws or aborts across the boundary.
 * A context is single-threaded (like the reference FEMMs, src/FEMMCorotBeamModule.jl:27-53)
 * and every call is synchronous unless stated otherwise.
 * There is NO CPU fallback: without a CUDA device fsgpu_create fails.
 */
#ifndef FSGPU_H
#define FSGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fsgpu_ctx fsgpu_ctx;

enum fsgpu_status {
  FSGPU_OK = 0,
  FSGPU_ERR_ARG = 1,        /* bad argument / size mismatch */
  FSGPU_ERR_CUDA = 2,       /* CUDA runtime error (message has the detail) */
  FSGPU_ERR_STATE = 3,      /* missing prerequisite, e.g. geometry not associated
                               (`@assert self._associatedgeometry`, src/FEMMShellT3FFModule.jl:643) */
  FSGPU_ERR_DOF_RANGE = 4,  /* dof < 1 or > nalldofs (FinEtools assemble! error) */
  FSGPU_ERR_SINGULAR = 5    /* "Singular metric matrix" (src/FEMMShellQ4RSModule.jl:308-311) */
};

/* Assembler semantics (FinEtools.AssemblyModule; in-tree
 * src/AssemblyModule.jl:20-53 for CSR_SYMM). */
enum fsgpu_target {
  FSGPU_SPARSE = 0,       /* SysmatAssemblerSparse: nall x nall, all entries incl. zeros */
  FSGPU_SPARSE_SYMM = 1,  /* SysmatAssemblerSparseSymm (default): S+S', exact zeros dropped */
  FSGPU_SPARSE_DIAG = 2,  /* SysmatAssemblerSparseDiag: diagonal entries only */
  FSGPU_FFBLOCK = 3,      /* SysmatAssemblerFFBlock(nfree): [1:nfree,1:nfree] of SPARSE */
  FSGPU_FFBLOCK_DIAG = 4, /* SysmatAssemblerFFBlock(SysmatAssemblerSparseDiag(), nfree, nfree) */
  FSGPU_CSR_SYMM = 5      /* SysmatAssemblerSparseCSRSymm: SPARSE returned as CSR
                             (colptr := rowptr, rowval := colval) */
};

/* Scalar parameters of the shell operators (mutable FEMM fields,
 * src/FEMMShellT3FFModule.jl:101-112, src/FEMMShellQ4RSModule.jl:75-85). */
typedef struct fsgpu_shell_params {
  double Dps[9];   /* plane-stress moduli, row-major 3x3 (_shell_material_stiffness) */
  double Dt[4];    /* transverse-shear moduli 2x2, NOT yet multiplied by 5/6 */
  double rho;      /* mass density (mass operator) */
  double stab_alpha; /* stab_fun(t,h) = t^2/(t^2 + alpha h^2); ignored if stab_factor set */
  double drilling_stiffness_scale;
  int32_t transv_shear_formulation; /* T3FF only: 0 = AVERAGE_B (default), 1 = AVERAGE_K */
  int32_t reserved;
} fsgpu_shell_params;

/* Scalar parameters of the corotational beam (FEMMCorotBeam.material). */
typedef struct fsgpu_beam_params {
  double E, nu, rho;
  int32_t mass_type; /* MASS_TYPE_* 0..3, src/FEMMCorotBeamModule.jl:144-147 */
  int32_t reserved;
} fsgpu_beam_params;

/* ---- context ------------------------------------------------------------------- */
int fsgpu_create(fsgpu_ctx** ctx, int device);
int fsgpu_destroy(fsgpu_ctx* ctx);
const char* fsgpu_last_error(void);
int fsgpu_version(void);
/* pinned host buffers (optional; make the result copies run at full PCIe rate) */
int fsgpu_host_alloc(void** p, int64_t bytes);
int fsgpu_host_free(void* p);
/* all work of a context goes to this stream (default: the legacy default stream) */
int fsgpu_set_stream(fsgpu_ctx* ctx, void* cuda_stream);
int fsgpu_sync(fsgpu_ctx* ctx);
/* number of kernels this context has launched so far */
int64_t fsgpu_launch_count(fsgpu_ctx* ctx);
/* bytes this context has moved device -> host so far (large patterns cross PCIe in compact form, so this is
 * less than the size of the host arrays fsgpu_fetch_matrix fills) */
int fsgpu_d2h_bytes(fsgpu_ctx* ctx, int64_t* bytes);
/* device time (CUDA events on the context's stream) of the element kernel of the last
 * matrix operator -- the number the roofline fraction is computed from */
int fsgpu_last_kernel_ms(fsgpu_ctx* ctx, double* ms);
/* on != 0: T3FF / T3FFComp stiffness uses the owner-computes tile kernel when the mesh allows it
 * (no atomics: bitwise reproducible values, every entry written once; schedule-driven, ~1.5x slower than the
 * RED.ADD scatter on B200).  Takes effect at the next fsgpu_symbolic.  Default: off, or the
 * environment variable FSGPU_TILE=1 at fsgpu_create.  The reference's serial loop is
 * deterministic by construction (src/FEMMShellT3FFModule.jl:664-733). */
int fsgpu_set_deterministic(fsgpu_ctx* ctx, int on);
/* scatter path of the last shell stiffness operator: 0 = per-entry slot map + RED.ADD,
 * 1 = run-structured addressing + RED.ADD, 2 = owner-computes tile kernel, -1 = none yet */
int fsgpu_scatter_path(fsgpu_ctx* ctx, int* path);
/* measurement support: FP64 FMA peak (TFLOP/s) and copy bandwidth (GB/s, read+write) of the
 * device, from micro-kernels -- the FP64 roofline denominator (not in MEASURED_PEAKS.json) */
int fsgpu_measure_peaks(fsgpu_ctx* ctx, double* fp64_tflops, double* copy_gbs);

/* ---- data hand-over (host pointers; copied to the device) ---------------------- */
/* fes.conn + geom0.values.  nnpe: 2 (L2 beam), 3 (T3), 4 (Q4).
 * (src/FEMMShellT3FFModule.jl:671, src/FEMMCorotBeamModule.jl:1133) */
int fsgpu_set_mesh(fsgpu_ctx* ctx, int32_t nnpe, int64_t nelem, const int64_t* conn, int64_t nnodes, const double* xyz);
/* dchi.dofnums, nfreedofs(dchi), nalldofs(dchi) (src/FEMMShellT3FFModule.jl:667,732) */
int fsgpu_set_dofnums(fsgpu_ctx* ctx, const int64_t* dofnums, int64_t nfree, int64_t nall);
/* femm._normals / femm._normal_valid as left by associategeometry!
 * (src/FEMMShellT3FFModule.jl:570-616); marks the geometry as associated. */
int fsgpu_set_normals(fsgpu_ctx* ctx, const double* normals, const uint8_t* valid);
/* OR compute them on the device with the default (isoparametric) csys:
 * T3: unweighted element normals; Q4: Jacobian-weighted (src/FEMMShellQ4RSModule.jl:472-525).
 * fixed_dir != NULL: use that direction instead (cartesian layup csys of the composite FEMMs,
 * src/FEMMShellT3FFCompModule.jl:489-543).  accumulate != 0 reproduces the homogeneous T3FF
 * quirk of not resetting the arrays. */
int fsgpu_associategeometry(fsgpu_ctx* ctx, double threshold_angle_deg, const double* fixed_dir, int32_t accumulate);
/* general csys (cylindrical, spherical, any CSys callback): dirs = 3 x nnpe x nelem column-major, the third column of
 * the csys matrix the host glue evaluated at every node of every element (`_compute_nodal_normal!`,
 * src/FEMMShellT3FFCompModule.jl:203-207,509); it replaces the element normal in the accumulation (and, for the Q4
 * elements, in the validity check, src/FEMMShellQ4RSModule.jl:508-519) */
int fsgpu_associategeometry_dirs(fsgpu_ctx* ctx, double threshold_angle_deg, const double* dirs, int32_t accumulate);
/* Built-in csys kinds evaluated ON THE DEVICE (the CSys callbacks of the reference's examples; only the kind, an
 * origin and an axis cross the boundary; origin NULL = (0,0,0); the axis is normalised):
 *   FSGPU_CSYS_CYLINDRICAL  e3 = radial from the axis, e2 = axis, e1 = e2 x e3
 *                           (examples/shells/dynamics/homogeneous/explicit/clamp_cyl_expl_examples.jl:62-68, axis = y)
 *   FSGPU_CSYS_SPHERICAL    e3 = radial from the origin, e1 = normalize(axis x e3), e2 = e3 x e1
 *                           (examples/shells/statics/homogeneous/hemisphere/hemisphere_examples.jl:31-39, axis = z)
 *   FSGPU_CSYS_NORMAL_AXIS  e3 = surface normal t1 x t2, e2 = axis, e1 = e2 x e3
 *                           (examples/shells/statics/homogeneous/miscellaneous/pressurized_cylinder_free_examples.jl:16-23)
 * fsgpu_associategeometry_csys: nodal normals from the csys evaluated at every node of every element.
 * fsgpu_set_layup_csys (after fsgpu_set_layup, whose csmat argument it replaces): layup csys matrices per element
 * (T3FFComp: at the centroid, src/FEMMShellT3FFCompModule.jl:617) or per element and integration point (Q4RSComp: with
 * the shape-function values Ns[j] as the location, as the reference passes them, src/FEMMShellQ4RSCompModule.jl:929). */
enum fsgpu_csys_kind { FSGPU_CSYS_CYLINDRICAL = 1, FSGPU_CSYS_SPHERICAL = 2, FSGPU_CSYS_NORMAL_AXIS = 3 };
int fsgpu_associategeometry_csys(fsgpu_ctx* ctx, double threshold_angle_deg, int32_t kind, const double* origin,
                                 const double* axis, int32_t accumulate);
int fsgpu_set_layup_csys(fsgpu_ctx* ctx, int32_t kind, const double* origin, const double* axis);
/* the same in two halves for element-partitioned runs: after _accumulate the host sums the
 * interface-node entries of *dev_sums ([nnodes][3] doubles, device) across ranks; after _finish it
 * min-combines the validity flags, 4th component of *dev_normals4 ([nnodes][4] doubles, device) */
int fsgpu_normals_accumulate(fsgpu_ctx* ctx, const double* fixed_dir, int32_t accumulate, double** dev_sums);
int fsgpu_normals_finish(fsgpu_ctx* ctx, double threshold_angle_deg, const double* fixed_dir, double** dev_normals4);
int fsgpu_get_normals(fsgpu_ctx* ctx, double* normals, uint8_t* valid);
/* integdomain.otherdimension evaluated by the host glue: n = 1 (uniform), nelem (T3: at the
 * centroid) or nelem*npts (Q4: per integration point, element-major).
 * (src/FEMMShellT3FFModule.jl:677, src/FEMMShellQ4RSModule.jl:924) */
int fsgpu_set_thickness(fsgpu_ctx* ctx, const double* t, int64_t n);
/* optional user stab_fun values evaluated on the host, per element (overrides stab_alpha) */
int fsgpu_set_stab_factor(fsgpu_ctx* ctx, const double* f, int64_t n);
int fsgpu_element_sizes(fsgpu_ctx* ctx, double* h); /* h per element (T3 sqrt(2Ae), Q4 quirk diameter) */
/* Q4 integration rule as data: npts points (xi, eta, w). */
int fsgpu_set_rule(fsgpu_ctx* ctx, int32_t npts, const double* xi, const double* eta, const double* w);
/* Composite layup groups (src/FEMMShellT3FFCompModule.jl:600-605): per group 34 doubles
 * [A(9) B(9) D(9) H(4) thickness mass_density moi_density] row-major matrices, from
 * laminate_stiffnesses!/laminate_transverse_stiffness!/laminate_inertia!;
 * group_of_elem Int64 1-based (femm._layup_group_lookup); csmat: layup csys matrix,
 * column-major 3x3, ncs = 1 (constant), nelem (T3 centroid) or nelem*npts (Q4, App. B.9). */
int fsgpu_set_layup(fsgpu_ctx* ctx, int32_t ngroups, const double* group_data, const int64_t* group_of_elem,
                    const double* csmat, int64_t ncs);
/* FESetL2Beam per-element section arrays (src/FESetL2BeamModule.jl:20-31); x1x2 flattened 3 x nelem. */
int fsgpu_set_beam_sections(fsgpu_ctx* ctx, const double* A, const double* I1, const double* I2, const double* I3,
                            const double* J, const double* A2s, const double* A3s, const double* x1x2);
/* u1.values, Rfield1.values (live for the beam, dead arguments for the shells, App. B.11) */
int fsgpu_set_state(fsgpu_ctx* ctx, const double* u1, const double* Rfield1);

/* ---- symbolic phase (once per mesh + dofnums + target) --------------------------- */
/* startassembly! equivalent: builds the CSC pattern, bit-exact to `sparse`, and the
 * element-entry -> nzval slot map.  For SPARSE_SYMM nnz is an upper bound until an
 * operator has run. */
int fsgpu_symbolic(fsgpu_ctx* ctx, int32_t target, int64_t* nrows, int64_t* ncols, int64_t* nnz);

/* ---- operators: one call per reference operator --------------------------------- */
/* Result stays on the device; fetch it with fsgpu_fetch_matrix / fsgpu_fetch_vector. */
int fsgpu_t3ff_stiffness(fsgpu_ctx* ctx, const fsgpu_shell_params* p);     /* src/FEMMShellT3FFModule.jl:635-736 */
int fsgpu_t3ff_mass(fsgpu_ctx* ctx, const fsgpu_shell_params* p);          /* :757-811 */
int fsgpu_q4rs_stiffness(fsgpu_ctx* ctx, const fsgpu_shell_params* p);     /* src/FEMMShellQ4RSModule.jl:877-947 */
int fsgpu_q4rs_mass(fsgpu_ctx* ctx, const fsgpu_shell_params* p);          /* :968-1022 */
int fsgpu_t3ffcomp_stiffness(fsgpu_ctx* ctx, const fsgpu_shell_params* p); /* src/FEMMShellT3FFCompModule.jl:561-689 */
int fsgpu_t3ffcomp_mass(fsgpu_ctx* ctx, const fsgpu_shell_params* p);      /* :710-770 */
int fsgpu_q4rscomp_stiffness(fsgpu_ctx* ctx, const fsgpu_shell_params* p); /* src/FEMMShellQ4RSCompModule.jl:861-958 */
int fsgpu_q4rscomp_mass(fsgpu_ctx* ctx, const fsgpu_shell_params* p);      /* :979-1038 */
int fsgpu_corotbeam_stiffness(fsgpu_ctx* ctx, const fsgpu_beam_params* p);    /* src/FEMMCorotBeamModule.jl:972-1023 */
int fsgpu_corotbeam_geostiffness(fsgpu_ctx* ctx, const fsgpu_beam_params* p); /* :1042-1094 */
int fsgpu_corotbeam_mass(fsgpu_ctx* ctx, const fsgpu_beam_params* p);         /* :810-868 */
/* v1.values (nnodes x 6 column-major), needed by gyroscopic only */
int fsgpu_set_velocity(fsgpu_ctx* ctx, const double* v1);
int fsgpu_corotbeam_gyroscopic(fsgpu_ctx* ctx, const fsgpu_beam_params* p);   /* :883-952 */
/* distribloads_global (:1186-1247): uniform force per unit length in global components, 3 values
 * (nforce = 1) or 3 x nelem (nforce = nelem); result is a vector like restoringforce */
int fsgpu_corotbeam_distribloads(fsgpu_ctx* ctx, const fsgpu_beam_params* p, const double* force, int64_t nforce,
                                 int32_t nfree_only);
/* vector operators; nfree_only != 0: SysvecAssemblerFBlock(nfree), else SysvecAssembler */
int fsgpu_corotbeam_restoringforce(fsgpu_ctx* ctx, const fsgpu_beam_params* p, int32_t nfree_only); /* :1112-1159 */
/* lumped shell mass as a diagonal VECTOR over all dofs (the diag(M) the explicit loop uses,
 * examples/.../plate_expl_examples.jl:69-71); kind 3 = T3FF, 4 = Q4RS, 13 = T3FFComp, 14 = Q4RSComp */
int fsgpu_shell_mass_diag(fsgpu_ctx* ctx, const fsgpu_shell_params* p, int32_t kind, int32_t nfree_only);
/* inspectintegpoints, batched (no per-point host callback): stress resultants of the T3FF (kind 3, one point
 * per element) / Q4RS (kind 4, one per integration point) shells and their laminated variants (kind 13 / 14:
 * A, B, D, H of fsgpu_set_layup rotated into the element frame, B-coupling included)
 * (src/FEMMShellT3FFModule.jl:850-962, src/FEMMShellQ4RSModule.jl:1061-1170,
 *  src/FEMMShellT3FFCompModule.jl:809-943, src/FEMMShellQ4RSCompModule.jl:1061-1210 -- the latter applies the
 *  global->element transformation twice as written; replicated).
 * quantity: 1 = bending moments (m11, m22, m12), 2 = transverse shear forces (q1, q2, 0),
 * 3 = membrane forces (n11, n22, n12), in the output csys: ncs = 0 -> the default (homogeneous: the element
 * triad; laminated: the layup csys), else ncs = 1 | nelem | nelem*npts column-major 3x3 matrices.
 * u: nnodes x 6 column-major (the displacement/rotation field); out: 3 x npts x nelem. */
int fsgpu_shell_resultants(fsgpu_ctx* ctx, const fsgpu_shell_params* p, int32_t kind, int32_t quantity, const double* u,
                           const double* outputcsys, int64_t ncs, double* out);
/* fieldfromintegpoints (FinEtools FEMMBaseModule, nodevalmethod = :invdistance; test/test_shell_resultants.jl:123-126): the
 * same resultants averaged to the nodes on the device, weights 1 / (squared distance node - integration point); same
 * arguments as fsgpu_shell_resultants; out: nnodes x 3 column-major (the three components of the quantity as nodal fields) */
int fsgpu_shell_nodal_field(fsgpu_ctx* ctx, const fsgpu_shell_params* p, int32_t kind, int32_t quantity, const double* u,
                            const double* outputcsys, int64_t ncs, double* out);
/* R <- exp(dtheta) R per node (src/RotUtilModule.jl:29-42); dchi_values nnodes x 6 column-major */
int fsgpu_update_rotation_field(fsgpu_ctx* ctx, const double* dchi_values, double* Rfield_out);

/* parity / debug: raw element matrices, n x n x nelem column-major (Julia elmat per element).
 * op: 0 stiffness, 1 mass, 2 geostiffness (beam), 3 gyroscopic (beam).  kind as in fsgpu_shell_mass_diag, 2 = beam. */
int fsgpu_element_matrices(fsgpu_ctx* ctx, int32_t kind, int32_t op, const void* params, double* out);
/* element vectors of restoringforce, 12 x nelem */
int fsgpu_element_vectors(fsgpu_ctx* ctx, const fsgpu_beam_params* p, double* out);

/* ---- results -------------------------------------------------------------------- */
int fsgpu_result_size(fsgpu_ctx* ctx, int64_t* nrows, int64_t* ncols, int64_t* nnz);
/* makematrix! equivalent; any pointer may be NULL to skip that array.  The arrays may be pinned (fsgpu_host_alloc:
 * direct DMA, fastest) or ordinary pageable memory (a Julia Vector): large pageable arrays are filled through the
 * library's pinned staging rings by host threads (FSGPU_HOST_THREADS, FSGPU_VALUE_THREADS), INTEGRATION.md section 6 */
int fsgpu_fetch_matrix(fsgpu_ctx* ctx, int64_t* colptr, int64_t* rowval, double* nzval);
/* One triangle of a square CSC result, diagonal included: uplo = 'L' (rows >= column) or 'U' (rows <= column).
 * For consumers that read one triangle of a symmetric operator (Julia: cholesky(Symmetric(K, :L)), the
 * SysmatAssemblerSparseSymm convention of src/AssemblyModule.jl / FinEtools keeps i >= j only): half the bytes
 * cross PCIe.  Same two-call convention: size query, then the caller's arrays are filled. */
int fsgpu_result_size_uplo(fsgpu_ctx* ctx, int32_t uplo, int64_t* nnz);
int fsgpu_fetch_matrix_uplo(fsgpu_ctx* ctx, int32_t uplo, int64_t* colptr, int64_t* rowval, double* nzval);
int fsgpu_fetch_vector(fsgpu_ctx* ctx, double* out, int64_t n);
/* device-resident handles (int32 0-based pattern, float64 values) for chaining on the GPU */
int fsgpu_result_device(fsgpu_ctx* ctx, const int32_t** colptr0, const int32_t** rowval0, const double** nzval);
int fsgpu_vector_device(fsgpu_ctx* ctx, const double** v, int64_t* n);
/* multi-GPU gathering of assembled blocks (option A: each rank assembles every element
 * that touches the nodes whose columns it owns, so its owned columns are complete and no partial sums are
 * exchanged).  Writes columns [col_lo, col_hi) (0-based, half open, this context's numbering) of the result as
 * pieces of the GLOBAL SparseMatrixCSC into caller-owned DEVICE buffers -- normally this rank's slice of the
 * gathered arrays, so the host's collective (NCCL broadcast / all-gather) runs in place:
 *   colcount_dev [col_hi-col_lo] Int64 stored entries per column (global colptr = 1 + running sum),
 *   rowval_dev   [*nnz_block]    Int64 1-based global rows = row_map_dev[local 0-based row]
 *                                (row_map_dev: device Int64 [nrows]; NULL = identity + 1),
 *   nzval_dev    [*nnz_block]    Float64.
 * Any output pointer may be NULL; with all three NULL the call only reports *nnz_block.
 * Targets SPARSE, FFBLOCK and their DIAG forms (a column of SPARSE_SYMM needs mirror entries of columns
 * other ranks own; gather SPARSE and symmetrise on the host side instead). */
int fsgpu_result_block(fsgpu_ctx* ctx, int64_t col_lo, int64_t col_hi, const int64_t* row_map_dev, int64_t* nnz_block,
                       int64_t* colcount_dev, int64_t* rowval_dev, double* nzval_dev);

/* ---- standalone COO -> CSC (Julia `sparse(I,J,V,m,n)`; makematrix! of any assembler) --- */
/* two calls: (1) colptr/rowval/nzval NULL -> *nnz; (2) fill caller-owned arrays. */
int fsgpu_coo_to_csc(fsgpu_ctx* ctx, int64_t m, int64_t n, int64_t ntriples, const int64_t* I, const int64_t* J,
                     const double* V, int64_t* nnz, int64_t* colptr, int64_t* rowval, double* nzval);

/* ---- explicit central-difference loop (examples/.../plate_expl_examples.jl:61-94) -- */
typedef struct fsgpu_explicit fsgpu_explicit;
/* K: CSR of the free-free block (Int64 1-based, as SparseMatricesCSR.sparsecsr gives it) and
 * diag(M); both copied to the device.  c_scale = ksi*2*omegad (C = c_scale * diag(M)). */
int fsgpu_explicit_create(fsgpu_explicit** h, fsgpu_ctx* ctx, int64_t n, const int64_t* rowptr, const int64_t* colval,
                          const double* nzval, const double* mdiag, double c_scale, double dt);
/* same, adopting device-resident results: K = last matrix result of `ctx` (must be FFBLOCK),
 * M = last vector result (fsgpu_shell_mass_diag with nfree_only) */
int fsgpu_explicit_create_from_ctx(fsgpu_explicit** h, fsgpu_ctx* ctx, double c_scale, double dt);
/* storage layout of the loop's stiffness (measurement support): rows, stored entries, the number of runs of
 * consecutive rows that share one column pattern (the 6 dofs of a shell node), and the column indices one SpMV
 * actually reads (one pattern per run): traffic per step = 8 nnz + 4 index_entries + vector passes */
int fsgpu_explicit_layout(fsgpu_explicit* h, int64_t* nrows, int64_t* nnz, int64_t* nruns, int64_t* index_entries);
int fsgpu_explicit_destroy(fsgpu_explicit* h);
int fsgpu_explicit_set_state(fsgpu_explicit* h, const double* U0, const double* V0);
/* new time step / damping factor after the handle was created (dt = 0.9 * 2/omega_max * ... is only known after
 * fsgpu_explicit_omega_max, examples/.../spherical_cap_expl_examples.jl:166-173): C and invMC are rebuilt */
int fsgpu_explicit_set_timestep(fsgpu_explicit* h, double c_scale, double dt);
/* constant load vector F0 scaled by a per-step factor table (force!(F,t) closures of the
 * examples are constant or windowed sines): F(t_k) = fscale[k] * F0; fscale NULL -> 1 */
int fsgpu_explicit_set_load(fsgpu_explicit* h, const double* F0);
/* A0 = invMC .* F(0) (plate_expl_examples.jl:81) */
int fsgpu_explicit_start(fsgpu_explicit* h, double fscale0);
/* advance nsteps; fscale[nsteps] or NULL */
int fsgpu_explicit_step(fsgpu_explicit* h, int64_t nsteps, const double* fscale);
int fsgpu_explicit_get_state(fsgpu_explicit* h, double* U, double* V, double* A);
/* y = K x on the device copy (ThreadedSparseCSR.bmul! equivalent) */
int fsgpu_explicit_spmv(fsgpu_explicit* h, const double* x, double* y);
/* largest eigenvalue of K x = lambda M x by power iteration (pwr_largest,
 * examples/.../spherical_cap_expl_examples.jl:166) */
int fsgpu_explicit_omega_max(fsgpu_explicit* h, int32_t maxit, double* lambda_max);
/* kinetic energy 1/2 V' M V (peek closures) */
int fsgpu_explicit_kinetic_energy(fsgpu_explicit* h, double* ke);
/* multi-GPU: device pointers of U, V, A, F-scratch for halo exchange by the host's NCCL plumbing.  U is the CURRENT
 * displacement buffer: _step / _step_begin alternate between two buffers, ask again after every advancing call. */
int fsgpu_explicit_device_state(fsgpu_explicit* h, double** U, double** V, double** A, double** E);
/* split step for element-partitioned runs: (1) U update + E = K_local U, (2) after the host has
 * summed interface entries of E across ranks: finish the step */
int fsgpu_explicit_step_begin(fsgpu_explicit* h);
int fsgpu_explicit_step_end(fsgpu_explicit* h, double fscale);


/* ---- row-partitioned explicit loop: ONE global mesh, one rank (process or context) per GPU ------------
 * SURVEY section 8(e); loop: examples/shells/dynamics/homogeneous/explicit/plate_expl_examples.jl:61-94 with the
 * node numbering of :119-121,147 (any numbering works; a banded one keeps the interfaces short).
 * Rank r owns the contiguous global rows [bounds[r], bounds[r+1]) (0-based) of K_ff, M and of all state vectors.
 * Its context has assembled, on a LOCAL mesh that contains every element touching a node with an owned row
 * (interface elements are assembled by both neighbours, no partial sums are exchanged), the FFBLOCK stiffness and
 * the nfree_only lumped-mass vector; local rows are numbered by an order-preserving map, loc2glob[k] = global 1-based
 * row of local row k (NULL: identity), and the own rows are the local rows [row_lo, row_hi).
 * Columns of other ranks become HALO entries of the displacement vector.  All peer-visible state of a rank lives in
 * one device allocation (the window) that the other ranks map -- cudaIpc across processes, the plain pointer for
 * contexts of one process -- and the fused step kernel writes the boundary entries of the next displacements
 * straight into the neighbours' windows over NVLink and raises a flag there; the neighbours' boundary rows wait
 * for that flag, their interior rows do not.  No host round trip, no NCCL call, no separate exchange launch per step.
 *
 * Bootstrap (like an NCCL unique id, but all-to-all): every rank calls _export, the host all-gathers the blobs
 * (MPI.Allgather, torch.distributed.all_gather, a shared file ...) in rank order and hands them to _connect.
 * After that the fsgpu_explicit_* calls below (set_state, set_load, start, step, get_state, spmv, omega_max,
 * kinetic_energy, destroy) are COLLECTIVE: every rank calls them in the same order; vectors are the rank's own
 * rows.  omega_max and kinetic_energy return the global value on every rank (partial sums are combined in rank
 * order: identical bits everywhere).  A wait for a peer gives up after FSGPU_PEER_TIMEOUT_MS (default 20 000)
 * and the call returns FSGPU_ERR_STATE instead of hanging.  If the context still uses the legacy default
 * stream the handle gives it a private non-blocking one (kernels that wait for peers must not serialise with
 * other work of the process). */
#define FSGPU_EXPLICIT_BLOB_BYTES 1024
int fsgpu_explicit_create_dist(fsgpu_explicit** h, fsgpu_ctx* ctx, int32_t rank, int32_t world, int64_t row_lo,
                               int64_t row_hi, const int64_t* loc2glob, const int64_t* bounds, double c_scale, double dt);
int fsgpu_explicit_export(fsgpu_explicit* h, void* blob /* FSGPU_EXPLICIT_BLOB_BYTES */);
int fsgpu_explicit_connect(fsgpu_explicit* h, const void* blobs /* world x FSGPU_EXPLICIT_BLOB_BYTES, rank order */);
/* partition facts (measurement support): own rows, halo entries, entries pushed per step, row runs that read halo
 * entries (they are scheduled first), neighbouring ranks */
int fsgpu_explicit_dist_info(fsgpu_explicit* h, int64_t* n_own, int64_t* n_halo, int64_t* n_push, int64_t* boundary_runs,
                             int32_t* npeers);

#ifdef __cplusplus
}
#endif
#endif /* FSGPU_H */
