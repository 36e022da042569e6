// Shared between the shell element kernels (fsgpu_elements.cu) and the tile kernel (fsgpu_tile.cu).
#pragma once
#include "fsgpu_internal.cuh"
#include "fsgpu_math.cuh"

namespace fsk {
using namespace fsm;
using fs::Rule;

struct ShellArgs {
  const int32_t* conn;
  const double4* xyz;
  const double4* nrm;
  const double* thick;
  int64_t nthick;
  const double* stabf;
  int64_t nstab;
  int64_t nelem;
  double Dps[9], Dt[4];  // Dt already x 5/6
  HomogFactors hf;       // LDL' factors of Dps and Dt (host)
  double rho, alpha, drill;
  // composite
  const double* gdata;
  const int32_t* gof;
  const double* cs;
  int64_t ncs;
  Rule rule;
  int32_t* flag;
  // Q4RSComp: factored laminate constitutive data per (element, integration point), [nelem][npts][24] doubles =
  // strictly-lower L6 (15, row-major), sqrt of the 6 + 2 pivots (8), L2 (1); written by k_q4_laminate_prep
  const double* lam;
};

__device__ __forceinline__ double4 ldg4(const double4* p) {
  const double2* q = reinterpret_cast<const double2*>(p);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ V3 ld3(const double4* p, int i) {
  const double4 v = ldg4(p + i);
  return v3(v.x, v.y, v.z);
}


__device__ __forceinline__ void build_constit_t3(const ShellArgs& P, int64_t e, const Triad& E, double Ae, double shear_scale,
                                                 bool comp, Constit& C) {
  const double h2 = 2 * Ae;  // h^2, h = sqrt(2 Ae)
  if (comp) {
    const double* gd = P.gdata + (size_t)__ldg(P.gof + e) * 34;
    const double t = gd[31];
    const double stab = P.nstab ? __ldg(P.stabf + e) : t * t * fs_rcp(t * t + P.alpha * h2);
    double m, n;
    layup_angle(E, P.cs + (P.ncs == 1 ? 0 : e * 9), m, n);
    constit_laminate(gd, gd + 9, gd + 18, gd + 27, m, n, Ae, stab * Ae * shear_scale, C);
  } else {
    const double t = P.nthick == 1 ? __ldg(P.thick) : __ldg(P.thick + e);
    const double stab = P.nstab ? __ldg(P.stabf + e) : t * t * fs_rcp(t * t + P.alpha * h2);
    constit_homogeneous(P.Dps, P.Dt, t * Ae, (t * t * t) * (1.0 / 12.0) * Ae, t * stab * Ae * shear_scale, C);
  }
}


// host: fills the kernel argument block from the context + operator parameters
int shell_args(fsgpu_ctx* c, const fsgpu_shell_params* p, int nnpe, bool comp, bool need_normals, ShellArgs& A);

// T3 fast path: per-warp emission plan (fsgpu_elements.cu), built by the symbolic phase once per mesh
int t3_build_plan(fsgpu_ctx* c);

// T3 tile (owner-computes, atomics-free) path: symbolic data and launch (fsgpu_tile.cu)
int tile_symbolic(fsgpu_ctx* c);
int launch_t3_tile(fsgpu_ctx* c, const ShellArgs& A, bool comp);

}  // namespace fsk
