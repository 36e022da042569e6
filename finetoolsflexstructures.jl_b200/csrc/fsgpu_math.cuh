// Element-level FP64 math for the shell / beam kernels (sm_100a).
//
// Formulation (not a transcription of the reference's dense 18x18 / 24x24 products):
// every shell element matrix is built as
//        K = sum_s d_s * b_s (x) b_s        (+ drilling term)
// where b_s (one row per generalized strain component and integration point) is the
// strain-displacement row expressed directly in GLOBAL dofs, i.e. the reference's
// B * T_ae * T_ga (src/FEMMShellQ4RSModule.jl:926-936; for T3FF the two QtEQ transforms
// of src/FEMMShellT3FFModule.jl:712,730 folded into B), and the constitutive matrix is
// folded in through a unit-lower LDL^T factorisation (D = L diag(d) L^T, b <- L^T b).
// The zero structure of T_ae (src/FEMMShellT3FFModule.jl:421-463) is exploited
// analytically: the drilling-consistency coupling collapses to two 8x3 matrices P1, P2.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FS_HD __host__ __device__ __forceinline__
#else
#define FS_HD inline
#endif

namespace fsm {

struct V3 {
  double x, y, z;
};
FS_HD V3 v3(double x, double y, double z) { return V3{x, y, z}; }
FS_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
FS_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
FS_HD V3 operator*(double s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
FS_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
FS_HD V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// Reciprocal, reciprocal square root and square root without the library routines' special-case paths (denormals,
// infinities; the operands here are element sizes and moduli): one MUFU seed + two Newton steps, <= 1-2 ulp.  The
// division/sqrt sequences were ~10 % of the warp-instructions of the shell stiffness kernels.  Host build (the CPU
// check of these formulas, tests/hostmath) and -DFS_PRECISE_DIV use the IEEE operations.
#if defined(__CUDA_ARCH__) && !defined(FS_PRECISE_DIV)
__device__ __forceinline__ double fs_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ double fs_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  double e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-h * y, y, 0.5);
  return fma(y, e, y);
}
__device__ __forceinline__ double fs_sqrt(double x) {
  const double y = fs_rsqrt(x);
  double s = x * y;
  s = fma(fma(-s, s, x), 0.5 * y, s);
  return x > 0.0 ? s : 0.0;
}
#else
FS_HD double fs_rcp(double x) { return 1.0 / x; }
FS_HD double fs_rsqrt(double x) { return 1.0 / sqrt(x); }
FS_HD double fs_sqrt(double x) { return sqrt(x); }
#endif
FS_HD double norm(V3 a) { return fs_sqrt(dot(a, a)); }

// Element triad: e1 along the first tangent, e3 = e1 x t2 normalised, e2 = e3 x e1.
// (src/FEMMShellT3FFModule.jl:283-305, src/FEMMShellQ4RSModule.jl:248-263)
struct Triad {
  V3 e1, e2, e3;
};
FS_HD Triad element_triad(V3 t1, V3 t2) {
  Triad E;
  E.e1 = fs_rsqrt(dot(t1, t1)) * t1;
  V3 n = cross(E.e1, t2);
  E.e3 = fs_rsqrt(dot(n, n)) * n;
  E.e2 = cross(E.e3, E.e1);
  return E;
}

// Nodal triad A (nodal basis vectors in columns, element-basis components): rotation by
// the VECTOR r = e3 x n_e, whose length is sin(theta) (reference quirk),
// identity when |r| <= 1e-12.  a[r][c].   (src/FEMMShellT3FFModule.jl:355-388)
struct M3 {
  double a[3][3];
};
FS_HD M3 nodal_triad(const Triad& E, V3 nk, bool valid) {
  double nx = 0.0, ny = 0.0;
  if (valid) {
    nx = dot(E.e1, nk);
    ny = dot(E.e2, nk);
  }
  // r = (0,0,1) x (nx,ny,nz) = (-ny, nx, 0)
  double rx = -ny, ry = nx;
  const double r2 = rx * rx + ry * ry;
  M3 A;
  if (r2 > 1.0e-24) {  // |r| > 1e-12
    const double inr = fs_rsqrt(r2), nr = r2 * inr;
    double ux = rx * inr, uy = ry * inr;
    double s, c;
#if defined(__CUDA_ARCH__)
    // (measured: Taylor polynomials without range reduction, 2 x 10 dependent FMAs, are SLOWER than the library sincos --
    // T3 3.36 against 3.11 ms, Q4 2.26 against 2.23 ms: the chains are latency, not throughput)
    sincos(nr, &s, &c);
#else
    s = sin(nr);
    c = cos(nr);
#endif
    // c (I - u u') + s skew(u) + u u'
    double uxx = ux * ux, uxy = ux * uy, uyy = uy * uy;
    A.a[0][0] = c * (1.0 - uxx) + uxx;
    A.a[0][1] = c * (-uxy) + uxy;
    A.a[0][2] = s * uy;
    A.a[1][0] = c * (-uxy) + uxy;
    A.a[1][1] = c * (1.0 - uyy) + uyy;
    A.a[1][2] = -s * ux;
    A.a[2][0] = -s * uy;
    A.a[2][1] = s * ux;
    A.a[2][2] = c;
  } else {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) A.a[i][j] = (i == j) ? 1.0 : 0.0;
  }
  return A;
}

// G = A' E'  (global -> nodal, 3x3).  (src/FEMMShellT3FFModule.jl:398-419)
FS_HD M3 global_to_nodal(const M3& A, const Triad& E) {
  const double Et[3][3] = {{E.e1.x, E.e1.y, E.e1.z}, {E.e2.x, E.e2.y, E.e2.z}, {E.e3.x, E.e3.y, E.e3.z}};
  M3 G;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) G.a[i][j] = A.a[0][i] * Et[0][j] + A.a[1][i] * Et[1][j] + A.a[2][i] * Et[2][j];
  return G;
}

// Unit-lower LDL' of a symmetric NxN matrix, no pivoting.  L is stored below the
// diagonal of `a` on return, `d` the pivots.
template <int N>
FS_HD void ldlt(double (&a)[N][N], double (&d)[N]) {
  for (int j = 0; j < N; ++j) {
    double dj = a[j][j];
    for (int k = 0; k < j; ++k) dj -= a[j][k] * a[j][k] * d[k];
    d[j] = dj;
    const double idj = (dj != 0.0) ? fs_rcp(dj) : 0.0;
    for (int i = j + 1; i < N; ++i) {
      double v = a[i][j];
      for (int k = 0; k < j; ++k) v -= a[i][k] * a[j][k] * d[k];
      a[i][j] = v * idj;
    }
  }
}

// Constitutive data of one element / integration point in factored form.
//  rows 0..5 : membrane (3) + curvature (3) strains, D6 = L6 diag(d6) L6'
//  rows 6..7 : transverse shear,               D2 = L2 diag(d2) L2'
// Homogeneous shells: D6 = blockdiag(cm Dps, cb Dps) -> L6 is block diagonal.
struct Constit {
  double L6[6][6];  // strictly-lower part used
  double d6[6];
  double L2;  // single off-diagonal entry L2[1][0]
  double d2[2];
};

// Plane-stress rotation of a symmetric 3x3 resultant matrix, s = Tbar' X Tbar with
// Tbar = Tinv(m,n)'  (src/CompositeLayupModule.jl:364-378,410-421; called at
// src/FEMMShellT3FFCompModule.jl:621-624).
FS_HD void rotate_ps(const double X[9], double m, double n, double (&out)[3][3]) {
  const double mm = m * m, nn = n * n, mn = m * n;
  // Tbar = Tinv' ; Tinv rows: [mm nn 2mn; nn mm -2mn; -mn mn mm-nn]
  const double Tb[3][3] = {{mm, nn, -mn}, {nn, mm, mn}, {2 * mn, -2 * mn, mm - nn}};
  double XT[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) XT[i][j] = X[i * 3 + 0] * Tb[0][j] + X[i * 3 + 1] * Tb[1][j] + X[i * 3 + 2] * Tb[2][j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[i][j] = Tb[0][i] * XT[0][j] + Tb[1][i] * XT[1][j] + Tb[2][i] * XT[2][j];
}
// Transverse shear: s = Tts' H Tts, Tts = [m -n; n m]  (src/CompositeLayupModule.jl:463-471).
FS_HD void rotate_ts(const double H[4], double m, double n, double (&out)[2][2]) {
  const double T[2][2] = {{m, -n}, {n, m}};
  double HT[2][2];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) HT[i][j] = H[i * 2 + 0] * T[0][j] + H[i * 2 + 1] * T[1][j];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) out[i][j] = T[0][i] * HT[0][j] + T[1][i] * HT[1][j];
}

// cos / sin of the layup-to-element angle (src/TransformerModule.jl:92-105).
// `cs` = layup csys matrix, row-major cs[r*3+c].  1-m^2 is clamped at 0.
FS_HD void layup_angle(const Triad& E, const double cs[9], double& m, double& n) {
  const V3 c1 = v3(cs[0], cs[3], cs[6]), c2 = v3(cs[1], cs[4], cs[7]);
  double M11 = dot(E.e1, c1), M21 = dot(E.e2, c1), M12 = dot(E.e1, c2), M22 = dot(E.e2, c2);
  const double n1 = fs_rsqrt(M11 * M11 + M21 * M21), n2 = fs_rsqrt(M12 * M12 + M22 * M22);
  M11 *= n1;
  M21 *= n1;
  M12 *= n2;
  M22 *= n2;
  m = (M11 + M22) / 2;
  double nn = (M12 - M21) / 2;
  double q = 1.0 - m * m;
  q = q > 0.0 ? q : 0.0;
  n = (nn >= 0.0 ? 1.0 : -1.0) * fs_sqrt(q);
}

// Build the factored constitutive data.
//  homogeneous: Dps (3x3 sym, row-major), Dt (2x2, already x5/6), weights cm, cb, cs.
FS_HD void constit_homogeneous(const double Dps[9], const double Dt[4], double cm, double cb, double cs, Constit& C) {
  double a[3][3], d[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = Dps[i * 3 + j];
  ldlt<3>(a, d);
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) C.L6[i][j] = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < i; ++j) {
      C.L6[i][j] = a[i][j];
      C.L6[3 + i][3 + j] = a[i][j];
    }
  for (int i = 0; i < 3; ++i) {
    C.d6[i] = cm * d[i];
    C.d6[3 + i] = cb * d[i];
  }
  double h[2][2] = {{Dt[0], Dt[1]}, {Dt[2], Dt[3]}}, dd[2];
  ldlt<2>(h, dd);
  C.L2 = h[1][0];
  C.d2[0] = cs * dd[0];
  C.d2[1] = cs * dd[1];
}
//  laminate: A, B, D (3x3 row-major), H (2x2) of the layup group, rotated by (m, n);
//  weights: c (membrane/bending/coupling), cs (shear).
FS_HD void constit_laminate(const double A[9], const double B[9], const double D[9], const double H[4], double m, double n,
                            double c, double cs, Constit& C) {
  double sA[3][3], sB[3][3], sD[3][3], sH[2][2];
  rotate_ps(A, m, n, sA);
  rotate_ps(B, m, n, sB);
  rotate_ps(D, m, n, sD);
  rotate_ts(H, m, n, sH);
  double a[6][6];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      a[i][j] = sA[i][j];
      a[i][3 + j] = sB[i][j];
      a[3 + i][j] = sB[j][i];
      a[3 + i][3 + j] = sD[i][j];
    }
  // the reference keeps only the upper triangle of each product (complete_lt!), i.e. it
  // effectively symmetrises; use the symmetric part explicitly.
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < i; ++j) {
      double v = a[j][i];
      a[i][j] = v;
    }
  ldlt<6>(a, C.d6);
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) C.L6[i][j] = (j < i) ? a[i][j] : 0.0;
  for (int i = 0; i < 6; ++i) C.d6[i] *= c;
  double h[2][2] = {{sH[0][0], sH[0][1]}, {sH[0][1], sH[1][1]}}, dd[2];
  ldlt<2>(h, dd);
  C.L2 = h[1][0];
  C.d2[0] = cs * dd[0];
  C.d2[1] = cs * dd[1];
}

// ---------------------------------------------------------------------------------
// "B in global dofs" for ONE node of a shell element (shared by T3FF and Q4RS).
//
// Element-basis B of node l (cols u v w tx ty tz; rows 3 membrane, 3 curvature, 2 shear):
//   membrane  [gx 0 0 | 0 0 0; 0 gy 0 | 0 0 0; gy gx 0 | 0 0 0]
//   curvature [0 0 0 | 0 gx 0; 0 0 0 | -gy 0 0; 0 0 0 | -gx gy 0]
//   shear     [0 0 bs[r][0] | bs[r][1] bs[r][2] 0]
// (src/FEMMShellT3FFModule.jl:539-561, src/FEMMShellQ4RSModule.jl:548-578).
// With T_ae of src/FEMMShellT3FFModule.jl:421-463 and the block-diagonal T_ga = A'E' (:398-419):
//  * translations: B_l[:,0:3] A_l (A_l' E') = B_l[:,0:3] E'   (A_l is a rotation), plus the
//    drilling-consistency coupling  cpl_l = 1/2 (gx_l P2 - gy_l P1),  P1|2 = sum_m q_m (x) A_m[0|1,:],
//    q_m = B_m[:,tx] A_m[0][2]/A_m[2][2] + B_m[:,ty] A_m[1][2]/A_m[2][2]   (rows 3..7 only);
//  * rotations: B_l[:,tx:ty] R_l G_l[0:2,:],  R = A[0:2,0:2] - A[0:2,2] A[0:2,2]'/A[2][2].
// ---------------------------------------------------------------------------------
// rotation-dof columns of the element-basis B, rows 3..7 -> index 0..4
FS_HD void node_brot(double gx, double gy, const double (&bs)[2][3], double (&c3)[5], double (&c4)[5]) {
  c3[0] = 0.0;
  c4[0] = gx;
  c3[1] = -gy;
  c4[1] = 0.0;
  c3[2] = -gx;
  c4[2] = gy;
  c3[3] = bs[0][1];
  c4[3] = bs[0][2];
  c3[4] = bs[1][1];
  c4[4] = bs[1][2];
}
// this node's contribution to P1, P2 (5 x 3 each)
FS_HD void node_coupling_contrib(const M3& A, double gx, double gy, const double (&bs)[2][3], double (&p1)[5][3],
                                 double (&p2)[5][3]) {
  const double ia = fs_rcp(A.a[2][2]);
  const double m1 = ia * A.a[0][2], m2 = ia * A.a[1][2];
  double c3[5], c4[5];
  node_brot(gx, gy, bs, c3, c4);
  for (int r = 0; r < 5; ++r) {
    const double q = c3[r] * m1 + c4[r] * m2;
    for (int k = 0; k < 3; ++k) {
      p1[r][k] = q * A.a[0][k];
      p2[r][k] = q * A.a[1][k];
    }
  }
}
// the same in factored form: p1[r][k] = q[r] A[0][k], p2[r][k] = q[r] A[1][k] -- 5 + 6 numbers instead of 30, for
// kernels that exchange the contributions between lanes
FS_HD void node_coupling_factors(const M3& A, double gx, double gy, const double (&bs)[2][3], double (&q)[5]) {
  const double ia = fs_rcp(A.a[2][2]);
  const double m1 = ia * A.a[0][2], m2 = ia * A.a[1][2];
  double c3[5], c4[5];
  node_brot(gx, gy, bs, c3, c4);
  for (int r = 0; r < 5; ++r) q[r] = c3[r] * m1 + c4[r] * m2;
}
// 2x2 reduced rotation block
FS_HD void node_R(const M3& A, double (&R)[2][2]) {
  const double ia = fs_rcp(A.a[2][2]);
  for (int rw = 0; rw < 2; ++rw)
    for (int cl = 0; cl < 2; ++cl) R[rw][cl] = A.a[rw][cl] - ia * A.a[rw][2] * A.a[cl][2];
}
// nodal-basis rotation columns (theta_1, theta_2) of rows 3..7: brn[r][cl]
FS_HD void node_bt_rot(double gx, double gy, const double (&bs)[2][3], const double (&R)[2][2], double (&brn)[5][2]) {
  double c3[5], c4[5];
  node_brot(gx, gy, bs, c3, c4);
  for (int r = 0; r < 5; ++r)
    for (int cl = 0; cl < 2; ++cl) brn[r][cl] = c3[r] * R[0][cl] + c4[r] * R[1][cl];
}
// Unfolded global-dof strip of one node, by constitutive row group (the groups are
// independent, so a kernel can build, fold and store them one after the other).
FS_HD void strip_membrane(const Triad& E, double gx, double gy, double (&m)[3][6]) {
  const double e1[3] = {E.e1.x, E.e1.y, E.e1.z}, e2[3] = {E.e2.x, E.e2.y, E.e2.z};
  for (int c = 0; c < 3; ++c) {
    m[0][c] = gx * e1[c];
    m[1][c] = gy * e2[c];
    m[2][c] = gy * e1[c] + gx * e2[c];
    m[0][3 + c] = m[1][3 + c] = m[2][3 + c] = 0.0;
  }
}
// one of the rows 3..7 (r = 0..4): w = shear w-entry (0 for the curvature rows)
FS_HD void strip_row(const Triad& E, const M3& G, const double (&brn)[5][2], double gx, double gy, double w, const double (&P1)[5][3],
                     const double (&P2)[5][3], int r, double (&row)[6]) {
  const double e3[3] = {E.e3.x, E.e3.y, E.e3.z};
  const double hx = 0.5 * gx, hy = 0.5 * gy;
  double cp[3];
  for (int k = 0; k < 3; ++k) cp[k] = hx * P2[r][k] - hy * P1[r][k];
  for (int c = 0; c < 3; ++c) {
    row[c] = w * e3[c] + cp[0] * G.a[0][c] + cp[1] * G.a[1][c] + cp[2] * G.a[2][c];
    row[3 + c] = brn[r][0] * G.a[0][c] + brn[r][1] * G.a[1][c];
  }
}
// Whole strip bg[8][6].  P1, P2 are the element's summed coupling matrices.  Returns the nodal
// normal direction in global components (third row of G).
FS_HD V3 node_strip(const Triad& E, const M3& A, double gx, double gy, const double (&bs)[2][3], const double (&P1)[5][3],
                    const double (&P2)[5][3], double (&bg)[8][6]) {
  const M3 G = global_to_nodal(A, E);
  double R[2][2], brn[5][2];
  node_R(A, R);
  node_bt_rot(gx, gy, bs, R, brn);
  double m[3][6];
  strip_membrane(E, gx, gy, m);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 6; ++c) bg[r][c] = m[r][c];
  for (int r = 0; r < 5; ++r) strip_row(E, G, brn, gx, gy, r >= 3 ? bs[r - 3][0] : 0.0, P1, P2, r, bg[3 + r]);
  return v3(G.a[2][0], G.a[2][1], G.a[2][2]);
}

// Host-factored homogeneous constitutive data: Dps = L diag(dps) L' (unit lower, entries
// L10, L20, L21), Dt (x5/6) = Lt diag(dts) Lt'.
struct HomogFactors {
  double L10, L20, L21, dps[3], Lt, dts[2];
  double sdps[3], sdts[2];  // square roots of the pivots
};
FS_HD void fold3(const HomogFactors& H, double (&m)[3][6]) {
  for (int c = 0; c < 6; ++c) {
    m[0][c] += H.L10 * m[1][c] + H.L20 * m[2][c];
    m[1][c] += H.L21 * m[2][c];
  }
}

// homogeneous shell: b <- L' b with the host-factored block-diagonal L (membrane | bending | shear)
FS_HD void fold_homogeneous(const HomogFactors& H, double (&b)[8][6]) {
  for (int c = 0; c < 6; ++c) {
    b[0][c] += H.L10 * b[1][c] + H.L20 * b[2][c];
    b[1][c] += H.L21 * b[2][c];
    b[3][c] += H.L10 * b[4][c] + H.L20 * b[5][c];
    b[4][c] += H.L21 * b[5][c];
    b[6][c] += H.Lt * b[7][c];
  }
}
// node_kavg_part for a homogeneous shell: wb, ws = bending and shear weights of the element
FS_HD double node_kavg_part_h(const HomogFactors& H, double wb, double ws, const double (&brn)[5][2], bool shear_only) {
  double kb = 0.0, ks0 = 0.0, ks1 = 0.0;
  for (int cl = 0; cl < 2; ++cl) {
    const double f3 = brn[0][cl] + H.L10 * brn[1][cl] + H.L20 * brn[2][cl];
    const double f4 = brn[1][cl] + H.L21 * brn[2][cl];
    const double f5 = brn[2][cl];
    const double f6 = brn[3][cl] + H.Lt * brn[4][cl];
    const double f7 = brn[4][cl];
    kb += H.dps[0] * f3 * f3 + H.dps[1] * f4 * f4 + H.dps[2] * f5 * f5;
    ks0 += f6 * f6;
    ks1 += f7 * f7;
  }
  const double k = ws * (H.dts[0] * ks0 + H.dts[1] * ks1);
  return shear_only ? k : k + wb * kb;
}

// b <- L' b (rows), so that K = sum_s d_s b_s (x) b_s.
FS_HD void fold_constit(const Constit& C, double (&b)[8][6]) {
  for (int c = 0; c < 6; ++c) {
    for (int s = 0; s < 6; ++s) {
      double v = b[s][c];
      for (int t = s + 1; t < 6; ++t) v += C.L6[t][s] * b[t][c];
      b[s][c] = v;
    }
    b[6][c] += C.L2 * b[7][c];
  }
}
FS_HD double constit_d(const Constit& C, int s) { return s < 6 ? C.d6[s] : C.d2[s - 6]; }
// this node's share of T3FF's kavg numerator: sum_s d_s (bt'_s[theta1]^2 + bt'_s[theta2]^2) over the
// folded nodal-basis rotation columns (membrane rows are zero) -- src/FEMMShellT3FFModule.jl:714-722.
// shear_only: count only the two shear rows (extra AVERAGE_K orderings)
FS_HD double node_kavg_part(const Constit& C, const double (&brn)[5][2], bool shear_only) {
  double f[8][2];
  for (int cl = 0; cl < 2; ++cl) {
    f[0][cl] = f[1][cl] = f[2][cl] = 0.0;
    for (int r = 0; r < 5; ++r) f[3 + r][cl] = brn[r][cl];
    for (int s = 0; s < 6; ++s) {
      double v = f[s][cl];
      for (int t = s + 1; t < 6; ++t) v += C.L6[t][s] * f[t][cl];
      f[s][cl] = v;
    }
    f[6][cl] += C.L2 * f[7][cl];
  }
  double k = 0.0;
  for (int s = shear_only ? 6 : 0; s < 8; ++s) k += constit_d(C, s) * (f[s][0] * f[s][0] + f[s][1] * f[s][1]);
  return k;
}

// ---------------------------------------------------------------------------------
// T3FF geometry + DSG shear B (src/FEMMShellT3FFModule.jl:269-314,465-537)
// ---------------------------------------------------------------------------------
struct T3Geom {
  Triad E;
  double x1, y1, x2, y2;  // local coordinates of nodes 2, 3 (node 1 at the origin)
  double gN[3][2];
  double Ae;
};
FS_HD T3Geom t3_geometry(V3 X0, V3 X1, V3 X2) {
  T3Geom g;
  const V3 t1 = X1 - X0, t2 = X2 - X0;
  g.E = element_triad(t1, t2);
  g.x1 = dot(t1, g.E.e1);
  g.y1 = dot(t1, g.E.e2);
  g.x2 = dot(t2, g.E.e1);
  g.y2 = dot(t2, g.E.e2);
  const double a = g.x1, b = g.y1, c = g.x2, d = g.y2;
  const double J = a * d - b * c;
  const double iJ = fs_rcp(J);
  g.gN[0][0] = (b - d) * iJ;
  g.gN[1][0] = d * iJ;
  g.gN[2][0] = -b * iJ;
  g.gN[0][1] = (c - a) * iJ;
  g.gN[1][1] = -c * iJ;
  g.gN[2][1] = a * iJ;
  g.Ae = J / 2;
  return g;
}
// one DSG ordering (s,p,q), ADDED into bs[2][3][3] (cols: w, theta_x, theta_y)
FS_HD void t3_add_bs(const T3Geom& g, int s, int p, int q, double (&bs)[2][3][3]) {
  const double ex[3] = {0.0, g.x1, g.x2}, ey[3] = {0.0, g.y1, g.y2};
  const double a = ex[p] - ex[s], b = ey[p] - ey[s], c = ex[q] - ex[s], d = ey[q] - ey[s];
  const double Ae = g.Ae, m = fs_rcp(2 * Ae);
  bs[0][s][0] += m * (b - d);
  bs[0][s][2] += m * Ae;
  bs[1][s][0] += m * (c - a);
  bs[1][s][1] += m * (-Ae);
  bs[0][p][0] += m * d;
  bs[0][p][1] += m * (-b * d / 2);
  bs[0][p][2] += m * (a * d / 2);
  bs[1][p][0] += m * (-c);
  bs[1][p][1] += m * (b * c / 2);
  bs[1][p][2] += m * (-a * c / 2);
  bs[0][q][0] += m * (-b);
  bs[0][q][1] += m * (b * d / 2);
  bs[0][q][2] += m * (-b * c / 2);
  bs[1][q][0] += m * a;
  bs[1][q][1] += m * (-a * d / 2);
  bs[1][q][2] += m * (a * c / 2);
}

// DSG shear entries of ONE node l (averaged over the three cyclic orderings, or one ordering
// `only` = 0,1,2 for the AVERAGE_K formulation): out[2][3] (cols w, theta_x, theta_y)
FS_HD void t3_bs_node(const T3Geom& g, int l, int only, double (&out)[2][3]) {
  double bs[2][3][3];
  for (int r = 0; r < 2; ++r)
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) bs[r][a][c] = 0.0;
  if (only >= 0) {
    t3_add_bs(g, only, (only + 1) % 3, (only + 2) % 3, bs);
  } else {
    t3_add_bs(g, 0, 1, 2, bs);
    t3_add_bs(g, 1, 2, 0, bs);
    t3_add_bs(g, 2, 0, 1, bs);
    for (int r = 0; r < 2; ++r)
      for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 3; ++c) bs[r][a][c] *= (1.0 / 3);
  }
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 3; ++c) out[r][c] = l == 0 ? bs[r][0][c] : (l == 1 ? bs[r][1][c] : bs[r][2][c]);
}

// ---------------------------------------------------------------------------------
// Q4RS per-integration-point geometry (src/FEMMShellQ4RSModule.jl:287-323,580-608,
// 753-804, 861-870)
// ---------------------------------------------------------------------------------
struct Q4Geom {
  Triad E;
  double Jac;
  double gN[4][2];
  double ex[4], ey[4];  // centroid-relative coordinates projected on (e1, e2)
  int singular;
};
FS_HD void q4_shape_derivs(double xi, double eta, double (&dN)[4][2]) {
  dN[0][0] = -0.25 * (1 - eta);
  dN[0][1] = -0.25 * (1 - xi);
  dN[1][0] = 0.25 * (1 - eta);
  dN[1][1] = -0.25 * (1 + xi);
  dN[2][0] = 0.25 * (1 + eta);
  dN[2][1] = 0.25 * (1 + xi);
  dN[3][0] = -0.25 * (1 + eta);
  dN[3][1] = 0.25 * (1 - xi);
}
FS_HD Q4Geom q4_geometry(const V3 (&X)[4], double xi, double eta) {
  Q4Geom g;
  double dN[4][2];
  q4_shape_derivs(xi, eta, dN);
  V3 t1 = v3(0, 0, 0), t2 = v3(0, 0, 0), cen = v3(0, 0, 0);
  for (int a = 0; a < 4; ++a) {
    t1 = t1 + dN[a][0] * X[a];
    t2 = t2 + dN[a][1] * X[a];
    cen = cen + X[a];
  }
  cen = 0.25 * cen;
  {
    // element triad (as element_triad) and the surface Jacobian |t1 x t2| = |t1| |e1 x t2| from the same two
    // reciprocal square roots
    const double l1 = dot(t1, t1), i1 = fs_rsqrt(l1);
    g.E.e1 = i1 * t1;
    const V3 n = cross(g.E.e1, t2);
    const double ln = dot(n, n), in = fs_rsqrt(ln);
    g.E.e3 = in * n;
    g.E.e2 = cross(g.E.e3, g.E.e1);
    g.Jac = (l1 * i1) * (ln * in);
  }
  for (int a = 0; a < 4; ++a) {
    const V3 d = X[a] - cen;
    g.ex[a] = dot(d, g.E.e1);
    g.ey[a] = dot(d, g.E.e2);
  }
  // gradN_e = E2' J (J'J)^-1 gradNparams
  const double G11 = dot(t1, t1), G12 = dot(t1, t2), G22 = dot(t2, t2);
  const double det = G11 * G22 - G12 * G12;
  // Julia `detG ≈ 0.0` is isapprox with atol 0: true only for an exact zero
  g.singular = (det == 0.0) || !(det == det);
  const double idet = fs_rcp(det);
  const double i11 = G22 * idet, i12 = -G12 * idet, i22 = G11 * idet;
  const double j1e1 = dot(t1, g.E.e1), j1e2 = dot(t1, g.E.e2), j2e1 = dot(t2, g.E.e1), j2e2 = dot(t2, g.E.e2);
  for (int a = 0; a < 4; ++a) {
    const double p = i11 * dN[a][0] + i12 * dN[a][1];
    const double q = i12 * dN[a][0] + i22 * dN[a][1];
    g.gN[a][0] = j1e1 * p + j2e1 * q;
    g.gN[a][1] = j1e2 * p + j2e2 * q;
  }
  return g;
}
// MITC4 tying (Bathe-Dvorkin) shear B in factored form.  With the reference's symbols
// (src/FEMMShellQ4RSModule.jl:753-777): edge functional
//   e_ab(W,Tx,Ty) = (Wa-Wb)/2 + (Xa-Xb)/4 (Tya+Tyb) - (Ya-Yb)/4 (Txa+Txb),
//   g_rz = SC [(1+s) e_12 + (1-s) e_43],  g_sz = SA [(1+r) e_14 + (1-r) e_23],
//   g_xz = -(g_rz sb - g_sz sa),  g_yz = -(-g_rz cb + g_sz ca).
FS_HD void q4_mitc_bs(const Q4Geom& g, double r, double s, double (&bs)[2][4][3]) {
  const double* X = g.ex;
  const double* Y = g.ey;
  const double J11 = (X[0] * (s - 1) - X[1] * (s - 1) + X[2] * (s + 1) - X[3] * (s + 1)) / 4;
  const double J21 = (Y[0] * (s - 1) - Y[1] * (s - 1) + Y[2] * (s + 1) - Y[3] * (s + 1)) / 4;
  const double J12 = (X[0] * (r - 1) - X[1] * (r + 1) + X[2] * (r + 1) - X[3] * (r - 1)) / 4;
  const double J22 = (Y[0] * (r - 1) - Y[1] * (r + 1) + Y[2] * (r + 1) - Y[3] * (r - 1)) / 4;
  const double Aa = sqrt(J11 * J11 + J21 * J21), Bb = sqrt(J12 * J12 + J22 * J22);
  const double ca = J11 / Aa, sa = J21 / Aa, cb = J12 / Bb, sb = J22 / Bb;
  const double detJ = J11 * J22 - J12 * J21;
  const double Ax = X[0] - X[1] - X[2] + X[3], Ay = Y[0] - Y[1] - Y[2] + Y[3];
  const double Bx = X[0] - X[1] + X[2] - X[3], By = Y[0] - Y[1] + Y[2] - Y[3];
  const double Cx = X[0] + X[1] - X[2] - X[3], Cy = Y[0] + Y[1] - Y[2] - Y[3];
  const double SC = sqrt((Cx + r * Bx) * (Cx + r * Bx) + (Cy + r * By) * (Cy + r * By)) / (8 * detJ);
  const double SA = sqrt((Ax + s * Bx) * (Ax + s * Bx) + (Ay + s * By) * (Ay + s * By)) / (8 * detJ);
  // coefficient of g_rz and g_sz per node and dof (w, tx, ty)
  double crz[4][3], csz[4][3];
  for (int a = 0; a < 4; ++a)
    for (int c = 0; c < 3; ++c) crz[a][c] = csz[a][c] = 0.0;
  // edge (a,b) with weight wgt added into coefficient table `t`
#define FS_EDGE(t, a, b, wgt)                     \
  {                                               \
    const double w_ = (wgt);                      \
    const double dx_ = (X[a] - X[b]) / 4 * w_;    \
    const double dy_ = (Y[a] - Y[b]) / 4 * w_;    \
    t[a][0] += w_ / 2;                            \
    t[b][0] -= w_ / 2;                            \
    t[a][2] += dx_;                               \
    t[b][2] += dx_;                               \
    t[a][1] -= dy_;                               \
    t[b][1] -= dy_;                               \
  }
  FS_EDGE(crz, 0, 1, SC * (1 + s));
  FS_EDGE(crz, 3, 2, SC * (1 - s));
  FS_EDGE(csz, 0, 3, SA * (1 + r));
  FS_EDGE(csz, 1, 2, SA * (1 - r));
#undef FS_EDGE
  for (int a = 0; a < 4; ++a)
    for (int c = 0; c < 3; ++c) {
      bs[0][a][c] = -(crz[a][c] * sb - csz[a][c] * sa);
      bs[1][a][c] = -(-crz[a][c] * cb + csz[a][c] * ca);
    }
}

// MITC shear entries of ONE node a: out[2][3].  Same tying functionals as q4_mitc_bs, restricted
// to the two edges that meet at node a (r-edges (0,1),(3,2); s-edges (0,3),(1,2)).
FS_HD void q4_mitc_bs_node(const Q4Geom& g, double r, double s, int a, double (&out)[2][3]) {
  const double* X = g.ex;
  const double* Y = g.ey;
  const double J11 = (X[0] * (s - 1) - X[1] * (s - 1) + X[2] * (s + 1) - X[3] * (s + 1)) / 4;
  const double J21 = (Y[0] * (s - 1) - Y[1] * (s - 1) + Y[2] * (s + 1) - Y[3] * (s + 1)) / 4;
  const double J12 = (X[0] * (r - 1) - X[1] * (r + 1) + X[2] * (r + 1) - X[3] * (r - 1)) / 4;
  const double J22 = (Y[0] * (r - 1) - Y[1] * (r + 1) + Y[2] * (r + 1) - Y[3] * (r - 1)) / 4;
  const double iA = fs_rsqrt(J11 * J11 + J21 * J21), iB = fs_rsqrt(J12 * J12 + J22 * J22);
  const double ca = J11 * iA, sa = J21 * iA, cb = J12 * iB, sb = J22 * iB;
  const double i8d = fs_rcp(8 * (J11 * J22 - J12 * J21));
  const double Ax = X[0] - X[1] - X[2] + X[3], Ay = Y[0] - Y[1] - Y[2] + Y[3];
  const double Bx = X[0] - X[1] + X[2] - X[3], By = Y[0] - Y[1] + Y[2] - Y[3];
  const double Cx = X[0] + X[1] - X[2] - X[3], Cy = Y[0] + Y[1] - Y[2] - Y[3];
  const double SC = fs_sqrt((Cx + r * Bx) * (Cx + r * Bx) + (Cy + r * By) * (Cy + r * By)) * i8d;
  const double SA = fs_sqrt((Ax + s * Bx) * (Ax + s * Bx) + (Ay + s * By) * (Ay + s * By)) * i8d;
  // own coordinates and the partners across the r-edge / s-edge
  const bool lo = a < 2, out03 = (a == 0) || (a == 3);
  const double xa = a == 0 ? X[0] : (a == 1 ? X[1] : (a == 2 ? X[2] : X[3]));
  const double ya = a == 0 ? Y[0] : (a == 1 ? Y[1] : (a == 2 ? Y[2] : Y[3]));
  const double xr = a == 0 ? X[1] : (a == 1 ? X[0] : (a == 2 ? X[3] : X[2]));
  const double yr = a == 0 ? Y[1] : (a == 1 ? Y[0] : (a == 2 ? Y[3] : Y[2]));
  const double xs = a == 0 ? X[3] : (a == 1 ? X[2] : (a == 2 ? X[1] : X[0]));
  const double ys = a == 0 ? Y[3] : (a == 1 ? Y[2] : (a == 2 ? Y[1] : Y[0]));
  const double wr = SC * (lo ? (1 + s) : (1 - s)), sr = out03 ? 1.0 : -1.0;   // first node of (0,1) is 0, of (3,2) is 3
  const double ws = SA * (out03 ? (1 + r) : (1 - r)), ss = lo ? 1.0 : -1.0;    // first node of (0,3) is 0, of (1,2) is 1
  const double crz[3] = {sr * wr / 2, -(sr * (ya - yr)) / 4 * wr, (sr * (xa - xr)) / 4 * wr};
  const double csz[3] = {ss * ws / 2, -(ss * (ya - ys)) / 4 * ws, (ss * (xa - xs)) / 4 * ws};
  for (int c = 0; c < 3; ++c) {
    out[0][c] = -(crz[c] * sb - csz[c] * sa);
    out[1][c] = -(-crz[c] * cb + csz[c] * ca);
  }
}

// ---------------------------------------------------------------------------------
// Corotational beam (src/FEMMCorotBeamModule.jl:155-243, 672-801; FESetL2BeamModule.jl:108-127)
// ---------------------------------------------------------------------------------
struct BeamSec {
  double A, I1, I2, I3, J, A2s, A3s;
  V3 x1x2;
};
struct BeamKin {
  double L0, L1;
  Triad Ft;  // columns e1, e2, e3 of the current element frame
  double dN[6];
};
FS_HD Triad beam_frame(V3 chord, V3 x1x2, double& L) {
  Triad F;
  L = norm(chord);
  F.e1 = (1.0 / L) * chord;
  V3 n = cross(F.e1, x1x2);
  F.e3 = (1.0 / norm(n)) * n;
  F.e2 = cross(F.e3, F.e1);
  return F;
}
// R stored as the reference stores a row of Rfield: column-major 3x3 (R[c*3+r]).
FS_HD V3 rot_apply(const double R[9], V3 v) {
  return v3(R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z, R[2] * v.x + R[5] * v.y + R[8] * v.z);
}
FS_HD BeamKin beam_kinematics(V3 x0I, V3 x0J, V3 uI, V3 uJ, const double RI[9], const double RJ[9], V3 x1x2) {
  BeamKin k;
  const Triad F0 = beam_frame(x0J - x0I, x1x2, k.L0);
  // nodal cross-section frames FtI = RI F0, FtJ = RJ F0 (columns)
  const V3 I1 = rot_apply(RI, F0.e1), I2 = rot_apply(RI, F0.e2), I3 = rot_apply(RI, F0.e3);
  const V3 J1 = rot_apply(RJ, F0.e1), J2 = rot_apply(RJ, F0.e2), J3 = rot_apply(RJ, F0.e3);
  k.Ft = beam_frame((x0J + uJ) - (x0I + uI), I2 + J2, k.L1);
  const Triad& F = k.Ft;
  // L = Ft' * FtX : L[r][c] = e_r . X_c
  const double LI11 = dot(F.e1, I1), LI21 = dot(F.e2, I1), LI31 = dot(F.e3, I1);
  const double LI22 = dot(F.e2, I2), LI32 = dot(F.e3, I2), LI23 = dot(F.e2, I3), LI33 = dot(F.e3, I3);
  const double LJ11 = dot(F.e1, J1), LJ21 = dot(F.e2, J1), LJ31 = dot(F.e3, J1);
  const double LJ22 = dot(F.e2, J2), LJ32 = dot(F.e3, J2), LJ23 = dot(F.e2, J3), LJ33 = dot(F.e3, J3);
  k.dN[0] = k.L1 - k.L0;
  k.dN[5] = (LJ32 / LJ22 - LI32 / LI22 - LJ23 / LJ33 + LI23 / LI33) / 2;
  const double TH2I = -LI31 / LI11, TH2J = -LJ31 / LJ11, TH3I = LI21 / LI11, TH3J = LJ21 / LJ11;
  k.dN[1] = TH3I - TH3J;
  k.dN[2] = TH3I + TH3J;
  k.dN[3] = -TH2I + TH2J;
  k.dN[4] = -TH2I - TH2J;
  return k;
}
// Natural stiffness diagonal; Bernoulli iff A2s == Inf || A3s == Inf (App. B.10).
FS_HD void beam_natural_stiffness(double E, double G, const BeamSec& s, double L, double (&DN)[6]) {
  DN[0] = E * s.A / L;
  DN[1] = E * s.I3 / L;
  DN[3] = E * s.I2 / L;
  DN[5] = G * s.J / L;
  if (isinf(s.A2s) || isinf(s.A3s)) {
    DN[2] = 3 * E * s.I3 / L;
    DN[4] = 3 * E * s.I2 / L;
  } else {
    const double Phi3 = 12 * E * s.I3 / (G * s.A2s * L * L);
    const double Phi2 = 12 * E * s.I2 / (G * s.A3s * L * L);
    DN[2] = 3 * E * s.I3 / L / (1 + Phi3);
    DN[4] = 3 * E * s.I2 / L / (1 + Phi2);
  }
}
// aN (6x12), Argyris natural-mode matrix (src/FEMMCorotBeamModule.jl:265-290)
FS_HD void beam_aN(double L, double (&aN)[6][12]) {
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 12; ++j) aN[i][j] = 0.0;
  const double q = 2 / L;
  aN[0][0] = -1;
  aN[0][6] = 1;
  aN[1][5] = 1;
  aN[1][11] = -1;
  aN[2][1] = q;
  aN[2][5] = 1;
  aN[2][7] = -q;
  aN[2][11] = 1;
  aN[3][4] = -1;
  aN[3][10] = 1;
  aN[4][2] = q;
  aN[4][4] = -1;
  aN[4][8] = -q;
  aN[4][10] = -1;
  aN[5][3] = -1;
  aN[5][9] = 1;
}
// local 12x12 (upper triangle filled, then mirrored) geometric stiffness of Krenk
// (src/FEMMCorotBeamModule.jl:321-382)
FS_HD void beam_local_geo(const double (&PN)[6], double L, double (&S)[12][12]) {
  for (int i = 0; i < 12; ++i)
    for (int j = 0; j < 12; ++j) S[i][j] = 0.0;
  const double N = PN[0], S2 = -2 * PN[2] / L, S3 = -2 * PN[4] / L, M1 = PN[5];
  const double M2I = PN[3] + PN[4], M2J = PN[3] - PN[4], M3I = -(PN[1] + PN[2]), M3J = -(PN[1] - PN[2]);
  // [1:3,1:3] and [7:9,7:9]
  for (int o = 0; o <= 6; o += 6) {
    S[o + 0][o + 1] = -S2 / L;
    S[o + 0][o + 2] = -S3 / L;
    S[o + 1][o + 1] = N / L;
    S[o + 2][o + 2] = N / L;
  }
  // [1:3,4:6]
  S[1][3] = -M2I / L;
  S[1][4] = M1 / L;
  S[2][3] = -M3I / L;
  S[2][5] = M1 / L;
  // [1:3,7:9]
  S[0][7] = S2 / L;
  S[0][8] = S3 / L;
  S[1][6] = S2 / L;
  S[1][7] = -N / L;
  S[2][6] = S3 / L;
  S[2][8] = -N / L;
  // [1:3,10:12]
  S[1][9] = M2J / L;
  S[1][10] = -M1 / L;
  S[2][9] = M3J / L;
  S[2][11] = -M1 / L;
  // [4:6,4:6]
  S[3][4] = M3I / 2;
  S[3][5] = -M2I / 2;
  // [4:6,7:9]
  S[3][7] = M2I / L;
  S[3][8] = M3I / L;
  S[4][7] = -M1 / L;
  S[5][8] = -M1 / L;
  // [4:6,10:12]
  S[4][11] = M1 / 2;
  S[5][10] = -M1 / 2;
  // [7:9,10:12]
  S[7][9] = -M2J / L;
  S[7][10] = M1 / L;
  S[8][9] = -M3J / L;
  S[8][11] = M1 / L;
  // [10:12,10:12]
  S[9][10] = -M3J / 2;
  S[9][11] = M2J / 2;
  for (int i = 0; i < 12; ++i)
    for (int j = 0; j < i; ++j) S[i][j] = S[j][i];
}

// local 12x12 mass matrix, 4 formulations (src/FEMMCorotBeamModule.jl:384-554)
FS_HD void beam_local_mass(const BeamSec& s, double rho, double L, int mass_type, double (&M)[12][12]) {
  for (int i = 0; i < 12; ++i)
    for (int j = 0; j < 12; ++j) M[i][j] = 0.0;
  const double A = s.A, I1 = s.I1, I2 = s.I2, I3 = s.I3;
  if (mass_type == 0 || mass_type == 1) {
    const double c1 = rho * A * L;
    M[0][0] = c1 * (1.0 / 3);
    M[0][6] = c1 * (1.0 / 6);
    M[1][1] = c1 * (13.0 / 35);
    M[1][5] = c1 * (11 * L / 210);
    M[1][7] = c1 * (9.0 / 70);
    M[1][11] = c1 * (-13 * L / 420);
    M[2][2] = c1 * (13.0 / 35);
    M[2][4] = c1 * (-11 * L / 210);
    M[2][8] = c1 * (9.0 / 70);
    M[2][10] = c1 * (13 * L / 420);
    M[3][3] = c1 * (I1 / 3 / A);
    M[3][9] = c1 * (I1 / 6 / A);
    M[4][4] = c1 * (L * L / 105);
    M[4][8] = c1 * (-13 * L / 420);
    M[4][10] = c1 * (-(L * L) / 140);
    M[5][5] = c1 * (L * L / 105);
    M[5][7] = c1 * (13 * L / 420);
    M[5][11] = c1 * (-(L * L) / 140);
    M[6][6] = c1 * (1.0 / 3);
    M[7][7] = c1 * (13.0 / 35);
    M[7][11] = c1 * (-11 * L / 210);
    M[8][8] = c1 * (13.0 / 35);
    M[8][10] = c1 * (11 * L / 210);
    M[9][9] = c1 * (I1 / 3 / A);
    M[10][10] = c1 * (L * L / 105);
    M[11][11] = c1 * (L * L / 105);
    if (mass_type == 1) {
      const double c2 = rho / L;
      M[1][1] += c2 * (6.0 / 5 * I2);
      M[1][5] += c2 * (L / 10 * I2);
      M[1][7] += c2 * (-6.0 / 5 * I2);
      M[1][11] += c2 * (L / 10 * I2);
      M[2][2] += c2 * (6.0 / 5 * I3);
      M[2][4] += c2 * (-L / 10 * I3);
      M[2][8] += c2 * (-6.0 / 5 * I3);
      M[2][10] += c2 * (-L / 10 * I3);
      M[4][4] += c2 * (2 * L * L / 15 * I3);
      M[4][8] += c2 * (L / 10 * I3);
      M[4][10] += c2 * (-(L * L) / 30 * I3);
      M[5][5] += c2 * (2 * L * L / 15 * I2);
      M[5][7] += c2 * (-L / 10 * I2);
      M[5][11] += c2 * (-(L * L) / 30 * I2);
      M[7][7] += c2 * (6.0 / 5 * I2);
      M[7][11] += c2 * (-L / 10 * I2);
      M[8][8] += c2 * (6.0 / 5 * I3);
      M[8][10] += c2 * (L / 10 * I3);
      M[10][10] += c2 * (2 * L * L / 15 * I3);
      M[11][11] += c2 * (2 * L * L / 15 * I2);
    }
    for (int i = 0; i < 12; ++i)
      for (int j = 0; j < i; ++j) M[i][j] = M[j][i];
  } else {
    const double CA = A * rho * L / 2.0;
    double d[6] = {CA, CA, CA, 0.0, 0.0, 0.0};
    if (mass_type == 3) {
      d[3] = rho * I1 * L / 2.0;
      d[4] = rho * I2 * L / 2.0;
      d[5] = rho * I3 * L / 2.0;
    }
    for (int k = 0; k < 6; ++k) {
      M[k][k] = d[k];
      M[k + 6][k + 6] = d[k];
    }
  }
}

}  // namespace fsm
