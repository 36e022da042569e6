/* refport.c -- C restatement of the REFERENCE ALGORITHM AS WRITTEN (dense transformation
 * matrices, dense Q'EQ products, COO append, counting-sort COO->CSC, CSR SpMV), used as
 * the CPU baseline ("port") that bench.py times on the host cores next to the GPU path.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py): never linked into
 * libfsgpu.so, never on the product path.
 *
 * Follows (reference v3.6.4, paths relative to /root/reference):
 *   T3FF  stiffness  src/FEMMShellT3FFModule.jl:635-736 (+ helpers :269-561)
 *   Q4RS  stiffness  src/FEMMShellQ4RSModule.jl:877-947 (+ helpers :248-870)
 *   assembler        FinEtools SysmatAssemblerSparse + Julia `sparse`
 *   explicit loop    examples/shells/dynamics/homogeneous/explicit/plate_expl_examples.jl:61-94
 * The reference's element loops are serial; `nthreads` > 1 is the "generous to the CPU"
 * variant (element-parallel OpenMP) described in SURVEY section 8(d).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 24

static void zero(double* a, int n) { memset(a, 0, sizeof(double) * n); }

/* C = A(m x k) * B(k x n), row-major, leading dims lda/ldb/ldc */
static void gemm(int m, int n, int k, const double* restrict A, int lda, const double* restrict B, int ldb,
                 double* restrict C, int ldc) {
  for (int i = 0; i < m; ++i) {
    double* restrict c = C + i * ldc;
    for (int j = 0; j < n; ++j) c[j] = 0;
    for (int p = 0; p < k; ++p) {
      const double a = A[i * lda + p];
      const double* restrict b = B + p * ldb;
      for (int j = 0; j < n; ++j) c[j] += a * b[j];
    }
  }
}
/* E <- Q' (E Q)   (TransformerQtEQ, src/TransformerModule.jl:34-42) */
static void qteq(int n, double* E, const double* Q) {
  double buf[MAXN * MAXN], out[MAXN * MAXN];
  gemm(n, n, n, E, n, Q, n, buf, n);
  for (int i = 0; i < n; ++i) {
    double* restrict o = out + i * n;
    for (int j = 0; j < n; ++j) o[j] = 0;
    for (int p = 0; p < n; ++p) {
      const double q = Q[p * n + i];
      const double* restrict b = buf + p * n;
      for (int j = 0; j < n; ++j) o[j] += q * b[j];
    }
  }
  memcpy(E, out, sizeof(double) * n * n);
}
/* upper triangle of Ke += c * B' D B  (add_btdb_ut_only!) ; B is nb x n, D nb x nb */
static void add_btdb_ut(int n, int nb, double* Ke, const double* B, double c, const double* D) {
  double DB[3 * MAXN];
  gemm(nb, n, nb, D, nb, B, n, DB, n);
  for (int i = 0; i < n; ++i)
    for (int p = 0; p < nb; ++p) {
      const double b = c * B[p * n + i];
      const double* restrict db = DB + p * n;
      double* restrict k = Ke + i * n;
      for (int j = i; j < n; ++j) k[j] += b * db[j];
    }
}
static void complete_lt(int n, double* Ke) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) Ke[i * n + j] = Ke[j * n + i];
}
static void e_g(const double* J /*3x2 row-major*/, double* E /*3x3 row-major, columns = triad*/) {
  double e1[3], e2[3], e3[3], nr;
  for (int k = 0; k < 3; ++k) e1[k] = J[k * 2];
  nr = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
  for (int k = 0; k < 3; ++k) e1[k] /= nr;
  e3[0] = -e1[2] * J[1 * 2 + 1] + e1[1] * J[2 * 2 + 1];
  e3[1] = e1[2] * J[0 * 2 + 1] - e1[0] * J[2 * 2 + 1];
  e3[2] = -e1[1] * J[0 * 2 + 1] + e1[0] * J[1 * 2 + 1];
  nr = sqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]);
  for (int k = 0; k < 3; ++k) e3[k] /= nr;
  e2[0] = -e3[2] * e1[1] + e3[1] * e1[2];
  e2[1] = e3[2] * e1[0] - e3[0] * e1[2];
  e2[2] = -e3[1] * e1[0] + e3[0] * e1[1];
  for (int k = 0; k < 3; ++k) {
    E[k * 3 + 0] = e1[k];
    E[k * 3 + 1] = e2[k];
    E[k * 3 + 2] = e3[k];
  }
}
static void rotmat3(const double* a, double* R) {
  double na = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  double n[3] = {a[0] / na, a[1] / na, a[2] / na}, c = cos(na), s = sin(na);
  double K[9] = {0, -n[2], n[1], n[2], 0, -n[0], -n[1], n[0], 0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = c * ((i == j) - n[i] * n[j]) + s * K[i * 3 + j] + n[i] * n[j];
}
/* nodal triads A_Es[k] (3x3 row-major) (src/FEMMShellT3FFModule.jl:355-388) */
static void nodal_triads(int nn, const double* E, const double* normals, const uint8_t* valid, const int64_t* c,
                         int64_t nnodes, double* A, int* nvalid) {
  for (int k = 0; k < nn; ++k) {
    int64_t nd = c[k] - 1;
    double nk[3] = {normals[nd], normals[nnodes + nd], normals[2 * nnodes + nd]}, ne[3];
    nvalid[k] = valid[nd] != 0;
    if (nvalid[k]) {
      for (int i = 0; i < 3; ++i) ne[i] = E[0 * 3 + i] * nk[0] + E[1 * 3 + i] * nk[1] + E[2 * 3 + i] * nk[2];
    } else {
      ne[0] = ne[1] = 0;
      ne[2] = 1;
    }
    double r[3] = {-ne[1], ne[0], 0};
    if (sqrt(r[0] * r[0] + r[1] * r[1]) > 1.0e-12) {
      rotmat3(r, A + 9 * k);
    } else {
      zero(A + 9 * k, 9);
      A[9 * k] = A[9 * k + 4] = A[9 * k + 8] = 1;
    }
  }
}
static void transf_g_to_a(int nn, const double* A, const double* E, double* T) {
  int n = 6 * nn;
  zero(T, n * n);
  for (int k = 0; k < nn; ++k) {
    double blk[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int p = 0; p < 3; ++p) s += A[9 * k + p * 3 + i] * E[j * 3 + p];
        blk[i * 3 + j] = s;
      }
    for (int h = 0; h < 2; ++h)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[(6 * k + 3 * h + i) * n + 6 * k + 3 * h + j] = blk[i * 3 + j];
  }
}
static void transf_a_to_e(int nn, const double* A, const double* gN /*nn x 2*/, double* T) {
  int n = 6 * nn;
  zero(T, n * n);
  for (int i = 0; i < nn; ++i) {
    const double* Ai = A + 9 * i;
    int ro = 6 * i;
    double a33 = Ai[8];
    for (int cl = 0; cl < 3; ++cl)
      for (int rw = 0; rw < 3; ++rw) T[(ro + rw) * n + ro + cl] = Ai[rw * 3 + cl];
    for (int cl = 0; cl < 2; ++cl)
      for (int rw = 0; rw < 2; ++rw) T[(ro + 3 + rw) * n + ro + 3 + cl] = Ai[rw * 3 + cl] - (1 / a33) * Ai[rw * 3 + 2] * Ai[cl * 3 + 2];
    double m1 = (1 / a33) * Ai[2], m2 = (1 / a33) * Ai[5];
    for (int j = 0; j < nn; ++j)
      for (int k = 0; k < 3; ++k) {
        double a3 = 0.5 * (Ai[3 + k] * gN[j * 2] - Ai[k] * gN[j * 2 + 1]);
        T[(ro + 3) * n + 6 * j + k] += m1 * a3;
        T[(ro + 4) * n + 6 * j + k] += m2 * a3;
      }
  }
}
static void bm_mat(int nn, const double* gN, double* B) {
  int n = 6 * nn;
  zero(B, 3 * n);
  for (int i = 0; i < nn; ++i) {
    B[0 * n + 6 * i] = gN[i * 2];
    B[1 * n + 6 * i + 1] = gN[i * 2 + 1];
    B[2 * n + 6 * i] = gN[i * 2 + 1];
    B[2 * n + 6 * i + 1] = gN[i * 2];
  }
}
static void bb_mat(int nn, const double* gN, double* B) {
  int n = 6 * nn;
  zero(B, 3 * n);
  for (int i = 0; i < nn; ++i) {
    B[0 * n + 6 * i + 4] = gN[i * 2];
    B[1 * n + 6 * i + 3] = -gN[i * 2 + 1];
    B[2 * n + 6 * i + 3] = -gN[i * 2];
    B[2 * n + 6 * i + 4] = gN[i * 2 + 1];
  }
}
static void t3_add_bs(double* Bs, const double* ec, double Ae, int s, int p, int q) {
  double a = ec[p * 2] - ec[s * 2], b = ec[p * 2 + 1] - ec[s * 2 + 1], c = ec[q * 2] - ec[s * 2], d = ec[q * 2 + 1] - ec[s * 2 + 1];
  double m = 1.0 / 2 / Ae;
  double* r0 = Bs;
  double* r1 = Bs + 18;
  int co = s * 6;
  r0[co + 2] += m * (b - d);
  r0[co + 4] += m * Ae;
  r1[co + 2] += m * (c - a);
  r1[co + 3] += m * (-Ae);
  co = p * 6;
  r0[co + 2] += m * d;
  r0[co + 3] += m * (-b * d / 2);
  r0[co + 4] += m * (a * d / 2);
  r1[co + 2] += m * (-c);
  r1[co + 3] += m * (b * c / 2);
  r1[co + 4] += m * (-a * c / 2);
  co = q * 6;
  r0[co + 2] += m * (-b);
  r0[co + 3] += m * (b * d / 2);
  r0[co + 4] += m * (-b * c / 2);
  r1[co + 2] += m * a;
  r1[co + 3] += m * (-a * d / 2);
  r1[co + 4] += m * (a * c / 2);
}

/* One T3FF element stiffness (18x18 row-major) -- src/FEMMShellT3FFModule.jl:670-730 */
static void t3ff_element(const int64_t* c, const double* xyz, int64_t nnodes, const double* normals, const uint8_t* valid,
                         const double* Dps, const double* Dt56, double t, double alpha, double drill, double* Ke) {
  double X[9], J0[6], E[9], ec[6] = {0}, gN[6];
  for (int k = 0; k < 3; ++k)
    for (int d = 0; d < 3; ++d) X[k * 3 + d] = xyz[d * nnodes + c[k] - 1];
  for (int d = 0; d < 3; ++d) {
    J0[d * 2] = X[3 + d] - X[d];
    J0[d * 2 + 1] = X[6 + d] - X[d];
  }
  e_g(J0, E);
  for (int q = 0; q < 2; ++q)
    for (int w = 0; w < 2; ++w) {
      double s = 0;
      for (int d = 0; d < 3; ++d) s += J0[d * 2 + q] * E[d * 3 + w];
      ec[(q + 1) * 2 + w] = s;
    }
  double a = ec[2], b = ec[3], cc = ec[4], d = ec[5], J = a * d - b * cc, Ae = J / 2;
  gN[0] = (b - d) / J;
  gN[2] = d / J;
  gN[4] = -b / J;
  gN[1] = (cc - a) / J;
  gN[3] = -cc / J;
  gN[5] = a / J;
  double B[3 * 18], Bs[2 * 18];
  zero(Ke, 324);
  bm_mat(3, gN, B);
  add_btdb_ut(18, 3, Ke, B, t * Ae, Dps);
  bb_mat(3, gN, B);
  add_btdb_ut(18, 3, Ke, B, t * t * t / 12 * Ae, Dps);
  double h = sqrt(2 * Ae), stab = t * t / (t * t + alpha * h * h);
  zero(Bs, 36);
  t3_add_bs(Bs, ec, Ae, 0, 1, 2);
  t3_add_bs(Bs, ec, Ae, 1, 2, 0);
  t3_add_bs(Bs, ec, Ae, 2, 0, 1);
  for (int i = 0; i < 36; ++i) Bs[i] *= (1.0 / 3);
  add_btdb_ut(18, 2, Ke, Bs, t * stab * Ae, Dt56);
  complete_lt(18, Ke);
  double A[27], T[324];
  int nv[3];
  nodal_triads(3, E, normals, valid, c, nnodes, A, nv);
  transf_a_to_e(3, A, gN, T);
  qteq(18, Ke, T);
  double kavg = (Ke[3 * 18 + 3] + Ke[9 * 18 + 9] + Ke[15 * 18 + 15] + Ke[4 * 18 + 4] + Ke[10 * 18 + 10] + Ke[16 * 18 + 16]) / 6 * drill;
  for (int k = 0; k < 3; ++k)
    if (nv[k]) Ke[(6 * k + 5) * 18 + 6 * k + 5] += kavg;
  transf_g_to_a(3, A, E, T);
  qteq(18, Ke, T);
}

static void q4_shape(double xi, double eta, double* N, double* dN) {
  N[0] = 0.25 * (1 - xi) * (1 - eta);
  N[1] = 0.25 * (1 + xi) * (1 - eta);
  N[2] = 0.25 * (1 + xi) * (1 + eta);
  N[3] = 0.25 * (1 - xi) * (1 + eta);
  dN[0] = -0.25 * (1 - eta);
  dN[1] = -0.25 * (1 - xi);
  dN[2] = 0.25 * (1 - eta);
  dN[3] = -0.25 * (1 + xi);
  dN[4] = 0.25 * (1 + eta);
  dN[5] = 0.25 * (1 + xi);
  dN[6] = -0.25 * (1 + eta);
  dN[7] = 0.25 * (1 - xi);
}
/* MITC4 shear B (2 x 24): the tying strains of src/FEMMShellQ4RSModule.jl:628-752 */
static void q4_bs(const double* ec, double r, double s, double* Bs) {
  const double X1 = ec[0], Y1 = ec[1], X2 = ec[2], Y2 = ec[3], X3 = ec[4], Y3 = ec[5], X4 = ec[6], Y4 = ec[7];
  double J11 = (X1 * (s - 1) - X2 * (s - 1) + X3 * (s + 1) - X4 * (s + 1)) / 4;
  double J21 = (Y1 * (s - 1) - Y2 * (s - 1) + Y3 * (s + 1) - Y4 * (s + 1)) / 4;
  double J12 = (X1 * (r - 1) - X2 * (r + 1) + X3 * (r + 1) - X4 * (r - 1)) / 4;
  double J22 = (Y1 * (r - 1) - Y2 * (r + 1) + Y3 * (r + 1) - Y4 * (r - 1)) / 4;
  double Aa = sqrt(J11 * J11 + J21 * J21), Bb = sqrt(J12 * J12 + J22 * J22);
  double ca = J11 / Aa, sa = J21 / Aa, cb = J12 / Bb, sb = J22 / Bb, detJ = J11 * J22 - J12 * J21;
  double Ax = X1 - X2 - X3 + X4, Ay = Y1 - Y2 - Y3 + Y4, Bx = X1 - X2 + X3 - X4, By = Y1 - Y2 + Y3 - Y4;
  double Cx = X1 + X2 - X3 - X4, Cy = Y1 + Y2 - Y3 - Y4;
  double SC = sqrt((Cx + r * Bx) * (Cx + r * Bx) + (Cy + r * By) * (Cy + r * By)) / 8 / detJ;
  double SA = sqrt((Ax + s * Bx) * (Ax + s * Bx) + (Ay + s * By) * (Ay + s * By)) / 8 / detJ;
  double crz[12] = {0}, csz[12] = {0}; /* [node][w,tx,ty] */
  const int ed[4][2] = {{0, 1}, {3, 2}, {0, 3}, {1, 2}};
  const double wg[4] = {SC * (1 + s), SC * (1 - s), SA * (1 + r), SA * (1 - r)};
  for (int k = 0; k < 4; ++k) {
    double* t = k < 2 ? crz : csz;
    int a = ed[k][0], b = ed[k][1];
    double dx = (ec[2 * a] - ec[2 * b]) / 4 * wg[k], dy = (ec[2 * a + 1] - ec[2 * b + 1]) / 4 * wg[k];
    t[3 * a] += wg[k] / 2;
    t[3 * b] -= wg[k] / 2;
    t[3 * a + 2] += dx;
    t[3 * b + 2] += dx;
    t[3 * a + 1] -= dy;
    t[3 * b + 1] -= dy;
  }
  zero(Bs, 48);
  for (int a = 0; a < 4; ++a)
    for (int cc = 0; cc < 3; ++cc) {
      Bs[6 * a + 2 + cc] = -(crz[3 * a + cc] * sb - csz[3 * a + cc] * sa);
      Bs[24 + 6 * a + 2 + cc] = -(-crz[3 * a + cc] * cb + csz[3 * a + cc] * ca);
    }
}
/* One Q4RS element stiffness (24x24 row-major) -- src/FEMMShellQ4RSModule.jl:914-941 */
static int q4rs_element(const int64_t* c, const double* xyz, int64_t nnodes, const double* normals, const uint8_t* valid,
                        const double* Dps, const double* Dt56, double t, double alpha, double drill, int npts,
                        const double* pc, const double* w, double* Ke) {
  double X[12];
  for (int k = 0; k < 4; ++k)
    for (int d = 0; d < 3; ++d) X[k * 3 + d] = xyz[d * nnodes + c[k] - 1];
  double md = 0;
  for (int k = 1; k < 4; ++k) {
    double s = 0;
    for (int d = 0; d < 3; ++d) s += (X[k * 3 + d] - X[d]) * (X[k * 3 + d] - X[d]);
    if (s > md) md = s;
  }
  double h = sqrt(md);
  zero(Ke, 576);
  for (int j = 0; j < npts; ++j) {
    double N[4], dN[8], J[6] = {0}, E[9], ec[8], gN[8], cen[3] = {0};
    q4_shape(pc[2 * j], pc[2 * j + 1], N, dN);
    for (int a = 0; a < 4; ++a)
      for (int d = 0; d < 3; ++d) {
        J[d * 2] += X[a * 3 + d] * dN[a * 2];
        J[d * 2 + 1] += X[a * 3 + d] * dN[a * 2 + 1];
        cen[d] += X[a * 3 + d] / 4;
      }
    double cr[3] = {J[2] * J[5] - J[4] * J[3], J[4] * J[1] - J[0] * J[5], J[0] * J[3] - J[2] * J[1]};
    double Jac = sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
    e_g(J, E);
    for (int a = 0; a < 4; ++a)
      for (int k = 0; k < 2; ++k) {
        double s = 0;
        for (int d = 0; d < 3; ++d) s += (X[a * 3 + d] - cen[d]) * E[d * 3 + k];
        ec[a * 2 + k] = s;
      }
    double G11 = 0, G12 = 0, G22 = 0;
    for (int d = 0; d < 3; ++d) {
      G11 += J[d * 2] * J[d * 2];
      G12 += J[d * 2] * J[d * 2 + 1];
      G22 += J[d * 2 + 1] * J[d * 2 + 1];
    }
    double det = G11 * G22 - G12 * G12;
    if (det == 0.0) return 1;
    double i11 = G22 / det, i12 = -G12 / det, i22 = G11 / det;
    for (int a = 0; a < 4; ++a) {
      double p = i11 * dN[a * 2] + i12 * dN[a * 2 + 1], q = i12 * dN[a * 2] + i22 * dN[a * 2 + 1], g3[3];
      for (int d = 0; d < 3; ++d) g3[d] = J[d * 2] * p + J[d * 2 + 1] * q;
      for (int k = 0; k < 2; ++k) gN[a * 2 + k] = E[0 * 3 + k] * g3[0] + E[1 * 3 + k] * g3[1] + E[2 * 3 + k] * g3[2];
    }
    double A[36], Tga[576], Tae[576], T[576], tB[3 * 24], B[3 * 24];
    int nv[4];
    nodal_triads(4, E, normals, valid, c, nnodes, A, nv);
    transf_g_to_a(4, A, E, Tga);
    transf_a_to_e(4, A, gN, Tae);
    gemm(24, 24, 24, Tae, 24, Tga, 24, T, 24);
    bm_mat(4, gN, tB);
    gemm(3, 24, 24, tB, 24, T, 24, B, 24);
    add_btdb_ut(24, 3, Ke, B, t * Jac * w[j], Dps);
    bb_mat(4, gN, tB);
    gemm(3, 24, 24, tB, 24, T, 24, B, 24);
    add_btdb_ut(24, 3, Ke, B, (t * t * t / 12.0) * Jac * w[j], Dps);
    q4_bs(ec, pc[2 * j], pc[2 * j + 1], tB);
    gemm(2, 24, 24, tB, 24, T, 24, B, 24);
    add_btdb_ut(24, 2, Ke, B, t * (t * t / (t * t + alpha * h * h)) * Jac * w[j], Dt56);
  }
  complete_lt(24, Ke);
  /* drilling (src/FEMMShellQ4RSModule.jl:807-859) */
  if (drill != 0.0) {
    double tsum = 0;
    int cnt = 0;
    for (int k = 0; k < 4; ++k) {
      int64_t nd = c[k] - 1;
      double n[3] = {normals[nd], normals[nnodes + nd], normals[2 * nnodes + nd]}, nl = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      if (!valid[nd] || nl == 0.0) continue;
      double P[9], KP[9], tr = 0;
      for (int i = 0; i < 3; ++i)
        for (int jj = 0; jj < 3; ++jj) P[i * 3 + jj] = (i == jj) - n[i] / nl * n[jj] / nl;
      for (int i = 0; i < 3; ++i)
        for (int jj = 0; jj < 3; ++jj) {
          double s = 0;
          for (int p = 0; p < 3; ++p) s += Ke[(6 * k + 3 + i) * 24 + 6 * k + 3 + p] * P[p * 3 + jj];
          KP[i * 3 + jj] = s;
        }
      for (int i = 0; i < 3; ++i)
        for (int p = 0; p < 3; ++p) tr += P[i * 3 + p] * KP[p * 3 + i];
      tsum += tr / 2 > 0 ? tr / 2 : 0;
      cnt++;
    }
    double kavg = cnt ? tsum / cnt * drill : 0.0;
    if (kavg != 0.0)
      for (int k = 0; k < 4; ++k) {
        int64_t nd = c[k] - 1;
        double n[3] = {normals[nd], normals[nnodes + nd], normals[2 * nnodes + nd]};
        if (!valid[nd] || (n[0] == 0 && n[1] == 0 && n[2] == 0)) continue;
        for (int i = 0; i < 3; ++i)
          for (int jj = 0; jj < 3; ++jj) Ke[(6 * k + 3 + i) * 24 + 6 * k + 3 + jj] += kavg * n[i] * n[jj];
      }
  }
  return 0;
}

/* ---- exported entry points ------------------------------------------------------- */
/* Element loop + SysmatAssemblerSparse.assemble! (j outer, i inner), writing COO triples.
 * kind 3 = T3FF, 4 = Q4RS.  conn: nnpe x nelem (1-based), xyz/normals nnodes x 3 col-major,
 * dofnums nnodes x 6 col-major.  I, J, V must hold nelem * n * n entries.  Returns 0 / 1. */
int ref_shell_stiffness_coo(int kind, int64_t nelem, const int64_t* conn, int64_t nnodes, const double* xyz,
                            const double* normals, const uint8_t* valid, const int64_t* dofnums, const double* Dps,
                            const double* Dt, double t, double alpha, double drill, int npts, const double* pc,
                            const double* w, int nthreads, int64_t* I, int64_t* J, double* V) {
  const int nn = kind, n = 6 * nn;
  double Dt56[4];
  for (int i = 0; i < 4; ++i) Dt56[i] = Dt[i] * (5.0 / 6.0);
  int bad = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int64_t e = 0; e < nelem; ++e) {
    double Ke[MAXN * MAXN];
    const int64_t* c = conn + e * nn;
    if (nn == 3)
      t3ff_element(c, xyz, nnodes, normals, valid, Dps, Dt56, t, alpha, drill, Ke);
    else
      bad |= q4rs_element(c, xyz, nnodes, normals, valid, Dps, Dt56, t, alpha, drill, npts, pc, w, Ke);
    int64_t dn[MAXN];
    for (int k = 0; k < nn; ++k)
      for (int d = 0; d < 6; ++d) dn[6 * k + d] = dofnums[d * nnodes + c[k] - 1];
    int64_t p = e * n * n;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i, ++p) {
        I[p] = dn[i];
        J[p] = dn[j];
        V[p] = Ke[i * n + j];
      }
  }
  return bad;
}
/* raw element matrices, n x n x nelem column-major (parity check of the port itself) */
int ref_shell_stiffness_elmats(int kind, int64_t nelem, const int64_t* conn, int64_t nnodes, const double* xyz,
                               const double* normals, const uint8_t* valid, const double* Dps, const double* Dt, double t,
                               double alpha, double drill, int npts, const double* pc, const double* w, double* out) {
  const int nn = kind, n = 6 * nn;
  double Dt56[4];
  for (int i = 0; i < 4; ++i) Dt56[i] = Dt[i] * (5.0 / 6.0);
  int bad = 0;
  for (int64_t e = 0; e < nelem; ++e) {
    double Ke[MAXN * MAXN];
    const int64_t* c = conn + e * nn;
    if (nn == 3)
      t3ff_element(c, xyz, nnodes, normals, valid, Dps, Dt56, t, alpha, drill, Ke);
    else
      bad |= q4rs_element(c, xyz, nnodes, normals, valid, Dps, Dt56, t, alpha, drill, npts, pc, w, Ke);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) out[e * n * n + j * n + i] = Ke[i * n + j];
  }
  return bad;
}

/* Julia `sparse(I,J,V,m,n)`: counting sort by column, rows sorted within a column,
 * duplicates summed in input order, zeros kept.  Two-pass: pass colptr == NULL to get nnz.
 * Only entries with row <= mlim and col <= nlim are kept (matrix_blocked(...)[:ff]). */
int64_t ref_coo_to_csc(int64_t nt, const int64_t* I, const int64_t* J, const double* V, int64_t m, int64_t n, int64_t mlim,
                       int64_t nlim, int64_t* colptr, int64_t* rowval, double* nzval) {
  /* two stable counting sorts (by row, then by column), as Julia's `sparse!` does via its
   * CSR intermediate: O(nt + m + n), duplicates stay in input order */
  int64_t* rcnt = (int64_t*)calloc((size_t)m + 2, sizeof(int64_t));
  int64_t* ccnt = (int64_t*)calloc((size_t)n + 2, sizeof(int64_t));
  int64_t kept = 0;
  for (int64_t p = 0; p < nt; ++p)
    if (I[p] <= mlim && J[p] <= nlim) {
      rcnt[I[p] + 1]++;
      ccnt[J[p] + 1]++;
      ++kept;
    }
  for (int64_t r = 1; r <= m + 1; ++r) rcnt[r] += rcnt[r - 1];
  for (int64_t c = 1; c <= n + 1; ++c) ccnt[c] += ccnt[c - 1];
  int64_t* byrow = (int64_t*)malloc(sizeof(int64_t) * (size_t)(kept + 1));
  int64_t* ord = (int64_t*)malloc(sizeof(int64_t) * (size_t)(kept + 1));
  for (int64_t p = 0; p < nt; ++p)
    if (I[p] <= mlim && J[p] <= nlim) byrow[rcnt[I[p]]++] = p;
  int64_t* cpos = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 2));
  memcpy(cpos, ccnt, sizeof(int64_t) * (size_t)(n + 2));
  for (int64_t q = 0; q < kept; ++q) ord[cpos[J[byrow[q]]]++] = byrow[q];
  int64_t nnz = 0;
  if (colptr) colptr[0] = 1;
  for (int64_t c = 1; c <= nlim; ++c) {
    const int64_t a = ccnt[c], b = ccnt[c + 1];
    for (int64_t i = a; i < b;) {
      const int64_t r = I[ord[i]];
      double s = 0;
      while (i < b && I[ord[i]] == r) s += V[ord[i++]];
      if (colptr) {
        rowval[nnz] = r;
        nzval[nnz] = s;
      }
      ++nnz;
    }
    if (colptr) colptr[c] = nnz + 1;
  }
  free(cpos);
  free(ord);
  free(byrow);
  free(ccnt);
  free(rcnt);
  return nnz;
}

/* y = K x, CSR with 1-based Int64 indices (ThreadedSparseCSR.bmul!: row-parallel) */
void ref_csr_spmv(int64_t n, const int64_t* rowptr, const int64_t* colval, const double* nz, const double* x, double* y,
                  int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double s = 0;
    for (int64_t p = rowptr[i] - 1; p < rowptr[i + 1] - 1; ++p) s += nz[p] * x[colval[p] - 1];
    y[i] = s;
  }
}
/* nsteps of `_cd_loop!` (plate_expl_examples.jl:83-93) with F(t) = fscale[k] * F0 */
void ref_explicit_steps(int64_t n, const int64_t* rowptr, const int64_t* colval, const double* nz, const double* M,
                        double c_scale, double dt, const double* F0, const double* fscale, int64_t nsteps, double* U,
                        double* V, double* A, int nthreads) {
  double* E = (double*)malloc(sizeof(double) * (size_t)n);
  const double dt2_2 = dt * dt / 2, dt_2 = dt / 2;
  for (int64_t s = 0; s < nsteps; ++s) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) U[i] += dt * V[i] + dt2_2 * A[i];
    ref_csr_spmv(n, rowptr, colval, nz, U, E, nthreads);
    const double fs = fscale ? fscale[s] : 1.0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      const double C = c_scale * M[i], inv = 1.0 / (M[i] + dt_2 * C);
      double F = (F0 ? fs * F0[i] : 0.0) - (E[i] + C * (V[i] + dt_2 * A[i]));
      V[i] += dt_2 * A[i];
      A[i] = inv * F;
      V[i] += dt_2 * A[i];
    }
  }
  free(E);
}
