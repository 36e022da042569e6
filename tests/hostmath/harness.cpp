// TEST-ONLY host build of csrc/fsgpu_math.cuh (the __host__ __device__ element math the
// CUDA kernels are made of), so the formulas can be checked against the oracle on a box
// without a GPU.  Not part of libfsgpu.so, never used by the product path.
#include <cstring>

#include "../../finetoolsflexstructures.jl_b200/csrc/fsgpu_math.cuh"
using namespace fsm;

static void block_accumulate(int nrows, const double* d, const double (*bi)[6], const double (*bj)[6], double (&acc)[6][6]) {
  for (int s = 0; s < nrows; ++s)
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) acc[r][c] += bi[s][r] * (d[s] * bj[s][c]);
}

extern "C" {

// out: 18x18 column-major.  group: 34 doubles (A B D H t md mi) or NULL; cs: row-major 3x3
void hm_t3_elmat(const double* X, const double* nrm, const unsigned char* valid, const double* Dps, const double* Dt56,
                 double t, double alpha, double drill, int sheark, const double* group, const double* cs, double* out) {
  const T3Geom g = t3_geometry(v3(X[0], X[1], X[2]), v3(X[3], X[4], X[5]), v3(X[6], X[7], X[8]));
  M3 A[3];
  for (int l = 0; l < 3; ++l) A[l] = nodal_triad(g.E, v3(nrm[3 * l], nrm[3 * l + 1], nrm[3 * l + 2]), valid[l] != 0);
  const double Ae = g.Ae, h = sqrt(2 * Ae);
  const double sk = sheark ? 1.0 / 3 : 1.0;
  Constit C;
  if (group) {
    const double tl = group[31];
    const double stab = tl * tl / (tl * tl + alpha * h * h);
    double m, n;
    layup_angle(g.E, cs, m, n);
    constit_laminate(group, group + 9, group + 18, group + 27, m, n, Ae, stab * Ae * sk, C);
  } else {
    const double stab = t * t / (t * t + alpha * h * h);
    constit_homogeneous(Dps, Dt56, t * Ae, t * t * t / 12 * Ae, t * stab * Ae * sk, C);
  }
  const int nsets = sheark ? 3 : 1;
  const int NR = sheark ? 12 : 8;
  double b[3][12][6], d[12], kpart[3] = {0, 0, 0};
  V3 gdir[3];
  for (int set = 0; set < nsets; ++set) {
    double bs[3][2][3], P1[5][3] = {}, P2[5][3] = {};
    for (int l = 0; l < 3; ++l) {
      t3_bs_node(g, l, sheark ? set : -1, bs[l]);
      double p1[5][3], p2[5][3];
      node_coupling_contrib(A[l], g.gN[l][0], g.gN[l][1], bs[l], p1, p2);
      for (int r = 0; r < 5; ++r)
        for (int k = 0; k < 3; ++k) {
          P1[r][k] += p1[r][k];
          P2[r][k] += p2[r][k];
        }
    }
    for (int j = 0; j < 3; ++j) {
      double R[2][2], brn[5][2], bg[8][6];
      node_R(A[j], R);
      node_bt_rot(g.gN[j][0], g.gN[j][1], bs[j], R, brn);
      kpart[j] += node_kavg_part(C, brn, set > 0);
      gdir[j] = node_strip(g.E, A[j], g.gN[j][0], g.gN[j][1], bs[j], P1, P2, bg);
      fold_constit(C, bg);
      if (set == 0) {
        for (int s = 0; s < 8; ++s) {
          d[s] = constit_d(C, s);
          for (int c = 0; c < 6; ++c) b[j][s][c] = bg[s][c];
        }
      } else {
        for (int s = 0; s < 2; ++s) {
          d[6 + 2 * set + s] = constit_d(C, 6 + s);
          for (int c = 0; c < 6; ++c) b[j][6 + 2 * set + s][c] = bg[6 + s][c];
        }
      }
    }
  }
  const double kavg = (kpart[0] + kpart[1] + kpart[2]) / 6 * drill;
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) {
      double acc[6][6] = {};
      block_accumulate(NR, d, b[i], b[j], acc);
      if (i == j && valid[j]) {
        const double gg[3] = {gdir[j].x, gdir[j].y, gdir[j].z};
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) acc[3 + r][3 + c] += kavg * gg[r] * gg[c];
      }
      for (int c = 0; c < 6; ++c)
        for (int r = 0; r < 6; ++r) out[(j * 6 + c) * 18 + i * 6 + r] = acc[r][c];
    }
}

// out: 24x24 column-major; thickness per gp t[npts]; cs: [npts][9] row-major or NULL
int hm_q4_elmat(const double* Xin, const double* nrm, const unsigned char* valid, const double* Dps, const double* Dt56,
                const double* t, double alpha, double drill, int npts, const double* xi, const double* eta, const double* w,
                const double* group, const double* cs, double* out) {
  V3 X[4];
  for (int a = 0; a < 4; ++a) X[a] = v3(Xin[3 * a], Xin[3 * a + 1], Xin[3 * a + 2]);
  double md = 0;
  for (int a = 1; a < 4; ++a) {
    V3 dd = X[a] - X[0];
    md = fmax(md, dot(dd, dd));
  }
  const double hq = sqrt(md);
  static double K[24][24];
  memset(K, 0, sizeof K);
  int singular = 0;
  for (int gp = 0; gp < npts; ++gp) {
    const Q4Geom g = q4_geometry(X, xi[gp], eta[gp]);
    singular |= g.singular;
    M3 A[4];
    double bs[4][2][3], P1[5][3] = {}, P2[5][3] = {};
    for (int a = 0; a < 4; ++a) {
      A[a] = nodal_triad(g.E, v3(nrm[3 * a], nrm[3 * a + 1], nrm[3 * a + 2]), valid[a] != 0);
      q4_mitc_bs_node(g, xi[gp], eta[gp], a, bs[a]);
      double p1[5][3], p2[5][3];
      node_coupling_contrib(A[a], g.gN[a][0], g.gN[a][1], bs[a], p1, p2);
      for (int r = 0; r < 5; ++r)
        for (int k = 0; k < 3; ++k) {
          P1[r][k] += p1[r][k];
          P2[r][k] += p2[r][k];
        }
    }
    Constit C;
    const double jw = g.Jac * w[gp];
    if (group) {
      const double tl = group[31];
      const double stab = tl * tl / (tl * tl + alpha * hq * hq);
      double m, n;
      layup_angle(g.E, cs + 9 * gp, m, n);
      constit_laminate(group, group + 9, group + 18, group + 27, m, n, jw, stab * jw, C);
    } else {
      const double tt = t[gp];
      const double stab = tt * tt / (tt * tt + alpha * hq * hq);
      constit_homogeneous(Dps, Dt56, tt * jw, tt * tt * tt / 12.0 * jw, tt * stab * jw, C);
    }
    double b[4][8][6], d[8];
    for (int j = 0; j < 4; ++j) {
      node_strip(g.E, A[j], g.gN[j][0], g.gN[j][1], bs[j], P1, P2, b[j]);
      fold_constit(C, b[j]);
    }
    for (int s = 0; s < 8; ++s) d[s] = constit_d(C, s);
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 4; ++i) {
        double acc[6][6] = {};
        block_accumulate(8, d, b[i], b[j], acc);
        for (int r = 0; r < 6; ++r)
          for (int c = 0; c < 6; ++c) K[i * 6 + r][j * 6 + c] += acc[r][c];
      }
  }
  // drilling
  double tsum = 0;
  int cnt = 0, ok[4] = {0, 0, 0, 0};
  for (int k = 0; k < 4; ++k) {
    const double* n4 = nrm + 3 * k;
    const double nl = sqrt(n4[0] * n4[0] + n4[1] * n4[1] + n4[2] * n4[2]);
    if (valid[k] && nl != 0.0) {
      ok[k] = 1;
      double Pm[3][3], KP[3][3];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Pm[r][c] = (r == c ? 1.0 : 0.0) - n4[r] / nl * n4[c] / nl;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          KP[r][c] = 0;
          for (int q = 0; q < 3; ++q) KP[r][c] += K[6 * k + 3 + r][6 * k + 3 + q] * Pm[q][c];
        }
      double tr = 0;
      for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) tr += Pm[r][q] * KP[q][r];
      tsum += fmax(0.0, tr / 2);
      cnt++;
    }
  }
  if (drill != 0.0 && cnt > 0) {
    const double kavg = tsum / cnt * drill;
    if (kavg != 0.0)
      for (int k = 0; k < 4; ++k)
        if (ok[k])
          for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) K[6 * k + 3 + r][6 * k + 3 + c] += kavg * nrm[3 * k + r] * nrm[3 * k + c];
  }
  for (int c = 0; c < 24; ++c)
    for (int r = 0; r < 24; ++r) out[c * 24 + r] = K[r][c];
  return singular;
}

// beam: op 0 stiffness, 1 mass, 2 geo -> out 12x12 column-major; op 3: restoring force -> out[12]
void hm_beam(const double* x0, const double* u, const double* RI, const double* RJ, const double* sec, double E, double nu,
             double rho, int mass_type, int op, double* out) {
  BeamSec s{sec[0], sec[1], sec[2], sec[3], sec[4], sec[5], sec[6], v3(sec[7], sec[8], sec[9])};
  const BeamKin k = beam_kinematics(v3(x0[0], x0[1], x0[2]), v3(x0[3], x0[4], x0[5]), v3(u[0], u[1], u[2]),
                                    v3(u[3], u[4], u[5]), RI, RJ, s.x1x2);
  const double G = E / 2 / (1 + nu);
  const double F[3][3] = {{k.Ft.e1.x, k.Ft.e2.x, k.Ft.e3.x}, {k.Ft.e1.y, k.Ft.e2.y, k.Ft.e3.y}, {k.Ft.e1.z, k.Ft.e2.z, k.Ft.e3.z}};
  double DN[6], aN[6][12];
  beam_natural_stiffness(E, G, s, k.L1, DN);
  beam_aN(k.L1, aN);
  if (op == 3) {
    for (int b = 0; b < 4; ++b)
      for (int r = 0; r < 3; ++r) {
        double v = 0;
        for (int a = 0; a < 3; ++a) {
          double lf = 0;
          for (int m = 0; m < 6; ++m) lf += aN[m][b * 3 + a] * (DN[m] * k.dN[m]);
          v += F[r][a] * (-lf);
        }
        out[b * 3 + r] = v;
      }
    return;
  }
  double Kl[12][12];
  if (op == 0) {
    for (int p = 0; p < 12; ++p)
      for (int q = 0; q < 12; ++q) {
        double v = 0;
        for (int m = 0; m < 6; ++m) v += aN[m][p] * DN[m] * aN[m][q];
        Kl[p][q] = v;
      }
  } else if (op == 2) {
    double PN[6];
    for (int m = 0; m < 6; ++m) PN[m] = DN[m] * k.dN[m];
    beam_local_geo(PN, k.L1, Kl);
  } else {
    beam_local_mass(s, rho, k.L0, mass_type, Kl);
  }
  for (int bp = 0; bp < 4; ++bp)
    for (int bq = 0; bq < 4; ++bq)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          double v = 0;
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) v += F[r][a] * Kl[bp * 3 + a][bq * 3 + b] * F[c][b];
          out[(bq * 3 + c) * 12 + bp * 3 + r] = v;
        }
}
}
