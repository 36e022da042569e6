// libfsgpu element kernels: T3FF / Q4RS shells (homogeneous + layered), corotational
// beam; scatter of the element matrices into the CSC value array through the slot map.
//
// Thread mapping (v1):
//   T3 : 3 lanes per element (lane = element node j), 10 elements per warp.  Each lane
//        builds its node's 8x6 global-dof B strip (folded with the LDL' factor of D),
//        publishes it through shared memory and then forms the block column
//        K[:, node j] = sum_s d_s b_i[s]' b_j[s] one 6x6 block at a time.
//   Q4 : 16 lanes per element (half warp).  Setup role lane = (integration point g, node j)
//        builds the B strip of node j at point g; product role lane = (i, j) accumulates
//        the 6x6 block K_ij over all 8*npts strain rows held in shared memory.
//   L2 : 4 lanes per element, lane = 6x6 block (I, J).
// All arithmetic is FP64 on the CUDA cores; the scatter uses RED.ADD.F64 into L2.
#include <cub/cub.cuh>

#include "fsgpu_internal.cuh"
#include "fsgpu_math.cuh"
#include "fsgpu_shell.cuh"

using namespace fs;
using namespace fsm;
using namespace fsk;

#ifndef FS_T3_MINB
#define FS_T3_MINB 6
#endif
#ifndef FS_Q4_MINB
#define FS_Q4_MINB 4
#endif
// warps per CTA (the warps of a CTA are independent: the CTA size only sets the granularity at which resident slots are
// refilled).  T3: 2 warps x 6 CTAs per SM 2.817 ms against 2.840 (4 x 3), 2.867 (1 x 12), 3.86 (8 x 1: 186 registers);
// Q4: 2.078 (2 x 8), 2.076 (4 x 4), 2.148 (1 x 15), 2.314 (8 x 2)
#ifndef FS_T3_WPB
#define FS_T3_WPB 2
#endif
#ifndef FS_Q4_WPB
#define FS_Q4_WPB 4
#endif
#ifndef FS_T3_PLAN_LD
#define FS_T3_PLAN_LD 38
#endif

// measurement aid (never defined in the shipped build): -DFS_NO_RED keeps all the arithmetic and addressing
// but issues no RED, which gives the compute-side time of the scatter kernels
#ifdef FS_NO_RED
#define FS_RED_GUARD &&(nelem < 0)
#else
#define FS_RED_GUARD
#endif

namespace {

// RED.ADD.F64 under a predicate instead of a branch.  `if (ok) atomicAdd(p, v)` compiles to BSSY / BRA / 4 instructions of
// 64-bit address arithmetic / REDG / BSYNC per add (a quarter of the T3 kernel's instructions); this is one IMAD.WIDE for the
// address (base pointer + 32-bit entry index) and one predicated REDG.  The address is formed unconditionally and only used
// when `ok`.
__device__ __forceinline__ void red_add(double* base, int idx, double v, bool ok) {
#ifdef FS_NO_RED
  ok = false;
#endif
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .u64 a;\n\tsetp.ne.b32 q, %3, 0;\n\tmad.wide.s32 a, %1, 8, %0;\n\t@q red.global.add.f64 [a], %2;\n\t}"
      :
      : "l"(base), "r"(idx), "d"(v), "r"((int)ok)
      : "memory");
}
// unpredicated form: exactly one IMAD.WIDE (entry address = base pointer + 8 * index) and one REDG
__device__ __forceinline__ void red_plain(double* base, int idx, double v) {
#ifndef FS_NO_RED
  asm volatile("{\n\t.reg .u64 a;\n\tmad.wide.s32 a, %1, 8, %0;\n\tred.global.add.f64 [a], %2;\n\t}" ::"l"(base), "r"(idx), "d"(v) : "memory");
#endif
}

// ---- emitters -----------------------------------------------------------------------
// t = e*nnpe + j (element-major, column node j), i = row node, r/c = local dof in the 6x6 block
// Every emitter takes one finished 6x6 block (row node i, column node j of element e).  The
// addressing data is split so that kernels can fetch it early (cols: once per column node,
// rows: once per block) and overlap the load latency with arithmetic.
struct BlockRef {
  int64_t e;
  int i, j;
};
struct EmitScatter {  // generic path: per-entry slot map, any dof numbering
  static constexpr bool kCoop = false;
  static constexpr bool kDenseCoop = false;
  double* nz;
  const int32_t* slot;
  int64_t plane;  // nelem * nnpe
  int nnpe;
  struct Cols {};
  struct Rows {};
  __device__ __forceinline__ Cols cols(int) const { return Cols{}; }
  __device__ __forceinline__ Rows rows(int64_t, int, int, int) const { return Rows{}; }
  __device__ __forceinline__ void block(const BlockRef& b, const Cols&, const Rows&, const double (&a)[6][6]) const {
    const int64_t t = b.e * nnpe + b.j;
    int sl[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) sl[k] = __ldg(slot + ((int64_t)(k * nnpe + b.i)) * plane + t);
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
      for (int r = 0; r < 6; ++r)
        if (sl[c * 6 + r] >= 0) atomicAdd(nz + sl[c * 6 + r], a[r][c]);
  }
};
// Cooperative (warp-transposed) emission: every lane publishes its 6x6 block and its addressing
// in shared memory, then the warp walks the 32 x 36 entries so that consecutive lanes add
// consecutive rows of one column -- RED.ADD.F64 requests that share 32 B sectors (measured:
// 266 G RED/s coalesced vs 194 G RED/s one-lane-per-sector, scripts/micro/red_bench.cu).
constexpr int COOP_DBL = (2 * 32 * 8) / 2;   // T3: colb[32][8] + raw[32][8] (ints); blocks are staged over the dead strips
struct EmitRuns {  // fast path: rows of a node form <= 2 consecutive runs in every column
  static constexpr bool kCoop = true;
  static constexpr bool kDenseCoop = false;
  double* nz;
  const int32_t* pairoff;
  const int32_t* nodeinfo;
  const int32_t* dof;
  const int32_t* colptr;
  int64_t nelem;
  int64_t nc;
  int nnpe;
  const int32_t* nodecol;  // [nnodes][8]: column starts of the node's 6 dofs (-1: not a column), nodeinfo, 0
  const unsigned* plan;  // T3: per-warp emission plan of the symbolic phase (k_t3_plan), or null
  struct Cols {
    int base[6];
  };
  struct Rows {
    int mA, mB, oA, oB;
  };
  // ---- Q4 path: addressing data travels global -> shared with cp.async while the product loop runs (no
  // registers held, no exposed load latency); layout per warp: colb[32][8], raw[32][4] = (nodeinfo, oA, oB, -)
  static constexpr int kAddrInts = 32 * 8 + 32 * 4;
  static constexpr int kStageLd = 38;  // doubles per staged 6x6 block: 36 + 2 (conflict-free reads, 16 B aligned)
  __device__ __forceinline__ void async_addr(int* addr, int lane, bool on, int nj, int ni, int64_t e, int i, int j) const {
    int* colb = addr + lane * 8;
    int* raw = addr + 32 * 8 + lane * 4;
    if (on) {
      const unsigned dc = (unsigned)__cvta_generic_to_shared(colb);
      const unsigned dr = (unsigned)__cvta_generic_to_shared(raw);
      const int32_t* sc = nodecol + (int64_t)nj * 8;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dc), "l"(sc) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dc + 16), "l"(sc + 4) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dr), "l"(nodecol + (int64_t)ni * 8 + 6) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dr + 4),
                   "l"(pairoff + ((int64_t)(i * 2 + 0) * nelem + e) * nnpe + j)
                   : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dr + 8),
                   "l"(pairoff + ((int64_t)(i * 2 + 1) * nelem + e) * nnpe + j)
                   : "memory");
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) colb[k] = -1;
      raw[0] = 0;
      raw[1] = raw[2] = -1;
    }
  }
  // stage every lane's full block in `scratch` (>= 32*kStageLd doubles + 6*32 ints, the dead strip area), then walk
  // the 32 x 36 entries with lane = (block sub-index, row): 6 consecutive lanes add 6 consecutive rows of one column
  __device__ __forceinline__ void coop_emit_full(double* scratch, const int* addr, int lane, const double (&a)[6][6]) const {
    asm volatile("cp.async.wait_all;" ::: "memory");
    double* stage = scratch;
    int* rowp = reinterpret_cast<int*>(scratch + 32 * kStageLd);
    const int* raw = addr + 32 * 8 + lane * 4;
    const int inf = raw[0], oA = raw[1], oB = raw[2];
    const int mA = inf & 63, mB = (inf >> 8) & 63;
    __syncwarp();  // every lane is done with the strips
    int ka = 0, kb = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      int p = -1;
      if ((mA >> r) & 1)
        p = oA >= 0 ? oA + ka++ : -1;
      else if ((mB >> r) & 1)
        p = oB >= 0 ? oB + kb++ : -1;
      rowp[r * 32 + lane] = p;
    }
    double2* st2 = reinterpret_cast<double2*>(stage + lane * kStageLd);
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
      for (int r = 0; r < 6; r += 2) st2[(c * 6 + r) >> 1] = make_double2(a[r][c], a[r + 1][c]);
    __syncwarp();
    const int sub = lane / 6, r = lane - sub * 6;
#ifndef FS_EMIT_UNROLLED
#pragma unroll 1
    for (int g = 0; g < 7; ++g) {
      const int o = g * 5 + sub;
      const bool in = lane < 30 && o < 32;
      const int oo = in ? o : 0;
      const int rp = rowp[r * 32 + oo];
      const int4 c0 = *reinterpret_cast<const int4*>(addr + oo * 8);
      const int2 c1 = *reinterpret_cast<const int2*>(addr + oo * 8 + 4);
      const bool act = in && rp >= 0;
      // one vote per round: when every active lane has all six columns in the pattern the REDs carry no predicates
      const bool fast = __all_sync(0xffffffffu, !act || (c0.x | c0.y | c0.z | c0.w | c1.x | c1.y) >= 0);
      if (!act) continue;
      const int cb[6] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y};
      const double* sv = stage + oo * kStageLd + r;
      double* nzr = nz + rp;
      if (fast) {
#pragma unroll
        for (int c = 0; c < 6; ++c) red_plain(nzr, cb[c], sv[c * 6]);
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) red_add(nzr, cb[c], sv[c * 6], cb[c] >= 0);
      }
    }
#else
    // (build flag; measured slower: Q4 3.30 against 3.20 ms, beam 0.79 against 0.74 ms)  branch-free and unrolled:
    // the shared-memory reads of all seven rounds are in flight together instead of one dependent chain per round
#pragma unroll
    for (int g = 0; g < 7; ++g) {
      const int o = g * 5 + sub;
      const bool in = lane < 30 && o < 32;
      const int oo = in ? o : 0;
      const int rp = rowp[r * 32 + oo];
      const int4 c0 = *reinterpret_cast<const int4*>(addr + oo * 8);
      const int2 c1 = *reinterpret_cast<const int2*>(addr + oo * 8 + 4);
      const int cb[6] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y};
      const double* sv = stage + oo * kStageLd + r;
      double v[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) v[c] = sv[c * 6];
      const bool ok = in && rp >= 0;
#pragma unroll
      for (int c = 0; c < 6; ++c)
        if (ok && cb[c] >= 0 FS_RED_GUARD) atomicAdd(nz + cb[c] + rp, v[c]);
    }
#endif
  }
  // ---- Q4 (DMMA kernel): the element matrix K_e (24 x 24, row stride kQ4KLd) is staged in shared memory by the
  // product stage; the addressing data is kept per NODE (ncol[8][8]: the nodecol rows of the warp's 2 x 4 nodes) and per
  // block (pr[32][2]: run offsets oA, oB of block (bi, bj)), fetched with cp.async while the setup / product stages run
  static constexpr int kQ4KLd = 26;
  __device__ __forceinline__ void q4_async_addr(int* ncol, int* pr, int lane, bool on, int nown, int64_t e) const {
    // ncol: lanes l16 < 4 of each half-warp fetch the nodecol row (32 B) of their node.  pr[i*2+w][half][j]: the run
    // offsets of the warp's two elements are 8 consecutive ints of pairoff[i][w][nelem][4] per (i, w): 16 lanes x 16 B
    // (scattered 4-byte cp.async cost one shared-memory wavefront per lane)
    const int l16 = lane & 15;
    if (l16 < 4) {
      int* c = ncol + ((lane >> 4) * 4 + l16) * 8;
      if (on) {
        const unsigned dc = (unsigned)__cvta_generic_to_shared(c);
        const int32_t* sc = nodecol + (int64_t)nown * 8;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dc), "l"(sc) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dc + 16), "l"(sc + 4) : "memory");
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = k < 6 ? -1 : 0;
      }
    }
    // lane -> (iw = l16 >> 1 .. , element of the pair): lanes 0..15 cover iw = 0..7 x 2 elements
    if (lane < 16) {
      const int iw = lane >> 1, hh = lane & 1;
      const int64_t e0 = e - (lane >> 4);  // (lane < 16: this lane's own element is the warp's first)
      int* d = pr + iw * 8 + hh * 4;
      if (e0 + hh < nelem) {
        const unsigned dr = (unsigned)__cvta_generic_to_shared(d);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dr), "l"(pairoff + ((int64_t)iw * nelem + e0 + hh) * 4) : "memory");
      } else {
        d[0] = d[1] = d[2] = d[3] = -1;
      }
    }
  }
  // Emission of the warp's two staged matrices.  Lane = one ROW of K_e (row node bi = lane / 6, dof r = lane % 6; lanes
  // 24..31 idle), fixed for the whole emission; the rounds walk (element h, column node bj) with compile-time indices.
  // Per round a lane reads its run offset, the six column starts of node bj (one broadcast read for the warp) and the
  // six values K[row][6 bj ..] (three 16-byte reads, conflict-free: consecutive rows are 13 sixteen-byte units apart),
  // and adds them: the 24 lanes of a RED instruction cover four 6-row runs of ONE column.  Against the earlier
  // 5-blocks-per-round mapping (lane = (block, row), 7 rounds) this is 8 rounds of ~30 instead of 7 of ~85 instructions.
  __device__ __forceinline__ void q4_emit_k(const double* k0, int kel, const int* ncol, const int* pr, int lane) const {
    if (lane >= 24) return;
    const int bi = lane / 6, r = lane - 6 * bi, below = (1 << r) - 1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // this row's place in the runs of its node (element h): run A, run B, or absent
      const int inf = ncol[(h * 4 + bi) * 8 + 6];
      const int mA = inf & 63, mB = (inf >> 8) & 63;
      const bool inA = (mA >> r) & 1, inB = (mB >> r) & 1;
      const int rank = __popc((inA ? mA : mB) & below);
      const int* po = pr + (bi * 2 + (inA ? 0 : 1)) * 8 + h * 4;  // run offsets of (row node bi, column node 0..3)
      const int4 o4 = *reinterpret_cast<const int4*>(po);
      const int off[4] = {o4.x, o4.y, o4.z, o4.w};
      const double* krow = k0 + h * kel + lane * kQ4KLd;
#pragma unroll
      for (int bj = 0; bj < 4; ++bj) {
        const bool act = (inA || inB) && off[bj] >= 0;
        const int4 c0 = *reinterpret_cast<const int4*>(ncol + (h * 4 + bj) * 8);
        const int2 c1 = *reinterpret_cast<const int2*>(ncol + (h * 4 + bj) * 8 + 4);
        // one vote per round: no constrained column, every row present -> unpredicated REDs
        const bool fast = __all_sync(0x00ffffffu, act && (c0.x | c0.y | c0.z | c0.w | c1.x | c1.y) >= 0);
        const double2* kv = reinterpret_cast<const double2*>(krow + 6 * bj);
        const double2 v0 = kv[0], v1 = kv[1], v2 = kv[2];
        double* nzr = nz + (off[bj] + rank);
        if (fast) {
          red_plain(nzr, c0.x, v0.x);
          red_plain(nzr, c0.y, v0.y);
          red_plain(nzr, c0.z, v1.x);
          red_plain(nzr, c0.w, v1.y);
          red_plain(nzr, c1.x, v2.x);
          red_plain(nzr, c1.y, v2.y);
        } else {
          red_add(nzr, c0.x, v0.x, act && c0.x >= 0);
          red_add(nzr, c0.y, v0.y, act && c0.y >= 0);
          red_add(nzr, c0.z, v1.x, act && c0.z >= 0);
          red_add(nzr, c0.w, v1.y, act && c0.w >= 0);
          red_add(nzr, c1.x, v2.x, act && c1.x >= 0);
          red_add(nzr, c1.y, v2.y, act && c1.y >= 0);
        }
      }
    }
  }
  // ---- T3 path: cooperative emission (optionally with IN-WARP MERGING, build flag FS_T3_MERGE).  A lane (element,
  // own node j) forms only two products: D = K_e[j, j] and X = K_e[next(j), j]; K_e[j, next(j)] = X' by symmetry.
  // With merging, blocks of different elements of the warp that land on the same matrix block (same node on the
  // diagonal; same edge, either orientation) are summed in shared memory and added once (194 instead of 324 RED per
  // element for a strip of 5 quads).
  //   Lanes: elements 0..4 on lanes 0..14, elements 5..9 on lanes 16..30 (lanes 15, 31 idle), so that no element
  //   straddles a half-warp: the product stage's reads of the neighbour lane's strip stay conflict-free.
  //   addr area per warp, structure of arrays (a cp.async / read of one plane by consecutive lanes is conflict-free):
  //   C0[32] int4 = column starts of dofs 0..3 of the own node; C1[32] int4 = (dof 4, dof 5, nodeinfo, leader list);
  //   P[6][32] = run offsets (oA, oB) of the three targets (j,j), (next,j), (j,next); M[2][32] = merge group mask, row node
  static constexpr int kT3C1 = 128, kT3P = 256, kT3M = 448;  // int offsets of the planes (512 ints = COOP_DBL doubles)
  static __device__ __forceinline__ int t3_lane_j(int lane) { return (lane & 15) % 3; }
  __device__ __forceinline__ void t3_async_addr(int* addr, int lane, bool on, int nj, int64_t e, int j, int jn) const {
    if (on) {
      const unsigned d0 = (unsigned)__cvta_generic_to_shared(addr + lane * 4);
      const unsigned d1 = (unsigned)__cvta_generic_to_shared(addr + kT3C1 + lane * 4);
      const int32_t* sc = nodecol + (int64_t)nj * 8;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(sc) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d1), "l"(sc + 4) : "memory");
      const unsigned dp = (unsigned)__cvta_generic_to_shared(addr + kT3P + lane);
      const int ii[3] = {j, jn, j}, jj[3] = {j, j, jn};
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int w = 0; w < 2; ++w)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dp + (t * 2 + w) * 128),
                       "l"(pairoff + ((int64_t)(ii[t] * 2 + w) * nelem + e) * 3 + jj[t])
                       : "memory");
    } else {
      *reinterpret_cast<int4*>(addr + lane * 4) = make_int4(-1, -1, -1, -1);
      *reinterpret_cast<int4*>(addr + kT3C1 + lane * 4) = make_int4(-1, -1, 0, 0);
#pragma unroll
      for (int k = 0; k < 6; ++k) addr[kT3P + k * 32 + lane] = -1;
    }
  }
  // position (relative to the column start) of dof r of a node with run masks `inf` and run offsets oA / oB
  static __device__ __forceinline__ int row_pos(int inf, int oA, int oB, int r) {
    // branch-free: the row is in run A, in run B, or absent
    const int mA = inf & 63, mB = (inf >> 8) & 63, below = (1 << r) - 1;
    const bool inA = (mA >> r) & 1, inB = (mB >> r) & 1;
    const int off = inA ? oA : oB, m = inA ? mA : mB;
    const int p = off + __popc(m & below);
    return ((inA || inB) && off >= 0) ? p : -1;
  }
  // merge groups of the warp: lanes with equal keys; the lowest lane of a group is its leader.  Writes the group
  // mask to M[0][lane] and the compact leader list to C1[k].w; returns the number of leaders.
  __device__ __forceinline__ int t3_groups(int* addr, int lane, bool on, unsigned long long key) const {
    const unsigned full = 0xffffffffu;
    const unsigned grp = __match_any_sync(full, on ? key : (0xffffffff00000000ull | (unsigned)lane));
    const bool leader = on && (__ffs(grp) - 1 == lane);
    const unsigned leaders = __ballot_sync(full, leader);
    addr[kT3M + lane] = (int)grp;
    if (leader) addr[kT3C1 + __popc(leaders & ((1u << lane) - 1)) * 4 + 3] = lane;
    return __popc(leaders);
  }
  // diagonal pass: `d` = upper triangle of the symmetric K_e[j, j] (row-major, r <= c); lanes with the same own node
  // are merged.  `stage`: 32 x kStageLd doubles (the dead strips).
  template <bool MERGE>
  __device__ __forceinline__ void t3_emit_diag(double* stage, int* addr, int lane, bool on, int nj, const double (&d)[21]) const {
    const int nlead = MERGE ? t3_groups(addr, lane, on, (unsigned long long)(unsigned)nj) : 30;
    {
      double2* st2 = reinterpret_cast<double2*>(stage + lane * kStageLd);
#pragma unroll
      for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int q = 0; q < 6; q += 2) st2[(c * 6 + q) >> 1] = make_double2(d[tri(q, c)], d[tri(q + 1, c)]);
    }
    __syncwarp();
    const int sub = lane / 6, r = lane - sub * 6;
#pragma unroll 1
    for (int g = 0; g * 5 < nlead; ++g) {
      const int k = g * 5 + sub;
      const bool in = lane < 30 && k < nlead;
      const int o = in ? (MERGE ? addr[kT3C1 + k * 4 + 3] : k + (k >= 15)) : 0;  // lane that staged the block
      const int4 c0 = *reinterpret_cast<const int4*>(addr + o * 4);
      const int4 c1 = *reinterpret_cast<const int4*>(addr + kT3C1 + o * 4);
      const int rp = row_pos(c1.z, addr[kT3P + o], addr[kT3P + 32 + o], r);
      const bool act = in && rp >= 0;
      // one vote per round: when every active lane has all six columns in the pattern the REDs carry no predicates
      const bool fast = __all_sync(0xffffffffu, !act || (c0.x | c0.y | c0.z | c0.w | c1.x | c1.y) >= 0);
      if (!act) continue;
      const int cb[6] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y};
      const double* sp = stage + o * kStageLd + r;
      double v[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) v[c] = sp[c * 6];
      if (MERGE) {
        unsigned mm = (unsigned)addr[kT3M + o];
        for (mm &= mm - 1; mm; mm &= mm - 1) {
          const double* sq = stage + (__ffs(mm) - 1) * kStageLd + r;
#pragma unroll
          for (int c = 0; c < 6; ++c) v[c] += sq[c * 6];
        }
      }
      double* nzr = nz + rp;
      if (fast) {
#pragma unroll
        for (int c = 0; c < 6; ++c) red_plain(nzr, cb[c], v[c]);
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) red_add(nzr, cb[c], v[c], cb[c] >= 0);
      }
    }
  }
  // index of (r, c) in the row-major upper triangle of a symmetric 6x6 block
  static __host__ __device__ constexpr int tri(int r, int c) {
    return r <= c ? r * 6 - r * (r - 1) / 2 + (c - r) : c * 6 - c * (c - 1) / 2 + (r - c);
  }
  // edge pass: `a` = K_e[next(j), j] (rows: node nnext, columns: own node nj).  Lanes on the same edge are merged,
  // the sum S goes to block (row, col) and S' to block (col, row).  `stage`: 32 x kStageLd doubles (the dead strips).
  template <bool MERGE>
  __device__ __forceinline__ void t3_emit_edge(double* stage, int* addr, int lane, bool on, int nj, int nnext,
                                               const double (&a)[6][6]) const {
    const unsigned lo = (unsigned)min(nj, nnext), hi = (unsigned)max(nj, nnext);
    __syncwarp();  // every lane is done with the strips and with the diagonal pass
    int nlead = 30;
    if (MERGE) {
      addr[kT3M + 32 + lane] = nnext;
      nlead = t3_groups(addr, lane, on, ((unsigned long long)lo << 32) | hi);
    }
    double2* st2 = reinterpret_cast<double2*>(stage + lane * kStageLd);
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
      for (int q = 0; q < 6; q += 2) st2[(c * 6 + q) >> 1] = make_double2(a[q][c], a[q + 1][c]);
    __syncwarp();
    const int sub = lane / 6, r = lane - sub * 6;
#pragma unroll 1
    for (int g = 0; g * 5 < nlead; ++g) {
      const int k = g * 5 + sub;
      const bool in = lane < 30 && k < nlead;
      const int o = in ? (MERGE ? addr[kT3C1 + k * 4 + 3] : k + (k >= 15)) : 0;
      const int jo = t3_lane_j(o);
      const int ln = o - jo + (jo == 2 ? 0 : jo + 1);  // lane whose own node is this block's row node
      const int4 co0 = *reinterpret_cast<const int4*>(addr + o * 4);
      const int4 co1 = *reinterpret_cast<const int4*>(addr + kT3C1 + o * 4);
      const int4 cl0 = *reinterpret_cast<const int4*>(addr + ln * 4);
      const int4 cl1 = *reinterpret_cast<const int4*>(addr + kT3C1 + ln * 4);
      // direct target (row node, own node) and transposed target (own node, row node)
      const int rpd = row_pos(cl1.z, addr[kT3P + 64 + o], addr[kT3P + 96 + o], r);
      const int rpt = row_pos(co1.z, addr[kT3P + 128 + o], addr[kT3P + 160 + o], r);
      const bool fast = __all_sync(0xffffffffu, !in || ((co0.x | co0.y | co0.z | co0.w | co1.x | co1.y | cl0.x | cl0.y | cl0.z | cl0.w |
                                                           cl1.x | cl1.y | rpd | rpt) >= 0));
      if (!in) continue;
      const double* so = stage + o * kStageLd;
      double vd[6], vt[6];  // S[r][c] and S[c][r]
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        vd[c] = so[c * 6 + r];
        vt[c] = so[r * 6 + c];
      }
      if (MERGE) {
        const int rowO = addr[kT3M + 32 + o];
        unsigned mm = (unsigned)addr[kT3M + o];
        for (mm &= mm - 1; mm; mm &= mm - 1) {
          const int p = __ffs(mm) - 1;
          const bool same = addr[kT3M + 32 + p] == rowO;
          const double* sp = stage + p * kStageLd;
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            const double x = sp[c * 6 + r], y = sp[r * 6 + c];
            vd[c] += same ? x : y;
            vt[c] += same ? y : x;
          }
        }
      }
      const int cbd[6] = {co0.x, co0.y, co0.z, co0.w, co1.x, co1.y};
      const int cbt[6] = {cl0.x, cl0.y, cl0.z, cl0.w, cl1.x, cl1.y};
      double *nzd = nz + rpd, *nzt = nz + rpt;
      if (fast) {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          red_plain(nzd, cbd[c], vd[c]);
          red_plain(nzt, cbt[c], vt[c]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          red_add(nzd, cbd[c], vd[c], cbd[c] >= 0 && rpd >= 0);
          red_add(nzt, cbt[c], vt[c], cbt[c] >= 0 && rpt >= 0);
        }
      }
    }
  }
  // ---- T3 emission driven by the per-warp PLAN of the symbolic phase (k_t3_plan).  The plan lists the UNIQUE matrix
  // blocks (row node, column node) the warp's ten elements touch, each with the lanes that contribute to it, sorted by
  // column node: a strip of five quads has 12 + 42 unique blocks instead of 30 + 60 contributions (40 % fewer REDs and L2
  // reduction sectors), and the five blocks of a RED instruction mostly lie in the columns of one or two nodes.
  //   Three lists: DIAGONAL blocks (contributors: D = K_e[j, j] of the lanes whose own node it is), LOWER blocks (row node >
  //   column node) and UPPER blocks of the element edges.  A lane stages ONE edge block Y = K_e[max, min] of its edge
  //   (own node, next node), i.e. X or X' depending on the orientation; a lower block is the sum of its contributors' Y
  //   read by rows (16-byte reads), the upper block of the same edge the sum of the same Y read by columns.
  //   descriptor (32 bits): [0:5) lane whose own node is the column node, [5:10) lane whose own node is the row node,
  //   [10:12) P-plane pair of the run offsets (0: (j,j), 1: (next,j), 2: (j,next)), [12:17) lane that holds them,
  //   [17:32) contributor lanes (5 bits each, three for a diagonal block, two for an edge block; 31 = none: the slot of
  //   idle lane 31 is staged as zeros).  Blocks with more contributors are split over several descriptors (RED adds).
  //   per warp: word 0 = nD | nL << 8 | nU << 16, then the descriptors of the three lists.
  // doubles per staged block of the plan-driven emission (measured: 38 2.69 ms, 40 3.08, 42 2.68, 44 2.73, 46 2.70)
  static constexpr int kT3PlanLd = FS_T3_PLAN_LD;
  static constexpr int kT3PlanStride = 92;  // 32-bit words per warp (1 + at most 30 + 30 + 30, one pad: 16-byte multiple)
  __device__ __forceinline__ void t3_plan_round(const int* addr, unsigned ds, bool in, int r, int4& c0, int4& c1, int& pos,
                                                bool& act) const {
    const int cl = (int)(ds & 31), rl = (int)((ds >> 5) & 31), ot = (int)((ds >> 10) & 3), ol = (int)((ds >> 12) & 31);
    c0 = *reinterpret_cast<const int4*>(addr + cl * 4);
    c1 = *reinterpret_cast<const int4*>(addr + kT3C1 + cl * 4);
    const int inf = addr[kT3C1 + rl * 4 + 2];
    const int mA = inf & 63, mB = (inf >> 8) & 63;
    const bool inA = (mA >> r) & 1, inB = (mB >> r) & 1;
    const int off = addr[kT3P + ot * 64 + (inA ? 0 : 32) + ol];
    pos = off + __popc((inA ? mA : mB) & ((1 << r) - 1));
    act = in && (inA || inB) && off >= 0;
  }
  __device__ __forceinline__ void t3_plan_red(const int4& c0, const int4& c1, bool act, int pos, const double (&v)[6]) const {
    const bool fast = __all_sync(0xffffffffu, !act || (c0.x | c0.y | c0.z | c0.w | c1.x | c1.y) >= 0);
    if (!act) return;
    double* nzr = nz + pos;
    if (fast) {
      red_plain(nzr, c0.x, v[0]);
      red_plain(nzr, c0.y, v[1]);
      red_plain(nzr, c0.z, v[2]);
      red_plain(nzr, c0.w, v[3]);
      red_plain(nzr, c1.x, v[4]);
      red_plain(nzr, c1.y, v[5]);
    } else {
      red_add(nzr, c0.x, v[0], c0.x >= 0);
      red_add(nzr, c0.y, v[1], c0.y >= 0);
      red_add(nzr, c0.z, v[2], c0.z >= 0);
      red_add(nzr, c0.w, v[3], c0.w >= 0);
      red_add(nzr, c1.x, v[4], c1.x >= 0);
      red_add(nzr, c1.y, v[5], c1.y >= 0);
    }
  }
  // sum of row r of the blocks staged by the (up to three) contributor lanes of descriptor ds
  // (NC contributor fields are read unconditionally, an empty one as the zero block of lane 31.  Measured alternatives:
  // a warp vote per field that skips fields no block of the round uses, 2.84 ms; the blocks of a list ordered by their
  // number of contributors so that whole rounds skip fields, 2.83 ms -- it breaks the ordering by column node; both 2.92;
  // against 2.71 ms for this form in the same build)
  template <int NC>
  static __device__ __forceinline__ void t3_plan_rows(const double* stage, unsigned ds, int r, double (&v)[6]) {
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int src = (int)((ds >> (17 + 5 * k)) & 31);
      const double2* sp = reinterpret_cast<const double2*>(stage + src * kT3PlanLd + r * 6);
      const double2 a0 = sp[0], a1 = sp[1], a2 = sp[2];
      if (k == 0) {
        v[0] = a0.x, v[1] = a0.y, v[2] = a1.x, v[3] = a1.y, v[4] = a2.x, v[5] = a2.y;
      } else {
        v[0] += a0.x, v[1] += a0.y, v[2] += a1.x, v[3] += a1.y, v[4] += a2.x, v[5] += a2.y;
      }
    }
  }
  // `on`: the lane holds blocks of an element (idle lanes stage zeros: lane 31 is the plan's "no contributor");
  // `swap`: the own node is the larger one of the lane's edge (own node, next node)
  __device__ __forceinline__ void t3_emit_plan(double* stage, const int* addr, const unsigned* pl, int lane, bool on, bool swap,
                                               const double (&d)[21], const double (&a)[6][6]) const {
    const int slot = lane / 6, r = lane - 6 * slot;
    const unsigned hdr = pl[0];
    const int nD = (int)(hdr & 255), nL = (int)((hdr >> 8) & 255), nU = (int)((hdr >> 16) & 255);
    double2* st2 = reinterpret_cast<double2*>(stage + lane * kT3PlanLd);
    // ---- diagonal blocks (idle lanes stage zeros: lane 31's slot is the plan's "no contributor")
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
      for (int q = 0; q < 6; q += 2)
        st2[(c * 6 + q) >> 1] = on ? make_double2(d[tri(q, c)], d[tri(q + 1, c)]) : make_double2(0.0, 0.0);
    __syncwarp();
#pragma unroll 1
    for (int g = 0; g * 5 < nD; ++g) {
      const int b = g * 5 + slot;
      const bool in = lane < 30 && b < nD;
      const unsigned ds = pl[1 + (in ? b : 0)];
      int4 c0, c1;
      int pos;
      bool act;
      t3_plan_round(addr, ds, in, r, c0, c1, pos, act);
      double v[6];
      t3_plan_rows<3>(stage, ds, r, v);
      t3_plan_red(c0, c1, act, pos, v);
    }
    __syncwarp();
    // ---- edge blocks: Y = K_e[max, min] row-major (X = K_e[next, own] or its transpose)
    // (measured: applying the orientation to the store index instead -- 36 scalar stores, no selects -- and skipping the
    // idle lanes' stores is slower, 2.93 against 2.68 ms: more spills in the product stage)
#pragma unroll
    for (int q = 0; q < 6; ++q)
#pragma unroll
      for (int c = 0; c < 6; c += 2) {
        const double y0 = swap ? a[c][q] : a[q][c], y1 = swap ? a[c + 1][q] : a[q][c + 1];
        st2[(q * 6 + c) >> 1] = on ? make_double2(y0, y1) : make_double2(0.0, 0.0);
      }
    __syncwarp();
#pragma unroll 1
    for (int g = 0; g * 5 < nL; ++g) {  // lower blocks: rows of the summed Y
      const int b = g * 5 + slot;
      const bool in = lane < 30 && b < nL;
      const unsigned ds = pl[1 + nD + (in ? b : 0)];
      int4 c0, c1;
      int pos;
      bool act;
      t3_plan_round(addr, ds, in, r, c0, c1, pos, act);
      double v[6];
      t3_plan_rows<2>(stage, ds, r, v);
      t3_plan_red(c0, c1, act, pos, v);
    }
#pragma unroll 1
    for (int g = 0; g * 5 < nU; ++g) {  // upper blocks: columns of the summed Y
      const int b = g * 5 + slot;
      const bool in = lane < 30 && b < nU;
      const unsigned ds = pl[1 + nD + nL + (in ? b : 0)];
      int4 c0, c1;
      int pos;
      bool act;
      t3_plan_round(addr, ds, in, r, c0, c1, pos, act);
      double v[6];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int src = (int)((ds >> (17 + 5 * k)) & 31);
        const double* sp = stage + src * kT3PlanLd + r;
#pragma unroll
        for (int c = 0; c < 6; ++c) v[c] = k == 0 ? sp[c * 6] : v[c] + sp[c * 6];
      }
      t3_plan_red(c0, c1, act, pos, v);
    }
  }
};
struct EmitDense {
  static constexpr bool kCoop = false;
  static constexpr bool kDenseCoop = true;  // the Q4 kernel writes its staged K_e with the whole warp, coalesced
  double* out;     // [nelem][n][n] column-major per element (fsgpu_element_matrices), or
  int nnpe;
  int blockmajor;  // [nelem][nnpe(j)][nnpe(i)][6 c][6 r]: every 6x6 block contiguous (the order-fixed gather path reads blocks)
  // the warp's two staged matrices (24 x 24, row stride ld, bitwise symmetric: K[a][b] is read as K[b][a] so that the
  // fastest index runs along a staged row) -> global, 16 bytes per lane and store
  __device__ __forceinline__ void q4_dense_from_k(const double* k0, int kel, int ld, int64_t e0, int64_t nelem, int lane) const {
#pragma unroll 3
    for (int j2 = lane; j2 < 2 * 288; j2 += 32) {
      const int h = j2 >= 288, k2 = j2 - 288 * h;  // k2: index of the double2 within the element
      if (e0 + h >= nelem) continue;
      int src;
      if (blockmajor) {
        const int blk = k2 / 18, w = k2 - 18 * blk, bj = blk >> 2, bi = blk & 3, c = w / 3, r = 2 * (w - 3 * c);
        src = (6 * bj + c) * ld + 6 * bi + r;
      } else {
        const int col = k2 / 12, row = 2 * (k2 - 12 * col);
        src = col * ld + row;
      }
      reinterpret_cast<double2*>(out + (e0 + h) * 576)[k2] = *reinterpret_cast<const double2*>(k0 + h * kel + src);
    }
  }
  struct Cols {};
  struct Rows {};
  __device__ __forceinline__ Cols cols(int) const { return Cols{}; }
  __device__ __forceinline__ Rows rows(int64_t, int, int, int) const { return Rows{}; }
  __device__ __forceinline__ void block(const BlockRef& b, const Cols&, const Rows&, const double (&a)[6][6]) const {
    const int n = 6 * nnpe;
    double* o = out + b.e * n * n;
    if (blockmajor) {
      double2* o2 = reinterpret_cast<double2*>(o + (b.j * nnpe + b.i) * 36);
#pragma unroll
      for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int r = 0; r < 6; r += 2) o2[(c * 6 + r) >> 1] = make_double2(a[r][c], a[r + 1][c]);
      return;
    }
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
      for (int r = 0; r < 6; ++r) o[(int64_t)(b.j * 6 + c) * n + (b.i * 6 + r)] = a[r][c];
  }
};

// =====================================================================================
// T3FF / T3FFComp stiffness
// =====================================================================================
constexpr int T3_EPW = 10;  // elements per warp (3 lanes each; lanes 30, 31 idle)
// Which emission a T3 instantiation uses: the plan-driven one (EmitRuns::t3_emit_plan: 2.82 -> 2.68 ms on C4) for the
// homogeneous shell; the laminated kernel keeps one block per lane and pass (its setup stage leaves fewer registers: with the
// plan-driven emission it spills 432 instead of 208 bytes and C3 goes from 1.68 to 1.78 ms).  -DFS_T3_OLD_EMIT: never the plan.
#ifdef FS_T3_OLD_EMIT
template <bool COMP>
constexpr bool kT3Plan = false;
#else
template <bool COMP>
constexpr bool kT3Plan = !COMP;
#endif
// per-warp shared memory of k_t3_stiffness: [strips | staged blocks] [addressing planes] [emission plan]
__host__ __device__ constexpr int t3_stage_doubles(bool sheark, bool) { return (sheark ? 12 : 8) * 6 * 32; }
__host__ __device__ constexpr int t3_warp_doubles(bool sheark, bool coop) {
  return t3_stage_doubles(sheark, coop) + (coop ? COOP_DBL + EmitRuns::kT3PlanStride / 2 : 0);
}

template <bool COMP, bool SHEARK, class Emit>
__global__ void __launch_bounds__(32 * FS_T3_WPB, FS_T3_MINB) k_t3_stiffness(ShellArgs P, Emit emit) {
  constexpr int NR = SHEARK ? 12 : 8;
  constexpr int SA = t3_stage_doubles(SHEARK, Emit::kCoop);  // strips, later the staged blocks
  constexpr int WARP_DBL = t3_warp_doubles(SHEARK, Emit::kCoop);
  extern __shared__ double smem[];  // per warp: strips [NR*6][32], rows pre-scaled by sqrt(d_s) (+ coop scratch)
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* sw = smem + (size_t)wib * WARP_DBL;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  // elements 0..4 on lanes 0..14, elements 5..9 on lanes 16..30 (no element straddles a half-warp)
  const int l15 = lane & 15, el = (lane >> 4) * 5 + l15 / 3, j = l15 % 3;
  const int64_t e = warp * T3_EPW + el;
  const bool active = (l15 < 15) && (e < P.nelem);
  const unsigned full = 0xffffffffu;
  const int base = lane - j;

  double kpart = 0.0;  // this node's share of the nodal-basis bending diagonal
  V3 gdir = v3(0, 0, 0);
  bool validj = false;
  int nn[3] = {0, 0, 0};
  T3Geom g;
  M3 A;
  Constit C;
  double wm = 0.0, wb = 0.0, ws = 0.0;  // homogeneous weights
  if (active) {
    const int32_t* cn = P.conn + e * 3;
    nn[0] = __ldg(cn);
    nn[1] = __ldg(cn + 1);
    nn[2] = __ldg(cn + 2);
  }
  if constexpr (Emit::kCoop) {
    // addressing data: requested first (cp.async into its own shared-memory area): the element index is dead before
    // the register-heavy passes, and the loads overlap all of them (3.29 -> 3.20 ms on C4)
    const int jn0 = j == 2 ? 0 : j + 1;
    emit.t3_async_addr(reinterpret_cast<int*>(sw + SA), lane, active, j == 0 ? nn[0] : (j == 1 ? nn[1] : nn[2]), e, j, jn0);
    if constexpr (!kT3Plan<COMP>) {
    } else if (warp * T3_EPW < P.nelem) {  // this warp's emission plan: 23 chunks of 16 bytes
      const unsigned* src = emit.plan + warp * EmitRuns::kT3PlanStride;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(sw + SA + COOP_DBL);
      if (lane < EmitRuns::kT3PlanStride / 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + lane * 16), "l"(src + lane * 4) : "memory");
    } else if (lane == 0) {  // a warp of the last CTA without elements: an empty plan (there is no entry for it)
      *reinterpret_cast<unsigned*>(sw + SA + COOP_DBL) = 0u;
    }
  }
  if (active) {
    g = t3_geometry(ld3(P.xyz, nn[0]), ld3(P.xyz, nn[1]), ld3(P.xyz, nn[2]));
    const double4 nv = ldg4(P.nrm + (j == 0 ? nn[0] : (j == 1 ? nn[1] : nn[2])));
    validj = nv.w != 0.0;
    A = nodal_triad(g.E, v3(nv.x, nv.y, nv.z), validj);
    if constexpr (COMP) {
      build_constit_t3(P, e, g.E, g.Ae, SHEARK ? (1.0 / 3) : 1.0, true, C);
    } else {
      // homogeneous shell: the LDL' factors of Dps and Dt come from the host (P.hf), only the three
      // weights (membrane, bending, shear) depend on the element
      const double t = P.nthick == 1 ? __ldg(P.thick) : __ldg(P.thick + e);
      const double h2 = 2 * g.Ae;  // h^2, h = sqrt(2 Ae)
      const double stab = P.nstab ? __ldg(P.stabf + e) : t * t * fs_rcp(t * t + P.alpha * h2);
      wm = t * g.Ae;
      wb = (t * t * t) * (1.0 / 12.0) * g.Ae;
      ws = t * stab * g.Ae * (SHEARK ? (1.0 / 3) : 1.0);
    }
  }
  constexpr int NSETS = SHEARK ? 3 : 1;
#pragma unroll
  for (int set = 0; set < NSETS; ++set) {
    double bs[2][3];
    double qf[5], af[6];  // this node's coupling contribution in factored form: p1 = qf (x) A[0,:], p2 = qf (x) A[1,:]
    double gx = 0.0, gy = 0.0;
    if (active) {
      gx = j == 0 ? g.gN[0][0] : (j == 1 ? g.gN[1][0] : g.gN[2][0]);
      gy = j == 0 ? g.gN[0][1] : (j == 1 ? g.gN[1][1] : g.gN[2][1]);
      t3_bs_node(g, j, SHEARK ? set : -1, bs);
      node_coupling_factors(A, gx, gy, bs, qf);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        af[k] = A.a[0][k];
        af[3 + k] = A.a[1][k];
      }
    } else {
      for (int r = 0; r < 5; ++r) qf[r] = 0.0;
      for (int k = 0; k < 6; ++k) af[k] = 0.0;
      for (int r = 0; r < 2; ++r)
        for (int k = 0; k < 3; ++k) bs[r][k] = 0.0;
    }
    // sum the coupling matrices over the element's three lanes, in node order on every lane (all three lanes get
    // bitwise the same P1, P2); the factors travel (11 numbers per node) instead of the 30 products
    double P1[5][3], P2[5][3];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) P1[r][k] = P2[r][k] = 0.0;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      const int srcl = (base + l) & 31;
      double ql[5], al[6];
#pragma unroll
      for (int r = 0; r < 5; ++r) ql[r] = __shfl_sync(full, qf[r], srcl);
#pragma unroll
      for (int k = 0; k < 6; ++k) al[k] = __shfl_sync(full, af[k], srcl);
#pragma unroll
      for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          P1[r][k] = fma(ql[r], al[k], P1[r][k]);
          P2[r][k] = fma(ql[r], al[3 + k], P2[r][k]);
        }
    }
    if (active) {
      double R[2][2], brn[5][2];
      node_R(A, R);
      node_bt_rot(gx, gy, bs, R, brn);
      double bg[8][6];
      gdir = node_strip(g.E, A, gx, gy, bs, P1, P2, bg);
      // K = sum_s d_s b_s (x) b_s = sum_s (sqrt(d_s) b_s) (x) (sqrt(d_s) b_s): d_s > 0 for a positive
      // definite constitutive matrix (a negative pivot raises flag[2] -> FSGPU_ERR_ARG)
      double q[8];
      if constexpr (COMP) {
        kpart += node_kavg_part(C, brn, set > 0);
        fold_constit(C, bg);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const double d = constit_d(C, s);
          if (d < 0.0) atomicExch(P.flag + 2, 1);
          q[s] = fs_sqrt(d);
        }
      } else {
        kpart += node_kavg_part_h(P.hf, wb, ws, brn, set > 0);
        fold_homogeneous(P.hf, bg);
        // sqrt(wb) = sqrt(wm) t / sqrt(12), sqrt(ws) = sqrt(wm) sqrt(stab): see the Q4 setup pass
        // (three independent square roots: deriving qb, qs from qm as the Q4 setup pass does keeps two more doubles live
        // through the register-bound T3 setup and measured slower, 3.36 against 3.18 ms)
        const double qm = fs_sqrt(wm), qb = fs_sqrt(wb), qs = fs_sqrt(ws);
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          q[s] = qm * P.hf.sdps[s];
          q[3 + s] = qb * P.hf.sdps[s];
        }
        q[6] = qs * P.hf.sdts[0];
        q[7] = qs * P.hf.sdts[1];
      }
      if (set == 0) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) sw[(s * 6 + cc) * 32 + lane] = q[s] * bg[s][cc];
      } else {
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) sw[((6 + 2 * set + s) * 6 + cc) * 32 + lane] = q[6 + s] * bg[6 + s][cc];
      }
    }
  }
  // kavg = mean of the 6 bending diagonals * scale (src/FEMMShellT3FFModule.jl:714-722)
  double ksum = 0.0;
#pragma unroll
  for (int l = 0; l < 3; ++l) ksum += __shfl_sync(full, kpart, (base + l) & 31);
  const double kavg = ksum * (1.0 / 6) * P.drill;
  __syncwarp();
  const int nj = j == 0 ? nn[0] : (j == 1 ? nn[1] : nn[2]);
  if constexpr (Emit::kCoop) {
    // Symmetric form with in-warp merging: this lane forms D = K_e[j, j] and X = K_e[next(j), j] only.
    const int jn = j == 2 ? 0 : j + 1;
    const int nnext = jn == 0 ? nn[0] : (jn == 1 ? nn[1] : nn[2]);
    int* addr = reinterpret_cast<int*>(sw + SA);
    // D is symmetric: 21 accumulators (row-major upper triangle); X is a full block.  One pass over the strain
    // rows feeds both (the own row b_j is loaded once).
    double dd[21], acc[6][6];
#pragma unroll
    for (int k = 0; k < 21; ++k) dd[k] = 0.0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) acc[r][cc] = 0.0;
    // (the idle lanes 15 and 31 would read lane 16's / lane 0's column: a bank conflict for their whole half-warp)
    const int src = l15 < 15 ? base + jn : lane;
#pragma unroll
    for (int s = 0; s < NR; ++s) {
      // membrane rows of a homogeneous shell have no rotation columns in global dofs
      constexpr int kMemb = COMP ? 0 : 3;
      const int w = s < kMemb ? 3 : 6;
      double bi[6], bj[6];
#pragma unroll
      for (int r = 0; r < 6; ++r)
        if (r < w) bi[r] = sw[(s * 6 + r) * 32 + src];
#pragma unroll
      for (int cc = 0; cc < 6; ++cc)
        if (cc < w) bj[cc] = sw[(s * 6 + cc) * 32 + lane];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int cc = 0; cc < 6; ++cc)
          if (r < w && cc < w) {
            acc[r][cc] = fma(bi[r], bj[cc], acc[r][cc]);
            if (r <= cc) dd[Emit::tri(r, cc)] = fma(bj[r], bj[cc], dd[Emit::tri(r, cc)]);
          }
    }
    if (validj) {
      // drilling stiffness kavg on the nodal normal direction (nodal dof 6), rotated to global
      const double gg[3] = {gdir.x, gdir.y, gdir.z};
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = r; cc < 3; ++cc) dd[Emit::tri(3 + r, 3 + cc)] += kavg * gg[r] * gg[cc];
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();  // every lane is done with the strips; the addressing data has landed
#ifdef FS_T3_MERGE
    constexpr bool kMerge = true;
#else
    constexpr bool kMerge = false;
#endif
    if constexpr (kT3Plan<COMP>) {
      emit.t3_emit_plan(sw, addr, reinterpret_cast<const unsigned*>(sw + SA + COOP_DBL), lane, l15 < 15, nj > nnext, dd, acc);
    } else {
      emit.template t3_emit_diag<kMerge>(sw, addr, lane, active, nj, dd);
      emit.template t3_emit_edge<kMerge>(sw, addr, lane, active, nj, nnext, acc);
    }
  } else {
    if (!active) return;
    typename Emit::Cols ecols = emit.cols(nj);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      typename Emit::Rows erows = emit.rows(e, i, j, nn[i]);
      double acc[6][6];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) acc[r][cc] = 0.0;
      const int src = base + i;
#pragma unroll
      for (int s = 0; s < NR; ++s) {
        double bi[6], bj[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) bi[r] = sw[(s * 6 + r) * 32 + src];
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) bj[cc] = sw[(s * 6 + cc) * 32 + lane];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) acc[r][cc] = fma(bi[r], bj[cc], acc[r][cc]);
      }
      if (i == j && validj) {
        // drilling stiffness kavg on the nodal normal direction (nodal dof 6), rotated to global
        const double gg[3] = {gdir.x, gdir.y, gdir.z};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) acc[3 + r][3 + cc] += kavg * gg[r] * gg[cc];
      }
      emit.block(BlockRef{e, i, j}, ecols, erows, acc);
    }
  }
}

// =====================================================================================
// Q4RS / Q4RSComp stiffness
// =====================================================================================
// Layout of the strips S (one row per generalised strain and integration point, 24 global dofs per row) in shared
// memory, chosen for the DMMA product K_e = S' S:
//   row base = 24 * row + 4 * (row >> 1)  (832 doubles per element and chunk of 4 points), so that
//   * the m8n8k4 fragment loads (lane l reads S[4q + l%4][8I + l/4]) are conflict-free: the four rows of a k-step
//     start at 0, 8, 4, 12 (mod 16 doubles);
//   * the setup pass's 16-byte stores are conflict-free: lanes of odd integration points write their rows pairwise
//     swapped (s ^ 1), which moves them by 8 doubles (mod 16) against the even points -- the product sums over all
//     rows, their order is free.
constexpr int Q4S_EL = 832;
__host__ __device__ constexpr int q4s_rowoff(int s) { return 24 * s + 4 * (s >> 1); }  // rows of one point: 208 doubles
struct Q4Row {
  double* base;  // element strips + 208 * g4 + 6 * jn
  int od;        // 24 for odd points
  __device__ __forceinline__ double* operator()(int s) const { return base + q4s_rowoff(s) + ((s & 1) ? -od : od); }
};
__device__ __forceinline__ void q4_store_row(double* d, const double (&v)[6]) {
  double2* d2 = reinterpret_cast<double2*>(d);
  d2[0] = make_double2(v[0], v[1]);
  d2[1] = make_double2(v[2], v[3]);
  d2[2] = make_double2(v[4], v[5]);
}

// One setup pass: lane (g4, jn) builds the folded strip of node jn at integration point gp
// (own triad / own shear entries only; the coupling matrices are summed over the four lanes of
// the same point by a shuffle butterfly) and stores it in the half-warp's shared tile.
template <bool COMP>
__device__ __forceinline__ void q4_setup_pass(const ShellArgs& P, bool on, int64_t e, int gp, int g4, int jn, const V3 (&X)[4],
                                              const double4& nvown, double hq2, const double* gd, double* sb_) {
  const unsigned full = 0xffffffffu;
  const Q4Row row{sb_ + 208 * g4 + 6 * jn, (g4 & 1) * 24};
  double p1[5][3], p2[5][3], bs[2][3];
  Q4Geom g;
  M3 A;
  double xi = 0.0, eta = 0.0, w = 0.0, gx = 0.0, gy = 0.0;
  if (on) {
    xi = P.rule.xi[gp];
    eta = P.rule.eta[gp];
    w = P.rule.w[gp];
    g = q4_geometry(X, xi, eta);
    if (g.singular) atomicExch(P.flag, 1);
    A = nodal_triad(g.E, v3(nvown.x, nvown.y, nvown.z), nvown.w != 0.0);
    gx = jn == 0 ? g.gN[0][0] : (jn == 1 ? g.gN[1][0] : (jn == 2 ? g.gN[2][0] : g.gN[3][0]));
    gy = jn == 0 ? g.gN[0][1] : (jn == 1 ? g.gN[1][1] : (jn == 2 ? g.gN[2][1] : g.gN[3][1]));
    q4_mitc_bs_node(g, xi, eta, jn, bs);
    node_coupling_contrib(A, gx, gy, bs, p1, p2);
  } else {
    for (int r = 0; r < 5; ++r)
      for (int k = 0; k < 3; ++k) p1[r][k] = p2[r][k] = 0.0;
  }
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      p1[r][k] += __shfl_xor_sync(full, p1[r][k], 1);
      p2[r][k] += __shfl_xor_sync(full, p2[r][k], 1);
      p1[r][k] += __shfl_xor_sync(full, p1[r][k], 2);
      p2[r][k] += __shfl_xor_sync(full, p2[r][k], 2);
    }
  if (on) {
    const double jw = g.Jac * w;
    const int npts = P.rule.npts;
    if (COMP) {
      // the factored constitutive data of this point (rotation of A, B, D, H into the element frame, 6 x 6 LDL', square
      // roots of the pivots: ~480 FP64 instructions, the same for the four lanes of a point) comes from
      // k_q4_laminate_prep, one thread per point
      const double2* lp = reinterpret_cast<const double2*>(P.lam + (e * npts + gp) * 24);
      double lv[24];
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const double2 v2 = __ldg(lp + k);
        lv[2 * k] = v2.x;
        lv[2 * k + 1] = v2.y;
      }
      Constit C;
#pragma unroll
      for (int i = 0, k = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) C.L6[i][j] = j < i ? lv[k++] : 0.0;
      C.L2 = lv[23];
      double bg[8][6];
      node_strip(g.E, A, gx, gy, bs, p1, p2, bg);
      fold_constit(C, bg);
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        double v[6];
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) v[cc] = lv[15 + s] * bg[s][cc];
        q4_store_row(row(s), v);
      }
    } else {
      const double t = P.nthick == 1 ? __ldg(P.thick) : (P.nthick == P.nelem ? __ldg(P.thick + e) : __ldg(P.thick + e * npts + gp));
      // rows are pre-scaled by sqrt(d_s): sqrt(c) * sqrt(dps) with sqrt(dps), sqrt(dts) from the host;
      // sqrt(t jw), sqrt(t^3/12 jw) = sqrt(t jw) t / sqrt(12), sqrt(t stab jw) = sqrt(t jw) sqrt(stab) with
      // sqrt(stab) = t / sqrt(t^2 + alpha h^2): one square root and one reciprocal square root (hq2 = h^2)
      const double qm = fs_sqrt(t * jw), qb = qm * (t * 0.28867513459481288225), 
                   qs = qm * (P.nstab ? fs_sqrt(__ldg(P.stabf + e)) : t * fs_rsqrt(t * t + P.alpha * hq2));
      {
        double m[3][6];
        strip_membrane(g.E, gx, gy, m);
        fold3(P.hf, m);
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          double v[6];
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) v[cc] = (qm * P.hf.sdps[s]) * m[s][cc];
          q4_store_row(row(s), v);
        }
      }
      const M3 G = global_to_nodal(A, g.E);
      double R[2][2], brn[5][2];
      node_R(A, R);
      node_bt_rot(gx, gy, bs, R, brn);
      {
        double m[3][6];
#pragma unroll
        for (int r = 0; r < 3; ++r) strip_row(g.E, G, brn, gx, gy, 0.0, p1, p2, r, m[r]);
        fold3(P.hf, m);
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          double v[6];
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) v[cc] = (qb * P.hf.sdps[s]) * m[s][cc];
          q4_store_row(row(3 + s), v);
        }
      }
      {
        double r6[6], r7[6], v6[6], v7[6];
        strip_row(g.E, G, brn, gx, gy, bs[0][0], p1, p2, 3, r6);
        strip_row(g.E, G, brn, gx, gy, bs[1][0], p1, p2, 4, r7);
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {
          v6[cc] = (qs * P.hf.sdts[0]) * (r6[cc] + P.hf.Lt * r7[cc]);
          v7[cc] = (qs * P.hf.sdts[1]) * r7[cc];
        }
        q4_store_row(row(6), v6);
        q4_store_row(row(7), v7);
      }
    }
  } else {
    const double z[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int s = 0; s < 8; ++s) q4_store_row(row(s), z);
  }
}

// Q4RSComp: factored constitutive data per (element, integration point) -> P.lam (see ShellArgs::lam).  The layup angle
// (src/TransformerModule.jl:92-105) needs the element triad at the point, the weights the surface Jacobian; laminate
// thickness enters only through the stabilisation factor (src/FEMMShellQ4RSCompModule.jl:918-935).
__global__ void k_q4_laminate_prep(ShellArgs P, double* __restrict__ lam) {
  const int npts = P.rule.npts;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (tid >= P.nelem * npts) return;
  const int64_t e = tid / npts;
  const int gp = (int)(tid - e * npts);
  V3 X[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) X[a] = ld3(P.xyz, __ldg(P.conn + e * 4 + a));
  double md = 0.0;  // quirk: "diameter" = max distance from node 1 (src/FEMMShellQ4RSModule.jl:861-870)
  for (int a = 1; a < 4; ++a) {
    const V3 d = X[a] - X[0];
    md = fmax(md, dot(d, d));
  }
  const Q4Geom g = q4_geometry(X, P.rule.xi[gp], P.rule.eta[gp]);
  const double jw = g.Jac * P.rule.w[gp];
  const double* gd = P.gdata + (size_t)__ldg(P.gof + e) * 34;
  const double t = gd[31];
  const double stab = P.nstab ? __ldg(P.stabf + e) : t * t * fs_rcp(t * t + P.alpha * md);
  double m, n;
  const int64_t ci = P.ncs == 1 ? 0 : (P.ncs == P.nelem ? e : e * npts + gp);
  layup_angle(g.E, P.cs + ci * 9, m, n);
  Constit C;
  constit_laminate(gd, gd + 9, gd + 18, gd + 27, m, n, jw, stab * jw, C);
  double* o = lam + tid * 24;
  int k = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < i; ++j) o[k++] = C.L6[i][j];
  for (int s = 0; s < 8; ++s) {
    const double d = constit_d(C, s);
    if (d < 0.0) atomicExch(P.flag + 2, 1);  // not positive definite -> FSGPU_ERR_ARG
    o[15 + s] = fs_sqrt(d);
  }
  o[23] = C.L2;
}

// K_e += S' S on the FP64 tensor-core path (DMMA.8x8x4, the same peak as DFMA on B200 but one instruction per 256 FMA
// and two fragment registers per operand: profiles/r02_dmma_microbench.txt).  The whole warp works on its two
// elements: per k-step of 4 strain rows three fragment loads (column blocks 0..7, 8..15, 16..23; the A fragment of
// block I is the B fragment of block I) and the six upper tiles (0,0) (0,1) (0,2) (1,1) (1,2) (2,2).
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void q4_mma_product(const double* wbase, int lane, double (&acc)[2][6][2]) {
  const int t = lane & 3;
  const double* p = wbase + 24 * t + 4 * (t >> 1) + (lane >> 2);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const double* ph = p + h * Q4S_EL + 104 * q;
      const double f0 = ph[0], f1 = ph[8], f2 = ph[16];
      dmma884(acc[h][0], f0, f0);
      dmma884(acc[h][1], f0, f1);
      dmma884(acc[h][2], f0, f2);
      dmma884(acc[h][3], f1, f1);
      dmma884(acc[h][4], f1, f2);
      dmma884(acc[h][5], f2, f2);
    }
  }
}
// accumulator fragments -> the full symmetric K_e (24 x 24, row stride kQ4KLd) over the element's dead strips.  Lane l
// holds K[8I + l/4][8J + 2 (l%4) + {0, 1}] of tile (I, J); off-diagonal tiles are mirrored.
__device__ __forceinline__ void q4_stage_k(double* wbase, int lane, const double (&acc)[2][6][2]) {
  constexpr int LD = EmitRuns::kQ4KLd;
  const int t = lane & 3, g = lane >> 2;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    double* K = wbase + h * Q4S_EL;
#pragma unroll
    for (int ti = 0, k = 0; ti < 3; ++ti)
#pragma unroll
      for (int tj = ti; tj < 3; ++tj, ++k) {
        *reinterpret_cast<double2*>(K + (8 * ti + g) * LD + 8 * tj + 2 * t) = make_double2(acc[h][k][0], acc[h][k][1]);
        if (ti != tj) {
          K[(8 * tj + 2 * t) * LD + 8 * ti + g] = acc[h][k][0];
          K[(8 * tj + 2 * t + 1) * LD + 8 * ti + g] = acc[h][k][1];
        }
      }
  }
}

__device__ __forceinline__ void q4_async_normal(const ShellArgs& P, double* slot, bool on, int node) {
  if (on) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(slot);
    const double4* src = P.nrm + node;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + 16), "l"(reinterpret_cast<const char*>(src) + 16) : "memory");
  }
}

// per warp: strips of its two elements (later their staged matrices) + node addressing ncol[8][8] + block run offsets
// pr[32][2] (ints) + 8 nodal normals
constexpr int Q4_WARP_DBL = 2 * Q4S_EL + 32 + 32 + 8 * 4;

// Stages: (1) setup, roles (integration point g4, node jn) per half-warp: strips into shared memory; (2) product on the
// DMMA path, whole warp, both elements; rules with more than 4 points repeat (1)-(2) in chunks of 4 points with the
// accumulators carried; (3) accumulators -> staged K_e; (4) drilling stiffness on the staged matrix (lanes 0..3 of each
// half-warp = the element's nodes); (5) emission from the staged matrix.
template <bool COMP, bool CHUNKED, class Emit>
__global__ void __launch_bounds__(32 * FS_Q4_WPB, FS_Q4_MINB) k_q4_stiffness(ShellArgs P, Emit emit) {
  extern __shared__ double smem[];
  constexpr int WARP_DBL = Q4_WARP_DBL;
  constexpr int LD = EmitRuns::kQ4KLd;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int half = lane >> 4, l16 = lane & 15;
  double* wbase = smem + (size_t)wib * WARP_DBL;
  double* sb_ = wbase + half * Q4S_EL;
  int* ncol = reinterpret_cast<int*>(wbase + 2 * Q4S_EL);
  int* pr = ncol + 64;
  double* nslot = wbase + 2 * Q4S_EL + 64 + (half * 4 + (l16 & 3)) * 4;  // normal of node (l16 & 3) of this half's element
  const int64_t e = ((int64_t)blockIdx.x * (blockDim.x >> 5) + wib) * 2 + half;
  const bool active = e < P.nelem;
  const int g4 = l16 >> 2, jn = l16 & 3;  // setup role
  const int bi = l16 >> 2, bj = l16 & 3;  // block role (addressing, non-cooperative emitters)
  const unsigned full = 0xffffffffu;

  V3 X[4];
  int nn[4] = {0, 0, 0, 0};
  double4 nvown = make_double4(0, 0, 1, 0);
  double hq = 0.0;
  const double* gd = nullptr;
  if (active) {
    const int32_t* cn = P.conn + e * 4;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      nn[a] = __ldg(cn + a);
      X[a] = ld3(P.xyz, nn[a]);
    }
    nvown = ldg4(P.nrm + (jn == 0 ? nn[0] : (jn == 1 ? nn[1] : (jn == 2 ? nn[2] : nn[3]))));
    // quirk: "diameter" = max distance from node 1 (src/FEMMShellQ4RSModule.jl:861-870)
    double md = 0.0;
    for (int a = 1; a < 4; ++a) {
      const V3 d = X[a] - X[0];
      md = fmax(md, dot(d, d));
    }
    hq = md;  // h^2: only the square enters the stabilisation factor
    if (COMP) gd = P.gdata + (size_t)__ldg(P.gof + e) * 34;
  }
  const int nbi = bi == 0 ? nn[0] : (bi == 1 ? nn[1] : (bi == 2 ? nn[2] : nn[3]));
  const int nbj = bj == 0 ? nn[0] : (bj == 1 ? nn[1] : (bj == 2 ? nn[2] : nn[3]));
  typename Emit::Cols ecols;
  typename Emit::Rows erows;
  // addressing data and the nodal normals of the drilling stage: requested first (cp.async into their own areas), the
  // loads overlap every stage up to the emission
  if constexpr (Emit::kCoop) {
    emit.q4_async_addr(ncol, pr, lane, active, nbj, e);  // lanes l16 < 4: bj = l16, nbj = node l16
  } else if (active) {
    ecols = emit.cols(nbj);
    erows = emit.rows(e, bi, bj, nbi);
  }
  q4_async_normal(P, nslot, active && l16 < 4, nbj);
  double acc[2][6][2];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[h][k][0] = acc[h][k][1] = 0.0;
  const int npts = P.rule.npts;
  if constexpr (!CHUNKED) {
    q4_setup_pass<COMP>(P, active && g4 < npts, e, g4, g4, jn, X, nvown, hq, gd, sb_);
    __syncwarp();
    q4_mma_product(wbase, lane, acc);
  } else {
    for (int chunk = 0; chunk * 4 < npts; ++chunk) {
      const int gp = chunk * 4 + g4;
      q4_setup_pass<COMP>(P, active && gp < npts, e, gp, g4, jn, X, nvown, hq, gd, sb_);
      __syncwarp();
      q4_mma_product(wbase, lane, acc);
      __syncwarp();
    }
  }
  __syncwarp();  // every lane is done with the strips
  q4_stage_k(wbase, lane, acc);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();

  // drilling stiffness (src/FEMMShellQ4RSModule.jl:807-859) on the rotational diagonal blocks of the staged matrix:
  // lane l16 = k < 4 of each half-warp handles node k
  double* Ke = wbase + half * Q4S_EL;
  double tang = 0.0;
  int ok = 0;
  double nvec[3] = {0, 0, 0};
  double* krr = Ke + (6 * (l16 & 3) + 3) * LD + 6 * (l16 & 3) + 3;
  if (active && l16 < 4) {
    const double4 n4 = *reinterpret_cast<const double4*>(nslot);
    const double nl2 = n4.x * n4.x + n4.y * n4.y + n4.z * n4.z;
    if (n4.w != 0.0 && nl2 != 0.0) {
      ok = 1;
      nvec[0] = n4.x;
      nvec[1] = n4.y;
      nvec[2] = n4.z;
      const double inl = fs_rsqrt(nl2);
      const double nh[3] = {n4.x * inl, n4.y * inl, n4.z * inl};
      // tr(P Krr P) with P = I - n n' (idempotent): tr(Krr P) = tr(Krr) - n' Krr n
      double tr = krr[0] + krr[LD + 1] + krr[2 * LD + 2];
#pragma unroll
      for (int r = 0; r < 3; ++r) tr -= nh[r] * (krr[r * LD] * nh[0] + krr[r * LD + 1] * nh[1] + krr[r * LD + 2] * nh[2]);
      tang = fmax(0.0, tr / 2.0);
    }
  }
  double tsum = tang;
  int cnt = ok;
  tsum += __shfl_xor_sync(full, tsum, 1);
  cnt += __shfl_xor_sync(full, cnt, 1);
  tsum += __shfl_xor_sync(full, tsum, 2);
  cnt += __shfl_xor_sync(full, cnt, 2);
  if (P.drill != 0.0 && cnt > 0 && ok) {
    const double kavg = tsum * (cnt == 4 ? 0.25 : fs_rcp((double)cnt)) * P.drill;
    if (kavg != 0.0) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) krr[r * LD + cc] += kavg * (nvec[r] * nvec[cc]);
    }
  }
  __syncwarp();
  if constexpr (Emit::kCoop) {
    emit.q4_emit_k(wbase, Q4S_EL, ncol, pr, lane);
  } else if constexpr (Emit::kDenseCoop) {
    emit.q4_dense_from_k(wbase, Q4S_EL, LD, e - half, P.nelem, lane);
  } else {
    if (!active) return;
    double a[6][6];
    const double* kb = Ke + (6 * bi) * LD + 6 * bj;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) a[r][cc] = kb[r * LD + cc];
    emit.block(BlockRef{e, bi, bj}, ecols, erows, a);
  }
}

// =====================================================================================
// lumped shell mass: thread per (element, node)
// =====================================================================================
// mode: 0 = matrix via diagslot, 1 = vector over dofs [0, limit)
template <int NNPE, bool COMP>
__global__ void k_shell_mass(ShellArgs P, const int32_t* __restrict__ dof, const int32_t* __restrict__ diagslot, double* out,
                             int mode, int64_t limit) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= P.nelem * NNPE) return;
  const int64_t e = t / NNPE;
  const int k = (int)(t % NNPE);
  const int32_t* cn = P.conn + e * NNPE;
  double tmass = 0.0, rmass = 0.0;
  double md = 0.0, mi = 0.0;
  if (COMP) {
    const double* gd = P.gdata + (size_t)P.gof[e] * 34;
    md = gd[32];
    mi = gd[33];
  }
  if (NNPE == 3) {
    const T3Geom g = t3_geometry(ld3(P.xyz, cn[0]), ld3(P.xyz, cn[1]), ld3(P.xyz, cn[2]));
    if (COMP) {
      tmass = md * (g.Ae / 3);
      rmass = mi * (g.Ae / 3);
    } else {
      const double th = P.nthick == 1 ? P.thick[0] : P.thick[e];
      tmass = P.rho * (th * g.Ae) / 3;
      rmass = P.rho * (th * th * th / 12 * g.Ae) / 3;
    }
  } else {
    V3 X[4] = {ld3(P.xyz, cn[0]), ld3(P.xyz, cn[1]), ld3(P.xyz, cn[2]), ld3(P.xyz, cn[3 % NNPE])};
    for (int gp = 0; gp < P.rule.npts; ++gp) {
      double dN[4][2];
      q4_shape_derivs(P.rule.xi[gp], P.rule.eta[gp], dN);
      V3 t1 = v3(0, 0, 0), t2 = v3(0, 0, 0);
      for (int a = 0; a < 4; ++a) {
        t1 = t1 + dN[a][0] * X[a];
        t2 = t2 + dN[a][1] * X[a];
      }
      const double Jac = norm(cross(t1, t2));
      if (COMP) {
        tmass += md * Jac * P.rule.w[gp];
        rmass += mi * Jac * P.rule.w[gp];
      } else {
        const double th = P.nthick == 1 ? P.thick[0] : (P.nthick == P.nelem ? P.thick[e] : P.thick[e * P.rule.npts + gp]);
        tmass += P.rho * th * Jac * P.rule.w[gp];
        rmass += P.rho * th * th * th / 12 * Jac * P.rule.w[gp];
      }
    }
    tmass = tmass / 4;
    rmass = rmass / 4;
  }
  const int32_t* dn = dof + (int64_t)cn[k] * 6;
  for (int d = 0; d < 6; ++d) {
    const double v = d < 3 ? tmass : rmass;
    const int32_t dd = dn[d];
    if (mode == 0) {
      const int s = diagslot[dd];
      if (s >= 0) atomicAdd(out + s, v);
    } else if (dd < limit) {
      atomicAdd(out + dd, v);
    }
  }
}
// dense element mass matrices (parity/debug)
__global__ void k_shell_mass_dense_from_diag(const double* __restrict__ dvals /*[nelem*nnpe][2]*/, int nnpe, int64_t nelem,
                                             double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int n = 6 * nnpe;
  if (i >= nelem * n) return;
  const int64_t e = i / n;
  const int q = (int)(i % n);
  const double* dv = dvals + (e * nnpe + q / 6) * 2;
  out[e * n * n + (int64_t)q * n + q] = (q % 6) < 3 ? dv[0] : dv[1];
}
template <int NNPE, bool COMP>
__global__ void k_shell_mass_pairs(ShellArgs P, double* __restrict__ dvals) {
  // (tmass, rmass) per element node, for the dense debug output
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= P.nelem * NNPE) return;
  // reuse k_shell_mass arithmetic through a tiny trick: vector mode into a 2-entry scratch is not
  // possible, so the arithmetic is restated compactly here.
  const int64_t e = t / NNPE;
  const int32_t* cn = P.conn + e * NNPE;
  double tmass = 0.0, rmass = 0.0, md = 0.0, mi = 0.0;
  if (COMP) {
    const double* gd = P.gdata + (size_t)P.gof[e] * 34;
    md = gd[32];
    mi = gd[33];
  }
  if (NNPE == 3) {
    const T3Geom g = t3_geometry(ld3(P.xyz, cn[0]), ld3(P.xyz, cn[1]), ld3(P.xyz, cn[2]));
    if (COMP) {
      tmass = md * (g.Ae / 3);
      rmass = mi * (g.Ae / 3);
    } else {
      const double th = P.nthick == 1 ? P.thick[0] : P.thick[e];
      tmass = P.rho * (th * g.Ae) / 3;
      rmass = P.rho * (th * th * th / 12 * g.Ae) / 3;
    }
  } else {
    V3 X[4] = {ld3(P.xyz, cn[0]), ld3(P.xyz, cn[1]), ld3(P.xyz, cn[2]), ld3(P.xyz, cn[3 % NNPE])};
    for (int gp = 0; gp < P.rule.npts; ++gp) {
      double dN[4][2];
      q4_shape_derivs(P.rule.xi[gp], P.rule.eta[gp], dN);
      V3 t1 = v3(0, 0, 0), t2 = v3(0, 0, 0);
      for (int a = 0; a < 4; ++a) {
        t1 = t1 + dN[a][0] * X[a];
        t2 = t2 + dN[a][1] * X[a];
      }
      const double Jac = norm(cross(t1, t2));
      if (COMP) {
        tmass += md * Jac * P.rule.w[gp];
        rmass += mi * Jac * P.rule.w[gp];
      } else {
        const double th = P.nthick == 1 ? P.thick[0] : (P.nthick == P.nelem ? P.thick[e] : P.thick[e * P.rule.npts + gp]);
        tmass += P.rho * th * Jac * P.rule.w[gp];
        rmass += P.rho * th * th * th / 12 * Jac * P.rule.w[gp];
      }
    }
    tmass = tmass / 4;
    rmass = rmass / 4;
  }
  dvals[t * 2] = tmass;
  dvals[t * 2 + 1] = rmass;
}

// =====================================================================================
// inspectintegpoints, batched: stress resultants per element (T3) / per integration point (Q4)
// in the output csys (src/FEMMShellT3FFModule.jl:850-962, src/FEMMShellQ4RSModule.jl:1061-1170).
// strains = (B T_ae T_ga) u_e, i.e. the unfolded global-dof strips applied to the nodal dofs.
// quant: 1 bending moment, 2 transverse shear, 3 membrane force.  out[(e*npts + gp)*3 + k].
// =====================================================================================
__device__ __forceinline__ void resultant_out(const ShellArgs& P, int quant, const double (&st)[8], double t, double stab,
                                              const Triad& E, const double* ocs, double* out) {
  double m = 1.0, n = 0.0;
  double cs[9];
  if (ocs) {
    for (int q = 0; q < 9; ++q) cs[q] = ocs[q];
  } else {  // default: the material csys = element triad itself (isoparametric!)
    cs[0] = E.e1.x; cs[3] = E.e1.y; cs[6] = E.e1.z;
    cs[1] = E.e2.x; cs[4] = E.e2.y; cs[7] = E.e2.z;
    cs[2] = E.e3.x; cs[5] = E.e3.y; cs[8] = E.e3.z;
  }
  layup_angle(E, cs, m, n);
  if (quant == 2) {
    const double f0 = t * stab * (P.Dt[0] * st[6] + P.Dt[1] * st[7]);
    const double f1 = t * stab * (P.Dt[2] * st[6] + P.Dt[3] * st[7]);
    // fo = o2' * frc, o2 = [m n; -n m]
    out[0] = m * f0 - n * f1;
    out[1] = n * f0 + m * f1;
    out[2] = 0.0;
    return;
  }
  const int o = quant == 1 ? 3 : 0;
  const double c = quant == 1 ? (t * t * t) * (1.0 / 12.0) : t;
  double v[3];
  for (int i = 0; i < 3; ++i) v[i] = c * (P.Dps[i * 3] * st[o] + P.Dps[i * 3 + 1] * st[o + 1] + P.Dps[i * 3 + 2] * st[o + 2]);
  // mo = o2' [v0 v2; v2 v1] o2
  const double M00 = v[0], M11 = v[1], M01 = v[2];
  const double a00 = m * M00 - n * M01, a01 = m * M01 - n * M11;  // (o2' M) row 0
  const double a10 = n * M00 + m * M01, a11 = n * M01 + m * M11;  // row 1
  out[0] = a00 * m - a01 * n;
  out[1] = a10 * n + a11 * m;
  out[2] = a00 * n + a01 * m;
}

// Laminated shells (src/FEMMShellT3FFCompModule.jl:892-937, src/FEMMShellQ4RSCompModule.jl:1164-1203):
// mom = sB eps + sD kappa, frc = sA eps + sB kappa, shear = stab_fun sH gamma with the layup group's A, B, D, H
// rotated into the element frame; the default output csys is the layup's.  lcs / ocs: row-major 3x3.
__device__ __forceinline__ void resultant_out_laminate(int quant, const double (&st)[8], double stab, const Triad& E,
                                                       const double* gd, const double* lcs, const double* ocs, double* out) {
  double lm, ln, m, n;
  layup_angle(E, lcs, lm, ln);
  layup_angle(E, ocs ? ocs : lcs, m, n);
  if (quant == 2) {
    double sH[2][2];
    rotate_ts(gd + 27, lm, ln, sH);
    const double f0 = stab * (sH[0][0] * st[6] + sH[0][1] * st[7]);
    const double f1 = stab * (sH[1][0] * st[6] + sH[1][1] * st[7]);
    out[0] = m * f0 - n * f1;
    out[1] = n * f0 + m * f1;
    out[2] = 0.0;
    return;
  }
  double sX[3][3], sB[3][3];
  rotate_ps(gd + 9, lm, ln, sB);
  rotate_ps(quant == 1 ? gd + 18 : gd, lm, ln, sX);  // D pairs with the curvatures, A with the membrane strains
  const int ox = quant == 1 ? 3 : 0, ob = quant == 1 ? 0 : 3;
  double v[3];
  for (int i = 0; i < 3; ++i)
    v[i] = sX[i][0] * st[ox] + sX[i][1] * st[ox + 1] + sX[i][2] * st[ox + 2] + sB[i][0] * st[ob] + sB[i][1] * st[ob + 1] +
           sB[i][2] * st[ob + 2];
  const double M00 = v[0], M11 = v[1], M01 = v[2];
  const double a00 = m * M00 - n * M01, a01 = m * M01 - n * M11;
  const double a10 = n * M00 + m * M01, a11 = n * M01 + m * M11;
  out[0] = a00 * m - a01 * n;
  out[1] = a10 * n + a11 * m;
  out[2] = a00 * n + a01 * m;
}

template <bool COMP>
__global__ void k_t3_resultants(ShellArgs P, const double* __restrict__ u, int quant, const double* __restrict__ ocs, int64_t nocs,
                                double* __restrict__ out) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= P.nelem) return;
  const int32_t* cn = P.conn + e * 3;
  const int nn[3] = {cn[0], cn[1], cn[2]};
  const T3Geom g = t3_geometry(ld3(P.xyz, nn[0]), ld3(P.xyz, nn[1]), ld3(P.xyz, nn[2]));
  double st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double P1[5][3], P2[5][3];
  for (int r = 0; r < 5; ++r)
    for (int k = 0; k < 3; ++k) P1[r][k] = P2[r][k] = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int l = 0; l < 3; ++l) {
      const double4 nv = ldg4(P.nrm + nn[l]);
      const M3 A = nodal_triad(g.E, v3(nv.x, nv.y, nv.z), nv.w != 0.0);
      double bs[2][3];
      t3_bs_node(g, l, -1, bs);
      if (pass == 0) {
        double p1[5][3], p2[5][3];
        node_coupling_contrib(A, g.gN[l][0], g.gN[l][1], bs, p1, p2);
        for (int r = 0; r < 5; ++r)
          for (int k = 0; k < 3; ++k) {
            P1[r][k] += p1[r][k];
            P2[r][k] += p2[r][k];
          }
      } else {
        double bg[8][6];
        node_strip(g.E, A, g.gN[l][0], g.gN[l][1], bs, P1, P2, bg);
        const double* ul = u + (int64_t)nn[l] * 6;
        for (int s = 0; s < 8; ++s)
          for (int c = 0; c < 6; ++c) st[s] += bg[s][c] * ul[c];
      }
    }
  }
  const double h = sqrt(2 * g.Ae);
  const double* oc = nocs == 0 ? nullptr : ocs + (nocs == 1 ? 0 : e * 9);
  if (COMP) {
    const double* gd = P.gdata + (size_t)P.gof[e] * 34;
    const double t = gd[31];
    const double stab = P.nstab ? P.stabf[e] : t * t / (t * t + P.alpha * h * h);
    resultant_out_laminate(quant, st, stab, g.E, gd, P.cs + (P.ncs == 1 ? 0 : e * 9), oc, out + e * 3);
  } else {
    const double t = P.nthick == 1 ? P.thick[0] : P.thick[e];
    const double stab = P.nstab ? P.stabf[e] : t * t / (t * t + P.alpha * h * h);
    resultant_out(P, quant, st, t, stab, g.E, oc, out + e * 3);
  }
}

template <bool COMP>
__global__ void k_q4_resultants(ShellArgs P, const double* __restrict__ u, int quant, const double* __restrict__ ocs, int64_t nocs,
                                double* __restrict__ out) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int npts = P.rule.npts;
  if (tid >= P.nelem * npts) return;
  const int64_t e = tid / npts;
  const int gp = (int)(tid % npts);
  const int32_t* cn = P.conn + e * 4;
  const int nn[4] = {cn[0], cn[1], cn[2], cn[3]};
  V3 X[4];
  for (int a = 0; a < 4; ++a) X[a] = ld3(P.xyz, nn[a]);
  double md = 0.0;
  for (int a = 1; a < 4; ++a) {
    const V3 d = X[a] - X[0];
    md = fmax(md, dot(d, d));
  }
  const double hq = sqrt(md);
  const double xi = P.rule.xi[gp], eta = P.rule.eta[gp];
  const Q4Geom g = q4_geometry(X, xi, eta);
  double st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double P1[5][3], P2[5][3];
  for (int r = 0; r < 5; ++r)
    for (int k = 0; k < 3; ++k) P1[r][k] = P2[r][k] = 0.0;
  // FEMMShellQ4RSComp as written applies T = T_ae T_ga twice: `edisp_e = T edisp` (src/FEMMShellQ4RSCompModule.jl:1163)
  // is then multiplied by B matrices that already carry T (:1181-1184, :1198) -- SURVEY App. B.9, replicated.
  // ut[l] = (T u)[6l .. 6l+5] from the definitions of T_ga (src/FEMMShellT3FFModule.jl:398-419) and T_ae (:421-463).
  double ut[4][6];
  if (COMP) {
    M3 Ak[4];
    double un[4][6];
    for (int l = 0; l < 4; ++l) {
      const double4 nv = ldg4(P.nrm + nn[l]);
      Ak[l] = nodal_triad(g.E, v3(nv.x, nv.y, nv.z), nv.w != 0.0);
      const M3 G = global_to_nodal(Ak[l], g.E);
      const double* ul = u + (int64_t)nn[l] * 6;
      for (int i = 0; i < 3; ++i) {
        un[l][i] = G.a[i][0] * ul[0] + G.a[i][1] * ul[1] + G.a[i][2] * ul[2];
        un[l][3 + i] = G.a[i][0] * ul[3] + G.a[i][1] * ul[4] + G.a[i][2] * ul[5];
      }
    }
    for (int l = 0; l < 4; ++l) {
      const M3& A = Ak[l];
      for (int i = 0; i < 3; ++i) ut[l][i] = A.a[i][0] * un[l][0] + A.a[i][1] * un[l][1] + A.a[i][2] * un[l][2];
      double R[2][2];
      node_R(A, R);
      double cpl = 0.0;  // sum_j sum_k 1/2 (A[1][k] gx_j - A[0][k] gy_j) un_j[k]
      for (int j = 0; j < 4; ++j)
        for (int k = 0; k < 3; ++k) cpl += 0.5 * (A.a[1][k] * g.gN[j][0] - A.a[0][k] * g.gN[j][1]) * un[j][k];
      const double ia = 1.0 / A.a[2][2];
      ut[l][3] = R[0][0] * un[l][3] + R[0][1] * un[l][4] + ia * A.a[0][2] * cpl;
      ut[l][4] = R[1][0] * un[l][3] + R[1][1] * un[l][4] + ia * A.a[1][2] * cpl;
      ut[l][5] = 0.0;
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
    for (int l = 0; l < 4; ++l) {
      const double4 nv = ldg4(P.nrm + nn[l]);
      const M3 A = nodal_triad(g.E, v3(nv.x, nv.y, nv.z), nv.w != 0.0);
      double bs[2][3];
      q4_mitc_bs_node(g, xi, eta, l, bs);
      if (pass == 0) {
        double p1[5][3], p2[5][3];
        node_coupling_contrib(A, g.gN[l][0], g.gN[l][1], bs, p1, p2);
        for (int r = 0; r < 5; ++r)
          for (int k = 0; k < 3; ++k) {
            P1[r][k] += p1[r][k];
            P2[r][k] += p2[r][k];
          }
      } else {
        double bg[8][6];
        node_strip(g.E, A, g.gN[l][0], g.gN[l][1], bs, P1, P2, bg);
        const double* ul = COMP ? ut[l] : u + (int64_t)nn[l] * 6;
        for (int s = 0; s < 8; ++s)
          for (int c = 0; c < 6; ++c) st[s] += bg[s][c] * ul[c];
      }
    }
  }
  const double* oc = nocs == 0 ? nullptr : ocs + (nocs == 1 ? 0 : (nocs == P.nelem ? e : tid) * 9);
  if (COMP) {
    const double* gd = P.gdata + (size_t)P.gof[e] * 34;
    const double t = gd[31];
    const double stab = P.nstab ? P.stabf[e] : t * t / (t * t + P.alpha * hq * hq);
    const int64_t ci = P.ncs == 1 ? 0 : (P.ncs == P.nelem ? e : tid);
    resultant_out_laminate(quant, st, stab, g.E, gd, P.cs + ci * 9, oc, out + tid * 3);
  } else {
    const double t = P.nthick == 1 ? P.thick[0] : (P.nthick == P.nelem ? P.thick[e] : P.thick[e * npts + gp]);
    const double stab = P.nstab ? P.stabf[e] : t * t / (t * t + P.alpha * hq * hq);
    resultant_out(P, quant, st, t, stab, g.E, oc, out + tid * 3);
  }
}

__global__ void k_element_sizes(const int32_t* __restrict__ conn, const double4* __restrict__ xyz, int nnpe, int64_t nelem,
                                double* __restrict__ h) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nelem) return;
  const int32_t* cn = conn + e * nnpe;
  if (nnpe == 3) {
    const T3Geom g = t3_geometry(ld3(xyz, cn[0]), ld3(xyz, cn[1]), ld3(xyz, cn[2]));
    h[e] = sqrt(2 * g.Ae);
  } else if (nnpe == 4) {
    const V3 x0 = ld3(xyz, cn[0]);
    double md = 0.0;
    for (int a = 1; a < 4; ++a) {
      const V3 d = ld3(xyz, cn[a]) - x0;
      md = fmax(md, dot(d, d));
    }
    h[e] = sqrt(md);
  } else {
    h[e] = norm(ld3(xyz, cn[1]) - ld3(xyz, cn[0]));
  }
}

// =====================================================================================
// corotational beam
// =====================================================================================
struct BeamArgs {
  const double* v1;     // [nnodes][6] (gyroscopic only)
  const double* force;  // distributed load: 3 values or [nelem][3]
  int64_t nforce;
  const int32_t* conn;
  const double4* xyz;
  const double4* u1;
  const double* R1;
  const double* sec;
  int64_t nelem;
  double E, G, rho;
  int mass_type;
};
__device__ __forceinline__ void beam_load(const BeamArgs& P, int64_t e, BeamSec& s, BeamKin& k) {
  const int32_t* cn = P.conn + e * 2;
  const int nI = cn[0], nJ = cn[1];
  const double* sp = P.sec + e * 10;
  s.A = sp[0];
  s.I1 = sp[1];
  s.I2 = sp[2];
  s.I3 = sp[3];
  s.J = sp[4];
  s.A2s = sp[5];
  s.A3s = sp[6];
  s.x1x2 = v3(sp[7], sp[8], sp[9]);
  double RI[9], RJ[9];
  for (int q = 0; q < 9; ++q) {
    RI[q] = P.R1[(int64_t)nI * 9 + q];
    RJ[q] = P.R1[(int64_t)nJ * 9 + q];
  }
  k = beam_kinematics(ld3(P.xyz, nI), ld3(P.xyz, nJ), ld3(P.u1, nI), ld3(P.u1, nJ), RI, RJ, s.x1x2);
}
// op: 0 stiffness, 1 mass, 2 geostiffness.  4 lanes per element: lane = block (I, J).
template <class Emit>
__global__ void k_beam_matrix(BeamArgs P, int op, Emit emit) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (tid >= P.nelem * 4) return;
  const int64_t e = tid >> 2;
  const int bI = (int)(tid & 3) >> 1, bJ = (int)(tid & 1);
  BeamSec s;
  BeamKin k;
  beam_load(P, e, s, k);
  double Kl[6][6];
  if (op == 0) {
    double DN[6], aN[6][12];
    beam_natural_stiffness(P.E, P.G, s, k.L1, DN);
    beam_aN(k.L1, aN);
    for (int p = 0; p < 6; ++p)
      for (int q = 0; q < 6; ++q) {
        double v = 0.0;
        for (int m = 0; m < 6; ++m) v += aN[m][bI * 6 + p] * DN[m] * aN[m][bJ * 6 + q];
        Kl[p][q] = v;
      }
  } else if (op == 2) {
    double DN[6], PN[6], S[12][12];
    beam_natural_stiffness(P.E, P.G, s, k.L1, DN);
    for (int m = 0; m < 6; ++m) PN[m] = DN[m] * k.dN[m];
    beam_local_geo(PN, k.L1, S);
    for (int p = 0; p < 6; ++p)
      for (int q = 0; q < 6; ++q) Kl[p][q] = S[bI * 6 + p][bJ * 6 + q];
  } else {  // op 1 (mass) and op 3 (gyroscopic) start from the mass matrix
    double M[12][12];
    beam_local_mass(s, P.rho, k.L0, P.mass_type, M);
    for (int p = 0; p < 6; ++p)
      for (int q = 0; q < 6; ++q) Kl[p][q] = M[bI * 6 + p][bJ * 6 + q];
  }
  // global block = Tb Kl Tb', Tb = blkdiag(Ft, Ft), Ft[r][a] = component r of e_a
  const double F[3][3] = {{k.Ft.e1.x, k.Ft.e2.x, k.Ft.e3.x}, {k.Ft.e1.y, k.Ft.e2.y, k.Ft.e3.y}, {k.Ft.e1.z, k.Ft.e2.z, k.Ft.e3.z}};
  double Kg[6][6];
  for (int sp = 0; sp < 2; ++sp)
    for (int sq = 0; sq < 2; ++sq) {
      double tmp[3][3];
      for (int a = 0; a < 3; ++a)
        for (int cc = 0; cc < 3; ++cc)
          tmp[a][cc] = Kl[sp * 3 + a][sq * 3 + 0] * F[cc][0] + Kl[sp * 3 + a][sq * 3 + 1] * F[cc][1] + Kl[sp * 3 + a][sq * 3 + 2] * F[cc][2];
      for (int cc = 0; cc < 3; ++cc)
        for (int r = 0; r < 3; ++r) {
          Kg[sp * 3 + r][sq * 3 + cc] = F[r][0] * tmp[0][cc] + F[r][1] * tmp[1][cc] + F[r][2] * tmp[2][cc];
        }
    }
  const int32_t* cn = P.conn + e * 2;
  if (op == 3) {
    // gyroscopic: Ge = Omega~ M - M Omega~ with the block-diagonal skew matrix of the element spin
    // (src/FEMMCorotBeamModule.jl:934-947); Kg holds the global mass block here
    const double* vI = P.v1 + (int64_t)cn[0] * 6;
    const double* vJ = P.v1 + (int64_t)cn[1] * 6;
    // evel1f[n, a] = sum_k evel1[n, k] Ft[k, a]
    auto loc = [&](const double* v, int off, int a) { return v[off] * F[0][a] + v[off + 1] * F[1][a] + v[off + 2] * F[2][a]; };
    const double w1 = (loc(vI, 3, 0) + loc(vJ, 3, 0)) / 2;
    const double w2 = (loc(vI, 0, 2) - loc(vJ, 0, 2)) / k.L1;
    const double w3 = (loc(vJ, 0, 1) - loc(vI, 0, 1)) / k.L1;
    const double Om[3] = {w1 * F[0][0] + w2 * F[0][1] + w3 * F[0][2], w1 * F[1][0] + w2 * F[1][1] + w3 * F[1][2],
                          w1 * F[2][0] + w2 * F[2][1] + w3 * F[2][2]};
    const double OS[3][3] = {{0, -Om[2], Om[1]}, {Om[2], 0, -Om[0]}, {-Om[1], Om[0], 0}};
    double Gg[6][6];
    for (int sp = 0; sp < 2; ++sp)
      for (int sq = 0; sq < 2; ++sq)
        for (int r = 0; r < 3; ++r)
          for (int cc = 0; cc < 3; ++cc) {
            double v = 0.0;
            for (int m = 0; m < 3; ++m) v += OS[r][m] * Kg[sp * 3 + m][sq * 3 + cc] - Kg[sp * 3 + r][sq * 3 + m] * OS[m][cc];
            Gg[sp * 3 + r][sq * 3 + cc] = v;
          }
    emit.block(BlockRef{e, bI, bJ}, emit.cols(cn[bJ]), emit.rows(e, bI, bJ, cn[bI]), Gg);
    return;
  }
  emit.block(BlockRef{e, bI, bJ}, emit.cols(cn[bJ]), emit.rows(e, bI, bJ, cn[bI]), Kg);
}
// Fast-path variant: ONE lane per element forms the four 6x6 blocks one after the other with compile-time block
// indices (the sparse local matrices fold into registers, the kinematics are evaluated once instead of four times),
// and every block leaves through the warp-cooperative emission (coalesced RED rows).
template <int OP, int BI, int BJ>
__device__ __forceinline__ void beam_block_global(const BeamArgs& P, const BeamSec& s, const BeamKin& k, const double (&F)[3][3],
                                                  const double (&Om)[3], double (&Kg)[6][6]) {
  double Kl[6][6];
  if (OP == 0) {
    double DN[6], aN[6][12];
    beam_natural_stiffness(P.E, P.G, s, k.L1, DN);
    beam_aN(k.L1, aN);
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        double v = 0.0;
#pragma unroll
        for (int m = 0; m < 6; ++m) v += aN[m][BI * 6 + p] * DN[m] * aN[m][BJ * 6 + q];
        Kl[p][q] = v;
      }
  } else if (OP == 2) {
    double DN[6], PN[6], S[12][12];
    beam_natural_stiffness(P.E, P.G, s, k.L1, DN);
#pragma unroll
    for (int m = 0; m < 6; ++m) PN[m] = DN[m] * k.dN[m];
    beam_local_geo(PN, k.L1, S);
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int q = 0; q < 6; ++q) Kl[p][q] = S[BI * 6 + p][BJ * 6 + q];
  } else {
    double M[12][12];
    beam_local_mass(s, P.rho, k.L0, P.mass_type, M);
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int q = 0; q < 6; ++q) Kl[p][q] = M[BI * 6 + p][BJ * 6 + q];
  }
  double T[6][6];
#pragma unroll
  for (int sp = 0; sp < 2; ++sp)
#pragma unroll
    for (int sq = 0; sq < 2; ++sq) {
      double tmp[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
          tmp[a][cc] = Kl[sp * 3 + a][sq * 3 + 0] * F[cc][0] + Kl[sp * 3 + a][sq * 3 + 1] * F[cc][1] + Kl[sp * 3 + a][sq * 3 + 2] * F[cc][2];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int r = 0; r < 3; ++r) T[sp * 3 + r][sq * 3 + cc] = F[r][0] * tmp[0][cc] + F[r][1] * tmp[1][cc] + F[r][2] * tmp[2][cc];
    }
  if (OP == 3) {
    const double OS[3][3] = {{0, -Om[2], Om[1]}, {Om[2], 0, -Om[0]}, {-Om[1], Om[0], 0}};
#pragma unroll
    for (int sp = 0; sp < 2; ++sp)
#pragma unroll
      for (int sq = 0; sq < 2; ++sq)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            double v = 0.0;
#pragma unroll
            for (int m = 0; m < 3; ++m) v += OS[r][m] * T[sp * 3 + m][sq * 3 + cc] - T[sp * 3 + r][sq * 3 + m] * OS[m][cc];
            Kg[sp * 3 + r][sq * 3 + cc] = v;
          }
  } else {
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) Kg[r][cc] = T[r][cc];
  }
}
constexpr int BEAM_WARP_DBL = 32 * EmitRuns::kStageLd + (6 * 32) / 2 + 4 * (EmitRuns::kAddrInts / 2);
template <int OP>
__global__ void __launch_bounds__(128) k_beam_matrix_coop(BeamArgs P, EmitRuns emit) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* wbase = smem + (size_t)wib * BEAM_WARP_DBL;  // stage[32][kStageLd], rowp[6][32], 4 addressing areas
  int* addr0 = reinterpret_cast<int*>(wbase + 32 * EmitRuns::kStageLd + (6 * 32) / 2);
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool active = e < P.nelem;
  int cn[2] = {0, 0};
  if (active) {
    cn[0] = __ldg(P.conn + e * 2);
    cn[1] = __ldg(P.conn + e * 2 + 1);
  }
#pragma unroll
  for (int b = 0; b < 4; ++b)
    emit.async_addr(addr0 + b * EmitRuns::kAddrInts, lane, active, cn[b & 1], cn[b >> 1], e, b >> 1, b & 1);
  BeamSec s;
  BeamKin k;
  double F[3][3], Om[3] = {0, 0, 0};
  if (active) {
    beam_load(P, e, s, k);
    const double Fi[3][3] = {{k.Ft.e1.x, k.Ft.e2.x, k.Ft.e3.x}, {k.Ft.e1.y, k.Ft.e2.y, k.Ft.e3.y}, {k.Ft.e1.z, k.Ft.e2.z, k.Ft.e3.z}};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) F[r][cc] = Fi[r][cc];
    if (OP == 3) {
      // element spin (src/FEMMCorotBeamModule.jl:934-947)
      const double* vI = P.v1 + (int64_t)cn[0] * 6;
      const double* vJ = P.v1 + (int64_t)cn[1] * 6;
      auto loc = [&](const double* v, int off, int a) { return v[off] * F[0][a] + v[off + 1] * F[1][a] + v[off + 2] * F[2][a]; };
      const double w1 = (loc(vI, 3, 0) + loc(vJ, 3, 0)) / 2;
      const double w2 = (loc(vI, 0, 2) - loc(vJ, 0, 2)) / k.L1;
      const double w3 = (loc(vJ, 0, 1) - loc(vI, 0, 1)) / k.L1;
      Om[0] = w1 * F[0][0] + w2 * F[0][1] + w3 * F[0][2];
      Om[1] = w1 * F[1][0] + w2 * F[1][1] + w3 * F[1][2];
      Om[2] = w1 * F[2][0] + w2 * F[2][1] + w3 * F[2][2];
    }
  }
  double Kg[6][6];
#define BEAM_ROUND(BI, BJ)                                                                  \
  do {                                                                                      \
    if (active) {                                                                           \
      beam_block_global<OP, BI, BJ>(P, s, k, F, Om, Kg);                                     \
    } else {                                                                                \
      for (int r = 0; r < 6; ++r)                                                           \
        for (int cc = 0; cc < 6; ++cc) Kg[r][cc] = 0.0;                                     \
    }                                                                                       \
    emit.coop_emit_full(wbase, addr0 + (BI * 2 + BJ) * EmitRuns::kAddrInts, lane, Kg);      \
  } while (0)
  BEAM_ROUND(0, 0);
  BEAM_ROUND(0, 1);
  BEAM_ROUND(1, 0);
  BEAM_ROUND(1, 1);
#undef BEAM_ROUND
}
// restoring force: elvec = Te (-aN' DN dN)   (src/FEMMCorotBeamModule.jl:1132-1157)
// mode 0: restoring force; mode 1: consistent nodal loads of a uniform global force per unit length
// (distribloads_global, src/FEMMCorotBeamModule.jl:1186-1247)
__global__ void k_beam_restoring(BeamArgs P, const int32_t* __restrict__ dof, double* __restrict__ out, int64_t limit,
                                 double* __restrict__ elvec_out, int mode) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= P.nelem) return;
  BeamSec s;
  BeamKin k;
  beam_load(P, e, s, k);
  double LF[12];
  if (mode == 0) {
    double DN[6], aN[6][12];
    beam_natural_stiffness(P.E, P.G, s, k.L1, DN);
    beam_aN(k.L1, aN);
    for (int q = 0; q < 12; ++q) {
      double v = 0.0;
      for (int m = 0; m < 6; ++m) v += aN[m][q] * (DN[m] * k.dN[m]);
      LF[q] = -v;
    }
  } else {
    const double* fp = P.force + (P.nforce == 1 ? 0 : e * 3);
    const V3 fg = v3(fp[0], fp[1], fp[2]);
    const double l1 = dot(k.Ft.e1, fg), l2 = dot(k.Ft.e2, fg), l3 = dot(k.Ft.e3, fg), L0 = k.L0;
    LF[0] = LF[6] = l1 * L0 / 2;
    LF[1] = LF[7] = l2 * L0 / 2;
    LF[2] = LF[8] = l3 * L0 / 2;
    LF[3] = LF[9] = 0.0;
    LF[4] = -l3 * L0 * L0 / 12;
    LF[5] = l2 * L0 * L0 / 12;
    LF[10] = l3 * L0 * L0 / 12;
    LF[11] = -l2 * L0 * L0 / 12;
  }
  const double F[3][3] = {{k.Ft.e1.x, k.Ft.e2.x, k.Ft.e3.x}, {k.Ft.e1.y, k.Ft.e2.y, k.Ft.e3.y}, {k.Ft.e1.z, k.Ft.e2.z, k.Ft.e3.z}};
  const int32_t* cn = P.conn + e * 2;
  for (int b = 0; b < 4; ++b) {
    const int node = cn[b >> 1];
    for (int r = 0; r < 3; ++r) {
      const double v = F[r][0] * LF[b * 3] + F[r][1] * LF[b * 3 + 1] + F[r][2] * LF[b * 3 + 2];
      if (elvec_out) {
        elvec_out[e * 12 + b * 3 + r] = v;
      } else {
        const int32_t dd = dof[(int64_t)node * 6 + (b & 1) * 3 + r];
        if (dd < limit) atomicAdd(out + dd, v);
      }
    }
  }
}
// R <- exp(dtheta) R (src/RotUtilModule.jl:29-42); R row = column-major 3x3
__global__ void k_update_rotation(double* __restrict__ R, const double* __restrict__ dchi /*nnodes x 6 col-major*/,
                                  int64_t nnodes) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nnodes) return;
  const double ax = dchi[3 * nnodes + i], ay = dchi[4 * nnodes + i], az = dchi[5 * nnodes + i];
  const double na = sqrt(ax * ax + ay * ay + az * az);
  double Rd[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  if (na > 0.0) {
    const double nx = ax / na, ny = ay / na, nz = az / na;
    double s, c;
    sincos(na, &s, &c);
    const double n[3] = {nx, ny, nz};
    const double K[3][3] = {{0, -nz, ny}, {nz, 0, -nx}, {-ny, nx, 0}};
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc) Rd[r][cc] = c * ((r == cc ? 1.0 : 0.0) - n[r] * n[cc]) + s * K[r][cc] + n[r] * n[cc];
  }
  double* Rp = R + i * 9;
  double Ro[9];
  for (int cc = 0; cc < 3; ++cc)
    for (int r = 0; r < 3; ++r) Ro[cc * 3 + r] = Rd[r][0] * Rp[cc * 3 + 0] + Rd[r][1] * Rp[cc * 3 + 1] + Rd[r][2] * Rp[cc * 3 + 2];
  for (int q = 0; q < 9; ++q) Rp[q] = Ro[q];
}

__global__ void k_cm3_to_rm3(const double* __restrict__ in, double* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * 9) return;
  int64_t m = i / 9;
  int k = (int)(i % 9), r = k / 3, cc = k % 3;
  out[i] = in[m * 9 + cc * 3 + r];
}
__global__ void k_cols_to_rows(const double* __restrict__ in, double* __restrict__ out, int64_t nrows, int ncols) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nrows * ncols) return;
  int64_t r = i / ncols;
  int cc = (int)(i % ncols);
  out[i] = in[(int64_t)cc * nrows + r];
}
__global__ void k_rows_to_colmajor(const double* __restrict__ in, double* __restrict__ out, int64_t nrows, int ncols) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nrows * ncols) return;
  int64_t r = i / ncols;
  int cc = (int)(i % ncols);
  out[(int64_t)cc * nrows + r] = in[i];
}

// ---- host helpers -------------------------------------------------------------------
}  // namespace
namespace fsk {
int shell_args(fsgpu_ctx* c, const fsgpu_shell_params* p, int nnpe, bool comp, bool need_normals, ShellArgs& A) {
  FS_REQUIRE(p != nullptr, FSGPU_ERR_ARG, "null parameter block");
  FS_REQUIRE(c->nnpe == nnpe, FSGPU_ERR_STATE, "mesh has %d nodes per element, operator needs %d", c->nnpe, nnpe);
  if (need_normals)
    FS_REQUIRE(c->associated, FSGPU_ERR_STATE, "geometry not associated (call fsgpu_set_normals / fsgpu_associategeometry)");
  if (!comp) {
    FS_REQUIRE(c->nthick == 1 || c->nthick == c->nelem || (nnpe == 4 && c->nthick == c->nelem * c->rule.npts),
               FSGPU_ERR_STATE, "thickness not set (need 1, nelem%s values)", nnpe == 4 ? " or nelem*npts" : "");
  } else {
    FS_REQUIRE(c->ngroups >= 1, FSGPU_ERR_STATE, "layup groups not set");
    FS_REQUIRE(c->ncs == 1 || c->ncs == c->nelem || (nnpe == 4 && c->ncs == c->nelem * c->rule.npts), FSGPU_ERR_STATE,
               "layup csys array has the wrong length");
  }
  if (nnpe == 4) FS_REQUIRE(c->rule.npts >= 1, FSGPU_ERR_STATE, "integration rule not set");
  A.conn = c->conn.p;
  A.xyz = c->xyz.p;
  A.nrm = c->nrm.p;
  A.lam = nullptr;
  A.thick = c->thick.p;
  A.nthick = c->nthick;
  A.stabf = c->stabf.p;
  A.nstab = c->nstab;
  A.nelem = c->nelem;
  for (int i = 0; i < 9; ++i) A.Dps[i] = p->Dps[i];
  for (int i = 0; i < 4; ++i) A.Dt[i] = p->Dt[i] * (5.0 / 6.0);  // shear correction (src/FEMMShellT3FFModule.jl:658-659)
  {
    double a[3][3], d[3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) a[i][j] = A.Dps[i * 3 + j];
    ldlt<3>(a, d);
    A.hf.L10 = a[1][0];
    A.hf.L20 = a[2][0];
    A.hf.L21 = a[2][1];
    for (int i = 0; i < 3; ++i) {
      A.hf.dps[i] = d[i];
      A.hf.sdps[i] = sqrt(d[i]);
    }
    double h[2][2] = {{A.Dt[0], A.Dt[1]}, {A.Dt[2], A.Dt[3]}}, dd[2];
    ldlt<2>(h, dd);
    A.hf.Lt = h[1][0];
    A.hf.dts[0] = dd[0];
    A.hf.dts[1] = dd[1];
    A.hf.sdts[0] = sqrt(dd[0]);
    A.hf.sdts[1] = sqrt(dd[1]);
    if (!comp) FS_REQUIRE(d[0] > 0 && d[1] > 0 && d[2] > 0 && dd[0] > 0 && dd[1] > 0, FSGPU_ERR_ARG,
                          "the plane-stress / transverse-shear moduli are not positive definite");
  }
  A.rho = p->rho;
  A.alpha = p->stab_alpha;
  A.drill = p->drilling_stiffness_scale;
  A.gdata = c->group_data.p;
  A.gof = c->group_of.p;
  A.cs = c->csmat.p;
  A.ncs = c->ncs;
  A.rule = c->rule;
  FS_TRY(c->flag.ensure(8));
  FS_CUDA(cudaMemsetAsync(c->flag.p, 0, 8 * sizeof(int32_t), c->stream));
  A.flag = c->flag.p;
  return FSGPU_OK;
}

}  // namespace fsk
namespace {
template <class Emit>
int launch_t3(fsgpu_ctx* c, const ShellArgs& A, bool comp, bool sheark, Emit em) {
  const int wpb = FS_T3_WPB;
  const int64_t nwarps = (A.nelem + T3_EPW - 1) / T3_EPW;
  const int grid = (int)((nwarps + wpb - 1) / wpb);
  if (grid == 0) return FSGPU_OK;
  const size_t sm = (size_t)wpb * t3_warp_doubles(sheark, Emit::kCoop) * sizeof(double);
#define T3_GO(CO, SK)                                                                                         \
  do {                                                                                                        \
    FS_CUDA(cudaFuncSetAttribute(k_t3_stiffness<CO, SK, Emit>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    k_t3_stiffness<CO, SK, Emit><<<grid, wpb * 32, sm, c->stream>>>(A, em);                                   \
  } while (0)
  if (comp && sheark)
    T3_GO(true, true);
  else if (comp)
    T3_GO(true, false);
  else if (sheark)
    T3_GO(false, true);
  else
    T3_GO(false, false);
#undef T3_GO
  c->launches++;
  FS_CUDA(cudaGetLastError());
  return FSGPU_OK;
}
template <class Emit>
int launch_q4(fsgpu_ctx* c, const ShellArgs& A0, bool comp, Emit em) {
  const int wpb = FS_Q4_WPB;
  const int64_t nwarps = (A0.nelem + 1) / 2;
  const int grid = (int)((nwarps + wpb - 1) / wpb);
  if (grid == 0) return FSGPU_OK;
  ShellArgs A = A0;
  if (comp) {  // laminate constitutive data once per integration point instead of on each of its four lanes
    const int64_t npt = A.nelem * A.rule.npts;
    FS_TRY(c->lam_prep.ensure((size_t)npt * 24));
    k_q4_laminate_prep<<<grid_for(npt, 128), 128, 0, c->stream>>>(A, c->lam_prep.p);
    FS_CUDA(cudaGetLastError());
    c->launches++;
    A.lam = c->lam_prep.p;
  }
  const size_t sm = (size_t)wpb * Q4_WARP_DBL * sizeof(double);
#define Q4_GO(CO, CH)                                                                                          \
  do {                                                                                                         \
    FS_CUDA(cudaFuncSetAttribute(k_q4_stiffness<CO, CH, Emit>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    k_q4_stiffness<CO, CH, Emit><<<grid, wpb * 32, sm, c->stream>>>(A, em);                                    \
  } while (0)
  const bool chunked = A.rule.npts > 4;
  if (comp && chunked)
    Q4_GO(true, true);
  else if (comp)
    Q4_GO(true, false);
  else if (chunked)
    Q4_GO(false, true);
  else
    Q4_GO(false, false);
#undef Q4_GO
  c->launches++;
  FS_CUDA(cudaGetLastError());
  return FSGPU_OK;
}

// ---- emission plan of the T3 fast path (EmitRuns::t3_emit_plan) ----------------------------------------------------
// One thread per warp of ten elements.  Lane of (element el of the warp, node j): 16 (el / 5) + 3 (el % 5) + j, as in
// k_t3_stiffness.  Three sorted lists of (key, record) pairs -- diagonal blocks keyed by the node, edge blocks keyed by
// (column node, row node) once as the lower and once as the upper block of the edge -- are cut into descriptors of at
// most three contributor lanes with equal keys.
__device__ void t3_plan_insert(unsigned long long* key, unsigned* rec, int& m, unsigned long long k, unsigned r) {
  int p = m++;
  while (p > 0 && key[p - 1] > k) {  // insertion sort: the lists are short and nearly sorted
    key[p] = key[p - 1];
    rec[p] = rec[p - 1];
    --p;
  }
  key[p] = k;
  rec[p] = r;
}
// descriptors of one list, in key order (sorted by column node), at most maxc contributors each
__device__ int t3_plan_emit(const unsigned long long* key, const unsigned* rec, int m, int maxc, unsigned* out) {
  int nb = 0;
  for (int a = 0; a < m;) {
    int b = a;
    while (b < m && key[b] == key[a]) ++b;
    for (int q = a; q < b; q += maxc) {  // record: [0:17) placement fields of the descriptor, [17:22) contributor lane
      const int n = b - q < maxc ? b - q : maxc;
      unsigned d = rec[q] & 0x1ffffu;
      for (int k = 0; k < 3; ++k) d |= (k < n ? (rec[q + k] >> 17) & 31u : 31u) << (17 + 5 * k);
      out[nb++] = d;
    }
    a = b;
  }
  return nb;
}
__global__ void k_t3_plan(const int32_t* __restrict__ conn, int64_t nelem, int64_t nwarps, unsigned* __restrict__ plan) {
  const int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (w >= nwarps) return;
  unsigned long long kd[30], kl[30], ku[30];
  unsigned rd[30], rl[30], ru[30];
  int md = 0, ml = 0, mu = 0;
  for (int el = 0; el < T3_EPW; ++el) {
    const int64_t e = w * T3_EPW + el;
    if (e >= nelem) break;
    const int l0 = 16 * (el / 5) + 3 * (el % 5);
    int nd[3];
    for (int j = 0; j < 3; ++j) nd[j] = conn[e * 3 + j];
    for (int j = 0; j < 3; ++j) {
      const int jn = j == 2 ? 0 : j + 1;
      const unsigned L = (unsigned)(l0 + j), Ln = (unsigned)(l0 + jn);
      const unsigned long long u = (unsigned)nd[j], un = (unsigned)nd[jn];
      // record = column lane | row lane << 5 | P-plane pair << 10 | lane holding the offsets << 12 | contributor << 17
      t3_plan_insert(kd, rd, md, u, L | (L << 5) | (0u << 10) | (L << 12) | (L << 17));
      const bool up = un > u;  // the next node is the larger one: X = K_e[next, own] is the lower block
      const unsigned long long lo = up ? u : un, hi = up ? un : u;
      const unsigned Llo = up ? L : Ln, Lhi = up ? Ln : L;
      t3_plan_insert(kl, rl, ml, (lo << 32) | hi, Llo | (Lhi << 5) | ((up ? 1u : 2u) << 10) | (L << 12) | (L << 17));
      t3_plan_insert(ku, ru, mu, (hi << 32) | lo, Lhi | (Llo << 5) | ((up ? 2u : 1u) << 10) | (L << 12) | (L << 17));
    }
  }
  unsigned* out = plan + w * EmitRuns::kT3PlanStride;
  const int nD = t3_plan_emit(kd, rd, md, 3, out + 1);
  const int nL = t3_plan_emit(kl, rl, ml, 2, out + 1 + nD);
  const int nU = t3_plan_emit(ku, ru, mu, 2, out + 1 + nD + nL);
  out[0] = (unsigned)nD | ((unsigned)nL << 8) | ((unsigned)nU << 16);
  for (int k = 1 + nD + nL + nU; k < EmitRuns::kT3PlanStride; ++k) out[k] = 0;
}
}  // namespace
namespace fsk {
int t3_build_plan(fsgpu_ctx* c) {
  if (c->nnpe != 3) return FSGPU_OK;
  if (c->t3_plan_ok) return FSGPU_OK;  // depends on the connectivity only (reset by fsgpu_set_mesh)
  const int64_t nwarps = (c->nelem + T3_EPW - 1) / T3_EPW;
  if (nwarps == 0) return FSGPU_OK;
  FS_TRY(c->t3_plan.ensure((size_t)nwarps * EmitRuns::kT3PlanStride));
  k_t3_plan<<<grid_for(nwarps, 64), 64, 0, c->stream>>>(c->conn.p, c->nelem, nwarps, c->t3_plan.p);
  FS_CUDA(cudaGetLastError());
  c->launches++;
  c->t3_plan_ok = true;
  return FSGPU_OK;
}
}  // namespace fsk
namespace {

int begin_matrix(fsgpu_ctx* c) {
  FS_REQUIRE(c->target >= 0, FSGPU_ERR_STATE, "run fsgpu_symbolic before an operator (startassembly!)");
  FS_CUDA(cudaMemsetAsync(c->nzval.p, 0, ((size_t)c->pnnz + 1) * sizeof(double), c->stream));
  c->have_matrix = false;
  return FSGPU_OK;
}
EmitScatter scatter_of(fsgpu_ctx* c) { return EmitScatter{c->nzval.p, c->slot.p, c->nelem * c->nnpe, c->nnpe}; }
EmitRuns runs_of(fsgpu_ctx* c) {
  return EmitRuns{c->nzval.p, c->pairoff.p, c->nodeinfo.p, c->dof.p, c->colptr.p, c->nelem, c->pcols, c->nnpe, c->nodecol.p,
                  c->nnpe == 3 ? c->t3_plan.p : nullptr};
}

// ---- order-fixed (deterministic) assembly for every element kind: element matrices -> dense buffer, then one
// owner per matrix block sums its contributions in ascending element order and writes the block once ------------
// (fsgpu_set_deterministic; T3FF / T3FFComp prefer the tile kernel, which needs no intermediate buffer.)  The sums run
// in the order of the reference's serial element loop (src/FEMMShellQ4RSModule.jl:914-945, src/FEMMCorotBeamModule.jl:
// 993-1022): values are bitwise reproducible from run to run; no atomics, no clearing of the value array.
__global__ void k_det_keys(const int32_t* __restrict__ conn, int nnpe, int64_t nelem, int nb, uint64_t* __restrict__ keys,
                           int32_t* __restrict__ idx) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int nn2 = nnpe * nnpe;
  if (k >= nelem * nn2) return;
  const int64_t e = k / nn2;
  const int ij = (int)(k - e * nn2), i = ij / nnpe, j = ij - i * nnpe;
  keys[k] = ((uint64_t)conn[e * nnpe + j] << nb) | (uint64_t)conn[e * nnpe + i];  // (column node, row node)
  idx[k] = (int32_t)k;
}
__global__ void k_det_heads(const uint64_t* __restrict__ keys, int64_t n, unsigned char* __restrict__ head) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
// thread = (block b, column c of the block): sums the six rows over the block's contributions, in order
template <int NNPE>
__global__ void k_det_gather(const int32_t* __restrict__ bstart, const int32_t* __restrict__ contrib, int64_t nblocks, int64_t ncontrib,
                             const double* __restrict__ dense, const int32_t* __restrict__ conn, int64_t nelem,
                             const int32_t* __restrict__ nodecol, const int32_t* __restrict__ pairoff, double* __restrict__ nz) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t b = t / 6;
  const int c = (int)(t - b * 6);
  if (b >= nblocks) return;
  constexpr int nnpe = NNPE, nn2 = NNPE * NNPE, n = 6 * NNPE;  // compile-time: the index decoding is shifts / small multiplies
  const int k0 = bstart[b], k1 = (b + 1 < nblocks) ? bstart[b + 1] : (int)ncontrib;
  const int q0 = contrib[k0];
  const int e0 = q0 / nn2, i0 = (q0 - e0 * nn2) / nnpe, j0 = q0 - e0 * nn2 - i0 * nnpe;
  const int ni = conn[(int64_t)e0 * nnpe + i0], nj = conn[(int64_t)e0 * nnpe + j0];
  const int cb = nodecol[(int64_t)nj * 8 + c];
  if (cb < 0) return;
  double a[6] = {0, 0, 0, 0, 0, 0};
  for (int k = k0; k < k1; ++k) {
    const int q = contrib[k];
    const int e = q / nn2, i = (q - e * nn2) / nnpe, j = q - e * nn2 - i * nnpe;
    const double2* src = reinterpret_cast<const double2*>(dense + (int64_t)e * n * n + (j * nnpe + i) * 36 + c * 6);  // block-major
    const double2 v0 = src[0], v1 = src[1], v2 = src[2];
    a[0] += v0.x;
    a[1] += v0.y;
    a[2] += v1.x;
    a[3] += v1.y;
    a[4] += v2.x;
    a[5] += v2.y;
  }
  const int inf = nodecol[(int64_t)ni * 8 + 6];
  const int oA = pairoff[((int64_t)(i0 * 2 + 0) * nelem + e0) * nnpe + j0];
  const int oB = pairoff[((int64_t)(i0 * 2 + 1) * nelem + e0) * nnpe + j0];
  if ((inf & 63) == 63 && oA >= 0) {
    // all six rows of the node in one run (no constrained dof): 48 contiguous bytes, written as 16-byte pieces
    double* p = nz + cb + oA;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      double2* p2 = reinterpret_cast<double2*>(p);
      p2[0] = make_double2(a[0], a[1]);
      p2[1] = make_double2(a[2], a[3]);
      p2[2] = make_double2(a[4], a[5]);
    } else {
      p[0] = a[0];
      double2* p2 = reinterpret_cast<double2*>(p + 1);
      p2[0] = make_double2(a[1], a[2]);
      p2[1] = make_double2(a[3], a[4]);
      p[5] = a[5];
    }
    return;
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const int rp = EmitRuns::row_pos(inf, oA, oB, r);
    if (rp >= 0) nz[cb + rp] = a[r];
  }
}
// contribution lists per matrix block, built once per pattern (cached until the next fsgpu_symbolic)
int det_symbolic(fsgpu_ctx* c) {
  if (c->det_ready) return FSGPU_OK;
  const int nn2 = c->nnpe * c->nnpe;
  const int64_t nk = c->nelem * nn2;
  FS_REQUIRE(nk < ((int64_t)1 << 31), FSGPU_ERR_ARG, "deterministic assembly: too many element blocks for 32-bit lists");
  int nb = 1;
  while (((int64_t)1 << nb) < c->nnodes) ++nb;
  cudaStream_t st = c->stream;
  DBuf<uint64_t> k1, k2;
  DBuf<int32_t> i1;
  DBuf<unsigned char> head;
  DBuf<int32_t> nsel;
  FS_TRY(k1.ensure((size_t)nk + 1));
  FS_TRY(k2.ensure((size_t)nk + 1));
  FS_TRY(i1.ensure((size_t)nk + 1));
  FS_TRY(head.ensure((size_t)nk + 1));
  FS_TRY(nsel.ensure(1));
  FS_TRY(c->det_contrib.ensure((size_t)nk + 1));
  FS_TRY(c->det_bstart.ensure((size_t)nk + 1));
  c->det_nblocks = 0;
  c->det_ncontrib = nk;
  if (nk > 0) {
    k_det_keys<<<grid_for(nk, 256), 256, 0, st>>>(c->conn.p, c->nnpe, c->nelem, nb, k1.p, i1.p);
    size_t tb = 0, tb2 = 0;
    // stable: within a block the contributions keep ascending (element, i, j) order
    FS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k1.p, k2.p, i1.p, c->det_contrib.p, nk, 0, 2 * nb, st));
    cub::CountingInputIterator<int32_t> iota(0);
    FS_CUDA(cub::DeviceSelect::Flagged(nullptr, tb2, iota, head.p, c->det_bstart.p, nsel.p, nk, st));
    FS_TRY(c->tmp.ensure(tb > tb2 ? tb : tb2));
    FS_CUDA(cub::DeviceRadixSort::SortPairs(c->tmp.p, tb, k1.p, k2.p, i1.p, c->det_contrib.p, nk, 0, 2 * nb, st));
    k_det_heads<<<grid_for(nk, 256), 256, 0, st>>>(k2.p, nk, head.p);
    FS_CUDA(cub::DeviceSelect::Flagged(c->tmp.p, tb2, iota, head.p, c->det_bstart.p, nsel.p, nk, st));
    c->launches += 6;
    int32_t nbk = 0;
    FS_CUDA(cudaMemcpyAsync(&nbk, nsel.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaStreamSynchronize(st));
    c->det_nblocks = nbk;
  }
  c->det_ready = true;
  return FSGPU_OK;
}
int det_gather(fsgpu_ctx* c) {
  if (c->det_nblocks > 0) {
#define DET_GO(NN)                                                                                                                  \
  k_det_gather<NN><<<grid_for(c->det_nblocks * 6, 256), 256, 0, c->stream>>>(c->det_bstart.p, c->det_contrib.p, c->det_nblocks,        \
                                                                             c->det_ncontrib, c->det_dense.p, c->conn.p, c->nelem,     \
                                                                             c->nodecol.p, c->pairoff.p, c->nzval.p)
    if (c->nnpe == 2)
      DET_GO(2);
    else if (c->nnpe == 3)
      DET_GO(3);
    else
      DET_GO(4);
#undef DET_GO
    c->launches++;
    FS_CUDA(cudaGetLastError());
  }
  return FSGPU_OK;
}
int det_begin(fsgpu_ctx* c) {
  FS_REQUIRE(c->target >= 0, FSGPU_ERR_STATE, "run fsgpu_symbolic before an operator (startassembly!)");
  FS_TRY(det_symbolic(c));
  const int n = 6 * c->nnpe;
  FS_TRY(c->det_dense.ensure((size_t)c->nelem * n * n + 2));
  c->have_matrix = false;
  return FSGPU_OK;
}

int shell_stiffness(fsgpu_ctx* c, const fsgpu_shell_params* p, int nnpe, bool comp) {
  FS_TRY(check_ctx(c));
  ShellArgs A;
  FS_TRY(shell_args(c, p, nnpe, comp, true, A));
  const bool tile = nnpe == 3 && c->tile_ok && c->want_tile && p->transv_shear_formulation != 1;
  const bool gather = !tile && c->want_tile && c->fast;  // order-fixed assembly through the dense element matrices
  if (tile) {
    // owner-computes: every stored entry is written exactly once, no clearing needed
    FS_REQUIRE(c->target >= 0, FSGPU_ERR_STATE, "run fsgpu_symbolic before an operator (startassembly!)");
    c->have_matrix = false;
  } else if (gather) {
    FS_TRY(det_begin(c));
  } else {
    FS_TRY(begin_matrix(c));
  }
  c->last_path = tile ? 2 : (gather ? 3 : (c->fast ? 1 : 0));
  FS_TRY(time_begin(c));
  if (tile) {
    FS_TRY(launch_t3_tile(c, A, comp));
  } else if (gather) {
    EmitDense em{c->det_dense.p, nnpe, 1};
    if (nnpe == 3)
      FS_TRY(launch_t3(c, A, comp, p->transv_shear_formulation == 1, em));
    else
      FS_TRY(launch_q4(c, A, comp, em));
    FS_TRY(det_gather(c));
  } else if (nnpe == 3) {
    if (c->fast)
      FS_TRY(launch_t3(c, A, comp, p->transv_shear_formulation == 1, runs_of(c)));
    else
      FS_TRY(launch_t3(c, A, comp, p->transv_shear_formulation == 1, scatter_of(c)));
  } else {
    if (c->fast)
      FS_TRY(launch_q4(c, A, comp, runs_of(c)));
    else
      FS_TRY(launch_q4(c, A, comp, scatter_of(c)));
  }
  FS_TRY(time_end(c));
  int32_t f[3] = {0, 0, 0};
  FS_CUDA(cudaMemcpyAsync(f, c->flag.p, sizeof f, cudaMemcpyDeviceToHost, c->stream));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  FS_REQUIRE(f[0] == 0, FSGPU_ERR_SINGULAR, "Singular metric matrix in _gradN_e!");
  FS_REQUIRE(f[2] == 0, FSGPU_ERR_ARG, "laminate constitutive matrix is not positive definite");
  return finalize_matrix(c);
}

template <int NNPE, bool COMP>
void launch_mass(fsgpu_ctx* c, const ShellArgs& A, double* out, int mode, int64_t limit) {
  const int64_t n = A.nelem * NNPE;
  if (n == 0) return;
  k_shell_mass<NNPE, COMP><<<grid_for(n, 256), 256, 0, c->stream>>>(A, c->dof.p, c->diagslot.p, out, mode, limit);
  c->launches++;
}
int shell_mass_any(fsgpu_ctx* c, const fsgpu_shell_params* p, int nnpe, bool comp, double* out, int mode, int64_t limit) {
  ShellArgs A;
  FS_TRY(shell_args(c, p, nnpe, comp, true, A));
  if (nnpe == 3 && !comp) launch_mass<3, false>(c, A, out, mode, limit);
  if (nnpe == 3 && comp) launch_mass<3, true>(c, A, out, mode, limit);
  if (nnpe == 4 && !comp) launch_mass<4, false>(c, A, out, mode, limit);
  if (nnpe == 4 && comp) launch_mass<4, true>(c, A, out, mode, limit);
  FS_CUDA(cudaGetLastError());
  return FSGPU_OK;
}
int shell_mass(fsgpu_ctx* c, const fsgpu_shell_params* p, int nnpe, bool comp) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->have_dofs, FSGPU_ERR_STATE, "dofnums not set");
  FS_TRY(begin_matrix(c));
  FS_TRY(shell_mass_any(c, p, nnpe, comp, c->nzval.p, 0, 0));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return finalize_matrix(c);
}

int beam_args(fsgpu_ctx* c, const fsgpu_beam_params* p, BeamArgs& B) {
  FS_REQUIRE(p != nullptr, FSGPU_ERR_ARG, "null parameter block");
  FS_REQUIRE(c->nnpe == 2, FSGPU_ERR_STATE, "beam operators need an L2 mesh");
  FS_REQUIRE(c->have_sections, FSGPU_ERR_STATE, "beam sections not set");
  FS_REQUIRE(c->have_state, FSGPU_ERR_STATE, "u1 / Rfield1 not set");
  B.conn = c->conn.p;
  B.xyz = c->xyz.p;
  B.u1 = c->u1.p;
  B.R1 = c->R1.p;
  B.sec = c->sec.p;
  B.nelem = c->nelem;
  B.E = p->E;
  B.G = p->E / 2 / (1 + p->nu);
  B.rho = p->rho;
  B.mass_type = p->mass_type;
  B.v1 = c->v1.p;
  B.force = nullptr;
  B.nforce = 0;
  return FSGPU_OK;
}
int beam_matrix(fsgpu_ctx* c, const fsgpu_beam_params* p, int op) {
  FS_TRY(check_ctx(c));
  BeamArgs B;
  FS_TRY(beam_args(c, p, B));
  if (op == 3) FS_REQUIRE(c->have_velocity, FSGPU_ERR_STATE, "v1 not set (fsgpu_set_velocity)");
  const bool gather = c->want_tile && c->fast;
  if (gather)
    FS_TRY(det_begin(c));
  else
    FS_TRY(begin_matrix(c));
  c->last_path = gather ? 3 : (c->fast ? 1 : 0);
  const int64_t n = B.nelem * 4;
  FS_TRY(time_begin(c));
  if (n > 0) {
    if (gather) {
      k_beam_matrix<EmitDense><<<grid_for(n, 128), 128, 0, c->stream>>>(B, op, EmitDense{c->det_dense.p, 2, 1});
      FS_TRY(det_gather(c));
    } else if (c->fast) {
      const size_t sm = (size_t)4 * BEAM_WARP_DBL * sizeof(double);
      const int grid = grid_for(B.nelem, 128);
#define BEAM_GO(OPV)                                                                                                   \
  do {                                                                                                                 \
    FS_CUDA(cudaFuncSetAttribute(k_beam_matrix_coop<OPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));      \
    k_beam_matrix_coop<OPV><<<grid, 128, sm, c->stream>>>(B, runs_of(c));                                              \
  } while (0)
      if (op == 0)
        BEAM_GO(0);
      else if (op == 1)
        BEAM_GO(1);
      else if (op == 2)
        BEAM_GO(2);
      else
        BEAM_GO(3);
#undef BEAM_GO
    } else
      k_beam_matrix<EmitScatter><<<grid_for(n, 128), 128, 0, c->stream>>>(B, op, scatter_of(c));
    c->launches++;
  }
  FS_TRY(time_end(c));
  FS_CUDA(cudaGetLastError());
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return finalize_matrix(c);
}

}  // namespace

// ---- C ABI --------------------------------------------------------------------------
extern "C" int fsgpu_t3ff_stiffness(fsgpu_ctx* c, const fsgpu_shell_params* p) { return shell_stiffness(c, p, 3, false); }
extern "C" int fsgpu_t3ffcomp_stiffness(fsgpu_ctx* c, const fsgpu_shell_params* p) { return shell_stiffness(c, p, 3, true); }
extern "C" int fsgpu_q4rs_stiffness(fsgpu_ctx* c, const fsgpu_shell_params* p) { return shell_stiffness(c, p, 4, false); }
extern "C" int fsgpu_q4rscomp_stiffness(fsgpu_ctx* c, const fsgpu_shell_params* p) { return shell_stiffness(c, p, 4, true); }
extern "C" int fsgpu_t3ff_mass(fsgpu_ctx* c, const fsgpu_shell_params* p) { return shell_mass(c, p, 3, false); }
extern "C" int fsgpu_t3ffcomp_mass(fsgpu_ctx* c, const fsgpu_shell_params* p) { return shell_mass(c, p, 3, true); }
extern "C" int fsgpu_q4rs_mass(fsgpu_ctx* c, const fsgpu_shell_params* p) { return shell_mass(c, p, 4, false); }
extern "C" int fsgpu_q4rscomp_mass(fsgpu_ctx* c, const fsgpu_shell_params* p) { return shell_mass(c, p, 4, true); }
extern "C" int fsgpu_corotbeam_stiffness(fsgpu_ctx* c, const fsgpu_beam_params* p) { return beam_matrix(c, p, 0); }
extern "C" int fsgpu_corotbeam_mass(fsgpu_ctx* c, const fsgpu_beam_params* p) { return beam_matrix(c, p, 1); }
extern "C" int fsgpu_corotbeam_geostiffness(fsgpu_ctx* c, const fsgpu_beam_params* p) { return beam_matrix(c, p, 2); }
extern "C" int fsgpu_corotbeam_gyroscopic(fsgpu_ctx* c, const fsgpu_beam_params* p) { return beam_matrix(c, p, 3); }

extern "C" int fsgpu_set_velocity(fsgpu_ctx* c, const double* v1) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->nnpe > 0 && v1, FSGPU_ERR_ARG, "set the mesh first / null v1");
  const int64_t n = c->nnodes;
  FS_TRY(c->v1.ensure((size_t)n * 6 + 1));
  FS_TRY(c->tmp.ensure((size_t)n * 6 * sizeof(double)));
  FS_TRY(upload(c, c->tmp.p, v1, (size_t)n * 6 * sizeof(double)));
  k_cols_to_rows<<<grid_for(n * 6, 256), 256, 0, c->stream>>>((const double*)c->tmp.p, c->v1.p, n, 6);
  c->launches++;
  FS_CUDA(cudaStreamSynchronize(c->stream));
  c->have_velocity = true;
  return FSGPU_OK;
}

extern "C" int fsgpu_corotbeam_distribloads(fsgpu_ctx* c, const fsgpu_beam_params* p, const double* force, int64_t nforce,
                                            int32_t nfree_only) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->have_dofs, FSGPU_ERR_STATE, "dofnums not set");
  BeamArgs B;
  FS_TRY(beam_args(c, p, B));
  FS_REQUIRE(force && (nforce == 1 || nforce == B.nelem), FSGPU_ERR_ARG, "force must hold 3 or 3*nelem values");
  DBuf<double> df;
  FS_TRY(df.ensure((size_t)nforce * 3 + 1));
  FS_TRY(upload(c, df.p, force, (size_t)nforce * 3 * sizeof(double)));
  B.force = df.p;
  B.nforce = nforce;
  const int64_t n = nfree_only ? c->nfree : c->nall;
  FS_TRY(c->vec.ensure((size_t)n + 1));
  FS_CUDA(cudaMemsetAsync(c->vec.p, 0, ((size_t)n + 1) * sizeof(double), c->stream));
  if (B.nelem > 0) {
    k_beam_restoring<<<grid_for(B.nelem, 128), 128, 0, c->stream>>>(B, c->dof.p, c->vec.p, n, nullptr, 1);
    c->launches++;
  }
  FS_CUDA(cudaGetLastError());
  FS_CUDA(cudaStreamSynchronize(c->stream));
  c->have_vector = true;
  c->vlen = n;
  return FSGPU_OK;
}

extern "C" int fsgpu_shell_mass_diag(fsgpu_ctx* c, const fsgpu_shell_params* p, int32_t kind, int32_t nfree_only) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->have_dofs, FSGPU_ERR_STATE, "dofnums not set");
  FS_REQUIRE(kind == 3 || kind == 4 || kind == 13 || kind == 14, FSGPU_ERR_ARG, "kind must be 3, 4, 13 or 14");
  const int64_t n = nfree_only ? c->nfree : c->nall;
  FS_TRY(c->vec.ensure((size_t)n + 1));
  FS_CUDA(cudaMemsetAsync(c->vec.p, 0, ((size_t)n + 1) * sizeof(double), c->stream));
  FS_TRY(shell_mass_any(c, p, kind % 10, kind >= 10, c->vec.p, 1, n));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  c->have_vector = true;
  c->vlen = n;
  return FSGPU_OK;
}

extern "C" int fsgpu_corotbeam_restoringforce(fsgpu_ctx* c, const fsgpu_beam_params* p, int32_t nfree_only) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->have_dofs, FSGPU_ERR_STATE, "dofnums not set");
  BeamArgs B;
  FS_TRY(beam_args(c, p, B));
  const int64_t n = nfree_only ? c->nfree : c->nall;
  FS_TRY(c->vec.ensure((size_t)n + 1));
  FS_CUDA(cudaMemsetAsync(c->vec.p, 0, ((size_t)n + 1) * sizeof(double), c->stream));
  if (B.nelem > 0) {
    k_beam_restoring<<<grid_for(B.nelem, 128), 128, 0, c->stream>>>(B, c->dof.p, c->vec.p, n, nullptr, 0);
    c->launches++;
  }
  FS_CUDA(cudaGetLastError());
  FS_CUDA(cudaStreamSynchronize(c->stream));
  c->have_vector = true;
  c->vlen = n;
  return FSGPU_OK;
}

extern "C" int fsgpu_element_vectors(fsgpu_ctx* c, const fsgpu_beam_params* p, double* out) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(out, FSGPU_ERR_ARG, "null output");
  BeamArgs B;
  FS_TRY(beam_args(c, p, B));
  DBuf<double> d;
  FS_TRY(d.ensure((size_t)B.nelem * 12 + 1));
  if (B.nelem > 0) {
    k_beam_restoring<<<grid_for(B.nelem, 128), 128, 0, c->stream>>>(B, nullptr, nullptr, 0, d.p, 0);
    c->launches++;
  }
  FS_CUDA(cudaGetLastError());
  FS_TRY(download(c, out, d.p, (size_t)B.nelem * 12 * sizeof(double)));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return FSGPU_OK;
}

extern "C" int fsgpu_element_matrices(fsgpu_ctx* c, int32_t kind, int32_t op, const void* params, double* out) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(out && params, FSGPU_ERR_ARG, "null argument");
  const int nnpe = kind == 2 ? 2 : kind % 10;
  FS_REQUIRE(nnpe == c->nnpe, FSGPU_ERR_STATE, "element kind %d does not match the mesh (nnpe %d)", kind, c->nnpe);
  const int n = 6 * nnpe;
  DBuf<double> d;
  const size_t total = (size_t)c->nelem * n * n;
  FS_TRY(d.ensure(total + 1));
  FS_CUDA(cudaMemsetAsync(d.p, 0, (total + 1) * sizeof(double), c->stream));
  EmitDense em{d.p, nnpe, 0};
  if (kind == 2) {
    BeamArgs B;
    FS_TRY(beam_args(c, (const fsgpu_beam_params*)params, B));
    FS_REQUIRE(op >= 0 && op <= 3, FSGPU_ERR_ARG, "beam op must be 0 (stiffness), 1 (mass), 2 (geostiffness) or 3 (gyroscopic)");
    if (op == 3) FS_REQUIRE(c->have_velocity, FSGPU_ERR_STATE, "v1 not set (fsgpu_set_velocity)");
    if (B.nelem > 0) {
      k_beam_matrix<EmitDense><<<grid_for(B.nelem * 4, 128), 128, 0, c->stream>>>(B, op, em);
      c->launches++;
    }
  } else {
    const fsgpu_shell_params* p = (const fsgpu_shell_params*)params;
    const bool comp = kind >= 10;
    FS_REQUIRE(kind == 3 || kind == 4 || kind == 13 || kind == 14, FSGPU_ERR_ARG, "unknown element kind %d", kind);
    ShellArgs A;
    FS_TRY(shell_args(c, p, nnpe, comp, true, A));
    if (op == 0) {
      if (nnpe == 3)
        FS_TRY(launch_t3(c, A, comp, p->transv_shear_formulation == 1, em));
      else
        FS_TRY(launch_q4(c, A, comp, em));
    } else if (op == 1) {
      DBuf<double> pairs;
      const int64_t m = c->nelem * nnpe;
      FS_TRY(pairs.ensure((size_t)m * 2 + 1));
      if (m > 0) {
        if (nnpe == 3 && !comp) k_shell_mass_pairs<3, false><<<grid_for(m, 256), 256, 0, c->stream>>>(A, pairs.p);
        if (nnpe == 3 && comp) k_shell_mass_pairs<3, true><<<grid_for(m, 256), 256, 0, c->stream>>>(A, pairs.p);
        if (nnpe == 4 && !comp) k_shell_mass_pairs<4, false><<<grid_for(m, 256), 256, 0, c->stream>>>(A, pairs.p);
        if (nnpe == 4 && comp) k_shell_mass_pairs<4, true><<<grid_for(m, 256), 256, 0, c->stream>>>(A, pairs.p);
        k_shell_mass_dense_from_diag<<<grid_for(c->nelem * n, 256), 256, 0, c->stream>>>(pairs.p, nnpe, c->nelem, d.p);
        c->launches += 2;
      }
      FS_CUDA(cudaGetLastError());
      FS_CUDA(cudaStreamSynchronize(c->stream));
    } else {
      FS_REQUIRE(false, FSGPU_ERR_ARG, "shell op must be 0 (stiffness) or 1 (mass)");
    }
    int32_t f = 0;
    FS_CUDA(cudaMemcpyAsync(&f, c->flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(cudaStreamSynchronize(c->stream));
    FS_REQUIRE(f == 0, FSGPU_ERR_SINGULAR, "Singular metric matrix in _gradN_e!");
  }
  FS_CUDA(cudaGetLastError());
  FS_TRY(download(c, out, d.p, total * sizeof(double)));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return FSGPU_OK;
}

// resultants of every integration point on the device: dout [nelem][npts][3]
static int resultants_device(fsgpu_ctx* c, const fsgpu_shell_params* p, int32_t kind, int32_t quantity, const double* u,
                             const double* outputcsys, int64_t ncs, DBuf<double>& dout, int* npts_out) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(kind == 3 || kind == 4 || kind == 13 || kind == 14, FSGPU_ERR_ARG,
             "kind must be 3 (T3FF), 4 (Q4RS), 13 (T3FFComp) or 14 (Q4RSComp)");
  FS_REQUIRE(quantity >= 1 && quantity <= 3 && u, FSGPU_ERR_ARG, "quantity must be 1 (bending), 2 (shear) or 3 (membrane)");
  const bool comp = kind > 10;
  kind = comp ? kind - 10 : kind;
  ShellArgs A;
  FS_TRY(shell_args(c, p, kind, comp, true, A));
  const int npts = kind == 3 ? 1 : c->rule.npts;
  FS_REQUIRE(ncs == 0 || ncs == 1 || ncs == c->nelem || ncs == c->nelem * npts, FSGPU_ERR_ARG, "bad output csys count");
  const int64_t n = c->nnodes, nout = c->nelem * npts * 3;
  DBuf<double> du, dcs, tmpc;
  FS_TRY(du.ensure((size_t)n * 6 + 1));
  FS_TRY(c->tmp.ensure((size_t)n * 6 * sizeof(double)));
  FS_TRY(upload(c, c->tmp.p, u, (size_t)n * 6 * sizeof(double)));
  k_cols_to_rows<<<grid_for(n * 6, 256), 256, 0, c->stream>>>((const double*)c->tmp.p, du.p, n, 6);
  c->launches++;
  if (ncs > 0) {
    FS_REQUIRE(outputcsys, FSGPU_ERR_ARG, "null output csys");
    FS_TRY(dcs.ensure((size_t)ncs * 9));
    FS_TRY(tmpc.ensure((size_t)ncs * 9));
    FS_TRY(upload(c, tmpc.p, outputcsys, (size_t)ncs * 9 * sizeof(double)));
    // column-major 3x3 -> row-major
    k_cm3_to_rm3<<<grid_for(ncs * 9, 256), 256, 0, c->stream>>>(tmpc.p, dcs.p, ncs);
    c->launches++;
  }
  FS_TRY(dout.ensure((size_t)nout + 1));
  if (c->nelem > 0) {
    if (kind == 3 && !comp)
      k_t3_resultants<false><<<grid_for(c->nelem, 128), 128, 0, c->stream>>>(A, du.p, quantity, dcs.p, ncs, dout.p);
    else if (kind == 3)
      k_t3_resultants<true><<<grid_for(c->nelem, 128), 128, 0, c->stream>>>(A, du.p, quantity, dcs.p, ncs, dout.p);
    else if (!comp)
      k_q4_resultants<false><<<grid_for(c->nelem * npts, 128), 128, 0, c->stream>>>(A, du.p, quantity, dcs.p, ncs, dout.p);
    else
      k_q4_resultants<true><<<grid_for(c->nelem * npts, 128), 128, 0, c->stream>>>(A, du.p, quantity, dcs.p, ncs, dout.p);
    c->launches++;
  }
  FS_CUDA(cudaGetLastError());
  *npts_out = npts;
  return FSGPU_OK;
}
extern "C" int fsgpu_shell_resultants(fsgpu_ctx* c, const fsgpu_shell_params* p, int32_t kind, int32_t quantity, const double* u,
                                      const double* outputcsys, int64_t ncs, double* out) {
  FS_REQUIRE(out != nullptr, FSGPU_ERR_ARG, "null output");
  DBuf<double> dout;
  int npts = 0;
  FS_TRY(resultants_device(c, p, kind, quantity, u, outputcsys, ncs, dout, &npts));
  FS_TRY(download(c, out, dout.p, (size_t)c->nelem * npts * 3 * sizeof(double)));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return FSGPU_OK;
}

// FinEtools' fieldfromintegpoints with nodevalmethod = :invdistance (FEMMBaseModule, FinEtools 8.2.5; called on the shells by
// test/test_shell_resultants.jl:123): every integration point adds value / (d + dmin / 1e9) to the nodes of its element, d the
// SQUARED distance node - point (the centroid for the T3 shells, the integration point for the Q4 shells: the `loc` of the
// reference's inspectintegpoints), dmin the smallest positive d of the element; the nodal value is the weighted mean.
__global__ void k_nodal_invdist(const int32_t* __restrict__ conn, const double4* __restrict__ xyz, int nnpe, int npts, Rule rule,
                                int64_t nelem, const double* __restrict__ val, double* __restrict__ num, double* __restrict__ den) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (tid >= nelem * npts) return;
  const int64_t e = tid / npts;
  const int q = (int)(tid - e * npts);
  int nd[4];
  V3 X[4];
  for (int a = 0; a < nnpe; ++a) {
    nd[a] = conn[e * nnpe + a];
    X[a] = ld3(xyz, nd[a]);
  }
  V3 loc = v3(0, 0, 0);
  if (nnpe == 3) {
    loc = (1.0 / 3.0) * (X[0] + X[1] + X[2]);
  } else {
    const double xi = rule.xi[q], eta = rule.eta[q];
    const double N[4] = {0.25 * (1 - xi) * (1 - eta), 0.25 * (1 + xi) * (1 - eta), 0.25 * (1 + xi) * (1 + eta), 0.25 * (1 - xi) * (1 + eta)};
    for (int a = 0; a < 4; ++a) loc = loc + N[a] * X[a];
  }
  double d[4], dmin = 1.0e300;
  for (int a = 0; a < nnpe; ++a) {
    const V3 r = X[a] - loc;
    d[a] = dot(r, r);
    if (d[a] > 0.0 && d[a] < dmin) dmin = d[a];
  }
  dmin *= 1.0e-9;
  for (int a = 0; a < nnpe; ++a) {
    const double w = 1.0 / (d[a] + dmin);
    for (int k = 0; k < 3; ++k) atomicAdd(num + (int64_t)nd[a] * 3 + k, w * val[tid * 3 + k]);
    atomicAdd(den + nd[a], w);
  }
}
// nodal means, written column-major (nnodes x 3) as a NodalField's values
__global__ void k_nodal_mean(const double* __restrict__ num, const double* __restrict__ den, int64_t nnodes, double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nnodes * 3) return;
  const int64_t node = i / 3;
  const int k = (int)(i - node * 3);
  out[k * nnodes + node] = den[node] > 0.0 ? num[i] / den[node] : 0.0;
}
extern "C" int fsgpu_shell_nodal_field(fsgpu_ctx* c, const fsgpu_shell_params* p, int32_t kind, int32_t quantity, const double* u,
                                       const double* outputcsys, int64_t ncs, double* out) {
  FS_REQUIRE(out != nullptr, FSGPU_ERR_ARG, "null output");
  DBuf<double> dout, num, den, res;
  int npts = 0;
  FS_TRY(resultants_device(c, p, kind, quantity, u, outputcsys, ncs, dout, &npts));
  const int64_t n = c->nnodes;
  FS_TRY(num.ensure((size_t)n * 3 + 1));
  FS_TRY(den.ensure((size_t)n + 1));
  FS_TRY(res.ensure((size_t)n * 3 + 1));
  FS_CUDA(cudaMemsetAsync(num.p, 0, (size_t)n * 3 * sizeof(double), c->stream));
  FS_CUDA(cudaMemsetAsync(den.p, 0, (size_t)n * sizeof(double), c->stream));
  if (c->nelem > 0) {
    k_nodal_invdist<<<grid_for(c->nelem * npts, 128), 128, 0, c->stream>>>(c->conn.p, c->xyz.p, c->nnpe, npts, c->rule, c->nelem, dout.p,
                                                                          num.p, den.p);
    c->launches++;
  }
  if (n > 0) {
    k_nodal_mean<<<grid_for(n * 3, 256), 256, 0, c->stream>>>(num.p, den.p, n, res.p);
    c->launches++;
  }
  FS_CUDA(cudaGetLastError());
  FS_TRY(download(c, out, res.p, (size_t)n * 3 * sizeof(double)));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return FSGPU_OK;
}

extern "C" int fsgpu_element_sizes(fsgpu_ctx* c, double* h) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(h && c->nnpe > 0, FSGPU_ERR_ARG, "bad arguments");
  DBuf<double> d;
  FS_TRY(d.ensure((size_t)c->nelem + 1));
  if (c->nelem > 0) {
    k_element_sizes<<<grid_for(c->nelem, 256), 256, 0, c->stream>>>(c->conn.p, c->xyz.p, c->nnpe, c->nelem, d.p);
    c->launches++;
  }
  FS_TRY(download(c, h, d.p, (size_t)c->nelem * sizeof(double)));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return FSGPU_OK;
}

extern "C" int fsgpu_update_rotation_field(fsgpu_ctx* c, const double* dchi_values, double* Rfield_out) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->have_state, FSGPU_ERR_STATE, "Rfield1 not set");
  FS_REQUIRE(dchi_values, FSGPU_ERR_ARG, "null dchi");
  const int64_t n = c->nnodes;
  FS_TRY(c->tmp.ensure((size_t)n * 9 * sizeof(double)));
  FS_TRY(upload(c, c->tmp.p, dchi_values, (size_t)n * 6 * sizeof(double)));
  if (n > 0) {
    k_update_rotation<<<grid_for(n, 256), 256, 0, c->stream>>>(c->R1.p, (const double*)c->tmp.p, n);
    c->launches++;
  }
  FS_CUDA(cudaGetLastError());
  if (Rfield_out) {
    // back to nnodes x 9 column-major
    DBuf<double> o;
    FS_TRY(o.ensure((size_t)n * 9 + 1));
    // transpose [n][9] row-major -> column-major: treat as column-major 9 x n -> rows
    k_rows_to_colmajor<<<grid_for(n * 9, 256), 256, 0, c->stream>>>(c->R1.p, o.p, n, 9);
    c->launches++;
    FS_TRY(download(c, Rfield_out, o.p, (size_t)n * 9 * sizeof(double)));
  }
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return FSGPU_OK;
}
