// T3FF / T3FFComp stiffness, OWNER-COMPUTES tile kernel: no atomics, no clearing of the value
// array, every stored entry written exactly once (deterministic).
//
// Why: the scatter kernels are capped by the L2 atomic unit (RED.ADD.F64: ~266 G elements/s
// measured, scripts/micro/red_bench.cu) -- 324 REDs per T3 element put a 4.9 ms floor under the
// 4M-element mesh whereas HBM needs 0.7 ms.  Here a CTA owns a compact patch of NODES (consecutive
// in Morton order of the coordinates).  Phase 1 builds the folded global-dof strips of every element
// touching an owned node (halo elements are recomputed by the neighbouring tiles, ~1.3x setup work)
// into shared memory.  Phase 2 forms, for every owned column node a and every neighbour b, the
// complete 6x6 block K[b,a] = sum over the elements containing both -- all of them are in the tile --
// and stores it with plain stores through the run-structured addressing (fsgpu_core.cu).
//
// Requirements (checked in tile_symbolic, otherwise the RED kernel is used): T3 mesh, fast-path
// numbering, non-diagonal target, AVERAGE_B shear, every tile fits the shared-memory capacity.
#include <cub/cub.cuh>

#include "fsgpu_shell.cuh"

using namespace fs;
using namespace fsm;

namespace fsk {

constexpr int TILE_NO = 32;        // owned nodes per tile
constexpr int TILE_CAP = 144;      // max elements per tile
constexpr int TILE_THREADS = 512;  // 16 warps
constexpr int STRIP_LD = 49;       // doubles per (element, node) strip: 48 + 1 pad (odd: conflict-free)
constexpr int TILE_MAXDEG = 24;    // max node valence handled

// ------------------------------------------------------------------------------------
// symbolic part
// ------------------------------------------------------------------------------------
__device__ inline void atomic_min_d(double* a, double v) {
  unsigned long long* p = (unsigned long long*)a;
  unsigned long long old = *p;
  while (__longlong_as_double(old) > v) {
    unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}
__device__ inline void atomic_max_d(double* a, double v) {
  unsigned long long* p = (unsigned long long*)a;
  unsigned long long old = *p;
  while (__longlong_as_double(old) < v) {
    unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}
__global__ void k_bbox(const double4* __restrict__ xyz, int64_t n, double* __restrict__ mm) {
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double4 v = xyz[i];
    const double c[3] = {v.x, v.y, v.z};
    for (int d = 0; d < 3; ++d) {
      lo[d] = fmin(lo[d], c[d]);
      hi[d] = fmax(hi[d], c[d]);
    }
  }
  for (int d = 0; d < 3; ++d) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomic_min_d(mm + d, lo[d]);
      atomic_max_d(mm + 3 + d, hi[d]);
    }
  }
}
__device__ inline uint64_t spread3(uint64_t v) {  // 21 bits -> every third bit
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}
__global__ void k_morton(const double4* __restrict__ xyz, int64_t n, const double* __restrict__ mm, uint64_t* __restrict__ keys,
                         int32_t* __restrict__ ids) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 v = xyz[i];
  const double c[3] = {v.x, v.y, v.z};
  uint64_t q[3];
  // one common scale for the three directions keeps the cells cubic (compact patches)
  const double ext = fmax(fmax(mm[3] - mm[0], mm[4] - mm[1]), fmax(mm[5] - mm[2], 1e-300));
  for (int d = 0; d < 3; ++d) {
    double t = (c[d] - mm[d]) / ext;
    t = t < 0 ? 0 : (t > 1 ? 1 : t);
    q[d] = (uint64_t)(t * 2097151.0);
  }
  keys[i] = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);
  ids[i] = (int32_t)i;
}
__global__ void k_rank(const int32_t* __restrict__ morder, int64_t n, int32_t* __restrict__ mrank) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) mrank[morder[i]] = (int32_t)i;
}
// keys (hi << 32 | e) for every (element, node) incidence; hi = node id or tile of the node
__global__ void k_inc_keys(const int32_t* __restrict__ conn, int nnpe, int64_t nelem, const int32_t* __restrict__ mrank, int no,
                           uint64_t* __restrict__ keys) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nelem * nnpe) return;
  const int64_t e = i / nnpe;
  const int32_t node = conn[i];
  const uint64_t hi = mrank ? (uint64_t)(mrank[node] / no) : (uint64_t)node;
  keys[i] = (hi << 32) | (uint64_t)e;
}
__global__ void k_count_hi(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ cnt) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(cnt + (keys[i] >> 32), 1);
}
__global__ void k_low32(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)(keys[i] & 0xffffffffu);
}
__device__ inline int find_row_t(const int32_t* __restrict__ rowval, int lo, int hi, int32_t r) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (rowval[mid] < r)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}
// adjoff[p] / adjoff[nadj + p]: row offset (relative to the column start) of neighbour adj[p]'s
// run A / run B inside any included column of the node that owns adjacency entry p
__global__ void k_adj_offsets(const int32_t* __restrict__ adjptr, const int32_t* __restrict__ adj, const int32_t* __restrict__ dof,
                              const int32_t* __restrict__ info, const int32_t* __restrict__ colptr,
                              const int32_t* __restrict__ rowval, int64_t nnodes, int64_t nc, int64_t nadj,
                              int32_t* __restrict__ adjoff) {
  const int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a >= nnodes) return;
  int col = -1;
  for (int cc = 0; cc < 6; ++cc)
    if (dof[a * 6 + cc] < nc) {
      col = dof[a * 6 + cc];
      break;
    }
  for (int p = adjptr[a]; p < adjptr[a + 1]; ++p) {
    const int b = adj[p];
    const int inf = info[b];
    const int mA = inf & 63, mB = (inf >> 8) & 63;
    int oA = -1, oB = -1;
    if (col >= 0) {
      const int lo = colptr[col], hi = colptr[col + 1];
      if (mA) oA = find_row_t(rowval, lo, hi, dof[(int64_t)b * 6 + (__ffs(mA) - 1)]) - lo;
      if (mB) oB = find_row_t(rowval, lo, hi, dof[(int64_t)b * 6 + (__ffs(mB) - 1)]) - lo;
    }
    adjoff[p] = oA;
    adjoff[nadj + p] = oB;
  }
}
__global__ void k_max_i32(const int32_t* __restrict__ v, int64_t n, int32_t* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) atomicMax(out, v[i]);
}

#define TLAUNCH(ctx, kern, n, ...)                                             \
  do {                                                                         \
    if ((n) > 0) {                                                             \
      kern<<<fs::grid_for((n), 256), 256, 0, (ctx)->stream>>>(__VA_ARGS__);    \
      (ctx)->launches++;                                                       \
    }                                                                          \
  } while (0)

static int sort_keys(fsgpu_ctx* c, DBuf<uint64_t>& a, DBuf<uint64_t>& b, int64_t n, int end_bit) {
  size_t tb = 0;
  FS_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, a.p, b.p, n, 0, end_bit, c->stream));
  FS_TRY(c->tmp.ensure(tb));
  FS_CUDA(cub::DeviceRadixSort::SortKeys(c->tmp.p, tb, a.p, b.p, n, 0, end_bit, c->stream));
  c->launches += 4;
  return FSGPU_OK;
}
static int scan_i32(fsgpu_ctx* c, const int32_t* in, int32_t* out, int64_t n) {
  size_t tb = 0;
  FS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, c->stream));
  FS_TRY(c->tmp.ensure(tb));
  FS_CUDA(cub::DeviceScan::ExclusiveSum(c->tmp.p, tb, in, out, n, c->stream));
  c->launches += 2;
  return FSGPU_OK;
}

int tile_symbolic(fsgpu_ctx* c) {
  c->tile_ok = false;
  if (c->nnpe != 3 || !c->fast || c->nelem == 0) return FSGPU_OK;
  if (!c->want_tile) return FSGPU_OK;
  cudaStream_t st = c->stream;
  const int64_t nn = c->nnodes, ne = c->nelem, ninc = ne * 3;
  const int NO = TILE_NO;
  const int64_t ntiles = (nn + NO - 1) / NO;
  // (1) Morton order of the nodes
  DBuf<double> mm;
  DBuf<uint64_t> k1, k2;
  DBuf<int32_t> ids, mrank, cnt;
  FS_TRY(mm.ensure(8));
  const double init[6] = {1e300, 1e300, 1e300, -1e300, -1e300, -1e300};
  FS_CUDA(cudaMemcpyAsync(mm.p, init, sizeof init, cudaMemcpyHostToDevice, st));
  k_bbox<<<592, 256, 0, st>>>(c->xyz.p, nn, mm.p);
  c->launches++;
  const int64_t nk = ninc > nn ? ninc : nn;
  FS_TRY(k1.ensure((size_t)nk + 1));
  FS_TRY(k2.ensure((size_t)nk + 1));
  FS_TRY(ids.ensure((size_t)nn + 1));
  FS_TRY(c->morder.ensure((size_t)nn + 1));
  FS_TRY(mrank.ensure((size_t)nn + 1));
  TLAUNCH(c, k_morton, nn, c->xyz.p, nn, mm.p, k1.p, ids.p);
  {
    size_t tb = 0;
    FS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k1.p, k2.p, ids.p, c->morder.p, nn, 0, 63, st));
    FS_TRY(c->tmp.ensure(tb));
    FS_CUDA(cub::DeviceRadixSort::SortPairs(c->tmp.p, tb, k1.p, k2.p, ids.p, c->morder.p, nn, 0, 63, st));
    c->launches += 4;
  }
  TLAUNCH(c, k_rank, nn, c->morder.p, nn, mrank.p);
  int nbits = 1;
  while (((int64_t)1 << nbits) < (nn > ne ? nn : ne) + 1) ++nbits;
  // (2) node -> incident elements
  FS_TRY(cnt.ensure((size_t)nn + 2));
  FS_TRY(c->nel_ptr.ensure((size_t)nn + 2));
  FS_TRY(c->nel.ensure((size_t)ninc + 1));
  TLAUNCH(c, k_inc_keys, ninc, c->conn.p, 3, ne, (const int32_t*)nullptr, NO, k1.p);
  FS_TRY(sort_keys(c, k1, k2, ninc, 32 + nbits));
  FS_CUDA(cudaMemsetAsync(cnt.p, 0, ((size_t)nn + 2) * sizeof(int32_t), st));
  TLAUNCH(c, k_count_hi, ninc, k2.p, ninc, cnt.p);
  FS_TRY(scan_i32(c, cnt.p, c->nel_ptr.p, nn + 1));
  TLAUNCH(c, k_low32, ninc, k2.p, ninc, c->nel.p);
  // (3) tile -> elements touching its owned nodes (unique)
  DBuf<int64_t> nsel;
  FS_TRY(nsel.ensure(1));
  TLAUNCH(c, k_inc_keys, ninc, c->conn.p, 3, ne, mrank.p, NO, k1.p);
  FS_TRY(sort_keys(c, k1, k2, ninc, 32 + nbits));
  {
    size_t tb = 0;
    FS_CUDA(cub::DeviceSelect::Unique(nullptr, tb, k2.p, k1.p, nsel.p, ninc, st));
    FS_TRY(c->tmp.ensure(tb));
    FS_CUDA(cub::DeviceSelect::Unique(c->tmp.p, tb, k2.p, k1.p, nsel.p, ninc, st));
    c->launches += 2;
  }
  int64_t ntel = 0;
  FS_CUDA(cudaMemcpyAsync(&ntel, nsel.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  DBuf<int32_t> tcnt, tmax;
  FS_TRY(tcnt.ensure((size_t)ntiles + 2));
  FS_TRY(tmax.ensure(1));
  FS_TRY(c->tel_ptr.ensure((size_t)ntiles + 2));
  FS_TRY(c->tel.ensure((size_t)ntel + 1));
  FS_CUDA(cudaMemsetAsync(tcnt.p, 0, ((size_t)ntiles + 2) * sizeof(int32_t), st));
  FS_CUDA(cudaMemsetAsync(tmax.p, 0, sizeof(int32_t), st));
  TLAUNCH(c, k_count_hi, ntel, k1.p, ntel, tcnt.p);
  FS_TRY(scan_i32(c, tcnt.p, c->tel_ptr.p, ntiles + 1));
  TLAUNCH(c, k_low32, ntel, k1.p, ntel, c->tel.p);
  TLAUNCH(c, k_max_i32, ntiles, tcnt.p, ntiles, tmax.p);
  // max valence
  DBuf<int32_t> dmax;
  FS_TRY(dmax.ensure(1));
  FS_CUDA(cudaMemsetAsync(dmax.p, 0, sizeof(int32_t), st));
  TLAUNCH(c, k_max_i32, nn, cnt.p, nn, dmax.p);
  int32_t maxel = 0, maxdeg = 0;
  FS_CUDA(cudaMemcpyAsync(&maxel, tmax.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaMemcpyAsync(&maxdeg, dmax.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  if (maxel > TILE_CAP || maxdeg + 1 > TILE_MAXDEG) return FSGPU_OK;  // keep the RED kernel
  // (4) per-adjacency row offsets
  int32_t nadj32 = 0;
  FS_CUDA(cudaMemcpyAsync(&nadj32, c->adjptr.p + nn, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  const int64_t nadj = nadj32;
  FS_TRY(c->adjoff.ensure((size_t)2 * nadj + 1));
  TLAUNCH(c, k_adj_offsets, nn, c->adjptr.p, c->adj.p, c->dof.p, c->nodeinfo.p, c->colptr.p, c->rowval.p, nn, c->pcols, nadj,
          c->adjoff.p);
  FS_CUDA(cudaStreamSynchronize(st));
  c->nadj = nadj;
  c->tile_no = NO;
  c->tile_cap = TILE_CAP;
  c->ntiles = ntiles;
  c->tile_ok = true;
  return FSGPU_OK;
}

// ------------------------------------------------------------------------------------
// numeric kernel
// ------------------------------------------------------------------------------------
struct TileArgs {
  const int32_t* morder;
  const int32_t* tel_ptr;
  const int32_t* tel;
  const int32_t* nel_ptr;
  const int32_t* nel;
  const int32_t* adjptr;
  const int32_t* adj;
  const int32_t* adjoff;
  const int32_t* nodeinfo;
  const int32_t* dof;
  const int32_t* colptr;
  int64_t nnodes, nadj, nc;
  double* nz;
};

template <bool COMP>
__global__ void __launch_bounds__(TILE_THREADS, 1) k_t3_tile(ShellArgs P, TileArgs T) {
  extern __shared__ double smem[];
  double* strips = smem;                                   // [cap][3][STRIP_LD]
  double* kd = strips + TILE_CAP * 3 * STRIP_LD;           // [cap][3][4]: kavg*valid*scale, g
  int* sel = reinterpret_cast<int*>(kd + TILE_CAP * 3 * 4);  // [cap] tile elements (ascending)
  int* scon = sel + TILE_CAP;                              // [cap][3]
  int* own = scon + TILE_CAP * 3;                          // [NO] owned node ids
  int* pfx = own + TILE_NO;                                // [NO+1] prefix of off-diagonal pair counts
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = TILE_THREADS / 32;
  const int64_t tile = blockIdx.x;
  const int e0 = T.tel_ptr[tile], nelt = T.tel_ptr[tile + 1] - e0;
  const int64_t n0 = tile * TILE_NO;
  const int nown = (int)((T.nnodes - n0) < TILE_NO ? (T.nnodes - n0) : TILE_NO);
  for (int k = tid; k < nelt; k += TILE_THREADS) {
    const int e = T.tel[e0 + k];
    sel[k] = e;
    scon[k * 3] = P.conn[(int64_t)e * 3];
    scon[k * 3 + 1] = P.conn[(int64_t)e * 3 + 1];
    scon[k * 3 + 2] = P.conn[(int64_t)e * 3 + 2];
  }
  if (tid < TILE_NO) own[tid] = tid < nown ? T.morder[n0 + tid] : -1;
  __syncthreads();
  if (warp == 0) {
    // exclusive prefix of (deg - 1) over the owned nodes (deg counts the node itself)
    int run = 0;
    for (int b0 = 0; b0 < TILE_NO; b0 += 32) {
      const int l = b0 + lane;
      int v = 0;
      if (l < nown) {
        const int a = own[l];
        const int dg = T.adjptr[a + 1] - T.adjptr[a];
        v = dg > 0 ? dg - 1 : 0;
      }
      int inc = v;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (l < TILE_NO) pfx[l] = run + inc - v;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) pfx[TILE_NO] = run;
  }

  // ---------------- phase 1: strips of every tile element (3 lanes per element) ----------------
  const unsigned full = 0xffffffffu;
  const int el = lane / 3, j = lane - 3 * el;
  const int base = lane - j;
  for (int s0 = warp * 10; s0 < nelt; s0 += nwarps * 10) {
    const int slot = s0 + el;
    const bool active = (lane < 30) && (slot < nelt);
    double kpart = 0.0;
    V3 gdir = v3(0, 0, 0);
    bool validj = false;
    double p1[5][3], p2[5][3], bs[2][3];
    double gx = 0.0, gy = 0.0;
    T3Geom g;
    M3 A;
    Constit C;
    int64_t e = 0;
    if (active) {
      e = sel[slot];
      const int n0_ = scon[slot * 3], n1_ = scon[slot * 3 + 1], n2_ = scon[slot * 3 + 2];
      g = t3_geometry(ld3(P.xyz, n0_), ld3(P.xyz, n1_), ld3(P.xyz, n2_));
      const double4 nv = ldg4(P.nrm + (j == 0 ? n0_ : (j == 1 ? n1_ : n2_)));
      validj = nv.w != 0.0;
      A = nodal_triad(g.E, v3(nv.x, nv.y, nv.z), validj);
      build_constit_t3(P, e, g.E, g.Ae, 1.0, COMP, C);
      gx = j == 0 ? g.gN[0][0] : (j == 1 ? g.gN[1][0] : g.gN[2][0]);
      gy = j == 0 ? g.gN[0][1] : (j == 1 ? g.gN[1][1] : g.gN[2][1]);
      t3_bs_node(g, j, -1, bs);
      node_coupling_contrib(A, gx, gy, bs, p1, p2);
    } else {
      for (int r = 0; r < 5; ++r)
        for (int k = 0; k < 3; ++k) p1[r][k] = p2[r][k] = 0.0;
    }
    double P1[5][3], P2[5][3];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          s1 += __shfl_sync(full, p1[r][k], (base + l) & 31);
          s2 += __shfl_sync(full, p2[r][k], (base + l) & 31);
        }
        P1[r][k] = s1;
        P2[r][k] = s2;
      }
    if (active) {
      double R[2][2], brn[5][2];
      node_R(A, R);
      node_bt_rot(gx, gy, bs, R, brn);
      kpart = node_kavg_part(C, brn, false);
      double bg[8][6];
      gdir = node_strip(g.E, A, gx, gy, bs, P1, P2, bg);
      fold_constit(C, bg);
      double* dst = strips + (slot * 3 + j) * STRIP_LD;
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const double d = constit_d(C, s);
        if (d < 0.0) atomicExch(P.flag + 2, 1);
        const double q = sqrt(d);
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) dst[s * 6 + cc] = q * bg[s][cc];
      }
    }
    double ksum = 0.0;
#pragma unroll
    for (int l = 0; l < 3; ++l) ksum += __shfl_sync(full, kpart, (base + l) & 31);
    if (active) {
      double* k4 = kd + (slot * 3 + j) * 4;
      k4[0] = validj ? ksum / 6 * P.drill : 0.0;
      k4[1] = gdir.x;
      k4[2] = gdir.y;
      k4[3] = gdir.z;
    }
  }
  __syncthreads();

  // ---------------- phase 2: complete blocks K[b, a], a owned; 2 lanes per pair ----------------
  // pair list: [diagonal pairs (one per owned node) padded to a multiple of 16] [off-diagonal pairs]
  const int ndiag = (nown + 15) & ~15;
  const int noff = pfx[TILE_NO];
  const int npairs = ndiag + noff;
  for (int w0 = 0; w0 < 2 * npairs; w0 += TILE_THREADS) {
    const int item = w0 + tid;
    const int pair = item >> 1, part = item & 1;
    int l = -1, a = -1, b = -1, p = -1;  // owned index, column node, row node, adjacency entry
    if (pair < ndiag) {
      if (pair < nown) {
        l = pair;
        a = own[l];
        b = a;
      }
    } else if (pair < npairs) {
      const int q = pair - ndiag;
      int lo = 0, hi = nown;  // last l with pfx[l] <= q
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (pfx[mid] <= q)
          lo = mid;
        else
          hi = mid;
      }
      l = lo;
      a = own[l];
      int k = q - pfx[l];  // k-th off-diagonal neighbour: skip the diagonal entry
      const int p0 = T.adjptr[a], p1_ = T.adjptr[a + 1];
      // adjacency is ascending and contains a itself: entries before a keep their index
      int pa = p0;
      {
        int lo2 = p0, hi2 = p1_;
        while (lo2 < hi2) {
          const int mid = (lo2 + hi2) >> 1;
          if (T.adj[mid] < a)
            lo2 = mid + 1;
          else
            hi2 = mid;
        }
        pa = lo2;
      }
      p = p0 + k;
      if (p >= pa) ++p;
      b = T.adj[p];
    }
    if (a >= 0 && b == a) {
      int lo2 = T.adjptr[a], hi2 = T.adjptr[a + 1];
      if (hi2 == lo2) {
        a = -1;  // node without elements: nothing stored
      } else {
        while (lo2 < hi2) {
          const int mid = (lo2 + hi2) >> 1;
          if (T.adj[mid] < a)
            lo2 = mid + 1;
          else
            hi2 = mid;
        }
        p = lo2;
      }
    }
    double acc[6][6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) acc[r][cc] = 0.0;
    if (a >= 0) {
      int hit = 0;
      for (int q = T.nel_ptr[a]; q < T.nel_ptr[a + 1]; ++q) {
        const int e = T.nel[q];
        // slot of e in the tile list (ascending)
        int lo2 = 0, hi2 = nelt;
        while (lo2 < hi2) {
          const int mid = (lo2 + hi2) >> 1;
          if (sel[mid] < e)
            lo2 = mid + 1;
          else
            hi2 = mid;
        }
        const int slot = lo2;
        const int c0 = scon[slot * 3], c1 = scon[slot * 3 + 1], c2 = scon[slot * 3 + 2];
        const int jj = c0 == a ? 0 : (c1 == a ? 1 : 2);
        const int ii = c0 == b ? 0 : (c1 == b ? 1 : (c2 == b ? 2 : -1));
        if (ii < 0) continue;
        if (((hit++) & 1) != part) continue;
        const double* bi = strips + (slot * 3 + ii) * STRIP_LD;
        const double* bj = strips + (slot * 3 + jj) * STRIP_LD;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          double vi[6], vj[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) vi[r] = bi[s * 6 + r];
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) vj[cc] = bj[s * 6 + cc];
#pragma unroll
          for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) acc[r][cc] = fma(vi[r], vj[cc], acc[r][cc]);
        }
        if (b == a) {
          const double* k4 = kd + (slot * 3 + jj) * 4;
          const double kv = k4[0];
          const double gg[3] = {k4[1], k4[2], k4[3]};
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) acc[3 + r][3 + cc] += kv * gg[r] * gg[cc];
        }
      }
    }
    // combine the two lanes of the pair
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) acc[r][cc] += __shfl_xor_sync(full, acc[r][cc], 1);
    if (a >= 0 && part == 0) {
      const int inf = T.nodeinfo[b];
      const int mA = inf & 63, mB = (inf >> 8) & 63;
      const int oA = T.adjoff[p], oB = T.adjoff[T.nadj + p];
      const int32_t* da = T.dof + (int64_t)a * 6;
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        const int cd = da[cc];
        if (cd >= T.nc) continue;
        double* col = T.nz + T.colptr[cd];
        int ka = 0, kb = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          if ((mA >> r) & 1)
            col[oA + ka++] = acc[r][cc];
          else if ((mB >> r) & 1)
            col[oB + kb++] = acc[r][cc];
        }
      }
    }
  }
}

int launch_t3_tile(fsgpu_ctx* c, const ShellArgs& A, bool comp) {
  TileArgs T;
  T.morder = c->morder.p;
  T.tel_ptr = c->tel_ptr.p;
  T.tel = c->tel.p;
  T.nel_ptr = c->nel_ptr.p;
  T.nel = c->nel.p;
  T.adjptr = c->adjptr.p;
  T.adj = c->adj.p;
  T.adjoff = c->adjoff.p;
  T.nodeinfo = c->nodeinfo.p;
  T.dof = c->dof.p;
  T.colptr = c->colptr.p;
  T.nnodes = c->nnodes;
  T.nadj = c->nadj;
  T.nc = c->pcols;
  T.nz = c->nzval.p;
  const size_t sm = (size_t)(TILE_CAP * 3 * STRIP_LD + TILE_CAP * 3 * 4) * sizeof(double) +
                    (size_t)(TILE_CAP + TILE_CAP * 3 + TILE_NO + TILE_NO + 1 + 3) * sizeof(int);
  if (comp) {
    FS_CUDA(cudaFuncSetAttribute(k_t3_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_t3_tile<true><<<(unsigned)c->ntiles, TILE_THREADS, sm, c->stream>>>(A, T);
  } else {
    FS_CUDA(cudaFuncSetAttribute(k_t3_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_t3_tile<false><<<(unsigned)c->ntiles, TILE_THREADS, sm, c->stream>>>(A, T);
  }
  c->launches++;
  FS_CUDA(cudaGetLastError());
  return FSGPU_OK;
}

}  // namespace fsk
