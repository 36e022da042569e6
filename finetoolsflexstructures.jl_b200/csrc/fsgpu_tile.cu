// T3FF / T3FFComp stiffness, OWNER-COMPUTES tile kernel: no atomics, no clearing of the value
// array, every stored entry written exactly once (deterministic).
//
// Why: the scatter kernels are capped by the L2 atomic unit (RED.ADD.F64: ~266 G elements/s
// measured, scripts/micro/red_bench.cu) -- 324 REDs per T3 element put a 4.9 ms floor under the
// 4M-element mesh whereas HBM needs 0.7 ms.  Here a CTA owns a compact patch of NODES (consecutive
// in Morton order of the coordinates).
//   phase 1  builds the folded global-dof strips of every element touching an owned node (halo
//            elements are recomputed by the neighbouring tiles, ~1.5x setup work) in shared memory;
//   phase 2  runs a SCHEDULE built once per mesh by the symbolic phase: one lane per work item
//            (<= 2 element contributions to one block K[b, a], a owned), 32 items per warp, the
//            items of a block adjacent and never split between warps -> perfectly regular products;
//   phase 3  stages the partial blocks in shared memory (over the dead strips), sums the items of
//            every block in a fixed order and writes the block with plain stores, six consecutive
//            lanes covering six consecutive rows of a column (run-structured addressing, fsgpu_core.cu).
//
// Requirements (checked in tile_symbolic, otherwise the RED kernel is used): T3 mesh, fast-path
// numbering, non-diagonal target, AVERAGE_B shear, every tile fits the shared-memory / item capacity.
#include <cub/cub.cuh>

#include "fsgpu_shell.cuh"

using namespace fs;
using namespace fsm;

namespace fsk {

struct TileCfg {
  int no;   // owned nodes per tile
  int cap;  // max elements per tile
  int nw;   // warps per CTA (= 32-item chunks of the schedule per tile)
};
// cfg 0: two CTAs per SM (2 x 105 KB); cfg 1: one CTA per SM (207 KB), less halo recomputation
constexpr TileCfg kTileCfg[2] = {{16, 80, 6}, {32, 160, 12}};
constexpr int STRIP_LD = 50;     // doubles per (element, node) strip: 48 + 2 (16 B aligned rows; 8 consecutive strips
                                 // are conflict-free for 128-bit loads: 400 B = 100 banks = 4 mod 32)
constexpr int STAGE_LD = 38;     // doubles per staged 6x6 block (16 B aligned, conflict-free 128-bit stores)
constexpr int WARP_STAGE = 32 * STAGE_LD + 64 + 16;  // + info[32] (int4) + plist[32] (int), in doubles
constexpr int TILE_MAXDEG = 24;  // max node valence handled

// one work item of the phase-2 schedule (24 B).  Strip indices are 1 + (slot * 3 + local node), 0 = none.
struct TileItem {
  uint16_t s[4];  // (row strip, column strip) of contribution 0, of contribution 1
  int32_t oA, oB; // first item of a block: row offsets of node b's runs inside node a's columns
  int32_t inf;    // first item of a block: nodeinfo[b] | number of items << 16; 0 otherwise
  int32_t a;      // first item of a block: column node
};
static_assert(sizeof(TileItem) == 24, "TileItem layout");

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

// ---- phase-2 schedule --------------------------------------------------------------------------
// pcnt[p]: number of elements shared by node a and its neighbour adj[p] (all elements of a for p = a itself)
__global__ void k_pair_counts(const int32_t* __restrict__ adjptr, const int32_t* __restrict__ adj,
                              const int32_t* __restrict__ nel_ptr, const int32_t* __restrict__ nel,
                              const int32_t* __restrict__ conn, int64_t nnodes, int32_t* __restrict__ pcnt) {
  const int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a >= nnodes) return;
  for (int p = adjptr[a]; p < adjptr[a + 1]; ++p) {
    const int b = adj[p];
    int n = 0;
    for (int q = nel_ptr[a]; q < nel_ptr[a + 1]; ++q) {
      const int32_t* cn = conn + (int64_t)nel[q] * 3;
      n += (cn[0] == b || cn[1] == b || cn[2] == b) ? 1 : 0;
    }
    pcnt[p] = n;
  }
}
// one thread per tile: lay the blocks (a owned, b neighbour) out in 32-item chunks; a block's items
// (ceil(shared elements / 2)) are adjacent and never straddle a chunk.  itemoff[p]: tile-relative first item.
__global__ void k_tile_layout(const int32_t* __restrict__ morder, const int32_t* __restrict__ adjptr,
                              const int32_t* __restrict__ pcnt, int64_t nnodes, int no, int64_t ntiles,
                              int32_t* __restrict__ itemoff, int32_t* __restrict__ tile_items) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  int pos = 0;
  for (int l = 0; l < no; ++l) {
    const int64_t m = t * no + l;
    if (m >= nnodes) break;
    const int a = morder[m];
    for (int p = adjptr[a]; p < adjptr[a + 1]; ++p) {
      const int n = (pcnt[p] + 1) >> 1;
      if (n == 0) {
        itemoff[p] = -1;
        continue;
      }
      if ((pos & 31) + n > 32) pos = (pos + 31) & ~31;
      itemoff[p] = pos;
      pos += n;
    }
  }
  tile_items[t] = pos;
}
// one thread per node a: write the items of its blocks into its tile's schedule
__global__ void k_fill_items(const int32_t* __restrict__ mrank, const int32_t* __restrict__ adjptr,
                             const int32_t* __restrict__ adj, const int32_t* __restrict__ adjoff, int64_t nadj,
                             const int32_t* __restrict__ nodeinfo, const int32_t* __restrict__ nel_ptr,
                             const int32_t* __restrict__ nel, const int32_t* __restrict__ conn,
                             const int32_t* __restrict__ tel_ptr, const int32_t* __restrict__ tel,
                             const int32_t* __restrict__ itemoff, const int32_t* __restrict__ pcnt, int64_t nnodes, int no,
                             int cap_items, TileItem* __restrict__ items) {
  const int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a >= nnodes) return;
  const int64_t t = mrank[a] / no;
  const int e0 = tel_ptr[t], net = tel_ptr[t + 1] - e0;
  TileItem* base = items + t * cap_items;
  for (int p = adjptr[a]; p < adjptr[a + 1]; ++p) {
    if (itemoff[p] < 0) continue;
    const int b = adj[p];
    TileItem* it = base + itemoff[p];
    int k = 0;
    for (int q = nel_ptr[a]; q < nel_ptr[a + 1]; ++q) {
      const int e = nel[q];
      const int32_t* cn = conn + (int64_t)e * 3;
      const int jj = cn[0] == a ? 0 : (cn[1] == a ? 1 : 2);
      const int ii = cn[0] == b ? 0 : (cn[1] == b ? 1 : (cn[2] == b ? 2 : -1));
      if (ii < 0) continue;
      int lo = 0, hi = net;  // slot of e in the tile's element list (ascending)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tel[e0 + mid] < e)
          lo = mid + 1;
        else
          hi = mid;
      }
      it[k >> 1].s[(k & 1) * 2] = (uint16_t)(1 + lo * 3 + ii);
      it[k >> 1].s[(k & 1) * 2 + 1] = (uint16_t)(1 + lo * 3 + jj);
      ++k;
    }
    it[0].oA = adjoff[p];
    it[0].oB = adjoff[nadj + p];
    it[0].inf = (nodeinfo[b] & 0xffff) | (((pcnt[p] + 1) >> 1) << 16);
    it[0].a = (int32_t)a;
  }
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
  int cfg = 0;
  if (const char* ev = getenv("FSGPU_TILE_CFG")) cfg = ev[0] == '1' ? 1 : 0;
  const TileCfg tc = kTileCfg[cfg];
  const int NO = tc.no;
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
  if (maxel > tc.cap || maxdeg + 1 > TILE_MAXDEG) return FSGPU_OK;  // keep the RED kernel
  // (4) per-adjacency row offsets
  int32_t nadj32 = 0;
  FS_CUDA(cudaMemcpyAsync(&nadj32, c->adjptr.p + nn, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  const int64_t nadj = nadj32;
  FS_TRY(c->adjoff.ensure((size_t)2 * nadj + 1));
  TLAUNCH(c, k_adj_offsets, nn, c->adjptr.p, c->adj.p, c->dof.p, c->nodeinfo.p, c->colptr.p, c->rowval.p, nn, c->pcols, nadj,
          c->adjoff.p);
  // (5) phase-2 schedule
  const int cap_items = tc.nw * 32;
  DBuf<int32_t> pcnt, itemoff, titems, imax;
  FS_TRY(pcnt.ensure((size_t)nadj + 1));
  FS_TRY(itemoff.ensure((size_t)nadj + 1));
  FS_TRY(titems.ensure((size_t)ntiles + 1));
  FS_TRY(imax.ensure(1));
  TLAUNCH(c, k_pair_counts, nn, c->adjptr.p, c->adj.p, c->nel_ptr.p, c->nel.p, c->conn.p, nn, pcnt.p);
  TLAUNCH(c, k_tile_layout, ntiles, c->morder.p, c->adjptr.p, pcnt.p, nn, NO, ntiles, itemoff.p, titems.p);
  FS_CUDA(cudaMemsetAsync(imax.p, 0, sizeof(int32_t), st));
  TLAUNCH(c, k_max_i32, ntiles, titems.p, ntiles, imax.p);
  int32_t maxitems = 0;
  FS_CUDA(cudaMemcpyAsync(&maxitems, imax.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  if (maxitems > cap_items) return FSGPU_OK;  // keep the RED kernel
  FS_TRY(c->tile_items.ensure((size_t)ntiles * cap_items * sizeof(TileItem) + 16));
  FS_CUDA(cudaMemsetAsync(c->tile_items.p, 0, (size_t)ntiles * cap_items * sizeof(TileItem), st));
  TLAUNCH(c, k_fill_items, nn, mrank.p, c->adjptr.p, c->adj.p, c->adjoff.p, nadj, c->nodeinfo.p, c->nel_ptr.p, c->nel.p,
          c->conn.p, c->tel_ptr.p, c->tel.p, itemoff.p, pcnt.p, nn, NO, cap_items, reinterpret_cast<TileItem*>(c->tile_items.p));
  FS_CUDA(cudaStreamSynchronize(st));
  c->nadj = nadj;
  c->tile_no = NO;
  c->tile_cap = tc.cap;
  c->tile_cfg = cfg;
  c->ntiles = ntiles;
  c->tile_ok = true;
  return FSGPU_OK;
}

// ------------------------------------------------------------------------------------
// numeric kernel
// ------------------------------------------------------------------------------------
struct TileArgs {
  const int32_t* tel_ptr;
  const int32_t* tel;
  const TileItem* items;
  const int32_t* nodecol;
  int cap;  // element capacity of the shared-memory layout
  double* nz;
};

// position (relative to the column start) of dof r of a node with run masks `inf` and run offsets oA / oB
__device__ __forceinline__ int tile_row_pos(int inf, int oA, int oB, int r) {
  const int mA = inf & 63, mB = (inf >> 8) & 63, below = (1 << r) - 1;
  if ((mA >> r) & 1) return oA >= 0 ? oA + __popc(mA & below) : -1;
  if ((mB >> r) & 1) return oB >= 0 ? oB + __popc(mB & below) : -1;
  return -1;
}

template <int NW, bool COMP>
__global__ void __launch_bounds__(NW * 32, NW <= 6 ? 2 : 1) k_t3_tile(ShellArgs P, TileArgs T) {
  extern __shared__ __align__(16) double smem[];
  const int cap = T.cap;
  double* strips = smem;                                   // [cap][3][STRIP_LD]; phase 3: per-warp staging
  double* kd = strips + cap * 3 * STRIP_LD;                // [cap][3][4]: kavg*valid*scale, g
  int* sel = reinterpret_cast<int*>(kd + cap * 3 * 4);     // [cap] tile elements (ascending)
  int* scon = sel + cap;                                   // [cap][3]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tile = blockIdx.x;
  const int e0 = T.tel_ptr[tile], nelt = T.tel_ptr[tile + 1] - e0;
  for (int k = tid; k < nelt; k += NW * 32) {
    const int e = T.tel[e0 + k];
    sel[k] = e;
    scon[k * 3] = P.conn[(int64_t)e * 3];
    scon[k * 3 + 1] = P.conn[(int64_t)e * 3 + 1];
    scon[k * 3 + 2] = P.conn[(int64_t)e * 3 + 2];
  }
  // this lane's work item (consumed after phase 1)
  TileItem it;
  {
    const int2* src = reinterpret_cast<const int2*>(T.items + (tile * NW + warp) * 32 + lane);
    const int2 w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
    it.s[0] = (uint16_t)(w0.x & 0xffff);
    it.s[1] = (uint16_t)((unsigned)w0.x >> 16);
    it.s[2] = (uint16_t)(w0.y & 0xffff);
    it.s[3] = (uint16_t)((unsigned)w0.y >> 16);
    it.oA = w1.x;
    it.oB = w1.y;
    it.inf = w2.x;
    it.a = w2.y;
  }
  __syncthreads();

  // ---------------- phase 1: strips of every tile element (3 lanes per element) ----------------
  const unsigned full = 0xffffffffu;
  {
    const int el = lane / 3, j = lane - 3 * el;
    const int base = lane - j;
    for (int s0 = warp * 10; s0 < nelt; s0 += NW * 10) {
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
      if (active) {
        const int64_t e = sel[slot];
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
          for (int cc = 0; cc < 6; cc += 2)
            *reinterpret_cast<double2*>(dst + s * 6 + cc) = make_double2(q * bg[s][cc], q * bg[s][cc + 1]);
        }
      }
      double ksum = 0.0;
#pragma unroll
      for (int l = 0; l < 3; ++l) ksum += __shfl_sync(full, kpart, (base + l) & 31);
      if (active) {
        double* k4 = kd + (slot * 3 + j) * 4;
        *reinterpret_cast<double2*>(k4) = make_double2(validj ? ksum / 6 * P.drill : 0.0, gdir.x);
        *reinterpret_cast<double2*>(k4 + 2) = make_double2(gdir.y, gdir.z);
      }
    }
  }
  __syncthreads();

  // ---------------- phase 2: this lane's item = up to two contributions b_i' b_j to one block ----------------
  double acc[6][6];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) acc[r][cc] = 0.0;
#pragma unroll 1
  for (int t = 0; t < 2; ++t) {
    const int si = it.s[2 * t], sj = it.s[2 * t + 1];
    if (si == 0) break;
    const double* bi = strips + (si - 1) * STRIP_LD;
    const double* bj = strips + (sj - 1) * STRIP_LD;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (!COMP && s < 3) {
        // membrane rows of a homogeneous shell have no rotation columns in global dofs
        const double2 i01 = *reinterpret_cast<const double2*>(bi + s * 6);
        const double2 j01 = *reinterpret_cast<const double2*>(bj + s * 6);
        const double vi[3] = {i01.x, i01.y, bi[s * 6 + 2]};
        const double vj[3] = {j01.x, j01.y, bj[s * 6 + 2]};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) acc[r][cc] = fma(vi[r], vj[cc], acc[r][cc]);
      } else {
        double vi[6], vj[6];
#pragma unroll
        for (int r = 0; r < 6; r += 2) {
          const double2 a2 = *reinterpret_cast<const double2*>(bi + s * 6 + r);
          const double2 b2 = *reinterpret_cast<const double2*>(bj + s * 6 + r);
          vi[r] = a2.x;
          vi[r + 1] = a2.y;
          vj[r] = b2.x;
          vj[r + 1] = b2.y;
        }
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) acc[r][cc] = fma(vi[r], vj[cc], acc[r][cc]);
      }
    }
    if (si == sj) {
      // drilling stiffness kavg on the nodal normal direction (nodal dof 6), rotated to global
      const double2 k01 = *reinterpret_cast<const double2*>(kd + (sj - 1) * 4);
      const double2 k23 = *reinterpret_cast<const double2*>(kd + (sj - 1) * 4 + 2);
      const double gg[3] = {k01.y, k23.x, k23.y};
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) acc[3 + r][3 + cc] += k01.x * gg[r] * gg[cc];
    }
  }
  __syncthreads();  // every warp is done with the strips

  // ---------------- phase 3: stage, sum the items of each block, store ----------------
  double* stage = smem + (size_t)warp * WARP_STAGE;
  int4* info = reinterpret_cast<int4*>(stage + 32 * STAGE_LD);
  int* plist = reinterpret_cast<int*>(stage + 32 * STAGE_LD + 64);
  {
    double2* st2 = reinterpret_cast<double2*>(stage + lane * STAGE_LD);
#pragma unroll
    for (int cc = 0; cc < 6; ++cc)
#pragma unroll
      for (int r = 0; r < 6; r += 2) st2[(cc * 6 + r) >> 1] = make_double2(acc[r][cc], acc[r + 1][cc]);
  }
  info[lane] = make_int4(it.oA, it.oB, it.inf, it.a);
  const bool first = ((it.inf >> 16) & 15) > 0;
  const unsigned firsts = __ballot_sync(full, first);
  if (first) plist[__popc(firsts & ((1u << lane) - 1))] = lane;
  const int npairs = __popc(firsts);
  __syncwarp();
  const int sub = lane / 6, r = lane - sub * 6;
#pragma unroll 1
  for (int g = 0; g * 5 < npairs; ++g) {
    const int k = g * 5 + sub;
    if (lane >= 30 || k >= npairs) continue;
    const int o = plist[k];
    const int4 I = info[o];
    const int rp = tile_row_pos(I.z, I.x, I.y, r);
    if (rp < 0) continue;
    const int nit = (I.z >> 16) & 15;
    const int4 c0 = __ldg(reinterpret_cast<const int4*>(T.nodecol + (int64_t)I.w * 8));
    const int2 c1 = __ldg(reinterpret_cast<const int2*>(T.nodecol + (int64_t)I.w * 8 + 4));
    const int cb[6] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y};
    const double* sv = stage + o * STAGE_LD + r;
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) {
      if (cb[cc] < 0) continue;
      double v = sv[cc * 6];
      for (int t = 1; t < nit; ++t) v += sv[t * STAGE_LD + cc * 6];
      T.nz[cb[cc] + rp] = v;
    }
  }
}

template <int NW>
static int launch_tile_nw(fsgpu_ctx* c, const ShellArgs& A, const TileArgs& T, bool comp, size_t sm) {
  if (comp) {
    FS_CUDA(cudaFuncSetAttribute(k_t3_tile<NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_t3_tile<NW, true><<<(unsigned)c->ntiles, NW * 32, sm, c->stream>>>(A, T);
  } else {
    FS_CUDA(cudaFuncSetAttribute(k_t3_tile<NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_t3_tile<NW, false><<<(unsigned)c->ntiles, NW * 32, sm, c->stream>>>(A, T);
  }
  return FSGPU_OK;
}

int launch_t3_tile(fsgpu_ctx* c, const ShellArgs& A, bool comp) {
  const TileCfg tc = kTileCfg[c->tile_cfg];
  TileArgs T;
  T.tel_ptr = c->tel_ptr.p;
  T.tel = c->tel.p;
  T.items = reinterpret_cast<const TileItem*>(c->tile_items.p);
  T.nodecol = c->nodecol.p;
  T.cap = tc.cap;
  T.nz = c->nzval.p;
  const size_t sm = (size_t)(tc.cap * 3 * STRIP_LD + tc.cap * 3 * 4) * sizeof(double) + (size_t)(tc.cap * 4) * sizeof(int);
  // the per-warp staging areas of phase 3 live inside the (dead) strip area
  static_assert(kTileCfg[0].nw * WARP_STAGE <= kTileCfg[0].cap * 3 * STRIP_LD, "stage does not fit (cfg 0)");
  static_assert(kTileCfg[1].nw * WARP_STAGE <= kTileCfg[1].cap * 3 * STRIP_LD, "stage does not fit (cfg 1)");
  if (c->tile_cfg == 0)
    FS_TRY(launch_tile_nw<kTileCfg[0].nw>(c, A, T, comp, sm));
  else
    FS_TRY(launch_tile_nw<kTileCfg[1].nw>(c, A, T, comp, sm));
  c->launches++;
  FS_CUDA(cudaGetLastError());
  return FSGPU_OK;
}

}  // namespace fsk
