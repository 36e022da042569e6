// Internal definitions of libfsgpu (context, device buffers, error plumbing).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../../include/fsgpu.h"

namespace fs {

// ---- error handling ---------------------------------------------------------------
void set_error(const char* fmt, ...);

#define FS_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      fs::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__,    \
                    cudaGetErrorString(e_));                                                   \
      return FSGPU_ERR_CUDA;                                                                   \
    }                                                                                          \
  } while (0)

#define FS_TRY(call)            \
  do {                          \
    int rc_ = (call);           \
    if (rc_ != FSGPU_OK) return rc_; \
  } while (0)

#define FS_REQUIRE(cond, code, ...) \
  do {                              \
    if (!(cond)) {                  \
      fs::set_error(__VA_ARGS__);   \
      return (code);                \
    }                               \
  } while (0)

// ---- simple owning device buffer ---------------------------------------------------
template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;  // capacity in elements
  int ensure(size_t count) {
    if (count <= n && p) return FSGPU_OK;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    if (count == 0) return FSGPU_OK;
    FS_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
    n = count;
    return FSGPU_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DBuf() { release(); }
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
};

constexpr int kMaxGP = 9;  // Simpson13Rule(2) has 9 points

struct Rule {
  int npts = 0;
  double xi[kMaxGP], eta[kMaxGP], w[kMaxGP];
};
// built-in csys kind (FSGPU_CSYS_*; 0 = none) with its origin and unit axis
struct CsysK {
  int kind = 0;
  double o[3] = {0, 0, 0}, a[3] = {0, 0, 1};
};

}  // namespace fs

// The context: device-resident mesh, fields, pattern and results.
struct fsgpu_ctx {
  int device = 0;
  cudaStream_t stream = 0;
  int64_t launches = 0;
  int64_t d2h_bytes = 0;  // bytes moved device -> host by this context so far (measurement support)
  // device time of the dominant kernel of the last operator (CUDA events on `stream`)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  // fetch of large results: two pinned staging buffers for the compact row indices, second stream for the values
  void* ring[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ring_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t stream2 = nullptr;
  // values into a PAGEABLE host array: pinned staging ring, moved on by host threads
  void* vring[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t vring_ev[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_x = nullptr;
  // run-length form of the row indices of the current pattern (fetch of large results)
  const int32_t* rle_for = nullptr;  // device array it was built from
  int64_t rle_nnz = 0, rle_nruns = 0;  // rle_nruns == 0: not smaller than int32 entries
  fs::DBuf<unsigned char> rle_runs;  // [nruns + 1] int2 (position, first row), sentinel (nnz, 0)
  fs::DBuf<unsigned char> scr_rle_flag;
  fs::DBuf<int32_t> scr_rle_pos;

  // mesh
  int nnpe = 0;
  int64_t nelem = 0, nnodes = 0;
  fs::DBuf<int32_t> conn;     // [nelem][nnpe] 0-based
  fs::DBuf<double4> xyz;      // (x, y, z, 0) per node: one 32 B sector per gather
  // dofs
  int64_t nfree = 0, nall = 0;
  bool have_dofs = false;
  fs::DBuf<int32_t> dof;      // [nnodes][6] 0-based
  // normals
  bool associated = false;
  fs::DBuf<double4> nrm;      // (nx, ny, nz, valid ? 1 : 0)
  fs::DBuf<double> nacc;      // [nnodes][3] unnormalised normal sums (associategeometry, split form)
  bool nacc_keep = false;
  const double* ndirs = nullptr;  // per (element, node) csys normal directions during fsgpu_associategeometry_dirs
  fs::CsysK ncsys;                // built-in csys kind during fsgpu_associategeometry_csys
  // thickness / stab factor
  int64_t nthick = 0;
  fs::DBuf<double> thick;
  int64_t nstab = 0;
  fs::DBuf<double> stabf;
  fs::Rule rule;
  // layup
  int ngroups = 0;
  fs::DBuf<double> group_data;  // [ngroups][34]
  fs::DBuf<int32_t> group_of;   // [nelem] 0-based
  int64_t ncs = 0;
  fs::DBuf<double> csmat;       // [ncs][9] ROW-major
  fs::DBuf<double> lam_prep;    // Q4RSComp: [nelem][npts][24] factored constitutive data of the current operator call
  // beam
  bool have_sections = false;
  fs::DBuf<double> sec;         // [nelem][10]: A I1 I2 I3 J A2s A3s x y z
  bool have_state = false;
  fs::DBuf<double4> u1;         // (ux,uy,uz,0)
  fs::DBuf<double> R1;          // [nnodes][9] each column-major 3x3
  bool have_velocity = false;
  fs::DBuf<double> v1;          // [nnodes][6] generalized velocities (gyroscopic)

  // symbolic
  int target = -1;
  int64_t prows = 0, pcols = 0, pnnz = 0;  // pattern (before SYMM compaction)
  fs::DBuf<int32_t> colptr;   // [pcols+1] 0-based
  fs::DBuf<int32_t> rowval;   // [pnnz] 0-based
  fs::DBuf<int32_t> slot;     // [36][nnpe][nelem][nnpe]  (-1 = dropped)
  fs::DBuf<int32_t> diagslot; // [nall] slot of (d,d) or -1
  // run-structured addressing (fast path): the included dofs of every node form <= 2
  // consecutive ascending runs (free / prescribed), as FinEtools' numberdofs! produces
  bool fast = false;
  fs::DBuf<int32_t> nodecol;  // fast path: [nnodes][8] column starts of the node's dofs (-1 none), nodeinfo, 0
  fs::DBuf<int32_t> nodeinfo; // [nnodes] bits 0-5 run-A mask, bits 8-13 run-B mask
  fs::DBuf<int32_t> pairoff;  // [nnpe(i)][2][nelem][nnpe(j)] row offset of node i's runs in node j's columns
  // T3 fast path: per-warp emission plan (unique matrix blocks of the warp's ten elements with their contributors, sorted by
  // column node; fsk::t3_build_plan).  Depends on the connectivity only.
  fs::DBuf<unsigned> t3_plan;  // [nwarps][92]
  bool t3_plan_ok = false;
  // node adjacency (kept from the symbolic phase) and the T3 tile (owner-computes) data
  fs::DBuf<int32_t> adjptr, adj;  // CSR over nodes, neighbours ascending by node id
  // scratch of the symbolic phase and of fetch_matrix, kept between calls: cudaMalloc/cudaFree of
  // 100+ MB blocks costs milliseconds and made the end-to-end time vary from call to call
  fs::DBuf<uint64_t> scr_keys, scr_keys2;
  fs::DBuf<int32_t> scr_deg, scr_nodecnt;
  fs::DBuf<int64_t> scr_colcnt, scr_colptr64, scr_nsel, scr_wide;
  bool tile_ok = false;
  bool want_tile = false;         // fsgpu_set_deterministic / FSGPU_TILE=1: prefer the tile kernel
  int last_path = -1;             // scatter path of the last matrix operator (fsgpu_scatter_path)
  int64_t nadj = 0;               // entries of adj
  int tile_no = 0, tile_cap = 0;  // owned nodes per tile, max elements per tile
  int tile_cfg = 0;               // index into fsk::kTileCfg
  fs::DBuf<unsigned char> tile_items;  // [ntiles][nw*32] fsk::TileItem: phase-2 schedule of the tile kernel
  int64_t ntiles = 0;
  fs::DBuf<int32_t> morder;       // [nnodes] node ids in Morton order
  fs::DBuf<int32_t> nel_ptr, nel; // node -> incident elements (ascending)
  fs::DBuf<int32_t> tel_ptr, tel; // tile -> elements touching its owned nodes (ascending)
  fs::DBuf<int32_t> adjoff;       // [2][nadj]: row offsets of the neighbour's runs in the node's columns
  // order-fixed assembly (fsgpu_set_deterministic, gather path): contributions per matrix block, dense element matrices
  bool det_ready = false;
  int64_t det_nblocks = 0, det_ncontrib = 0;
  fs::DBuf<int32_t> det_bstart, det_contrib;
  fs::DBuf<double> det_dense;
  // result matrix
  bool have_matrix = false;
  int64_t rrows = 0, rcols = 0, rnnz = 0;
  fs::DBuf<double> nzval;     // [pnnz]
  bool compacted = false;     // SPARSE_SYMM after zero dropping: use c_* below
  fs::DBuf<int32_t> c_colptr, c_rowval;
  fs::DBuf<double> c_nzval;
  // one triangle of the result (fsgpu_fetch_matrix_uplo)
  fs::DBuf<int32_t> t_cnt, t_first, t_colptr, t_rowval;
  fs::DBuf<double> t_nzval;
  // COO -> CSC: result of the size query, kept for the fill call (two-call convention)
  const void* coo_key[3] = {nullptr, nullptr, nullptr};
  int64_t coo_dims[4] = {0, 0, 0, 0};
  fs::DBuf<int64_t> coo_row, coo_ptr, coo_I, coo_J, coo_start;
  fs::DBuf<double> coo_val, coo_V, coo_V2;
  // result vector
  bool have_vector = false;
  int64_t vlen = 0;
  fs::DBuf<double> vec;
  // scratch
  fs::DBuf<unsigned char> tmp, tmp2;
  fs::DBuf<int32_t> flag;     // device error flags
};

namespace fs {
int check_ctx(fsgpu_ctx* c);
inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }
// staging helpers
int upload(fsgpu_ctx* c, void* dst, const void* src, size_t bytes);
int download(fsgpu_ctx* c, void* dst, const void* src, size_t bytes);
// after an operator has filled nzval: SPARSE_SYMM symmetrisation + zero dropping, bookkeeping
int finalize_matrix(fsgpu_ctx* c);
int time_begin(fsgpu_ctx* c);
int time_end(fsgpu_ctx* c);
// CSC (0-based, device) -> CSR of the same matrix, by sorting entries on (row, col)
int csc_to_csr(fsgpu_ctx* c, const int32_t* colptr, const int32_t* rowval, const double* nz, int64_t nrows,
               int64_t ncols, int64_t nnz, DBuf<int32_t>& rowptr, DBuf<int32_t>& colval, DBuf<double>& val);
}  // namespace fs
