// libfsgpu core: context, data hand-over, nodal normals, symbolic phase (CSC pattern
// bit-exact to Julia `sparse` + element-entry slot map), result hand-back, COO->CSC.
#include <stdarg.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include <atomic>
#include <thread>
#include <vector>

#include <cub/cub.cuh>

#include "fsgpu_internal.cuh"
#include "fsgpu_math.cuh"
#include "fsgpu_shell.cuh"

namespace fs {

static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}

int check_ctx(fsgpu_ctx* c) {
  FS_REQUIRE(c != nullptr, FSGPU_ERR_ARG, "null context");
  FS_CUDA(cudaSetDevice(c->device));
  return FSGPU_OK;
}

int time_begin(fsgpu_ctx* c) {
  if (!c->ev0) {
    FS_CUDA(cudaEventCreate(&c->ev0));
    FS_CUDA(cudaEventCreate(&c->ev1));
  }
  FS_CUDA(cudaEventRecord(c->ev0, c->stream));
  return FSGPU_OK;
}
int time_end(fsgpu_ctx* c) {
  FS_CUDA(cudaEventRecord(c->ev1, c->stream));
  c->timed = true;
  return FSGPU_OK;
}

constexpr int kVRing = 3;
constexpr int64_t kVRingWords = (int64_t)4 << 20;  // 32 MB per staging buffer
static bool host_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}
static int value_threads() {
  // the host copy is the slower half of this path (16 vCPUs: 4 threads 98 ms, 8 threads 73 ms for 2.6 GB)
  int nth = (int)std::thread::hardware_concurrency() / 2;
  if (const char* lw = getenv("LOCAL_WORLD_SIZE")) {  // one process per GPU on this host: share the cores
    const int w = atoi(lw);
    if (w > 1) nth /= w;
  }
  if (nth > 8) nth = 8;
  if (const char* ev = getenv("FSGPU_VALUE_THREADS")) nth = atoi(ev);
  return nth < 1 ? 1 : (nth > 32 ? 32 : nth);
}
static cudaError_t ensure_vring(fsgpu_ctx* c) {
  cudaError_t err = cudaSuccess;
  for (int k = 0; k < kVRing && err == cudaSuccess; ++k)
    if (!c->vring[k]) {
      err = cudaMallocHost(&c->vring[k], (size_t)kVRingWords * 8);
      if (err == cudaSuccess) err = cudaEventCreateWithFlags(&c->vring_ev[k], cudaEventDisableTiming);
    }
  return err;
}
int upload(fsgpu_ctx* c, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return FSGPU_OK;
  // (pageable sources: the driver's own staging costs 2.5 - 7 ms for C2's 129 MB of inputs; a pinned ring filled by host
  // threads, as on the way back, measured 7 - 9 ms and was not kept)
  FS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return FSGPU_OK;
}
int download(fsgpu_ctx* c, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return FSGPU_OK;
  c->d2h_bytes += (int64_t)bytes;
  FS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  return FSGPU_OK;
}

// ------------------------------------------------------------------------------------
// conversion kernels (Julia layout -> device layout)
// ------------------------------------------------------------------------------------
__global__ void k_conn_convert(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t n, int64_t nnodes,
                               int32_t* __restrict__ flag) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t v = in[i];
  if (v < 1 || v > nnodes) {
    atomicExch(flag, 1);
    v = 1;
  }
  out[i] = (int32_t)(v - 1);
}
__global__ void k_pack3(const double* __restrict__ in, double4* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = make_double4(in[i], in[n + i], in[2 * n + i], 0.0);
}
__global__ void k_pack_normals(const double* __restrict__ in, const unsigned char* __restrict__ valid,
                               double4* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = make_double4(in[i], in[n + i], in[2 * n + i], valid[i] ? 1.0 : 0.0);
}
__global__ void k_unpack_normals(const double4* __restrict__ in, double* __restrict__ out, unsigned char* __restrict__ valid,
                                 int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 v = in[i];
  out[i] = v.x;
  out[n + i] = v.y;
  out[2 * n + i] = v.z;
  valid[i] = v.w != 0.0 ? 1 : 0;
}
__global__ void k_dof_convert(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t nnodes, int64_t nall,
                              int32_t* __restrict__ flag) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // i = node*6 + d
  if (i >= nnodes * 6) return;
  int64_t node = i / 6, d = i % 6;
  int64_t v = in[d * nnodes + node];
  if (v < 1 || v > nall) {
    atomicExch(flag, 1);
    v = 1;
  }
  out[i] = (int32_t)(v - 1);
}
__global__ void k_transpose_rows(const double* __restrict__ in, double* __restrict__ out, int64_t nrows, int ncols) {
  // column-major nrows x ncols  ->  row-major [nrows][ncols]
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nrows * ncols) return;
  int64_t r = i / ncols;
  int c = (int)(i % ncols);
  out[i] = in[(int64_t)c * nrows + r];
}
__global__ void k_csmat_convert(const double* __restrict__ in, double* __restrict__ out, int64_t n) {
  // each 3x3 column-major -> row-major
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * 9) return;
  int64_t m = i / 9;
  int k = (int)(i % 9), r = k / 3, cc = k % 3;
  out[i] = in[m * 9 + cc * 3 + r];
}
__global__ void k_i64_minus1_to_i32(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t n, int64_t lo,
                                    int64_t hi, int32_t* __restrict__ flag) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t v = in[i];
  if (v < lo || v > hi) {
    atomicExch(flag, 1);
    v = lo;
  }
  out[i] = (int32_t)(v - 1);
}
__global__ void k_i32_plus1_to_i64(const int32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (int64_t)in[i] + 1;
}

static int ensure_flag(fsgpu_ctx* c) {
  FS_TRY(c->flag.ensure(8));
  FS_CUDA(cudaMemsetAsync(c->flag.p, 0, 8 * sizeof(int32_t), c->stream));
  return FSGPU_OK;
}
static int read_flag(fsgpu_ctx* c, int idx, int32_t* v) {
  FS_CUDA(cudaMemcpyAsync(v, c->flag.p + idx, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return FSGPU_OK;
}

// stage a host array on the device scratch buffer `tmp`
static int stage(fsgpu_ctx* c, const void* host, size_t bytes, void** dev) {
  FS_TRY(c->tmp.ensure(bytes));
  FS_TRY(upload(c, c->tmp.p, host, bytes));
  *dev = c->tmp.p;
  return FSGPU_OK;
}

#define LAUNCH(ctx, kern, n, ...)                                              \
  do {                                                                         \
    if ((n) > 0) {                                                             \
      kern<<<fs::grid_for((n), 256), 256, 0, (ctx)->stream>>>(__VA_ARGS__);    \
      (ctx)->launches++;                                                       \
    }                                                                          \
  } while (0)

}  // namespace fs

using namespace fs;

// ------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------
extern "C" const char* fsgpu_last_error(void) { return fs::g_err.c_str(); }
extern "C" int fsgpu_version(void) { return 100; }

extern "C" int fsgpu_create(fsgpu_ctx** out, int device) {
  FS_REQUIRE(out != nullptr, FSGPU_ERR_ARG, "null output pointer");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); libfsgpu has no CPU fallback", cudaGetErrorString(e));
    return FSGPU_ERR_CUDA;
  }
  FS_REQUIRE(device >= 0 && device < n, FSGPU_ERR_ARG, "device %d out of range (have %d)", device, n);
  FS_CUDA(cudaSetDevice(device));
  fsgpu_ctx* c = new fsgpu_ctx();
  c->device = device;
  const char* tl = getenv("FSGPU_TILE");
  c->want_tile = tl && tl[0] && tl[0] != '0';
  *out = c;
  return FSGPU_OK;
}
extern "C" int fsgpu_destroy(fsgpu_ctx* c) {
  if (!c) return FSGPU_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  for (int k = 0; k < 4; ++k) {
    if (c->ring[k]) cudaFreeHost(c->ring[k]);
    if (c->ring_ev[k]) cudaEventDestroy(c->ring_ev[k]);
  }
  if (c->ev_x) cudaEventDestroy(c->ev_x);
  for (int k = 0; k < 3; ++k) {
    if (c->vring[k]) cudaFreeHost(c->vring[k]);
    if (c->vring_ev[k]) cudaEventDestroy(c->vring_ev[k]);
  }
  if (c->stream2) cudaStreamDestroy(c->stream2);
  delete c;
  return FSGPU_OK;
}
extern "C" int fsgpu_host_alloc(void** p, int64_t bytes) {
  FS_REQUIRE(p && bytes >= 0, FSGPU_ERR_ARG, "bad arguments");
  FS_CUDA(cudaMallocHost(p, (size_t)bytes));
  return FSGPU_OK;
}
extern "C" int fsgpu_host_free(void* p) {
  if (p) FS_CUDA(cudaFreeHost(p));
  return FSGPU_OK;
}
extern "C" int fsgpu_set_stream(fsgpu_ctx* c, void* s) {
  FS_TRY(check_ctx(c));
  c->stream = (cudaStream_t)s;
  return FSGPU_OK;
}
extern "C" int fsgpu_sync(fsgpu_ctx* c) {
  FS_TRY(check_ctx(c));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return FSGPU_OK;
}
extern "C" int64_t fsgpu_launch_count(fsgpu_ctx* c) { return c ? c->launches : 0; }
extern "C" int fsgpu_set_deterministic(fsgpu_ctx* c, int on) {
  FS_REQUIRE(c != nullptr, FSGPU_ERR_ARG, "null context");
  c->want_tile = on != 0;
  return FSGPU_OK;
}
extern "C" int fsgpu_scatter_path(fsgpu_ctx* c, int* path) {
  FS_REQUIRE(c != nullptr && path != nullptr, FSGPU_ERR_ARG, "null argument");
  *path = c->last_path;
  return FSGPU_OK;
}
extern "C" int fsgpu_last_kernel_ms(fsgpu_ctx* c, double* ms) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(ms && c->timed, FSGPU_ERR_STATE, "no timed operator has run");
  FS_CUDA(cudaEventSynchronize(c->ev1));
  float f = 0;
  FS_CUDA(cudaEventElapsedTime(&f, c->ev0, c->ev1));
  *ms = f;
  return FSGPU_OK;
}

// ------------------------------------------------------------------------------------
// data hand-over
// ------------------------------------------------------------------------------------
extern "C" int fsgpu_set_mesh(fsgpu_ctx* c, int32_t nnpe, int64_t nelem, const int64_t* conn, int64_t nnodes,
                              const double* xyz) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(nnpe >= 2 && nnpe <= 4, FSGPU_ERR_ARG, "nnpe must be 2, 3 or 4 (got %d)", nnpe);
  FS_REQUIRE(nelem >= 0 && nnodes >= 0 && nnodes < (int64_t)INT32_MAX / 8, FSGPU_ERR_ARG, "bad mesh sizes");
  FS_REQUIRE((conn || nelem == 0) && (xyz || nnodes == 0), FSGPU_ERR_ARG, "null mesh arrays");
  c->nnpe = nnpe;
  c->nelem = nelem;
  c->nnodes = nnodes;
  c->target = -1;
  c->have_matrix = c->have_vector = false;
  c->associated = false;
  c->t3_plan_ok = false;
  c->have_sections = c->have_state = false;
  c->ngroups = 0;
  c->nthick = c->nstab = 0;
  FS_TRY(ensure_flag(c));
  void* d;
  FS_TRY(c->conn.ensure((size_t)nelem * nnpe));
  FS_TRY(stage(c, conn, (size_t)nelem * nnpe * sizeof(int64_t), &d));
  LAUNCH(c, k_conn_convert, nelem * nnpe, (const int64_t*)d, c->conn.p, nelem * nnpe, nnodes, c->flag.p);
  FS_TRY(c->xyz.ensure((size_t)nnodes));
  FS_TRY(c->tmp2.ensure((size_t)nnodes * 3 * sizeof(double)));
  FS_TRY(upload(c, c->tmp2.p, xyz, (size_t)nnodes * 3 * sizeof(double)));
  LAUNCH(c, k_pack3, nnodes, (const double*)c->tmp2.p, c->xyz.p, nnodes);
  int32_t f;
  FS_TRY(read_flag(c, 0, &f));
  FS_REQUIRE(f == 0, FSGPU_ERR_ARG, "connectivity refers to a node outside 1..%lld", (long long)nnodes);
  return FSGPU_OK;
}

extern "C" int fsgpu_set_dofnums(fsgpu_ctx* c, const int64_t* dofnums, int64_t nfree, int64_t nall) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->nnpe > 0, FSGPU_ERR_STATE, "set the mesh first");
  FS_REQUIRE(dofnums != nullptr, FSGPU_ERR_ARG, "null dofnums");
  FS_REQUIRE(nfree >= 0 && nfree <= nall && nall < (int64_t)INT32_MAX, FSGPU_ERR_ARG, "bad nfree/nall");
  FS_TRY(ensure_flag(c));
  void* d;
  FS_TRY(c->dof.ensure((size_t)c->nnodes * 6));
  FS_TRY(stage(c, dofnums, (size_t)c->nnodes * 6 * sizeof(int64_t), &d));
  LAUNCH(c, k_dof_convert, c->nnodes * 6, (const int64_t*)d, c->dof.p, c->nnodes, nall, c->flag.p);
  int32_t f;
  FS_TRY(read_flag(c, 0, &f));
  // FinEtools assemble!: "Row degree of freedom < 1" / "> size"
  FS_REQUIRE(f == 0, FSGPU_ERR_DOF_RANGE, "degree of freedom < 1 or > nalldofs (%lld)", (long long)nall);
  c->nfree = nfree;
  c->nall = nall;
  c->have_dofs = true;
  c->target = -1;
  return FSGPU_OK;
}

extern "C" int fsgpu_set_normals(fsgpu_ctx* c, const double* normals, const uint8_t* valid) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->nnpe > 0, FSGPU_ERR_STATE, "set the mesh first");
  FS_REQUIRE(normals && valid, FSGPU_ERR_ARG, "null normals");
  const int64_t n = c->nnodes;
  FS_TRY(c->nrm.ensure((size_t)n));
  FS_TRY(c->tmp.ensure((size_t)n * 3 * sizeof(double)));
  FS_TRY(c->tmp2.ensure((size_t)n));
  FS_TRY(upload(c, c->tmp.p, normals, (size_t)n * 3 * sizeof(double)));
  FS_TRY(upload(c, c->tmp2.p, valid, (size_t)n));
  LAUNCH(c, k_pack_normals, n, (const double*)c->tmp.p, (const unsigned char*)c->tmp2.p, c->nrm.p, n);
  FS_CUDA(cudaStreamSynchronize(c->stream));
  c->associated = true;
  return FSGPU_OK;
}

extern "C" int fsgpu_get_normals(fsgpu_ctx* c, double* normals, uint8_t* valid) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->associated, FSGPU_ERR_STATE, "geometry not associated");
  const int64_t n = c->nnodes;
  FS_TRY(c->tmp.ensure((size_t)n * 3 * sizeof(double)));
  FS_TRY(c->tmp2.ensure((size_t)n));
  LAUNCH(c, k_unpack_normals, n, c->nrm.p, (double*)c->tmp.p, (unsigned char*)c->tmp2.p, n);
  FS_TRY(download(c, normals, c->tmp.p, (size_t)n * 3 * sizeof(double)));
  FS_TRY(download(c, valid, c->tmp2.p, (size_t)n));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return FSGPU_OK;
}

// ---- associategeometry! on the device ------------------------------------------------
// pass 1: accumulate (weighted) element normals at the nodes; pass 2: normalise;
// pass 3: mark nodes whose nodal normal deviates from an adjacent element normal.
__device__ inline fsm::V3 ldxyz(const double4* p, int i) {
  double4 v = p[i];
  return fsm::v3(v.x, v.y, v.z);
}
// Built-in coordinate-system kinds (include/fsgpu.h FSGPU_CSYS_*): the reference's CSys callbacks of its examples,
// evaluated on the device.  Columns e1, e2, e3 of the csys matrix at location X with surface normal nsurf.
__device__ inline void csys_eval(const CsysK& k, fsm::V3 X, fsm::V3 nsurf, fsm::V3& e1, fsm::V3& e2, fsm::V3& e3) {
  using namespace fsm;
  const V3 a = v3(k.a[0], k.a[1], k.a[2]);
  const V3 r = X - v3(k.o[0], k.o[1], k.o[2]);
  if (k.kind == FSGPU_CSYS_CYLINDRICAL) {
    // clamp_cyl_expl_examples.jl:62-68: r with its axial component removed, e2 = axis, e1 = e2 x e3
    const V3 q = r - dot(r, a) * a;
    e3 = (1.0 / sqrt(dot(q, q))) * q;
    e2 = a;
    e1 = cross(e2, e3);
  } else if (k.kind == FSGPU_CSYS_SPHERICAL) {
    // hemisphere_examples.jl:31-39: e3 radial, e1 = normalize(axis x e3), e2 = e3 x e1
    e3 = (1.0 / sqrt(dot(r, r))) * r;
    const V3 c = cross(a, e3);
    e1 = (1.0 / sqrt(dot(c, c))) * c;
    e2 = cross(e3, e1);
  } else {
    // FSGPU_CSYS_NORMAL_AXIS, pressurized_cylinder_free_examples.jl:16-23: e3 = surface normal, e2 = axis, e1 = e2 x e3
    e3 = nsurf;
    e2 = a;
    e1 = cross(e2, e3);
  }
}
__device__ inline fsm::V3 elem_normal_at(const double4* xyz, const int32_t* cn, int nnpe, int k, double& wgt) {
  using namespace fsm;
  if (nnpe == 3) {
    V3 a = ldxyz(xyz, cn[0]), b = ldxyz(xyz, cn[1]), cc = ldxyz(xyz, cn[2]);
    wgt = 1.0;  // T3: unweighted (src/FEMMShellT3FFModule.jl:583-586)
    return element_triad(b - a, cc - a).e3;
  }
  V3 X[4] = {ldxyz(xyz, cn[0]), ldxyz(xyz, cn[1]), ldxyz(xyz, cn[2]), ldxyz(xyz, cn[3])};
  const double px[4] = {-1, 1, 1, -1}, py[4] = {-1, -1, 1, 1};  // NodalTensorProductRule(2)
  Q4Geom g = q4_geometry(X, px[k], py[k]);
  wgt = g.Jac;  // Q4: Jacobian weighted (src/FEMMShellQ4RSModule.jl:489-494)
  return g.E.e3;
}
__global__ void k_normals_accumulate(const int32_t* __restrict__ conn, const double4* __restrict__ xyz, int nnpe,
                                     int64_t nelem, double* __restrict__ acc /*[nnodes][3]*/, int use_fixed, double fx,
                                     double fy, double fz, const double* __restrict__ dirs /*[nelem][nnpe][3] or null*/, CsysK ck) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nelem * nnpe) return;
  int64_t e = i / nnpe;
  int k = (int)(i % nnpe);
  const int32_t* cn = conn + e * nnpe;
  double w;
  fsm::V3 n = elem_normal_at(xyz, cn, nnpe, k, w);
  if (use_fixed) n = fsm::v3(fx, fy, fz);
  if (dirs) n = fsm::v3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]);
  if (ck.kind) {
    fsm::V3 e1, e2, e3;
    csys_eval(ck, ldxyz(xyz, cn[k]), n, e1, e2, e3);  // the csys evaluated AT THE NODE (`_compute_nodal_normal!`)
    n = e3;
  }
  double* a = acc + (int64_t)cn[k] * 3;
  atomicAdd(a + 0, w * n.x);
  atomicAdd(a + 1, w * n.y);
  atomicAdd(a + 2, w * n.z);
}
__global__ void k_normals_normalize(const double* __restrict__ acc, double4* __restrict__ nrm, int64_t n, int keep_valid) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = acc[3 * i], y = acc[3 * i + 1], z = acc[3 * i + 2];
  double nn = sqrt(x * x + y * y + z * z);
  if (nn > 0.0) {
    x /= nn;
    y /= nn;
    z /= nn;
  }
  double v = keep_valid ? nrm[i].w : 1.0;
  nrm[i] = make_double4(x, y, z, v);
}
__global__ void k_normals_validate(const int32_t* __restrict__ conn, const double4* __restrict__ xyz, int nnpe,
                                   int64_t nelem, double4* __restrict__ nrm, double limit, int use_fixed, double fx,
                                   double fy, double fz, int fixed_in_check, const double* __restrict__ dirs, CsysK ck) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nelem * nnpe) return;
  int64_t e = i / nnpe;
  int k = (int)(i % nnpe);
  const int32_t* cn = conn + e * nnpe;
  double w;
  fsm::V3 n = elem_normal_at(xyz, cn, nnpe, k, w);
  if (use_fixed && fixed_in_check) n = fsm::v3(fx, fy, fz);
  if (dirs && fixed_in_check) n = fsm::v3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]);
  if (ck.kind && fixed_in_check) {
    fsm::V3 e1, e2, e3;
    csys_eval(ck, ldxyz(xyz, cn[k]), n, e1, e2, e3);
    n = e3;
  }
  double4 nn = nrm[cn[k]];
  double nd = nn.x * n.x + nn.y * n.y + nn.z * n.z;
  if (nd < limit) nrm[cn[k]].w = 0.0;  // benign race: every writer stores 0
}
__global__ void k_unpack_acc(const double4* __restrict__ nrm, double* __restrict__ acc, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 v = nrm[i];
  acc[3 * i] = v.x;
  acc[3 * i + 1] = v.y;
  acc[3 * i + 2] = v.z;
}

// Split form for element-partitioned runs: (1) accumulate the (weighted) element normals at the
// nodes of THIS partition, (2) the host sums the interface-node entries across ranks (device
// pointer handed out), (3) normalise + validity pass; the host then min-combines the validity
// flags of interface nodes (the 4th component of the packed normals).
extern "C" int fsgpu_normals_accumulate(fsgpu_ctx* c, const double* fixed_dir, int32_t accumulate, double** dev_sums) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->nnpe == 3 || c->nnpe == 4, FSGPU_ERR_STATE, "associategeometry needs a T3 or Q4 mesh");
  const int64_t n = c->nnodes;
  const bool had = c->associated && c->nrm.p != nullptr;
  FS_TRY(c->nrm.ensure((size_t)n));
  FS_TRY(c->nacc.ensure((size_t)n * 3 + 1));
  double* acc = c->nacc.p;
  c->nacc_keep = accumulate && had;
  if (c->nacc_keep) {
    LAUNCH(c, k_unpack_acc, n, c->nrm.p, acc, n);
  } else {
    FS_CUDA(cudaMemsetAsync(acc, 0, (size_t)n * 3 * sizeof(double), c->stream));
  }
  const int uf = fixed_dir ? 1 : 0;
  const double fx = uf ? fixed_dir[0] : 0, fy = uf ? fixed_dir[1] : 0, fz = uf ? fixed_dir[2] : 0;
  LAUNCH(c, k_normals_accumulate, c->nelem * c->nnpe, c->conn.p, c->xyz.p, c->nnpe, c->nelem, acc, uf, fx, fy, fz, c->ndirs, c->ncsys);
  FS_CUDA(cudaStreamSynchronize(c->stream));
  if (dev_sums) *dev_sums = acc;
  return FSGPU_OK;
}
extern "C" int fsgpu_normals_finish(fsgpu_ctx* c, double threshold_angle_deg, const double* fixed_dir, double** dev_normals4) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->nacc.p != nullptr, FSGPU_ERR_STATE, "call fsgpu_normals_accumulate first");
  const int64_t n = c->nnodes;
  const int uf = fixed_dir ? 1 : 0;
  const double fx = uf ? fixed_dir[0] : 0, fy = uf ? fixed_dir[1] : 0, fz = uf ? fixed_dir[2] : 0;
  LAUNCH(c, k_normals_normalize, n, c->nacc.p, c->nrm.p, n, c->nacc_keep ? 1 : 0);
  const double s = sin(threshold_angle_deg / 180 * M_PI);
  const double ntol = 1 - sqrt(1 - s * s);
  // T3 (homogeneous and composite) checks against the ELEMENT normal (src/FEMMShellT3FFModule.jl:601-611,
  // ...CompModule.jl:528-538); Q4 checks against the csys normal (src/FEMMShellQ4RSModule.jl:508-519).
  const int fixed_in_check = (c->nnpe == 4) ? 1 : 0;
  LAUNCH(c, k_normals_validate, c->nelem * c->nnpe, c->conn.p, c->xyz.p, c->nnpe, c->nelem, c->nrm.p, 1 - ntol, uf, fx,
         fy, fz, fixed_in_check, c->ndirs, c->ncsys);
  FS_CUDA(cudaStreamSynchronize(c->stream));
  c->associated = true;
  if (dev_normals4) *dev_normals4 = reinterpret_cast<double*>(c->nrm.p);
  return FSGPU_OK;
}
extern "C" int fsgpu_associategeometry(fsgpu_ctx* c, double threshold_angle_deg, const double* fixed_dir,
                                       int32_t accumulate) {
  FS_TRY(fsgpu_normals_accumulate(c, fixed_dir, accumulate, nullptr));
  return fsgpu_normals_finish(c, threshold_angle_deg, fixed_dir, nullptr);
}

// General csys: the direction csmat(csys)[:, 3] evaluated by the host glue per element and node
// (`_compute_nodal_normal!`, src/FEMMShellT3FFCompModule.jl:203-207,509; src/FEMMShellQ4RSModule.jl:489-494).
extern "C" int fsgpu_associategeometry_dirs(fsgpu_ctx* c, double threshold_angle_deg, const double* dirs, int32_t accumulate) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(dirs != nullptr, FSGPU_ERR_ARG, "null direction array");
  FS_REQUIRE(c->nnpe == 3 || c->nnpe == 4, FSGPU_ERR_STATE, "associategeometry needs a T3 or Q4 mesh");
  const size_t bytes = (size_t)c->nelem * c->nnpe * 3 * sizeof(double);
  DBuf<double> d;
  FS_TRY(d.ensure((size_t)c->nelem * c->nnpe * 3 + 1));
  FS_TRY(upload(c, d.p, dirs, bytes));
  c->ndirs = d.p;
  int rc = fsgpu_normals_accumulate(c, nullptr, accumulate, nullptr);
  if (rc == FSGPU_OK) rc = fsgpu_normals_finish(c, threshold_angle_deg, nullptr, nullptr);
  c->ndirs = nullptr;
  return rc;
}

// Built-in csys kinds: nothing but the kind and two vectors crosses the boundary (C3: the host evaluation of the
// callback at 6 M element nodes took 1.2 s against 2 ms for the stiffness itself).
static int make_csys(int32_t kind, const double* origin, const double* axis, CsysK& k) {
  FS_REQUIRE(kind == FSGPU_CSYS_CYLINDRICAL || kind == FSGPU_CSYS_SPHERICAL || kind == FSGPU_CSYS_NORMAL_AXIS, FSGPU_ERR_ARG,
             "unknown csys kind %d", (int)kind);
  FS_REQUIRE(axis != nullptr, FSGPU_ERR_ARG, "the csys kinds need an axis");
  const double al = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
  FS_REQUIRE(al > 0.0, FSGPU_ERR_ARG, "zero csys axis");
  k.kind = kind;
  for (int i = 0; i < 3; ++i) {
    k.o[i] = origin ? origin[i] : 0.0;
    k.a[i] = axis[i] / al;
  }
  return FSGPU_OK;
}
extern "C" int fsgpu_associategeometry_csys(fsgpu_ctx* c, double threshold_angle_deg, int32_t kind, const double* origin,
                                            const double* axis, int32_t accumulate) {
  FS_TRY(check_ctx(c));
  CsysK k;
  FS_TRY(make_csys(kind, origin, axis, k));
  c->ncsys = k;
  int rc = fsgpu_normals_accumulate(c, nullptr, accumulate, nullptr);
  if (rc == FSGPU_OK) rc = fsgpu_normals_finish(c, threshold_angle_deg, nullptr, nullptr);
  c->ncsys.kind = 0;
  return rc;
}
// layup csys matrices (row-major 3x3) on the device: T3FFComp one per element at the centroid with J0
// (src/FEMMShellT3FFCompModule.jl:617); Q4RSComp one per element and integration point with the SHAPE-FUNCTION VALUES
// as the location (src/FEMMShellQ4RSCompModule.jl:929, SURVEY App. B.9 -- reproduced) and the point's Jacobian
__global__ void k_layup_csys(const int32_t* __restrict__ conn, const double4* __restrict__ xyz, int nnpe, int64_t nelem, fs::Rule rule,
                             CsysK ck, double* __restrict__ cs) {
  using namespace fsm;
  const int npts = nnpe == 3 ? 1 : rule.npts;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nelem * npts) return;
  const int64_t e = i / npts;
  const int j = (int)(i % npts);
  const int32_t* cn = conn + e * nnpe;
  V3 X, n;
  if (nnpe == 3) {
    const V3 a = ldxyz(xyz, cn[0]), b = ldxyz(xyz, cn[1]), cc = ldxyz(xyz, cn[2]);
    X = (1.0 / 3) * (a + b + cc);
    n = element_triad(b - a, cc - a).e3;
  } else {
    const V3 Xe[4] = {ldxyz(xyz, cn[0]), ldxyz(xyz, cn[1]), ldxyz(xyz, cn[2]), ldxyz(xyz, cn[3])};
    const double xi = rule.xi[j], eta = rule.eta[j];
    n = q4_geometry(Xe, xi, eta).E.e3;
    X = v3(0.25 * (1 - xi) * (1 - eta), 0.25 * (1 + xi) * (1 - eta), 0.25 * (1 + xi) * (1 + eta));  // Ns[j][1:3]
  }
  V3 e1, e2, e3;
  csys_eval(ck, X, n, e1, e2, e3);
  double* o = cs + i * 9;
  o[0] = e1.x; o[1] = e2.x; o[2] = e3.x;
  o[3] = e1.y; o[4] = e2.y; o[5] = e3.y;
  o[6] = e1.z; o[7] = e2.z; o[8] = e3.z;
}
extern "C" int fsgpu_set_layup_csys(fsgpu_ctx* c, int32_t kind, const double* origin, const double* axis) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->ngroups >= 1, FSGPU_ERR_STATE, "call fsgpu_set_layup first (group data; its csys argument is replaced)");
  FS_REQUIRE(c->nnpe == 3 || (c->nnpe == 4 && c->rule.npts >= 1), FSGPU_ERR_STATE, "needs a T3 mesh, or a Q4 mesh with its integration rule");
  CsysK k;
  FS_TRY(make_csys(kind, origin, axis, k));
  const int64_t ncs = c->nelem * (c->nnpe == 3 ? 1 : c->rule.npts);
  FS_TRY(c->csmat.ensure((size_t)ncs * 9 + 1));
  LAUNCH(c, k_layup_csys, ncs, c->conn.p, c->xyz.p, c->nnpe, c->nelem, c->rule, k, c->csmat.p);
  FS_CUDA(cudaStreamSynchronize(c->stream));
  c->ncs = ncs;
  return FSGPU_OK;
}

extern "C" int fsgpu_set_thickness(fsgpu_ctx* c, const double* t, int64_t n) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(t && n >= 1, FSGPU_ERR_ARG, "bad thickness array");
  FS_TRY(c->thick.ensure((size_t)n));
  FS_TRY(upload(c, c->thick.p, t, (size_t)n * sizeof(double)));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  c->nthick = n;
  return FSGPU_OK;
}
extern "C" int fsgpu_set_stab_factor(fsgpu_ctx* c, const double* f, int64_t n) {
  FS_TRY(check_ctx(c));
  if (!f || n == 0) {
    c->nstab = 0;
    return FSGPU_OK;
  }
  FS_REQUIRE(n == c->nelem, FSGPU_ERR_ARG, "stab factor array must have nelem entries");
  FS_TRY(c->stabf.ensure((size_t)n));
  FS_TRY(upload(c, c->stabf.p, f, (size_t)n * sizeof(double)));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  c->nstab = n;
  return FSGPU_OK;
}
extern "C" int fsgpu_set_rule(fsgpu_ctx* c, int32_t npts, const double* xi, const double* eta, const double* w) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(npts >= 1 && npts <= kMaxGP && xi && eta && w, FSGPU_ERR_ARG, "rule must have 1..%d points", kMaxGP);
  c->rule.npts = npts;
  for (int i = 0; i < npts; ++i) {
    c->rule.xi[i] = xi[i];
    c->rule.eta[i] = eta[i];
    c->rule.w[i] = w[i];
  }
  return FSGPU_OK;
}
extern "C" int fsgpu_set_layup(fsgpu_ctx* c, int32_t ngroups, const double* group_data, const int64_t* group_of_elem,
                               const double* csmat, int64_t ncs) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->nnpe > 0, FSGPU_ERR_STATE, "set the mesh first");
  FS_REQUIRE(ngroups >= 1 && group_data && csmat && ncs >= 1, FSGPU_ERR_ARG, "bad layup arguments");
  FS_REQUIRE(ngroups == 1 || group_of_elem, FSGPU_ERR_ARG, "group_of_elem required for more than one group");
  FS_TRY(ensure_flag(c));
  FS_TRY(c->group_data.ensure((size_t)ngroups * 34));
  FS_TRY(upload(c, c->group_data.p, group_data, (size_t)ngroups * 34 * sizeof(double)));
  FS_TRY(c->group_of.ensure((size_t)c->nelem));
  if (group_of_elem) {
    void* d;
    FS_TRY(stage(c, group_of_elem, (size_t)c->nelem * sizeof(int64_t), &d));
    LAUNCH(c, k_i64_minus1_to_i32, c->nelem, (const int64_t*)d, c->group_of.p, c->nelem, (int64_t)1, (int64_t)ngroups,
           c->flag.p);
  } else {
    FS_CUDA(cudaMemsetAsync(c->group_of.p, 0, (size_t)c->nelem * sizeof(int32_t), c->stream));
  }
  FS_TRY(c->csmat.ensure((size_t)ncs * 9));
  FS_TRY(c->tmp2.ensure((size_t)ncs * 9 * sizeof(double)));
  FS_TRY(upload(c, c->tmp2.p, csmat, (size_t)ncs * 9 * sizeof(double)));
  LAUNCH(c, k_csmat_convert, ncs * 9, (const double*)c->tmp2.p, c->csmat.p, ncs);
  int32_t f;
  FS_TRY(read_flag(c, 0, &f));
  FS_REQUIRE(f == 0, FSGPU_ERR_ARG, "layup group index outside 1..%d", ngroups);
  c->ngroups = ngroups;
  c->ncs = ncs;
  return FSGPU_OK;
}
__global__ void k_pack_sections(const double* A, const double* I1, const double* I2, const double* I3, const double* J,
                                const double* A2s, const double* A3s, const double* x1x2, double* out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* o = out + i * 10;
  o[0] = A[i];
  o[1] = I1[i];
  o[2] = I2[i];
  o[3] = I3[i];
  o[4] = J[i];
  o[5] = A2s[i];
  o[6] = A3s[i];
  o[7] = x1x2[3 * i];
  o[8] = x1x2[3 * i + 1];
  o[9] = x1x2[3 * i + 2];
}
extern "C" int fsgpu_set_beam_sections(fsgpu_ctx* c, const double* A, const double* I1, const double* I2, const double* I3,
                                       const double* J, const double* A2s, const double* A3s, const double* x1x2) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->nnpe == 2, FSGPU_ERR_STATE, "beam sections need an L2 mesh");
  FS_REQUIRE(A && I1 && I2 && I3 && J && A2s && A3s && x1x2, FSGPU_ERR_ARG, "null section array");
  const int64_t n = c->nelem;
  FS_TRY(c->sec.ensure((size_t)n * 10));
  FS_TRY(c->tmp.ensure((size_t)n * 10 * sizeof(double)));
  double* t = (double*)c->tmp.p;
  const double* src[7] = {A, I1, I2, I3, J, A2s, A3s};
  for (int k = 0; k < 7; ++k) FS_TRY(upload(c, t + k * n, src[k], (size_t)n * sizeof(double)));
  FS_TRY(upload(c, t + 7 * n, x1x2, (size_t)n * 3 * sizeof(double)));
  LAUNCH(c, k_pack_sections, n, t, t + n, t + 2 * n, t + 3 * n, t + 4 * n, t + 5 * n, t + 6 * n, t + 7 * n, c->sec.p, n);
  FS_CUDA(cudaStreamSynchronize(c->stream));
  c->have_sections = true;
  return FSGPU_OK;
}
extern "C" int fsgpu_set_state(fsgpu_ctx* c, const double* u1, const double* Rfield1) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->nnpe > 0, FSGPU_ERR_STATE, "set the mesh first");
  FS_REQUIRE(u1 && Rfield1, FSGPU_ERR_ARG, "null state arrays");
  const int64_t n = c->nnodes;
  FS_TRY(c->u1.ensure((size_t)n));
  FS_TRY(c->R1.ensure((size_t)n * 9));
  FS_TRY(c->tmp.ensure((size_t)n * 9 * sizeof(double)));
  FS_TRY(upload(c, c->tmp.p, u1, (size_t)n * 3 * sizeof(double)));
  LAUNCH(c, k_pack3, n, (const double*)c->tmp.p, c->u1.p, n);
  FS_CUDA(cudaStreamSynchronize(c->stream));
  FS_TRY(upload(c, c->tmp.p, Rfield1, (size_t)n * 9 * sizeof(double)));
  LAUNCH(c, k_transpose_rows, n * 9, (const double*)c->tmp.p, c->R1.p, n, 9);
  FS_CUDA(cudaStreamSynchronize(c->stream));
  c->have_state = true;
  return FSGPU_OK;
}

// ------------------------------------------------------------------------------------
// symbolic phase
// ------------------------------------------------------------------------------------
namespace fs {

struct TargetInfo {
  bool diag_only;
  int64_t nr, nc;  // matrix size; rows/cols with dof index >= nr/nc are dropped
};
static TargetInfo target_info(const fsgpu_ctx* c, int target) {
  switch (target) {
    case FSGPU_FFBLOCK:
      return {false, c->nfree, c->nfree};
    case FSGPU_FFBLOCK_DIAG:
      return {true, c->nfree, c->nfree};
    case FSGPU_SPARSE_DIAG:
      return {true, c->nall, c->nall};
    default:
      return {false, c->nall, c->nall};
  }
}

__global__ void k_pair_keys(const int32_t* __restrict__ conn, int nnpe, int64_t nelem, uint64_t* __restrict__ keys) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t n = nelem * nnpe * nnpe;
  if (i >= n) return;
  int64_t e = i / (nnpe * nnpe);
  int r = (int)(i % (nnpe * nnpe));
  uint64_t a = (uint64_t)conn[e * nnpe + r / nnpe], b = (uint64_t)conn[e * nnpe + r % nnpe];
  keys[i] = (a << 32) | b;
}
__global__ void k_adj_count(const uint64_t* __restrict__ ukeys, int64_t n, int32_t* __restrict__ deg) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  atomicAdd(deg + (int32_t)(ukeys[i] >> 32), 1);
}
__global__ void k_adj_fill(const uint64_t* __restrict__ ukeys, int64_t n, int32_t* __restrict__ adj) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  adj[i] = (int32_t)(ukeys[i] & 0xffffffffu);
}
// number of row-included dofs per node
__global__ void k_node_rowcount(const int32_t* __restrict__ dof, int64_t nnodes, int64_t nr, int32_t* __restrict__ cnt) {
  int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a >= nnodes) return;
  int n = 0;
  for (int d = 0; d < 6; ++d) n += (dof[a * 6 + d] < nr) ? 1 : 0;
  cnt[a] = n;
}
__global__ void k_col_count(const int32_t* __restrict__ dof, const int32_t* __restrict__ adjptr,
                            const int32_t* __restrict__ adj, const int32_t* __restrict__ nodecnt, int64_t nnodes,
                            int64_t nc, int diag_only, int64_t* __restrict__ colcnt) {
  int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a >= nnodes) return;
  const int b0 = adjptr[a], b1 = adjptr[a + 1];
  int64_t total = 0;
  if (diag_only) {
    total = (b1 > b0) ? 1 : 0;
  } else {
    for (int p = b0; p < b1; ++p) total += nodecnt[adj[p]];
  }
  for (int d = 0; d < 6; ++d) {
    int32_t cdof = dof[a * 6 + d];
    if (cdof < nc) colcnt[cdof] = total;
  }
}
__global__ void k_colptr_narrow(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (int32_t)in[i];
}
__global__ void k_fill_rowval(const int32_t* __restrict__ dof, const int32_t* __restrict__ adjptr,
                              const int32_t* __restrict__ adj, const int32_t* __restrict__ colptr, int64_t nnodes,
                              int64_t nr, int64_t nc, int diag_only, int32_t* __restrict__ rowval) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // node*6 + d
  if (i >= nnodes * 6) return;
  const int64_t a = i / 6;
  const int32_t cdof = dof[i];
  if (cdof >= nc) return;
  const int b0 = adjptr[a], b1 = adjptr[a + 1];
  if (b1 == b0) return;
  int p0 = colptr[cdof];
  if (diag_only) {
    rowval[p0] = cdof;
    return;
  }
  int p = p0;
  for (int q = b0; q < b1; ++q) {
    const int32_t* db = dof + (int64_t)adj[q] * 6;
    for (int d = 0; d < 6; ++d) {
      int32_t r = db[d];
      if (r < nr) {
        // insertion into the sorted prefix [p0, p)
        int k = p;
        while (k > p0 && rowval[k - 1] > r) {
          rowval[k] = rowval[k - 1];
          --k;
        }
        rowval[k] = r;
        ++p;
      }
    }
  }
}
__device__ inline int find_row(const int32_t* __restrict__ rowval, int lo, int hi, int32_t r) {
  // first position in [lo,hi) with rowval >= r; caller guarantees presence
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (rowval[mid] < r)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}
// slot[k][i][e][j], k = c*6 + r : nzval index of entry (row dof r of node i, col dof c of node j)
__global__ void k_slot_map(const int32_t* __restrict__ conn, const int32_t* __restrict__ dof,
                           const int32_t* __restrict__ colptr, const int32_t* __restrict__ rowval, int nnpe,
                           int64_t nelem, int64_t nr, int64_t nc, int diag_only, int32_t* __restrict__ slot) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // e*nnpe + j
  if (t >= nelem * nnpe) return;
  const int64_t e = t / nnpe;
  const int j = (int)(t % nnpe);
  const int32_t* cj = dof + (int64_t)conn[e * nnpe + j] * 6;
  const int64_t plane = nelem * nnpe;  // elements of one (k,i) plane
  for (int i = 0; i < nnpe; ++i) {
    const int32_t* ri = dof + (int64_t)conn[e * nnpe + i] * 6;
    for (int cc = 0; cc < 6; ++cc) {
      const int32_t cd = cj[cc];
      const bool cin = cd < nc;
      int lo = 0, hi = 0;
      if (cin) {
        lo = colptr[cd];
        hi = colptr[cd + 1];
      }
      for (int rr = 0; rr < 6; ++rr) {
        const int32_t rd = ri[rr];
        int s = -1;
        if (cin && rd < nr) {
          if (diag_only) {
            if (rd == cd) s = lo;
          } else {
            s = find_row(rowval, lo, hi, rd);
          }
        }
        slot[((int64_t)((cc * 6 + rr) * nnpe + i)) * plane + t] = s;
      }
    }
  }
}
// per node: masks of the included dofs in the free run (A) and the prescribed run (B);
// flag[1] is raised when a run is not consecutive-ascending in the local dof order
__global__ void k_node_info(const int32_t* __restrict__ dof, int64_t nnodes, int64_t nr, int64_t nfree,
                            int32_t* __restrict__ info, int32_t* __restrict__ flag) {
  int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a >= nnodes) return;
  int mA = 0, mB = 0, lastA = -2, lastB = -2, bad = 0;
  for (int d = 0; d < 6; ++d) {
    const int32_t v = dof[a * 6 + d];
    if (v >= nr) continue;
    if (v < nfree) {
      if (mA && v != lastA + 1) bad = 1;
      lastA = v;
      mA |= 1 << d;
    } else {
      if (mB && v != lastB + 1) bad = 1;
      lastB = v;
      mB |= 1 << d;
    }
  }
  info[a] = mA | (mB << 8);
  if (bad) atomicExch(flag + 1, 1);
}
// nodecol[a*8 + k], k < 6: colptr of the node's k-th dof (-1: not a column); [6] = nodeinfo; [7] = 0
__global__ void k_node_cols(const int32_t* __restrict__ dof, const int32_t* __restrict__ info,
                            const int32_t* __restrict__ colptr, int64_t nnodes, int64_t nc, int32_t* __restrict__ nodecol) {
  int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a >= nnodes) return;
  for (int k = 0; k < 6; ++k) {
    const int32_t d = dof[a * 6 + k];
    nodecol[a * 8 + k] = d < nc ? colptr[d] : -1;
  }
  nodecol[a * 8 + 6] = info[a];
  nodecol[a * 8 + 7] = 0;
}
// pairoff[((i*2 + which) * nelem + e) * nnpe + j]: position, relative to the column start, of the
// first run-A / run-B row of node i inside any included column of node j (-1: no such entries)
__global__ void k_pair_offsets(const int32_t* __restrict__ conn, const int32_t* __restrict__ dof,
                               const int32_t* __restrict__ info, const int32_t* __restrict__ colptr,
                               const int32_t* __restrict__ rowval, int nnpe, int64_t nelem, int64_t nc,
                               int32_t* __restrict__ pairoff) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // e*nnpe + j
  if (t >= nelem * nnpe) return;
  const int64_t e = t / nnpe;
  const int j = (int)(t % nnpe);
  const int32_t* cj = dof + (int64_t)conn[e * nnpe + j] * 6;
  int col = -1;
  for (int cc = 0; cc < 6; ++cc)
    if (cj[cc] < nc) {
      col = cj[cc];
      break;
    }
  for (int i = 0; i < nnpe; ++i) {
    const int ni = conn[e * nnpe + i];
    const int inf = info[ni];
    const int mA = inf & 63, mB = (inf >> 8) & 63;
    int oA = -1, oB = -1;
    if (col >= 0) {
      const int lo = colptr[col], hi = colptr[col + 1];
      if (mA) oA = find_row(rowval, lo, hi, dof[(int64_t)ni * 6 + (__ffs(mA) - 1)]) - lo;
      if (mB) oB = find_row(rowval, lo, hi, dof[(int64_t)ni * 6 + (__ffs(mB) - 1)]) - lo;
    }
    pairoff[((int64_t)(i * 2 + 0) * nelem + e) * nnpe + j] = oA;
    pairoff[((int64_t)(i * 2 + 1) * nelem + e) * nnpe + j] = oB;
  }
}
__global__ void k_diag_slot(const int32_t* __restrict__ colptr, const int32_t* __restrict__ rowval, int64_t nall,
                            int64_t nc, int32_t* __restrict__ diagslot) {
  int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (d >= nall) return;
  int s = -1;
  if (d < nc) {
    int lo = colptr[d], hi = colptr[d + 1];
    int p = find_row(rowval, lo, hi, (int32_t)d);
    if (p < hi && rowval[p] == (int32_t)d) s = p;
  }
  diagslot[d] = s;
}

}  // namespace fs

extern "C" int fsgpu_symbolic(fsgpu_ctx* c, int32_t target, int64_t* nrows, int64_t* ncols, int64_t* nnz) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->nnpe > 0 && c->have_dofs, FSGPU_ERR_STATE, "set mesh and dofnums before the symbolic phase");
  FS_REQUIRE(target >= 0 && target <= 5, FSGPU_ERR_ARG, "unknown assembler target %d", target);
  const TargetInfo ti = target_info(c, target);
  const int nnpe = c->nnpe;
  const int64_t ne = c->nelem, nn = c->nnodes;
  cudaStream_t st = c->stream;

  // (1) node adjacency from unique (a,b) pairs
  const int64_t npairs = ne * nnpe * nnpe;
  DBuf<uint64_t>&keys = c->scr_keys, &keys2 = c->scr_keys2;
  DBuf<int32_t>&deg = c->scr_deg, &nodecnt = c->scr_nodecnt;
  DBuf<int32_t>& adjptr = c->adjptr;
  DBuf<int32_t>& adj = c->adj;
  c->tile_ok = false;
  c->det_ready = false;
  DBuf<int64_t>&colcnt = c->scr_colcnt, &nsel = c->scr_nsel;
  FS_TRY(keys.ensure((size_t)npairs + 1));
  FS_TRY(keys2.ensure((size_t)npairs + 1));
  FS_TRY(nsel.ensure(1));
  LAUNCH(c, k_pair_keys, npairs, c->conn.p, nnpe, ne, keys.p);
  int nbits = 1;
  while (((int64_t)1 << nbits) < nn + 1) ++nbits;
  size_t tb = 0, tb2 = 0;
  if (npairs > 0) {
    FS_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, keys.p, keys2.p, npairs, 0, 32 + nbits, st));
    FS_CUDA(cub::DeviceSelect::Unique(nullptr, tb2, keys2.p, keys.p, nsel.p, npairs, st));
    FS_TRY(c->tmp.ensure(tb > tb2 ? tb : tb2));
    FS_CUDA(cub::DeviceRadixSort::SortKeys(c->tmp.p, tb, keys.p, keys2.p, npairs, 0, 32 + nbits, st));
    FS_CUDA(cub::DeviceSelect::Unique(c->tmp.p, tb2, keys2.p, keys.p, nsel.p, npairs, st));
    c->launches += 8;
  }
  int64_t nuniq = 0;
  if (npairs > 0) {
    FS_CUDA(cudaMemcpyAsync(&nuniq, nsel.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaStreamSynchronize(st));
  }
  FS_TRY(deg.ensure((size_t)nn + 1));
  FS_TRY(adjptr.ensure((size_t)nn + 1));
  FS_TRY(adj.ensure((size_t)nuniq + 1));
  FS_CUDA(cudaMemsetAsync(deg.p, 0, ((size_t)nn + 1) * sizeof(int32_t), st));
  LAUNCH(c, k_adj_count, nuniq, keys.p, nuniq, deg.p);
  FS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, deg.p, adjptr.p, nn + 1, st));
  FS_TRY(c->tmp.ensure(tb));
  FS_CUDA(cub::DeviceScan::ExclusiveSum(c->tmp.p, tb, deg.p, adjptr.p, nn + 1, st));
  c->launches += 2;
  LAUNCH(c, k_adj_fill, nuniq, keys.p, nuniq, adj.p);

  // (2) column counts -> colptr
  FS_TRY(nodecnt.ensure((size_t)nn + 1));
  LAUNCH(c, k_node_rowcount, nn, c->dof.p, nn, ti.nr, nodecnt.p);
  FS_TRY(colcnt.ensure((size_t)ti.nc + 1));
  FS_CUDA(cudaMemsetAsync(colcnt.p, 0, ((size_t)ti.nc + 1) * sizeof(int64_t), st));
  LAUNCH(c, k_col_count, nn, c->dof.p, adjptr.p, adj.p, nodecnt.p, nn, ti.nc, ti.diag_only ? 1 : 0, colcnt.p);
  DBuf<int64_t>& colptr64 = c->scr_colptr64;
  FS_TRY(colptr64.ensure((size_t)ti.nc + 1));
  FS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, colcnt.p, colptr64.p, ti.nc + 1, st));
  FS_TRY(c->tmp.ensure(tb));
  FS_CUDA(cub::DeviceScan::ExclusiveSum(c->tmp.p, tb, colcnt.p, colptr64.p, ti.nc + 1, st));
  c->launches += 2;
  int64_t total = 0;
  FS_CUDA(cudaMemcpyAsync(&total, colptr64.p + ti.nc, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  FS_REQUIRE(total < (int64_t)INT32_MAX, FSGPU_ERR_ARG, "pattern has %lld entries; the int32 slot map limit is 2^31-1",
             (long long)total);
  FS_TRY(c->colptr.ensure((size_t)ti.nc + 1));
  LAUNCH(c, k_colptr_narrow, ti.nc + 1, colptr64.p, c->colptr.p, ti.nc + 1);

  // (3) rowval, sorted ascending within each column
  FS_TRY(c->rowval.ensure((size_t)total + 1));
  LAUNCH(c, k_fill_rowval, nn * 6, c->dof.p, adjptr.p, adj.p, c->colptr.p, nn, ti.nr, ti.nc, ti.diag_only ? 1 : 0,
         c->rowval.p);

  // (4) addressing: run-structured fast path when every node's dofs form consecutive runs
  //     (FinEtools numbering), generic per-entry slot map otherwise / for diagonal targets
  FS_TRY(ensure_flag(c));
  FS_TRY(c->nodeinfo.ensure((size_t)nn + 1));
  LAUNCH(c, k_node_info, nn, c->dof.p, nn, ti.nr, c->nfree, c->nodeinfo.p, c->flag.p);
  int32_t notruns = 0;
  FS_TRY(read_flag(c, 1, &notruns));
  c->fast = !ti.diag_only && !notruns && getenv("FSGPU_FORCE_GENERIC") == nullptr;
  if (c->fast) {
    FS_TRY(c->pairoff.ensure((size_t)2 * nnpe * nnpe * ne + 1));
    LAUNCH(c, k_pair_offsets, ne * nnpe, c->conn.p, c->dof.p, c->nodeinfo.p, c->colptr.p, c->rowval.p, nnpe, ne, ti.nc,
           c->pairoff.p);
    FS_TRY(c->nodecol.ensure((size_t)8 * nn + 8));
    LAUNCH(c, k_node_cols, nn, c->dof.p, c->nodeinfo.p, c->colptr.p, nn, ti.nc, c->nodecol.p);
    c->slot.release();
    if (nnpe == 3) FS_TRY(fsk::t3_build_plan(c));
  } else {
    c->pairoff.release();
    FS_TRY(c->slot.ensure((size_t)36 * nnpe * nnpe * ne + 1));
    LAUNCH(c, k_slot_map, ne * nnpe, c->conn.p, c->dof.p, c->colptr.p, c->rowval.p, nnpe, ne, ti.nr, ti.nc,
           ti.diag_only ? 1 : 0, c->slot.p);
  }
  FS_TRY(c->diagslot.ensure((size_t)c->nall + 1));
  LAUNCH(c, k_diag_slot, c->nall, c->colptr.p, c->rowval.p, c->nall, ti.nc, c->diagslot.p);
  FS_TRY(c->nzval.ensure((size_t)total + 1));
  FS_CUDA(cudaStreamSynchronize(st));

  c->target = target;
  c->rle_for = nullptr;  // the cached run-length form belongs to the previous pattern
  c->prows = ti.nr;
  c->pcols = ti.nc;
  c->pnnz = total;
  // T3 meshes: data of the atomics-free owner-computes tile kernel (fsgpu_tile.cu)
  FS_TRY(fsk::tile_symbolic(c));
  c->have_matrix = false;
  c->compacted = false;
  if (nrows) *nrows = ti.nr;
  if (ncols) *ncols = ti.nc;
  if (nnz) *nnz = total;
  return FSGPU_OK;
}

// ------------------------------------------------------------------------------------
// result finalisation and hand-back
// ------------------------------------------------------------------------------------
namespace fs {

// SPARSE_SYMM: make the matrix exactly symmetric (the reference's S + S' is), then drop
// exact zeros (either sign) as Julia's sparse `+` does.
__global__ void k_symmetrize(const int32_t* __restrict__ colptr, const int32_t* __restrict__ rowval,
                             double* __restrict__ nz, int64_t ncols) {
  // one thread per stored entry in the strict upper triangle copies from the mirror entry
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  for (int p = colptr[c]; p < colptr[c + 1]; ++p) {
    int r = rowval[p];
    if (r >= c) break;  // rows ascending: strict upper part comes first
    // mirror entry (c, r) lives in column r
    int q = find_row(rowval, colptr[r], colptr[r + 1], (int32_t)c);
    nz[p] = nz[q];
  }
}
__global__ void k_nonzero_flags(const double* __restrict__ nz, int64_t n, int32_t* __restrict__ f) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  f[i] = (nz[i] != 0.0) ? 1 : 0;
}
__global__ void k_compact(const int32_t* __restrict__ pos, const int32_t* __restrict__ rowval, const double* __restrict__ nz,
                          int64_t n, int32_t* __restrict__ orow, double* __restrict__ onz) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (nz[i] != 0.0) {
    orow[pos[i]] = rowval[i];
    onz[pos[i]] = nz[i];
  }
}
__global__ void k_compact_colptr(const int32_t* __restrict__ pos, const int32_t* __restrict__ colptr, int64_t ncols,
                                 int32_t* __restrict__ ocolptr) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > ncols) return;
  ocolptr[i] = pos[colptr[i]];
}

int finalize_matrix(fsgpu_ctx* c) {
  c->have_matrix = true;
  c->rrows = c->prows;
  c->rcols = c->pcols;
  c->rnnz = c->pnnz;
  c->compacted = false;
  if (c->target != FSGPU_SPARSE_SYMM) return FSGPU_OK;
  cudaStream_t st = c->stream;
  const int64_t n = c->pnnz;
  LAUNCH(c, k_symmetrize, c->pcols, c->colptr.p, c->rowval.p, c->nzval.p, c->pcols);
  DBuf<int32_t> flags, pos;
  FS_TRY(flags.ensure((size_t)n + 1));
  FS_TRY(pos.ensure((size_t)n + 1));
  FS_CUDA(cudaMemsetAsync(flags.p + n, 0, sizeof(int32_t), st));
  LAUNCH(c, k_nonzero_flags, n, c->nzval.p, n, flags.p);
  size_t tb = 0;
  FS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, flags.p, pos.p, n + 1, st));
  FS_TRY(c->tmp.ensure(tb));
  FS_CUDA(cub::DeviceScan::ExclusiveSum(c->tmp.p, tb, flags.p, pos.p, n + 1, st));
  c->launches += 2;
  int32_t kept = 0;
  FS_CUDA(cudaMemcpyAsync(&kept, pos.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  FS_TRY(c->c_colptr.ensure((size_t)c->pcols + 1));
  FS_TRY(c->c_rowval.ensure((size_t)kept + 1));
  FS_TRY(c->c_nzval.ensure((size_t)kept + 1));
  LAUNCH(c, k_compact, n, pos.p, c->rowval.p, c->nzval.p, n, c->c_rowval.p, c->c_nzval.p);
  LAUNCH(c, k_compact_colptr, c->pcols + 1, pos.p, c->colptr.p, c->pcols, c->c_colptr.p);
  FS_CUDA(cudaStreamSynchronize(st));
  c->compacted = true;
  c->rnnz = kept;
  return FSGPU_OK;
}

}  // namespace fs

extern "C" int fsgpu_result_size(fsgpu_ctx* c, int64_t* nrows, int64_t* ncols, int64_t* nnz) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->have_matrix, FSGPU_ERR_STATE, "no matrix result available");
  if (nrows) *nrows = c->rrows;
  if (ncols) *ncols = c->rcols;
  if (nnz) *nnz = c->rnnz;
  return FSGPU_OK;
}


// ------------------------------------------------------------------------------------
// fetch of a large pattern: compact row indices over PCIe, expanded to Int64 1-based on the host
// ------------------------------------------------------------------------------------
namespace fs {
// Destination arrays are written once and not read here: non-temporal stores (no read-for-ownership traffic),
// 16 bytes at a time where the address allows it.
#if defined(__x86_64__)
static inline void store_nt(int64_t* __restrict__ dst, const int64_t* __restrict__ src, int64_t n) {
  int64_t i = 0;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) && n > 0) {
    _mm_stream_si64(reinterpret_cast<long long*>(dst), (long long)src[0]);
    i = 1;
  }
  for (; i + 1 < n; i += 2)
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i)));
  if (i < n) _mm_stream_si64(reinterpret_cast<long long*>(dst + i), (long long)src[i]);
}
#else
static inline void store_nt(int64_t* __restrict__ dst, const int64_t* __restrict__ src, int64_t n) {
  for (int64_t i = 0; i < n; ++i) dst[i] = src[i];
}
#endif
constexpr int kStageEntries = 1024;  // 8 KB per thread: stays in L1

static void widen_plus1(const int32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
  alignas(64) int64_t buf[kStageEntries];
  for (int64_t o = 0; o < n; o += kStageEntries) {
    const int64_t m = n - o < kStageEntries ? n - o : kStageEntries;
    for (int64_t i = 0; i < m; ++i) buf[i] = (int64_t)in[o + i] + 1;
    store_nt(out + o, buf, m);
  }
#if defined(__x86_64__)
  _mm_sfence();
#endif
}
// runs [k0, k1) of (position, first row) pairs; run k covers positions [runs[k].x, runs[k+1].x)
static void expand_runs(const int2* __restrict__ runs, int64_t k0, int64_t k1, int64_t* __restrict__ out) {
  alignas(64) int64_t buf[kStageEntries + 64];
  if (k0 >= k1) return;
  int64_t base = runs[k0].x;  // output position of buf[0]
  int64_t fill = 0;
  for (int64_t k = k0; k < k1; ++k) {
    int64_t len = (int64_t)runs[k + 1].x - runs[k].x;
    int64_t v = (int64_t)runs[k].y + 1;
    while (len > 0) {
      const int64_t room = kStageEntries - fill;
      const int64_t m = len < room ? len : room;
      for (int64_t i = 0; i < m; ++i) buf[fill + i] = v + i;
      fill += m;
      v += m;
      len -= m;
      if (fill == kStageEntries) {
        store_nt(out + base, buf, fill);
        base += fill;
        fill = 0;
      }
    }
  }
  if (fill) store_nt(out + base, buf, fill);
#if defined(__x86_64__)
  _mm_sfence();
#endif
}

// rows of a pattern as runs of consecutive indices: flag[i] = 1 where rv[i] != rv[i-1] + 1
__global__ void k_run_starts(const int32_t* __restrict__ rv, int64_t n, unsigned char* __restrict__ flag) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || rv[i] != rv[i - 1] + 1) ? 1 : 0;
}
__global__ void k_run_pairs(const int32_t* __restrict__ rv, const int32_t* __restrict__ pos, int64_t nruns, int64_t n,
                            int2* __restrict__ runs) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < nruns) runs[k] = make_int2(pos[k], rv[pos[k]]);
  if (k == nruns) runs[k] = make_int2((int)n, 0);  // sentinel
}

// ---- 8-byte words (values, wide indices) into a PAGEABLE host array ------------------------------------------
// A DMA copy into pageable memory is staged by the driver at ~19 GB/s (C2: 134 ms for the 2.6 GB of values, and a Julia
// `Vector` is pageable).  Here the pieces land in a pinned ring at link speed and host threads move them on with
// non-temporal stores while the next pieces are in flight.
static cudaError_t words_through_ring(fsgpu_ctx* c, const void* dev, void* host, int64_t n, cudaStream_t st) {
  cudaError_t err = ensure_vring(c);
  if (err != cudaSuccess) return err;
  const int64_t* src = static_cast<const int64_t*>(dev);
  int64_t* dst = static_cast<int64_t*>(host);
  int64_t piece = kVRingWords;
  if (const char* ev = getenv("FSGPU_VALUE_RING_WORDS")) {  // tests: force many pieces
    const int64_t v = atoll(ev);
    if (v >= 64 && v < piece) piece = v;
  }
  const int64_t npieces = (n + piece - 1) / piece;
  const int nth = value_threads();
  auto issue = [&](int64_t k) {
    const int64_t o = k * piece, m = n - o < piece ? n - o : piece;
    err = cudaMemcpyAsync(c->vring[k % kVRing], src + o, (size_t)m * 8, cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaEventRecord(c->vring_ev[k % kVRing], st);
  };
  for (int64_t k = 0; k < kVRing && k < npieces && err == cudaSuccess; ++k) issue(k);
  for (int64_t k = 0; k < npieces && err == cudaSuccess; ++k) {
    err = cudaEventSynchronize(c->vring_ev[k % kVRing]);
    if (err != cudaSuccess) break;
    const int64_t o = k * piece, m = n - o < piece ? n - o : piece;
    const int64_t* buf = static_cast<const int64_t*>(c->vring[k % kVRing]);
    auto work = [&, o, m, buf](int t) {
      const int64_t lo = m * t / nth, hi = m * (t + 1) / nth;
      // (measured on a 16-vCPU host: these 16-byte non-temporal stores 73 ms for 2.6 GB with 8 threads, glibc memcpy 92 ms,
      // software prefetch + stores 125 ms)
      store_nt(dst + o + lo, buf + lo, hi - lo);
#if defined(__x86_64__)
      _mm_sfence();
#endif
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nth; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    if (k + kVRing < npieces) issue(k + kVRing);  // this piece's buffer is free again
  }
  return err;
}

constexpr int kRing = 2;                            // double buffer
constexpr int64_t kRingBytes = (int64_t)128 << 20;  // per staging buffer (pinned)

// Builds (once per pattern) the run-length form of the row indices when it is the smaller one.
static int ensure_row_runs(fsgpu_ctx* c, const int32_t* rv, int64_t nnz) {
  if (c->rle_for == rv && c->rle_nnz == nnz) return FSGPU_OK;
  c->rle_for = nullptr;
  c->rle_nruns = 0;
  // scratch kept between calls: cudaMalloc/cudaFree of 200-300 MB blocks costs 1-10 ms, at times far more
  DBuf<unsigned char>& flag = c->scr_rle_flag;
  DBuf<int32_t>& pos = c->scr_rle_pos;
  DBuf<int64_t>& nsel = c->scr_nsel;
  FS_TRY(flag.ensure((size_t)nnz + 1));
  FS_TRY(nsel.ensure(1));
  LAUNCH(c, k_run_starts, nnz, rv, nnz, flag.p);
  // count first (cheap), then materialise only when the run form wins
  size_t tb = 0;
  FS_CUDA(cub::DeviceReduce::Sum(nullptr, tb, flag.p, nsel.p, nnz, c->stream));
  FS_TRY(c->tmp.ensure(tb));
  FS_CUDA(cub::DeviceReduce::Sum(c->tmp.p, tb, flag.p, nsel.p, nnz, c->stream));
  c->launches += 2;
  int64_t nruns = 0;
  FS_CUDA(cudaMemcpyAsync(&nruns, nsel.p, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  c->rle_for = rv;
  c->rle_nnz = nnz;
  if (nruns * 8 > nnz * 4 * 6 / 10) return FSGPU_OK;  // int32 entries are (nearly) as compact: rle_nruns stays 0
  FS_TRY(pos.ensure((size_t)nruns + 1));
  FS_TRY(c->rle_runs.ensure((size_t)(nruns + 1) * sizeof(int2) + 16));
  cub::CountingInputIterator<int32_t> iota(0);
  tb = 0;
  FS_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, iota, flag.p, pos.p, nsel.p, nnz, c->stream));
  FS_TRY(c->tmp.ensure(tb));
  FS_CUDA(cub::DeviceSelect::Flagged(c->tmp.p, tb, iota, flag.p, pos.p, nsel.p, nnz, c->stream));
  c->launches += 2;
  LAUNCH(c, k_run_pairs, nruns + 1, rv, pos.p, nruns, nnz, reinterpret_cast<int2*>(c->rle_runs.p));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  c->rle_nruns = nruns;
  return FSGPU_OK;
}

// mode 1: int32 entries widened by host threads; mode 2: (position, first row) runs expanded by host threads.
// The compact indices arrive in two alternating pinned staging buffers; while the host threads expand one piece the
// next one is in flight, and the values travel on a second stream the whole time.  No spinning: the threads of a
// piece are started when its data has arrived and joined before its buffer is reused.
static int fetch_rows_narrow(fsgpu_ctx* c, const int32_t* rv, int64_t nnz, int64_t* rowval, const double* nz, double* nzval,
                             int mode) {
  if (!c->ring[0]) {
    for (int k = 0; k < kRing; ++k) {
      FS_CUDA(cudaMallocHost(&c->ring[k], (size_t)kRingBytes + 64));
      FS_CUDA(cudaEventCreateWithFlags(&c->ring_ev[k], cudaEventDisableTiming));
    }
    FS_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    FS_CUDA(cudaEventCreateWithFlags(&c->ev_x, cudaEventDisableTiming));
  }
  const bool rle = mode == 2;
  if (rle) FS_TRY(ensure_row_runs(c, rv, nnz));
  const bool use_rle = rle && c->rle_nruns > 0;
  // units: int32 entries, or runs (int2)
  const int64_t nunits = use_rle ? c->rle_nruns : nnz;
  const int64_t unit_bytes = use_rle ? (int64_t)sizeof(int2) : (int64_t)sizeof(int32_t);
  int64_t per_chunk = kRingBytes / unit_bytes;
  if (const char* ev = getenv("FSGPU_FETCH_CHUNK_UNITS")) {  // tests: force many pieces
    const int64_t v = atoll(ev);
    if (v >= 16 && v < per_chunk) per_chunk = v;
  }
  const char* src = use_rle ? reinterpret_cast<const char*>(c->rle_runs.p) : reinterpret_cast<const char*>(rv);
  // Order on the link.  The device-to-host copies of both streams share one copy engine, which serves them in issue
  // order: a single 2.6 GB copy of the values issued first made the (small) run-length pieces wait for all of it, and
  // the host-side expansion (30 ms for C2) then ran after the transfer instead of under it.  So: the first row pieces
  // are issued before any values, and the values follow in pieces of 64 MB with at most two in flight, which lets
  // later row pieces slip in between them.
  cudaError_t nz_err = cudaSuccess;
  std::thread nz_thread;
  FS_CUDA(cudaEventRecord(c->ev_x, c->stream));
  FS_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_x, 0));
  std::atomic<int> rows_issued{0};
  auto start_values = [&] {
    if (!nzval) return;
    c->d2h_bytes += nnz * (int64_t)sizeof(double);
    // from a helper thread: with a pageable destination the copy call blocks its caller
    nz_thread = std::thread([&] {
      cudaSetDevice(c->device);
      while (rows_issued.load(std::memory_order_acquire) == 0) std::this_thread::yield();
      if (host_pageable(nzval)) {
        nz_err = words_through_ring(c, nz, nzval, nnz, c->stream2);
        return;
      }
      int64_t piece = (int64_t)8 << 20;  // entries (64 MB)
      int inflight = 2;
      if (const char* e = getenv("FSGPU_VALUE_PIECE_MB")) piece = (int64_t)(atoi(e) > 0 ? atoi(e) : 64) << 17;
      if (const char* e = getenv("FSGPU_VALUE_INFLIGHT")) inflight = atoi(e) >= 1 && atoi(e) <= 8 ? atoi(e) : 2;
      cudaEvent_t ev[8];
      for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      int64_t k = 0;
      for (int64_t o = 0; o < nnz && nz_err == cudaSuccess; o += piece, ++k) {
        if (k >= inflight) nz_err = cudaEventSynchronize(ev[k % inflight]);
        const int64_t m = nnz - o < piece ? nnz - o : piece;
        if (nz_err == cudaSuccess) nz_err = cudaMemcpyAsync(nzval + o, nz + o, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, c->stream2);
        if (nz_err == cudaSuccess) nz_err = cudaEventRecord(ev[k % inflight], c->stream2);
      }
      if (nz_err == cudaSuccess) nz_err = cudaStreamSynchronize(c->stream2);
      for (auto& e : ev) cudaEventDestroy(e);
    });
  };
  // measured on a 16-core host: 8 threads keep up with the PCIe link; more only compete with the copy-issuing
  // threads for cores (occasional 4x outliers at 16)
  int nth = (int)std::thread::hardware_concurrency() - 2;
  if (const char* lw = getenv("LOCAL_WORLD_SIZE")) {  // one process per GPU on this host: share the cores
    const int w = atoi(lw);
    if (w > 1) nth /= w;
  }
  if (nth > 8) nth = 8;
  if (const char* ev = getenv("FSGPU_HOST_THREADS")) nth = atoi(ev);
  nth = nth < 1 ? 1 : (nth > 32 ? 32 : nth);
  const int64_t nchunks = (nunits + per_chunk - 1) / per_chunk;
  cudaError_t err = cudaSuccess;
  auto issue = [&](int64_t k) {
    const int64_t o = k * per_chunk, m = nunits - o < per_chunk ? nunits - o : per_chunk;
    // runs: one extra entry (the next run's position, or the sentinel) closes the piece's last run
    const size_t bytes = (size_t)(m + (use_rle ? 1 : 0)) * unit_bytes;
    c->d2h_bytes += (int64_t)bytes;
    err = cudaMemcpyAsync(c->ring[k % kRing], src + o * unit_bytes, bytes, cudaMemcpyDeviceToHost, c->stream);
    if (err == cudaSuccess) err = cudaEventRecord(c->ring_ev[k % kRing], c->stream);
  };
  if (nchunks > 0) issue(0);
  if (nchunks > 1) issue(1);
  rows_issued.store(1, std::memory_order_release);
  start_values();
  for (int64_t k = 0; k < nchunks && err == cudaSuccess; ++k) {
    if (k >= 1 && k + 1 < nchunks) issue(k + 1);  // its buffer was released when piece k - 1 was joined
    if (err == cudaSuccess) err = cudaEventSynchronize(c->ring_ev[k % kRing]);
    if (err != cudaSuccess) break;
    const int64_t o = k * per_chunk, m = nunits - o < per_chunk ? nunits - o : per_chunk;
    const void* buf = c->ring[k % kRing];
    auto work = [&, o, m, buf](int t) {
      const int64_t lo = m * t / nth, hi = m * (t + 1) / nth;
      if (use_rle)
        expand_runs(static_cast<const int2*>(buf), lo, hi, rowval);  // positions are absolute
      else
        widen_plus1(static_cast<const int32_t*>(buf) + lo, rowval + o + lo, hi - lo);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nth; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
  }
  if (nz_thread.joinable()) nz_thread.join();
  FS_CUDA(err);
  FS_CUDA(nz_err);
  return FSGPU_OK;
}
}  // namespace fs

extern "C" int fsgpu_d2h_bytes(fsgpu_ctx* c, int64_t* bytes) {
  FS_REQUIRE(c != nullptr && bytes != nullptr, FSGPU_ERR_ARG, "null argument");
  *bytes = c->d2h_bytes;
  return FSGPU_OK;
}
namespace fs {
// device CSC (int32, 0-based) -> host arrays in the Julia layout (Int64, 1-based)
static int fetch_csc(fsgpu_ctx* c, const int32_t* cp, const int32_t* rv, const double* nz, int64_t nc, int64_t nnz, int64_t* colptr,
                     int64_t* rowval, double* nzval, bool transient_pattern) {
  DBuf<int64_t>& wide = c->scr_wide;
  const int64_t chunk = (int64_t)1 << 26;  // widen in chunks to bound the scratch
  {
    const int64_t need = nnz < chunk ? nnz : chunk;
    FS_TRY(wide.ensure((size_t)(need > nc + 1 ? need : nc + 1)));
  }
  if (colptr) {
    LAUNCH(c, k_i32_plus1_to_i64, nc + 1, cp, wide.p, nc + 1);
    if (nc >= ((int64_t)1 << 20) && host_pageable(colptr)) {
      c->d2h_bytes += (nc + 1) * (int64_t)sizeof(int64_t);
      FS_CUDA(words_through_ring(c, wide.p, colptr, nc + 1, c->stream));
    } else {
      FS_TRY(download(c, colptr, wide.p, ((size_t)nc + 1) * sizeof(int64_t)));
      FS_CUDA(cudaStreamSynchronize(c->stream));
    }
  }
  // Large results: the PCIe link (not the GPU) bounds this call, so the row indices cross it as the
  // device's int32 0-based array and are widened to Int64 1-based by host threads out of a pinned ring,
  // while the values travel on a second stream.  Small results: widen on the device (no thread start-up).
  int64_t narrow_min = (int64_t)1 << 22;
  if (const char* ev = getenv("FSGPU_FETCH_NARROW_MIN")) narrow_min = atoll(ev);  // < 0: never (tests)
  const bool narrow = rowval && narrow_min >= 0 && nnz >= narrow_min && nnz > 0;
  if (narrow) {
    // FSGPU_FETCH_MODE=entries: int32 entries; default: run-length form when it is the smaller one
    const char* fm = getenv("FSGPU_FETCH_MODE");
    const int mode = (fm && fm[0] == 'e') ? 1 : 2;
    // index arrays that change from call to call (compacted SparseSymm, CSR, triangles): no caching of their run form
    if (transient_pattern) c->rle_for = nullptr;
    FS_TRY(fetch_rows_narrow(c, rv, nnz, rowval, nzval ? nz : nullptr, nzval, mode));
    if (transient_pattern) c->rle_for = nullptr;
  } else {
    if (rowval && nnz > 0) {
      for (int64_t o = 0; o < nnz; o += chunk) {
        int64_t m = nnz - o < chunk ? nnz - o : chunk;
        LAUNCH(c, k_i32_plus1_to_i64, m, rv + o, wide.p, m);
        FS_TRY(download(c, rowval + o, wide.p, (size_t)m * sizeof(int64_t)));
        FS_CUDA(cudaStreamSynchronize(c->stream));
      }
    }
    if (nzval && nnz > 0) {
      if (nnz >= ((int64_t)1 << 20) && host_pageable(nzval)) {
        c->d2h_bytes += nnz * (int64_t)sizeof(double);
        FS_CUDA(words_through_ring(c, nz, nzval, nnz, c->stream));
      } else {
        FS_TRY(download(c, nzval, nz, (size_t)nnz * sizeof(double)));
        FS_CUDA(cudaStreamSynchronize(c->stream));
      }
    }
  }
  return FSGPU_OK;
}

// one triangle of a square result.  Rows ascend within a column, so the triangle of column j is a suffix
// (lower: rows >= j) or a prefix (upper: rows <= j) of its segment.
__global__ void k_tri_count(const int32_t* __restrict__ cp, const int32_t* __restrict__ rv, int64_t nc, int lower,
                            int32_t* __restrict__ cnt, int32_t* __restrict__ first) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= nc) return;
  int lo = cp[j], hi = cp[j + 1];
  const int b = lo, e = hi;
  const int key = lower ? (int)j : (int)j + 1;  // first position with row >= key
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (rv[mid] < key) lo = mid + 1; else hi = mid;
  }
  first[j] = lower ? lo : b;
  cnt[j] = lower ? e - lo : lo - b;
}
__global__ void k_tri_gather(const int32_t* __restrict__ tcp, const int32_t* __restrict__ first, const int32_t* __restrict__ rv,
                             const double* __restrict__ nz, int64_t nc, int32_t* __restrict__ trv, double* __restrict__ tnz) {
  const int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= nc) return;
  const int o = tcp[j], n = tcp[j + 1] - o, s = first[j];
  for (int k = lane; k < n; k += 32) {
    trv[o + k] = rv[s + k];
    tnz[o + k] = nz[s + k];
  }
}
static int build_triangle(fsgpu_ctx* c, int uplo, int64_t* tnnz) {
  FS_REQUIRE(c->have_matrix, FSGPU_ERR_STATE, "no matrix result available");
  FS_REQUIRE(uplo == 'L' || uplo == 'U', FSGPU_ERR_ARG, "uplo must be 'L' or 'U'");
  FS_REQUIRE(c->target != FSGPU_CSR_SYMM && c->rrows == c->rcols, FSGPU_ERR_STATE, "a triangle needs a square CSC result");
  const int32_t* cp = c->compacted ? c->c_colptr.p : c->colptr.p;
  const int32_t* rv = c->compacted ? c->c_rowval.p : c->rowval.p;
  const int64_t nc = c->rcols;
  FS_TRY(c->t_cnt.ensure((size_t)nc + 1));
  FS_TRY(c->t_first.ensure((size_t)nc + 1));
  FS_TRY(c->t_colptr.ensure((size_t)nc + 1));
  FS_CUDA(cudaMemsetAsync(c->t_cnt.p + nc, 0, sizeof(int32_t), c->stream));
  LAUNCH(c, k_tri_count, nc, cp, rv, nc, uplo == 'L' ? 1 : 0, c->t_cnt.p, c->t_first.p);
  size_t tb = 0;
  FS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, c->t_cnt.p, c->t_colptr.p, nc + 1, c->stream));
  FS_TRY(c->tmp.ensure(tb));
  FS_CUDA(cub::DeviceScan::ExclusiveSum(c->tmp.p, tb, c->t_cnt.p, c->t_colptr.p, nc + 1, c->stream));
  c->launches++;
  int32_t n32 = 0;
  FS_CUDA(cudaMemcpyAsync(&n32, c->t_colptr.p + nc, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  *tnnz = n32;
  return FSGPU_OK;
}
}  // namespace fs

extern "C" int fsgpu_fetch_matrix(fsgpu_ctx* c, int64_t* colptr, int64_t* rowval, double* nzval) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->have_matrix, FSGPU_ERR_STATE, "no matrix result available");
  const int32_t* cp = c->compacted ? c->c_colptr.p : c->colptr.p;
  const int32_t* rv = c->compacted ? c->c_rowval.p : c->rowval.p;
  const double* nz = c->compacted ? c->c_nzval.p : c->nzval.p;
  const int64_t nnz = c->rnnz, nc = c->rcols;
  DBuf<int32_t> rp2, cv2;
  DBuf<double> v2;
  if (c->target == FSGPU_CSR_SYMM) {
    // src/AssemblyModule.jl:47-53: findnz -> sparsecsr
    FS_TRY(csc_to_csr(c, cp, rv, nz, c->rrows, c->rcols, nnz, rp2, cv2, v2));
    cp = rp2.p;
    rv = cv2.p;
    nz = v2.p;
  }
  return fetch_csc(c, cp, rv, nz, nc, nnz, colptr, rowval, nzval, c->compacted || c->target == FSGPU_CSR_SYMM);
}

extern "C" int fsgpu_result_size_uplo(fsgpu_ctx* c, int32_t uplo, int64_t* nnz) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(nnz != nullptr, FSGPU_ERR_ARG, "null argument");
  return build_triangle(c, uplo, nnz);
}
extern "C" int fsgpu_fetch_matrix_uplo(fsgpu_ctx* c, int32_t uplo, int64_t* colptr, int64_t* rowval, double* nzval) {
  FS_TRY(check_ctx(c));
  int64_t tnnz = 0;
  FS_TRY(build_triangle(c, uplo, &tnnz));
  const int32_t* rv = c->compacted ? c->c_rowval.p : c->rowval.p;
  const double* nz = c->compacted ? c->c_nzval.p : c->nzval.p;
  const int64_t nc = c->rcols;
  FS_TRY(c->t_rowval.ensure((size_t)tnnz + 1));
  FS_TRY(c->t_nzval.ensure((size_t)tnnz + 1));
  if (nc > 0) {
    k_tri_gather<<<grid_for(nc * 32, 256), 256, 0, c->stream>>>(c->t_colptr.p, c->t_first.p, rv, nz, nc, c->t_rowval.p, c->t_nzval.p);
    c->launches++;
    FS_CUDA(cudaGetLastError());
  }
  return fetch_csc(c, c->t_colptr.p, c->t_rowval.p, c->t_nzval.p, nc, tnnz, colptr, rowval, nzval, true);
}

extern "C" int fsgpu_fetch_vector(fsgpu_ctx* c, double* out, int64_t n) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->have_vector, FSGPU_ERR_STATE, "no vector result available");
  FS_REQUIRE(out && n == c->vlen, FSGPU_ERR_ARG, "vector length mismatch (have %lld, asked %lld)", (long long)c->vlen,
             (long long)n);
  FS_TRY(download(c, out, c->vec.p, (size_t)n * sizeof(double)));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return FSGPU_OK;
}
extern "C" int fsgpu_result_device(fsgpu_ctx* c, const int32_t** colptr0, const int32_t** rowval0, const double** nzval) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->have_matrix, FSGPU_ERR_STATE, "no matrix result available");
  if (colptr0) *colptr0 = c->compacted ? c->c_colptr.p : c->colptr.p;
  if (rowval0) *rowval0 = c->compacted ? c->c_rowval.p : c->rowval.p;
  if (nzval) *nzval = c->compacted ? c->c_nzval.p : c->nzval.p;
  return FSGPU_OK;
}
// ------------------------------------------------------------------------------------
// column blocks of the result for multi-GPU gathering (owner computes the columns of its
// nodes): Julia-layout pieces (Int64, 1-based GLOBAL rows) written straight into the caller's device
// buffers -- normally this rank's slice of the gathered global arrays, so the collective runs in place.
// ------------------------------------------------------------------------------------
namespace fs {
__global__ void k_block_counts(const int32_t* __restrict__ colptr, int64_t lo, int64_t n, int64_t* __restrict__ cnt) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  cnt[i] = (int64_t)colptr[lo + i + 1] - (int64_t)colptr[lo + i];
}
__global__ void k_block_rows(const int32_t* __restrict__ rv, const int64_t* __restrict__ row_map, int64_t n,
                             int64_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t r = rv[i];
  out[i] = row_map ? row_map[r] : (int64_t)r + 1;
}
}  // namespace fs

extern "C" int fsgpu_result_block(fsgpu_ctx* c, int64_t col_lo, int64_t col_hi, const int64_t* row_map_dev,
                                  int64_t* nnz_block, int64_t* colcount_dev, int64_t* rowval_dev, double* nzval_dev) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->have_matrix, FSGPU_ERR_STATE, "no matrix result available");
  FS_REQUIRE(c->target != FSGPU_SPARSE_SYMM && c->target != FSGPU_CSR_SYMM, FSGPU_ERR_ARG,
             "column blocks are defined for the SPARSE / FFBLOCK / DIAG targets (SPARSE_SYMM needs the mirror "
             "columns of other ranks, CSR_SYMM is row-major)");
  FS_REQUIRE(col_lo >= 0 && col_lo <= col_hi && col_hi <= c->rcols, FSGPU_ERR_ARG,
             "column range [%lld, %lld) outside the result's %lld columns", (long long)col_lo, (long long)col_hi,
             (long long)c->rcols);
  const int32_t* cp = c->colptr.p;
  int32_t ends[2] = {0, 0};
  if (col_hi > col_lo) {
    FS_CUDA(cudaMemcpyAsync(&ends[0], cp + col_lo, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(cudaMemcpyAsync(&ends[1], cp + col_hi, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    FS_CUDA(cudaStreamSynchronize(c->stream));
  }
  const int64_t nb = (int64_t)ends[1] - (int64_t)ends[0];
  if (nnz_block) *nnz_block = nb;
  if (colcount_dev) LAUNCH(c, k_block_counts, col_hi - col_lo, cp, col_lo, col_hi - col_lo, colcount_dev);
  if (rowval_dev) LAUNCH(c, k_block_rows, nb, c->rowval.p + ends[0], row_map_dev, nb, rowval_dev);
  if (nzval_dev && nb > 0)
    FS_CUDA(cudaMemcpyAsync(nzval_dev, c->nzval.p + ends[0], (size_t)nb * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return FSGPU_OK;
}

extern "C" int fsgpu_vector_device(fsgpu_ctx* c, const double** v, int64_t* n) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(c->have_vector, FSGPU_ERR_STATE, "no vector result available");
  if (v) *v = c->vec.p;
  if (n) *n = c->vlen;
  return FSGPU_OK;
}

// ------------------------------------------------------------------------------------
// COO -> CSC (Julia `sparse(I,J,V,m,n)`): stable radix sort on (col,row), segmented sum
// in input order, explicit zeros retained.
// ------------------------------------------------------------------------------------
namespace fs {
// key = (column << rbits) | row with rbits = ceil(log2 m): only the significant ceil(log2 m) + ceil(log2 n) bits are sorted
__global__ void k_coo_keys(const int64_t* __restrict__ I, const int64_t* __restrict__ J, int64_t n, int64_t m, int64_t nc, int rbits,
                           uint64_t* __restrict__ keys, int32_t* __restrict__ flag) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t r = I[i], cc = J[i];
  if (r < 1 || r > m || cc < 1 || cc > nc) {
    atomicExch(flag, 1);
    r = cc = 1;
  }
  keys[i] = ((uint64_t)(cc - 1) << rbits) | (uint64_t)(r - 1);
}
__global__ void k_coo_heads(const uint64_t* __restrict__ keys, int64_t n, unsigned char* __restrict__ head) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
// One thread per stored entry: the duplicates of an entry are adjacent and in INPUT ORDER after the stable sort; they
// are added one after the other, left to right, exactly as Julia's `sparse` combines them (segments of a finite-element
// COO list have <= a few entries)
__global__ void k_coo_segments(const uint64_t* __restrict__ keys, const double* __restrict__ v, const int64_t* __restrict__ start,
                               int64_t nu, int64_t nt, int rbits, int64_t* __restrict__ rowval, double* __restrict__ nz,
                               int64_t* __restrict__ colcnt) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nu) return;
  const int64_t b = start[i], e = (i + 1 < nu) ? start[i + 1] : nt;
  double s = v[b];
  for (int64_t k = b + 1; k < e; ++k) s += v[k];
  nz[i] = s;
  const uint64_t key = keys[b];
  rowval[i] = (int64_t)(key & ((1ull << rbits) - 1)) + 1;
  atomicAdd((unsigned long long*)(colcnt + (key >> rbits)), 1ull);
}
__global__ void k_add1_i64(int64_t* p, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] += 1;
}
__global__ void k_csr_keys(const int32_t* __restrict__ colptr, const int32_t* __restrict__ rowval, int64_t ncols,
                           uint64_t* __restrict__ keys) {
  int64_t cidx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (cidx >= ncols) return;
  for (int p = colptr[cidx]; p < colptr[cidx + 1]; ++p) keys[p] = ((uint64_t)rowval[p] << 32) | (uint64_t)cidx;
}
__global__ void k_csr_unpack(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ colval,
                             int32_t* __restrict__ rowcnt) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  colval[i] = (int32_t)(keys[i] & 0xffffffffu);
  atomicAdd(rowcnt + (keys[i] >> 32), 1);
}
int csc_to_csr(fsgpu_ctx* c, const int32_t* colptr, const int32_t* rowval, const double* nz, int64_t nrows,
               int64_t ncols, int64_t nnz, DBuf<int32_t>& rowptr, DBuf<int32_t>& colval, DBuf<double>& val) {
  cudaStream_t st = c->stream;
  DBuf<uint64_t> k1, k2;
  DBuf<int32_t> cnt;
  FS_TRY(k1.ensure((size_t)nnz + 1));
  FS_TRY(k2.ensure((size_t)nnz + 1));
  FS_TRY(val.ensure((size_t)nnz + 1));
  FS_TRY(colval.ensure((size_t)nnz + 1));
  FS_TRY(rowptr.ensure((size_t)nrows + 1));
  FS_TRY(cnt.ensure((size_t)nrows + 1));
  LAUNCH(c, k_csr_keys, ncols, colptr, rowval, ncols, k1.p);
  size_t tb = 0;
  if (nnz > 0) {
    FS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k1.p, k2.p, nz, val.p, nnz, 0, 64, st));
    FS_TRY(c->tmp.ensure(tb));
    FS_CUDA(cub::DeviceRadixSort::SortPairs(c->tmp.p, tb, k1.p, k2.p, nz, val.p, nnz, 0, 64, st));
    c->launches += 8;
  }
  FS_CUDA(cudaMemsetAsync(cnt.p, 0, ((size_t)nrows + 1) * sizeof(int32_t), st));
  LAUNCH(c, k_csr_unpack, nnz, k2.p, nnz, colval.p, cnt.p);
  FS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, rowptr.p, nrows + 1, st));
  FS_TRY(c->tmp.ensure(tb));
  FS_CUDA(cub::DeviceScan::ExclusiveSum(c->tmp.p, tb, cnt.p, rowptr.p, nrows + 1, st));
  c->launches += 2;
  FS_CUDA(cudaStreamSynchronize(st));
  return FSGPU_OK;
}
}  // namespace fs

extern "C" int fsgpu_coo_to_csc(fsgpu_ctx* c, int64_t m, int64_t n, int64_t nt, const int64_t* I, const int64_t* J,
                                const double* V, int64_t* nnz_out, int64_t* colptr, int64_t* rowval, double* nzval) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(m >= 0 && n >= 0 && nt >= 0 && m < (int64_t)UINT32_MAX && n < (int64_t)UINT32_MAX, FSGPU_ERR_ARG, "bad sizes");
  FS_REQUIRE(nt == 0 || (I && J && V), FSGPU_ERR_ARG, "null triples");
  FS_REQUIRE(nnz_out, FSGPU_ERR_ARG, "null nnz");
  cudaStream_t st = c->stream;
  FS_TRY(ensure_flag(c));
  // two-call convention: the size query does the whole conversion and keeps the result on the device; the fill call
  // with the same arguments only downloads it
  DBuf<int64_t>&drow = c->coo_row, &dptr = c->coo_ptr;
  DBuf<double>& dVr = c->coo_val;
  const bool cached = c->coo_key[0] == (const void*)I && c->coo_key[1] == (const void*)J && c->coo_key[2] == (const void*)V &&
                      c->coo_dims[0] == m && c->coo_dims[1] == n && c->coo_dims[2] == nt && (colptr || rowval || nzval);
  if (cached) {
    const int64_t nu = c->coo_dims[3];
    *nnz_out = nu;
    if (colptr) FS_TRY(download(c, colptr, dptr.p, ((size_t)n + 1) * sizeof(int64_t)));
    if (rowval) FS_TRY(download(c, rowval, drow.p, (size_t)nu * sizeof(int64_t)));
    if (nzval) FS_TRY(download(c, nzval, dVr.p, (size_t)nu * sizeof(double)));
    FS_CUDA(cudaStreamSynchronize(st));
    c->coo_key[0] = c->coo_key[1] = c->coo_key[2] = nullptr;
    return FSGPU_OK;
  }
  c->coo_key[0] = c->coo_key[1] = c->coo_key[2] = nullptr;
  // scratch kept in the context between calls: cudaMalloc / cudaFree of GB-sized blocks cost 0.1 - 1 s on these boxes
  // and made the conversion time vary by 4x from call to call
  DBuf<int64_t>&dI = c->coo_I, &dJ = c->coo_J;
  DBuf<double>&dV = c->coo_V, &dV2 = c->coo_V2;
  DBuf<uint64_t>&k1 = c->scr_keys, &k2 = c->scr_keys2;
  DBuf<int64_t> nu_d, dcnt;
  FS_TRY(dI.ensure((size_t)nt + 1));
  FS_TRY(dJ.ensure((size_t)nt + 1));
  FS_TRY(dV.ensure((size_t)nt + 1));
  FS_TRY(dV2.ensure((size_t)nt + 1));
  FS_TRY(k1.ensure((size_t)nt + 1));
  FS_TRY(k2.ensure((size_t)nt + 1));
  FS_TRY(upload(c, dI.p, I, (size_t)nt * sizeof(int64_t)));
  FS_TRY(upload(c, dJ.p, J, (size_t)nt * sizeof(int64_t)));
  FS_TRY(upload(c, dV.p, V, (size_t)nt * sizeof(double)));
  int rbits = 1, cbits = 1;
  while (((int64_t)1 << rbits) < m) ++rbits;
  while (((int64_t)1 << cbits) < n) ++cbits;
  LAUNCH(c, k_coo_keys, nt, dI.p, dJ.p, nt, m, n, rbits, k1.p, c->flag.p);
  int32_t f;
  FS_TRY(read_flag(c, 0, &f));
  FS_REQUIRE(f == 0, FSGPU_ERR_DOF_RANGE, "row or column index outside the matrix");
  size_t tb = 0, tb2 = 0;
  int64_t nu = 0;
  DBuf<unsigned char>& head = c->scr_rle_flag;
  DBuf<int64_t>& start = c->coo_start;
  FS_TRY(head.ensure((size_t)nt + 1));
  FS_TRY(start.ensure((size_t)nt + 1));
  FS_TRY(dVr.ensure((size_t)nt + 1));
  FS_TRY(nu_d.ensure(1));
  if (nt > 0) {
    // LSD radix sort over the significant key bits only (C2-sized index spaces: 46 instead of 64 bits); it is stable:
    // equal (column, row) keys keep their input order
    FS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k1.p, k2.p, dV.p, dV2.p, nt, 0, rbits + cbits, st));
    cub::CountingInputIterator<int64_t> iota(0);
    FS_CUDA(cub::DeviceSelect::Flagged(nullptr, tb2, iota, head.p, start.p, nu_d.p, nt, st));
    FS_TRY(c->tmp.ensure(tb > tb2 ? tb : tb2));
    FS_CUDA(cub::DeviceRadixSort::SortPairs(c->tmp.p, tb, k1.p, k2.p, dV.p, dV2.p, nt, 0, rbits + cbits, st));
    LAUNCH(c, k_coo_heads, nt, k2.p, nt, head.p);
    FS_CUDA(cub::DeviceSelect::Flagged(c->tmp.p, tb2, iota, head.p, start.p, nu_d.p, nt, st));
    c->launches += 8;
    FS_CUDA(cudaMemcpyAsync(&nu, nu_d.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaStreamSynchronize(st));
  }
  *nnz_out = nu;
  FS_TRY(drow.ensure((size_t)nu + 1));
  FS_TRY(dcnt.ensure((size_t)n + 1));
  FS_TRY(dptr.ensure((size_t)n + 1));
  FS_CUDA(cudaMemsetAsync(dcnt.p, 0, ((size_t)n + 1) * sizeof(int64_t), st));
  LAUNCH(c, k_coo_segments, nu, k2.p, dV2.p, start.p, nu, nt, rbits, drow.p, dVr.p, dcnt.p);
  FS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, dcnt.p, dptr.p, n + 1, st));
  FS_TRY(c->tmp.ensure(tb));
  FS_CUDA(cub::DeviceScan::ExclusiveSum(c->tmp.p, tb, dcnt.p, dptr.p, n + 1, st));
  c->launches += 2;
  LAUNCH(c, k_add1_i64, n + 1, dptr.p, n + 1);
  if (!colptr && !rowval && !nzval) {
    FS_CUDA(cudaStreamSynchronize(st));
    c->coo_key[0] = I;
    c->coo_key[1] = J;
    c->coo_key[2] = V;
    c->coo_dims[0] = m;
    c->coo_dims[1] = n;
    c->coo_dims[2] = nt;
    c->coo_dims[3] = nu;
    return FSGPU_OK;
  }
  if (colptr) FS_TRY(download(c, colptr, dptr.p, ((size_t)n + 1) * sizeof(int64_t)));
  if (rowval) FS_TRY(download(c, rowval, drow.p, (size_t)nu * sizeof(int64_t)));
  if (nzval) FS_TRY(download(c, nzval, dVr.p, (size_t)nu * sizeof(double)));
  FS_CUDA(cudaStreamSynchronize(st));
  return FSGPU_OK;
}
