// libfsgpu explicit central-difference loop with lumped mass and mass-proportional
// damping -- the reference algorithm of
// examples/shells/dynamics/homogeneous/explicit/plate_expl_examples.jl:61-94 (`_cd_loop!`),
// statement order preserved.  The internal force is E = K U with the assembled free-free
// stiffness in CSR (ThreadedSparseCSR.bmul!, :87); the vector updates of :88-91 are fused
// into the SpMV epilogue so one step is two kernels (U update; SpMV + force/velocity/
// acceleration update).  HBM-bound: 8 B per stored entry (f64 value) + 4 B of column index per entry of
// every DISTINCT row pattern: consecutive rows with identical column patterns (the 6 dofs of a shell
// node -- found from the CSR arrays alone, no mesh knowledge) read one shared copy of the indices, which
// cuts the index traffic ~6x (12 -> ~8.7 B per entry).
#include <cub/cub.cuh>

#include <utility>

#include "fsgpu_internal.cuh"

using namespace fs;

struct fsgpu_explicit {
  fsgpu_ctx* ctx = nullptr;
  int64_t n = 0, nnz = 0;
  DBuf<int32_t> rowptr, colval;
  DBuf<int32_t> runs;   // [nruns + 1] first rows of the runs of <= 6 consecutive rows with one column pattern
  int64_t nruns = 0;
  int64_t index_entries = 0;  // column indices actually read per SpMV (one pattern per run)
  DBuf<double> val;
  DBuf<double> M, C, invMC, U, V, A, F0, E, X, Y;
  DBuf<double> Un;       // displacements of the NEXT step, written by the fused step's epilogue
  bool u_ahead = false;  // Un holds U + dt V + dt^2/2 A of the current (V, A): the next step swaps instead of updating
  double dt = 0, c_scale = 0;
  bool have_load = false;
};

namespace {

constexpr int LPR = 8;  // lanes per row

__global__ void k_setup_damping(const double* __restrict__ M, double c_scale, double dt, double* __restrict__ C,
                                double* __restrict__ invMC, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double c = c_scale * M[i];
  C[i] = c;
  invMC[i] = 1.0 / (M[i] + (dt / 2) * c);
}
// U += dt*V + dt^2/2*A   (:85)
__global__ void k_update_u(double* __restrict__ U, const double* __restrict__ V, const double* __restrict__ A, double dt,
                           double dt2_2, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  U[i] += dt * V[i] + dt2_2 * A[i];
}
// Row runs with one shared column pattern ("supernodes", at most SNR rows): LPR lanes walk the pattern once,
// every gathered x entry feeds all rows of the run.  sums[r] holds the row results on every lane of the group.
constexpr int SNR = 6;
__device__ __forceinline__ void run_dot(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colval,
                                        const double* __restrict__ val, const double* __restrict__ x, int64_t r0, int nr,
                                        int sub, double (&sums)[SNR]) {
  const int p0 = rowptr[r0];
  const int len = rowptr[r0 + 1] - p0;  // all rows of the run have this length
  const int32_t* cv = colval + p0;
#pragma unroll
  for (int r = 0; r < SNR; ++r) sums[r] = 0.0;
  if (nr == SNR) {
    for (int k = sub; k < len; k += LPR) {
      const double xv = __ldg(x + cv[k]);
      const double* v = val + p0 + k;
#pragma unroll
      for (int r = 0; r < SNR; ++r) sums[r] = fma(v[(int64_t)r * len], xv, sums[r]);
    }
  } else {
    for (int k = sub; k < len; k += LPR) {
      const double xv = __ldg(x + cv[k]);
      const double* v = val + p0 + k;
#pragma unroll
      for (int r = 0; r < SNR; ++r)
        if (r < nr) sums[r] = fma(v[(int64_t)r * len], xv, sums[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < SNR; ++r)
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sums[r] += __shfl_xor_sync(0xffffffffu, sums[r], o);
}
// lead[r] = r when row r starts a run (its pattern differs from row r - 1, or the run reached SNR rows), else 0
__global__ void k_same_pattern(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colval, int64_t n,
                               int32_t* __restrict__ same) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = g / LPR;
  const int sub = (int)(g % LPR);
  const bool ok = row < n && row > 0;
  int diff = 0;
  if (ok) {
    const int a0 = rowptr[row - 1], a1 = rowptr[row], b1 = rowptr[row + 1];
    if (a1 - a0 != b1 - a1) {
      diff = 1;
    } else {
      for (int p = sub; p < a1 - a0; p += LPR) diff |= colval[a0 + p] != colval[a1 + p];
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) diff |= __shfl_xor_sync(0xffffffffu, diff, o);
  if (row < n && sub == 0) same[row] = (row == 0 || diff) ? 0 : 1;
}
// run starts: row r starts a run when its pattern is new or SNR rows of the same pattern precede it
__global__ void k_run_flags(const int32_t* __restrict__ same, int64_t n, int32_t* __restrict__ flag) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  int back = 0;  // number of consecutive `same` rows ending at r
  while (back <= 8 * SNR && r - back >= 0 && same[r - back]) ++back;
  // rows with a longer history of equal patterns are rare (isolated elements): they start their own run
  flag[r] = (back > 8 * SNR) ? 1 : (back % SNR == 0 ? 1 : 0);
}
__global__ void k_run_entries(const int32_t* __restrict__ runs, int64_t nruns, const int32_t* __restrict__ rowptr,
                              unsigned long long* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  unsigned long long v = 0;
  if (i < nruns) v = (unsigned long long)(rowptr[runs[i] + 1] - rowptr[runs[i]]);
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}
__global__ void k_iota(int32_t* __restrict__ v, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int32_t)i;
}
// y = K x
__global__ void k_spmv(const int32_t* __restrict__ runs, int64_t nruns, const int32_t* __restrict__ rowptr,
                       const int32_t* __restrict__ colval, const double* __restrict__ val, const double* __restrict__ x,
                       double* __restrict__ y) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t run = g / LPR;
  const int sub = (int)(g % LPR);
  const bool ok = run < nruns;
  const int64_t r0 = runs[ok ? run : nruns - 1];
  const int nr = (int)(runs[(ok ? run : nruns - 1) + 1] - r0);
  double sums[SNR];
  run_dot(rowptr, colval, val, x, r0, nr, sub, sums);
#pragma unroll
  for (int r = 0; r < SNR; ++r)
    if (ok && sub == r && r < nr) y[r0 + r] = sums[r];
}
// E = K U, then :88-91 for the row: F = fs*F0 - (E + C (V + dt/2 A)); V += dt/2 A; A = invMC F; V += dt/2 A
__global__ void k_spmv_step(const int32_t* __restrict__ runs, int64_t nruns, const int32_t* __restrict__ rowptr,
                            const int32_t* __restrict__ colval, const double* __restrict__ val,
                            const double* __restrict__ U, const double* __restrict__ F0, double fs,
                            const double* __restrict__ C, const double* __restrict__ invMC, double* __restrict__ V,
                            double* __restrict__ A, double* __restrict__ E, double dt_2, double dt, double dt2_2,
                            double* __restrict__ Unext) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t run = g / LPR;
  const int sub = (int)(g % LPR);
  const bool ok = run < nruns;
  const int64_t r0 = runs[ok ? run : nruns - 1];
  const int nr = (int)(runs[(ok ? run : nruns - 1) + 1] - r0);
  double sums[SNR];
  run_dot(rowptr, colval, val, U, r0, nr, sub, sums);
  double e = 0.0;
#pragma unroll
  for (int r = 0; r < SNR; ++r)
    if (sub == r) e = sums[r];
  if (ok && sub < nr) {
    const int64_t row = r0 + sub;
    const double a0 = A[row];
    double v = V[row];
    double f = (F0 ? fs * F0[row] : 0.0) - (e + C[row] * (v + dt_2 * a0));
    v += dt_2 * a0;
    const double a1 = invMC[row] * f;
    v += dt_2 * a1;
    V[row] = v;
    A[row] = a1;
    E[row] = e;
    // the next step's first statement (:85, U += dt V + dt^2/2 A) for this row, into the other buffer: this
    // kernel still reads U of other rows
    Unext[row] = U[row] + dt * v + dt2_2 * a1;
  }
}
// second half alone (element-partitioned runs: E already summed across ranks)
__global__ void k_finish_step(const double* __restrict__ E, const double* __restrict__ F0, double fs,
                              const double* __restrict__ C, const double* __restrict__ invMC, double* __restrict__ V,
                              double* __restrict__ A, double dt_2, int64_t n) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= n) return;
  const double a0 = A[row];
  double v = V[row];
  const double f = (F0 ? fs * F0[row] : 0.0) - (E[row] + C[row] * (v + dt_2 * a0));
  v += dt_2 * a0;
  const double a1 = invMC[row] * f;
  v += dt_2 * a1;
  V[row] = v;
  A[row] = a1;
}
__global__ void k_start(const double* __restrict__ F0, double fs, const double* __restrict__ invMC, double* __restrict__ A,
                        int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  A[i] = invMC[i] * (F0 ? fs * F0[i] : 0.0);
}
__global__ void k_div(const double* __restrict__ y, const double* __restrict__ M, double* __restrict__ z, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) z[i] = y[i] / M[i];
}
__global__ void k_scale(double* __restrict__ x, const double* __restrict__ y, double s, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) x[i] = s * y[i];
}
__global__ void k_fill_start(double* __restrict__ x, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  // deterministic pseudo-random start vector in (-1, 1)
  uint64_t z = (uint64_t)i * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
  z ^= z >> 31;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 29;
  x[i] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}
// block-level reduction of sum_i w_i a_i b_i into out[0] (double atomics)
__global__ void k_wdot(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ w, int64_t n,
                       double* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += (w ? w[i] : 1.0) * a[i] * b[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

#define XL(h, kern, n, ...)                                                    \
  do {                                                                         \
    if ((n) > 0) {                                                             \
      kern<<<grid_for((n), 256), 256, 0, (h)->ctx->stream>>>(__VA_ARGS__);     \
      (h)->ctx->launches++;                                                    \
    }                                                                          \
  } while (0)

int wdot(fsgpu_explicit* h, const double* a, const double* b, const double* w, double* result) {
  fsgpu_ctx* c = h->ctx;
  FS_TRY(c->flag.ensure(8));
  double* acc = (double*)c->flag.p;  // 8 int32 = 4 doubles of scratch
  FS_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), c->stream));
  if (h->n > 0) {
    k_wdot<<<592, 256, 0, c->stream>>>(a, b, w, h->n, acc);
    c->launches++;
  }
  FS_CUDA(cudaMemcpyAsync(result, acc, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return FSGPU_OK;
}

int alloc_vectors(fsgpu_explicit* h) {
  const size_t n = (size_t)h->n + 1;
  FS_TRY(h->C.ensure(n));
  FS_TRY(h->invMC.ensure(n));
  FS_TRY(h->U.ensure(n));
  FS_TRY(h->V.ensure(n));
  FS_TRY(h->A.ensure(n));
  FS_TRY(h->F0.ensure(n));
  FS_TRY(h->E.ensure(n));
  cudaStream_t st = h->ctx->stream;
  FS_CUDA(cudaMemsetAsync(h->U.p, 0, n * sizeof(double), st));
  FS_CUDA(cudaMemsetAsync(h->V.p, 0, n * sizeof(double), st));
  FS_CUDA(cudaMemsetAsync(h->A.p, 0, n * sizeof(double), st));
  FS_CUDA(cudaMemsetAsync(h->E.p, 0, n * sizeof(double), st));
  XL(h, k_setup_damping, h->n, h->M.p, h->c_scale, h->dt, h->C.p, h->invMC.p, h->n);
  // runs of consecutive rows with one column pattern
  {
    DBuf<int32_t> same, flag, iota;
    DBuf<int64_t> nsel;
    FS_TRY(same.ensure(n));
    FS_TRY(flag.ensure(n));
    FS_TRY(iota.ensure(n));
    FS_TRY(nsel.ensure(1));
    FS_TRY(h->runs.ensure(n + 1));
    XL(h, k_same_pattern, h->n * LPR, h->rowptr.p, h->colval.p, h->n, same.p);
    XL(h, k_run_flags, h->n, same.p, h->n, flag.p);
    XL(h, k_iota, h->n, iota.p, h->n);
    h->nruns = 0;
    if (h->n > 0) {
      size_t tb = 0;
      FS_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, iota.p, flag.p, h->runs.p, nsel.p, h->n, st));
      FS_TRY(h->ctx->tmp.ensure(tb));
      FS_CUDA(cub::DeviceSelect::Flagged(h->ctx->tmp.p, tb, iota.p, flag.p, h->runs.p, nsel.p, h->n, st));
      h->ctx->launches += 2;
      FS_CUDA(cudaMemcpyAsync(&h->nruns, nsel.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      FS_CUDA(cudaStreamSynchronize(st));
      const int32_t last = (int32_t)h->n;
      FS_CUDA(cudaMemcpyAsync(h->runs.p + h->nruns, &last, sizeof(int32_t), cudaMemcpyHostToDevice, st));
      FS_CUDA(cudaMemsetAsync(nsel.p, 0, sizeof(int64_t), st));
      XL(h, k_run_entries, h->nruns, h->runs.p, h->nruns, h->rowptr.p, reinterpret_cast<unsigned long long*>(nsel.p));
      FS_CUDA(cudaMemcpyAsync(&h->index_entries, nsel.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    }
    FS_CUDA(cudaStreamSynchronize(st));
  }
  return FSGPU_OK;
}

__global__ void k_i64_to_i32_m1(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)(in[i] - 1);
}

}  // namespace

extern "C" int fsgpu_explicit_create(fsgpu_explicit** out, fsgpu_ctx* c, int64_t n, const int64_t* rowptr,
                                     const int64_t* colval, const double* nzval, const double* mdiag, double c_scale,
                                     double dt) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(out && rowptr && colval && nzval && mdiag && n >= 0, FSGPU_ERR_ARG, "bad arguments");
  const int64_t nnz = rowptr[n] - 1;
  FS_REQUIRE(nnz >= 0 && nnz < (int64_t)INT32_MAX, FSGPU_ERR_ARG, "bad rowptr");
  fsgpu_explicit* h = new fsgpu_explicit();
  h->ctx = c;
  h->n = n;
  h->nnz = nnz;
  h->dt = dt;
  h->c_scale = c_scale;
  int rc = FSGPU_OK;
  auto fail = [&](int code) {
    delete h;
    return code;
  };
  if ((rc = h->rowptr.ensure((size_t)n + 1))) return fail(rc);
  if ((rc = h->colval.ensure((size_t)nnz + 1))) return fail(rc);
  if ((rc = h->val.ensure((size_t)nnz + 1))) return fail(rc);
  if ((rc = h->M.ensure((size_t)n + 1))) return fail(rc);
  DBuf<int64_t> w;
  if ((rc = w.ensure((size_t)(nnz > n + 1 ? nnz : n + 1) + 1))) return fail(rc);
  if ((rc = upload(c, w.p, rowptr, ((size_t)n + 1) * sizeof(int64_t)))) return fail(rc);
  XL(h, k_i64_to_i32_m1, n + 1, w.p, h->rowptr.p, n + 1);
  cudaStreamSynchronize(c->stream);
  if ((rc = upload(c, w.p, colval, (size_t)nnz * sizeof(int64_t)))) return fail(rc);
  XL(h, k_i64_to_i32_m1, nnz, w.p, h->colval.p, nnz);
  if ((rc = upload(c, h->val.p, nzval, (size_t)nnz * sizeof(double)))) return fail(rc);
  if ((rc = upload(c, h->M.p, mdiag, (size_t)n * sizeof(double)))) return fail(rc);
  cudaStreamSynchronize(c->stream);
  if ((rc = alloc_vectors(h))) return fail(rc);
  *out = h;
  return FSGPU_OK;
}

extern "C" int fsgpu_explicit_create_from_ctx(fsgpu_explicit** out, fsgpu_ctx* c, double c_scale, double dt) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(out, FSGPU_ERR_ARG, "null output");
  FS_REQUIRE(c->have_matrix && c->target == FSGPU_FFBLOCK, FSGPU_ERR_STATE,
             "the context must hold an FFBLOCK stiffness result (SysmatAssemblerFFBlock)");
  FS_REQUIRE(c->have_vector && c->vlen == c->rrows, FSGPU_ERR_STATE,
             "the context must hold the lumped mass vector of the free dofs (fsgpu_shell_mass_diag, nfree_only)");
  fsgpu_explicit* h = new fsgpu_explicit();
  h->ctx = c;
  h->n = c->rrows;
  h->nnz = c->rnnz;
  h->dt = dt;
  h->c_scale = c_scale;
  int rc = csc_to_csr(c, c->colptr.p, c->rowval.p, c->nzval.p, c->rrows, c->rcols, c->rnnz, h->rowptr, h->colval, h->val);
  if (rc == FSGPU_OK) rc = h->M.ensure((size_t)h->n + 1);
  if (rc == FSGPU_OK) {
    cudaError_t e = cudaMemcpyAsync(h->M.p, c->vec.p, (size_t)h->n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
    if (e != cudaSuccess) {
      set_error("CUDA error: %s", cudaGetErrorString(e));
      rc = FSGPU_ERR_CUDA;
    }
  }
  if (rc == FSGPU_OK) rc = alloc_vectors(h);
  if (rc != FSGPU_OK) {
    delete h;
    return rc;
  }
  *out = h;
  return FSGPU_OK;
}

extern "C" int fsgpu_explicit_layout(fsgpu_explicit* h, int64_t* nrows, int64_t* nnz, int64_t* nruns, int64_t* index_entries) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  if (nrows) *nrows = h->n;
  if (nnz) *nnz = h->nnz;
  if (nruns) *nruns = h->nruns;
  if (index_entries) *index_entries = h->index_entries;
  return FSGPU_OK;
}

extern "C" int fsgpu_explicit_destroy(fsgpu_explicit* h) {
  if (!h) return FSGPU_OK;
  cudaSetDevice(h->ctx->device);
  cudaStreamSynchronize(h->ctx->stream);
  delete h;
  return FSGPU_OK;
}

extern "C" int fsgpu_explicit_set_state(fsgpu_explicit* h, const double* U0, const double* V0) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  const size_t b = (size_t)h->n * sizeof(double);
  h->u_ahead = false;
  if (U0) FS_TRY(upload(h->ctx, h->U.p, U0, b));
  if (V0) FS_TRY(upload(h->ctx, h->V.p, V0, b));
  FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_set_load(fsgpu_explicit* h, const double* F0) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  if (F0) {
    FS_TRY(upload(h->ctx, h->F0.p, F0, (size_t)h->n * sizeof(double)));
    FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  }
  h->have_load = F0 != nullptr;
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_start(fsgpu_explicit* h, double fscale0) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  h->u_ahead = false;  // the acceleration changes
  XL(h, k_start, h->n, h->have_load ? h->F0.p : nullptr, fscale0, h->invMC.p, h->A.p, h->n);
  FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_step(fsgpu_explicit* h, int64_t nsteps, const double* fscale) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  const double dt = h->dt;
  FS_TRY(h->Un.ensure((size_t)h->n + 1));
  for (int64_t s = 0; s < nsteps; ++s) {
    if (h->u_ahead) {
      std::swap(h->U.p, h->Un.p);
      std::swap(h->U.n, h->Un.n);
    } else {
      XL(h, k_update_u, h->n, h->U.p, h->V.p, h->A.p, dt, (dt * dt) / 2, h->n);
    }
    XL(h, k_spmv_step, h->nruns * LPR, h->runs.p, h->nruns, h->rowptr.p, h->colval.p, h->val.p, h->U.p,
       h->have_load ? h->F0.p : nullptr, fscale ? fscale[s] : 1.0, h->C.p, h->invMC.p, h->V.p, h->A.p, h->E.p, dt / 2, dt,
       (dt * dt) / 2, h->Un.p);
    h->u_ahead = true;
  }
  FS_CUDA(cudaGetLastError());
  FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_step_begin(fsgpu_explicit* h) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  const double dt = h->dt;
  if (h->u_ahead) {
    std::swap(h->U.p, h->Un.p);
    std::swap(h->U.n, h->Un.n);
    h->u_ahead = false;
  } else {
    XL(h, k_update_u, h->n, h->U.p, h->V.p, h->A.p, dt, (dt * dt) / 2, h->n);
  }
  XL(h, k_spmv, h->nruns * LPR, h->runs.p, h->nruns, h->rowptr.p, h->colval.p, h->val.p, h->U.p, h->E.p);
  FS_CUDA(cudaGetLastError());
  return FSGPU_OK;  // asynchronous: the host's exchange is enqueued on the same stream
}
extern "C" int fsgpu_explicit_step_end(fsgpu_explicit* h, double fscale) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  XL(h, k_finish_step, h->n, h->E.p, h->have_load ? h->F0.p : nullptr, fscale, h->C.p, h->invMC.p, h->V.p, h->A.p,
     h->dt / 2, h->n);
  FS_CUDA(cudaGetLastError());
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_get_state(fsgpu_explicit* h, double* U, double* V, double* A) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  const size_t b = (size_t)h->n * sizeof(double);
  if (U) FS_TRY(download(h->ctx, U, h->U.p, b));
  if (V) FS_TRY(download(h->ctx, V, h->V.p, b));
  if (A) FS_TRY(download(h->ctx, A, h->A.p, b));
  FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_device_state(fsgpu_explicit* h, double** U, double** V, double** A, double** E) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  if (U) *U = h->U.p;
  if (V) *V = h->V.p;
  if (A) *A = h->A.p;
  if (E) *E = h->E.p;
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_spmv(fsgpu_explicit* h, const double* x, double* y) {
  FS_REQUIRE(h && x && y, FSGPU_ERR_ARG, "null argument");
  FS_TRY(check_ctx(h->ctx));
  FS_TRY(h->X.ensure((size_t)h->n + 1));
  FS_TRY(h->Y.ensure((size_t)h->n + 1));
  FS_TRY(upload(h->ctx, h->X.p, x, (size_t)h->n * sizeof(double)));
  XL(h, k_spmv, h->nruns * LPR, h->runs.p, h->nruns, h->rowptr.p, h->colval.p, h->val.p, h->X.p, h->Y.p);
  FS_TRY(download(h->ctx, y, h->Y.p, (size_t)h->n * sizeof(double)));
  FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_omega_max(fsgpu_explicit* h, int32_t maxit, double* lambda_max) {
  FS_REQUIRE(h && lambda_max, FSGPU_ERR_ARG, "null argument");
  FS_TRY(check_ctx(h->ctx));
  FS_TRY(h->X.ensure((size_t)h->n + 1));
  FS_TRY(h->Y.ensure((size_t)h->n + 1));
  XL(h, k_fill_start, h->n, h->X.p, h->n);
  double lam = 0.0;
  for (int it = 0; it < maxit; ++it) {
    // y = M^-1 K x ; lambda = (x' M y) / (x' M x) ; x = y / |y|
    XL(h, k_spmv, h->nruns * LPR, h->runs.p, h->nruns, h->rowptr.p, h->colval.p, h->val.p, h->X.p, h->Y.p);
    XL(h, k_div, h->n, h->Y.p, h->M.p, h->Y.p, h->n);
    double xmy, xmx, yy;
    FS_TRY(wdot(h, h->X.p, h->Y.p, h->M.p, &xmy));
    FS_TRY(wdot(h, h->X.p, h->X.p, h->M.p, &xmx));
    FS_TRY(wdot(h, h->Y.p, h->Y.p, nullptr, &yy));
    lam = xmy / xmx;
    XL(h, k_scale, h->n, h->X.p, h->Y.p, 1.0 / sqrt(yy), h->n);
  }
  *lambda_max = lam;
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_kinetic_energy(fsgpu_explicit* h, double* ke) {
  FS_REQUIRE(h && ke, FSGPU_ERR_ARG, "null argument");
  FS_TRY(check_ctx(h->ctx));
  double s;
  FS_TRY(wdot(h, h->V.p, h->V.p, h->M.p, &s));
  *ke = 0.5 * s;
  return FSGPU_OK;
}

// ---- measurement support: FP64 FMA peak and copy bandwidth of this device ------------------
namespace {
__global__ void k_dfma_peak(double* out, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  const double x = 1.0000001, y = 1e-9 * threadIdx.x;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, x, y);
    a1 = fma(a1, x, y);
    a2 = fma(a2, x, y);
    a3 = fma(a3, x, y);
    a4 = fma(a4, x, y);
    a5 = fma(a5, x, y);
    a6 = fma(a6, x, y);
    a7 = fma(a7, x, y);
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.678) out[0] = s;  // keep the loop alive
}
__global__ void k_copy(const double2* __restrict__ a, double2* __restrict__ b, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] = a[i];
}
}  // namespace

extern "C" int fsgpu_measure_peaks(fsgpu_ctx* c, double* fp64_tflops, double* copy_gbs) {
  FS_TRY(check_ctx(c));
  cudaEvent_t e0, e1;
  FS_CUDA(cudaEventCreate(&e0));
  FS_CUDA(cudaEventCreate(&e1));
  DBuf<double> out;
  FS_TRY(out.ensure(8));
  int sms = 0;
  FS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  float best = 1e30f;
  const int iters = 1 << 15, blocks = sms * 8, threads = 256;
  for (int rep = 0; rep < 5; ++rep) {
    FS_CUDA(cudaEventRecord(e0, c->stream));
    k_dfma_peak<<<blocks, threads, 0, c->stream>>>(out.p, iters, 1.0);
    FS_CUDA(cudaEventRecord(e1, c->stream));
    FS_CUDA(cudaEventSynchronize(e1));
    float ms;
    FS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  c->launches += 5;
  if (fp64_tflops) *fp64_tflops = (double)blocks * threads * iters * 8 * 2 / (best * 1e-3) / 1e12;
  if (copy_gbs) {
    const int64_t n = (int64_t)1 << 26;  // 2 x 1 GiB of double2
    DBuf<double2> a, b;
    FS_TRY(a.ensure((size_t)n));
    FS_TRY(b.ensure((size_t)n));
    FS_CUDA(cudaMemsetAsync(a.p, 0, (size_t)n * sizeof(double2), c->stream));
    best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      FS_CUDA(cudaEventRecord(e0, c->stream));
      k_copy<<<sms * 16, 512, 0, c->stream>>>(a.p, b.p, n);
      FS_CUDA(cudaEventRecord(e1, c->stream));
      FS_CUDA(cudaEventSynchronize(e1));
      float ms;
      FS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    c->launches += 6;
    *copy_gbs = 2.0 * (double)n * sizeof(double2) / (best * 1e-3) / 1e9;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return FSGPU_OK;
}
