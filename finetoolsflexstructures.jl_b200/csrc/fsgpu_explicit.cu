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
#include <unistd.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "fsgpu_internal.cuh"

using namespace fs;

struct fsgpu_explicit {
  fsgpu_ctx* ctx = nullptr;
  int64_t n = 0, nnz = 0;
  int64_t row0 = 0;  // global index of the first row (row-partitioned runs)
  DBuf<int32_t> rowptr, colval;
  DBuf<int32_t> runs;   // [nruns + 1] first rows of the runs of <= 6 consecutive rows with one column pattern
  int64_t nruns = 0;
  int64_t index_entries = 0;  // column indices actually read per SpMV (one pattern per run)
  DBuf<double> val;
  DBuf<double> M, C, invMC, U, V, A, F0, E, X, Y;
  DBuf<double> Un;       // displacements of the NEXT step, written by the fused step's epilogue
  // current / next displacement buffers: U.p / Un.p, or -- row-partitioned runs -- two vectors of the peer-visible window
  double* Uc = nullptr;
  double* Unx = nullptr;
  bool u_ahead = false;  // Unx holds U + dt V + dt^2/2 A of the current (V, A): the next step swaps instead of updating
  struct Dist* dist = nullptr;  // row-partitioned run (fsgpu_explicit_create_dist): peers, window, halo lists
  double dt = 0, c_scale = 0;
  bool have_load = false;
};

// ---- row-partitioned runs (one rank per GPU) -------------------------------------------
// Every rank owns a contiguous block of rows of the global K_ff and holds, besides its own entries of the
// displacement vector, HALO entries (the columns its rows reference in other ranks' blocks).  All peer-visible
// state lives in one device allocation per rank, the WINDOW, which the other ranks map (cudaIpc across
// processes, the plain pointer inside one process) and write directly over NVLink:
//   [0, kCtrlBytes)   control words (uint64): step flags [sender rank], barrier flags, reduction flags (one 128 B
//                     line each), reduction values double[2][kMaxWorld][kRedN] at byte kRedOff
//   then kNVec vectors of `ext` doubles: displacement buffers 0 / 1 and the power-iteration vector, each
//                     [own rows | padding to a 128 B line | halo entries grouped by owner rank, ascending]
constexpr int kMaxWorld = 32;
constexpr int kMaxPeers = 16;
constexpr size_t kCtrlBytes = 32768;
// uint64 word offsets of the three flag arrays; the flag of sender rank p is word base + p * kFlagStride: every
// flag has its own 128 B line (one writer -- the sender --, one polling reader)
constexpr int kFlagStride = 16;
constexpr int kFlagStep = 0, kFlagBar = kMaxWorld * kFlagStride, kFlagRed = 2 * kMaxWorld * kFlagStride;
constexpr size_t kRedOff = 16384;
constexpr int kRedN = 8;
constexpr int kNVec = 3;
constexpr uint64_t kBlobMagic = 0x46534750555f5631ull;  // "FSGPU_V1"

struct DistDev {  // by-value kernel argument
  int me, world, npeers;
  int nb_ctas;  // CTAs [0, nb_ctas) of the fused step hold the boundary runs (rows that read halo entries)
  int n_push;
  unsigned long long wait_epoch;  // step flags of all halo peers must have reached this value before halo reads
  unsigned long long timeout_ns;
  unsigned char* ctrl[kMaxWorld];  // window base of every rank (ctrl[me] = own)
  int peer_rank[kMaxPeers];        // halo peers
  double* peer_vec[kMaxPeers];     // the vector of the peer's window this launch pushes into
  const int32_t* push_src;         // [n_push] own row
  const int32_t* push_dst;         // [n_push] entry of the peer's vector
  const int32_t* push_peer;        // [n_push] index into peer_rank / peer_vec
  const int32_t* order;            // [nruns] run handled by slot s: boundary runs first
  unsigned int* done;              // boundary CTAs finished (fused step)
  int* err;                        // set when a wait timed out
};

struct Dist {
  int rank = 0, world = 1;
  unsigned char* win = nullptr;
  size_t win_bytes = 0;
  int64_t n_own = 0, n_halo = 0, halo0 = 0, ext = 0;
  int64_t halo_base[kMaxWorld] = {}, halo_cnt[kMaxWorld] = {};  // my halo entries of rank p's rows
  int64_t peer_ext[kMaxWorld] = {};
  int64_t send_cnt[kMaxWorld] = {};
  std::vector<int32_t> push_src, push_peer, push_rank;  // host copies (push_dst is known at connect)
  int npeers = 0;
  int peer_rank[kMaxPeers] = {};
  unsigned char* ctrl[kMaxWorld] = {};
  bool ipc_opened[kMaxWorld] = {};
  bool connected = false;
  fs::DBuf<int32_t> d_push_src, d_push_dst, d_push_peer, order;
  fs::DBuf<unsigned int> done;
  fs::DBuf<int> err;
  fs::DBuf<double> red;  // [2 * kRedN] allreduce in / out
  // pinned host scratch for the few bytes the collective calls read back (reduction results, error flag): a copy
  // to PAGEABLE memory queued behind a kernel that waits for a peer holds driver-internal staging resources, and a
  // pageable copy of another rank of the same process then blocks behind it -- the peer never launches, deadlock
  double* hpin = nullptr;
  int n_push = 0, nb_ctas = 0;
  int64_t n_bruns = 0;
  unsigned long long epoch = 0, bar_epoch = 0, red_epoch = 0;
  int cur = 0;  // window vector that holds the current displacements
  unsigned long long timeout_ns = 20000000000ull;
  cudaStream_t own_stream = nullptr;
  double* vec(int p, int b) const { return (double*)(ctrl[p] + kCtrlBytes) + (size_t)b * (size_t)peer_ext[p]; }
};

struct DistBlob {  // what fsgpu_explicit_export writes; FSGPU_EXPLICIT_BLOB_BYTES in fsgpu.h
  uint64_t magic;
  int32_t rank, world;
  int64_t pid;
  uint64_t proc_token;
  int32_t device, pad;
  uint64_t devptr, win_bytes;
  int64_t n_own, ext;
  cudaIpcMemHandle_t ipc;
  int64_t halo_base[kMaxWorld], halo_cnt[kMaxWorld];
};
static_assert(sizeof(DistBlob) <= 1024, "blob size");

namespace {

constexpr int LPR = 8;  // lanes per row

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// wait until *p >= target (a peer's release store); gives up after timeout_ns and raises *err
__device__ __noinline__ bool spin_until(const unsigned long long* p, unsigned long long target, unsigned long long timeout_ns,
                                        int* err) {
  if (ld_acquire_sys(p) >= target) return true;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(p) < target) {
    if (*(volatile int*)err) return false;
    if (global_ns() - t0 > timeout_ns) {
      atomicExch(err, 1);
      return false;
    }
    __nanosleep(64);
  }
  return true;
}
// flag `base` (kFlagStep / kFlagBar / kFlagRed) of sender `from` in the window of rank `rank`
__device__ __forceinline__ unsigned long long* ctrl_flag(const DistDev& d, int rank, int base, int from) {
  return reinterpret_cast<unsigned long long*>(d.ctrl[rank]) + base + from * kFlagStride;
}
// one thread waits for the flags of a list of senders, one after the other (no divergent spinning inside a warp)
__device__ __forceinline__ void wait_flags(const DistDev& d, int base, const int* senders, int n, bool all_ranks,
                                           unsigned long long target) {
  for (int k = 0; k < n; ++k) {
    const int from = all_ranks ? k : senders[k];
    if (from == d.me) continue;
    if (!spin_until(ctrl_flag(d, d.me, base, from), target, d.timeout_ns, d.err)) return;
  }
}

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
// DIST (row-partitioned run): the first d.nb_ctas CTAs hold the boundary runs -- they wait for the peers' halo
// entries of U (step flags), and when the last of them is done that CTA writes the boundary entries of the NEXT
// displacements straight into the peers' windows and raises this rank's step flag there, while the interior CTAs
// are still working: the exchange rides under the interior rows.
template <bool DIST>
__global__ void k_spmv_step(const int32_t* __restrict__ runs, int64_t nruns, const int32_t* __restrict__ rowptr,
                            const int32_t* __restrict__ colval, const double* __restrict__ val,
                            const double* __restrict__ U, const double* __restrict__ F0, double fs,
                            const double* __restrict__ C, const double* __restrict__ invMC, double* __restrict__ V,
                            double* __restrict__ A, double* __restrict__ E, double dt_2, double dt, double dt2_2,
                            double* __restrict__ Unext, const DistDev d) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t slot = g / LPR;
  const int sub = (int)(g % LPR);
  const bool ok = slot < nruns;
  int64_t run = ok ? slot : nruns - 1;
  if (DIST) {
    if ((int)blockIdx.x < d.nb_ctas) {
      if (threadIdx.x == 0) wait_flags(d, kFlagStep, d.peer_rank, d.npeers, false, d.wait_epoch);
      __syncthreads();
    }
    run = d.order[run];
  }
  const int64_t r0 = runs[run];
  const int nr = (int)(runs[run + 1] - r0);
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
  if (DIST) {
    if ((int)blockIdx.x < d.nb_ctas) {
      __shared__ int s_last;
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) s_last = atomicAdd(d.done, 1u) == (unsigned)(d.nb_ctas - 1);
      __syncthreads();
      if (s_last) {
        __threadfence();
        for (int i = threadIdx.x; i < d.n_push; i += blockDim.x)
          d.peer_vec[d.push_peer[i]][d.push_dst[i]] = __ldcg(Unext + d.push_src[i]);
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < d.npeers)
          st_release_sys(ctrl_flag(d, d.peer_rank[threadIdx.x], kFlagStep, d.me), d.wait_epoch + 1);
        if (threadIdx.x == 0) *d.done = 0u;
      }
    }
  }
}
// all ranks: every rank's stream has reached this point (previous kernels of the stream are complete)
__global__ void k_dist_barrier(const DistDev d, unsigned long long value) {
  const int t = threadIdx.x;
  if (t < d.world && t != d.me) {
    __threadfence_system();
    st_release_sys(ctrl_flag(d, t, kFlagBar, d.me), value);
  }
  __syncthreads();
  if (t == 0) wait_flags(d, kFlagBar, nullptr, d.world, true, value);
}
// boundary entries of `vec` (own rows) into the halo entries of the peers' vectors, then the step flag
__global__ void k_dist_push(const DistDev d, const double* __restrict__ vec, unsigned long long value) {
  for (int i = threadIdx.x; i < d.n_push; i += blockDim.x) d.peer_vec[d.push_peer[i]][d.push_dst[i]] = vec[d.push_src[i]];
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < d.npeers) st_release_sys(ctrl_flag(d, d.peer_rank[threadIdx.x], kFlagStep, d.me), value);
}
// the halo entries announced by step flag d.wait_epoch have arrived
__global__ void k_dist_wait(const DistDev d) {
  if (threadIdx.x == 0) wait_flags(d, kFlagStep, d.peer_rank, d.npeers, false, d.wait_epoch);
}
// io[0..nv) <- sum (op 0) or max (op 1) over all ranks, summed in rank order: identical bits on every rank
__global__ void k_dist_allreduce(const DistDev d, double* __restrict__ io, int nv, unsigned long long r, int op) {
  const int t = threadIdx.x;
  const int slot = (int)(r & 1ull);
  if (t < d.world) {
    double* dst = reinterpret_cast<double*>(d.ctrl[t] + kRedOff) + ((size_t)slot * kMaxWorld + d.me) * kRedN;
    for (int k = 0; k < nv; ++k) dst[k] = io[k];
    __threadfence_system();
    if (t != d.me) st_release_sys(ctrl_flag(d, t, kFlagRed, d.me), r);
  }
  __syncthreads();
  if (t == 0) wait_flags(d, kFlagRed, nullptr, d.world, true, r);
  __syncthreads();
  if (t < nv) {
    const volatile double* src = reinterpret_cast<const volatile double*>(d.ctrl[d.me] + kRedOff) + (size_t)slot * kMaxWorld * kRedN;
    double s = src[t];
    for (int p = 1; p < d.world; ++p) {
      const double x = src[(size_t)p * kRedN + t];
      s = op == 0 ? s + x : (x > s ? x : s);
    }
    io[t] = s;
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
__global__ void k_fill_start(double* __restrict__ x, int64_t n, int64_t row0) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  // deterministic pseudo-random start vector in (-1, 1)
  uint64_t z = (uint64_t)(i + row0) * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
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

DistDev dist_dev(const fsgpu_explicit* h, int push_vec) {
  const Dist* D = h->dist;
  DistDev d;
  memset(&d, 0, sizeof d);
  d.me = D->rank;
  d.world = D->world;
  d.npeers = D->npeers;
  d.nb_ctas = D->nb_ctas;
  d.n_push = D->n_push;
  d.wait_epoch = D->epoch;
  d.timeout_ns = D->timeout_ns;
  for (int p = 0; p < D->world; ++p) d.ctrl[p] = D->ctrl[p];
  for (int k = 0; k < D->npeers; ++k) {
    d.peer_rank[k] = D->peer_rank[k];
    d.peer_vec[k] = push_vec >= 0 ? D->vec(D->peer_rank[k], push_vec) : nullptr;
  }
  d.push_src = D->d_push_src.p;
  d.push_dst = D->d_push_dst.p;
  d.push_peer = D->d_push_peer.p;
  d.order = D->order.p;
  d.done = D->done.p;
  d.err = D->err.p;
  return d;
}
int dist_check(fsgpu_explicit* h) {  // synchronises the stream
  Dist* D = h->dist;
  int* eh = reinterpret_cast<int*>(D->hpin + 16);
  FS_CUDA(cudaMemcpyAsync(eh, D->err.p, sizeof(int), cudaMemcpyDeviceToHost, h->ctx->stream));
  FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  if (*eh) {
    if (getenv("FSGPU_DIST_TRACE")) {
      static unsigned long long w[3 * kMaxWorld * kFlagStride];
      cudaMemcpy(w, D->win, sizeof w, cudaMemcpyDeviceToHost);
      fprintf(stderr, "[fsgpu rank %d] timeout: epoch %llu bar %llu red %llu |", D->rank, D->epoch, D->bar_epoch, D->red_epoch);
      for (int p = 0; p < D->world; ++p)
        fprintf(stderr, " from %d: step %llu bar %llu red %llu;", p, w[kFlagStep + p * kFlagStride], w[kFlagBar + p * kFlagStride],
                w[kFlagRed + p * kFlagStride]);
      fprintf(stderr, "\n");
    }
    cudaMemsetAsync(D->err.p, 0, sizeof(int), h->ctx->stream);
    cudaStreamSynchronize(h->ctx->stream);
    set_error("rank %d: waiting for a peer rank timed out (%.1f s): ranks out of step, or a peer has gone", D->rank,
              (double)D->timeout_ns * 1e-9);
    return FSGPU_ERR_STATE;
  }
  return FSGPU_OK;
}
int dist_require(fsgpu_explicit* h) {
  FS_REQUIRE(!h->dist || h->dist->connected, FSGPU_ERR_STATE,
             "row-partitioned run: call fsgpu_explicit_export / fsgpu_explicit_connect on every rank first");
  return FSGPU_OK;
}
// all ranks: barrier of the streams (asynchronous on this rank's stream)
int dist_barrier(fsgpu_explicit* h) {
  Dist* D = h->dist;
  if (D->world == 1) return FSGPU_OK;
  k_dist_barrier<<<1, kMaxWorld, 0, h->ctx->stream>>>(dist_dev(h, -1), ++D->bar_epoch);
  h->ctx->launches++;
  FS_CUDA(cudaGetLastError());
  return FSGPU_OK;
}
// halo entries of window vector b on all ranks <- the owners' current values; complete for kernels launched after
int dist_sync_halo(fsgpu_explicit* h, int b) {
  Dist* D = h->dist;
  if (D->world == 1) return FSGPU_OK;
  FS_TRY(dist_barrier(h));  // nobody still reads the halo entries this overwrites
  k_dist_push<<<1, 1024, 0, h->ctx->stream>>>(dist_dev(h, b), D->vec(D->rank, b), D->epoch + 1);
  D->epoch++;
  k_dist_wait<<<1, 32, 0, h->ctx->stream>>>(dist_dev(h, -1));
  h->ctx->launches += 2;
  FS_CUDA(cudaGetLastError());
  return FSGPU_OK;
}
// dev[0..nv) <- sum / max over all ranks, in place on device memory (asynchronous on this rank's stream)
int dist_allreduce_dev(fsgpu_explicit* h, double* dev, int nv, int op) {
  Dist* D = h->dist;
  if (D->world == 1) return FSGPU_OK;
  k_dist_allreduce<<<1, kMaxWorld, 0, h->ctx->stream>>>(dist_dev(h, -1), dev, nv, ++D->red_epoch, op);
  h->ctx->launches++;
  FS_CUDA(cudaGetLastError());
  return FSGPU_OK;
}

int wdot(fsgpu_explicit* h, const double* a, const double* b, const double* w, double* result) {
  fsgpu_ctx* c = h->ctx;
  FS_TRY(c->flag.ensure(8));
  double* acc = (double*)c->flag.p;  // 8 int32 = 4 doubles of scratch
  FS_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), c->stream));
  if (h->n > 0) {
    k_wdot<<<592, 256, 0, c->stream>>>(a, b, w, h->n, acc);
    c->launches++;
  }
  if (h->dist) {
    FS_TRY(dist_allreduce_dev(h, acc, 1, 0));  // partial sums of the row blocks, combined in rank order
    FS_CUDA(cudaMemcpyAsync(h->dist->hpin, acc, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    FS_TRY(dist_check(h));
    *result = h->dist->hpin[0];
    return FSGPU_OK;
  }
  FS_CUDA(cudaMemcpyAsync(result, acc, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  FS_CUDA(cudaStreamSynchronize(c->stream));
  return FSGPU_OK;
}

int alloc_vectors(fsgpu_explicit* h) {
  const size_t n = (size_t)h->n + 1;
  FS_TRY(h->C.ensure(n));
  FS_TRY(h->invMC.ensure(n));
  if (h->dist) {  // the displacement buffers are vectors 0 / 1 of the peer-visible window (zeroed at creation)
    h->Uc = h->dist->vec(h->dist->rank, 0);
    h->Unx = h->dist->vec(h->dist->rank, 1);
  } else {
    FS_TRY(h->U.ensure(n));
    h->Uc = h->U.p;
  }
  FS_TRY(h->V.ensure(n));
  FS_TRY(h->A.ensure(n));
  FS_TRY(h->F0.ensure(n));
  FS_TRY(h->E.ensure(n));
  cudaStream_t st = h->ctx->stream;
  if (!h->dist) FS_CUDA(cudaMemsetAsync(h->U.p, 0, n * sizeof(double), st));
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

// a malformed CSR would make the step kernel read out of bounds: rowptr must start at 0 and not decrease, every
// column index must lie in [0, n)  (0-based copies); flag[0] counts violations
__global__ void k_csr_validate(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colval, int64_t n, int64_t nnz,
                               int32_t* __restrict__ flag) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  bool bad = false;
  if (i < n) bad = rowptr[i] > rowptr[i + 1] || (i == 0 && rowptr[0] != 0) || rowptr[i + 1] > nnz;
  if (i < nnz) bad = bad || colval[i] < 0 || colval[i] >= n;
  if (bad) atomicAdd(flag, 1);
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
  {
    DBuf<int32_t> bad;
    int32_t nbad = 0;
    if ((rc = bad.ensure(1))) return fail(rc);
    cudaMemsetAsync(bad.p, 0, sizeof(int32_t), c->stream);
    XL(h, k_csr_validate, (n > nnz ? n : nnz), h->rowptr.p, h->colval.p, n, nnz, bad.p);
    cudaMemcpyAsync(&nbad, bad.p, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    if (nbad != 0) {
      set_error("malformed CSR: rowptr must be 1-based and non-decreasing, column indices in 1..%lld (%d violations)", (long long)n, (int)nbad);
      return fail(FSGPU_ERR_ARG);
    }
  }
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

namespace {
void dist_release(fsgpu_explicit* h) {
  Dist* D = h->dist;
  if (!D) return;
  for (int p = 0; p < D->world; ++p)
    if (D->ipc_opened[p] && D->ctrl[p]) cudaIpcCloseMemHandle(D->ctrl[p]);
  if (D->win) cudaFree(D->win);
  if (D->hpin) cudaFreeHost(D->hpin);
  if (D->own_stream) {
    if (h->ctx->stream == D->own_stream) h->ctx->stream = 0;
    cudaStreamDestroy(D->own_stream);
  }
  delete D;
  h->dist = nullptr;
}
}  // namespace

extern "C" int fsgpu_explicit_destroy(fsgpu_explicit* h) {
  if (!h) return FSGPU_OK;
  cudaSetDevice(h->ctx->device);
  if (h->dist && h->dist->connected && h->dist->world > 1) {
    // peers may still be writing their last halo entries into this rank's window: leave together
    h->dist->timeout_ns = std::min<unsigned long long>(h->dist->timeout_ns, 5000000000ull);
    dist_barrier(h);
  }
  cudaStreamSynchronize(h->ctx->stream);
  dist_release(h);
  delete h;
  return FSGPU_OK;
}

// ---- row-partitioned run ----------------------------------------------------------------------------
namespace {
// per own row: bit p set when the row has an entry in a column owned by rank p; used[col] = 1 for every such column
__global__ void k_dist_scan(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colval,
                            const signed char* __restrict__ owner, int64_t n, unsigned char* __restrict__ used,
                            uint32_t* __restrict__ rowmask) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = g / LPR;
  const int sub = (int)(g % LPR);
  uint32_t m = 0;
  if (row < n) {
    for (int p = rowptr[row] + sub; p < rowptr[row + 1]; p += LPR) {
      const int c = colval[p];
      const int o = owner[c];
      if (o >= 0) {
        used[c] = 1;
        m |= 1u << o;
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
  if (row < n && sub == 0) rowmask[row] = m;
}
__global__ void k_remap(int32_t* __restrict__ colval, const int32_t* __restrict__ colmap, int64_t nnz) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < nnz) colval[i] = colmap[colval[i]];
}
__global__ void k_shift_i32(const int32_t* __restrict__ in, int32_t* __restrict__ out, int32_t base, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] - base;
}
uint64_t process_token() {
  static uint64_t tok = 0;
  if (!tok) {
    tok = ((uint64_t)getpid() << 32) ^ (uint64_t)(uintptr_t)&tok ^ 0x9E3779B97F4A7C15ull;
    if (!tok) tok = 1;
  }
  return tok;
}
}  // namespace

extern "C" int fsgpu_explicit_create_dist(fsgpu_explicit** out, fsgpu_ctx* c, int32_t rank, int32_t world, int64_t row_lo,
                                          int64_t row_hi, const int64_t* loc2glob, const int64_t* bounds, double c_scale,
                                          double dt) {
  FS_TRY(check_ctx(c));
  FS_REQUIRE(out && bounds, FSGPU_ERR_ARG, "null argument");
  FS_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, FSGPU_ERR_ARG, "rank %d / world %d (at most %d ranks)",
             rank, world, kMaxWorld);
  FS_REQUIRE(c->have_matrix && c->target == FSGPU_FFBLOCK, FSGPU_ERR_STATE,
             "the context must hold an FFBLOCK stiffness result (SysmatAssemblerFFBlock)");
  FS_REQUIRE(c->have_vector && c->vlen == c->rrows, FSGPU_ERR_STATE,
             "the context must hold the lumped mass vector of the free dofs (fsgpu_shell_mass_diag, nfree_only)");
  const int64_t nfl = c->rrows;
  FS_REQUIRE(0 <= row_lo && row_lo < row_hi && row_hi <= nfl, FSGPU_ERR_ARG, "own rows [%lld, %lld) of %lld", (long long)row_lo,
             (long long)row_hi, (long long)nfl);
  for (int p = 0; p < world; ++p) FS_REQUIRE(bounds[p] <= bounds[p + 1], FSGPU_ERR_ARG, "bounds must ascend");
  FS_REQUIRE(bounds[rank + 1] - bounds[rank] == row_hi - row_lo, FSGPU_ERR_ARG,
             "bounds[rank+1] - bounds[rank] = %lld but %lld own rows", (long long)(bounds[rank + 1] - bounds[rank]),
             (long long)(row_hi - row_lo));
  const int64_t g_lo = loc2glob ? loc2glob[row_lo] - 1 : row_lo;
  FS_REQUIRE(g_lo == bounds[rank], FSGPU_ERR_ARG, "first own row is global row %lld, bounds[rank] = %lld", (long long)g_lo,
             (long long)bounds[rank]);
  const int64_t n_own = row_hi - row_lo;
  fsgpu_explicit* h = new fsgpu_explicit();
  h->ctx = c;
  Dist* D = h->dist = new Dist();
  D->rank = rank;
  D->world = world;
  if (const char* ev = getenv("FSGPU_PEER_TIMEOUT_MS")) D->timeout_ns = (unsigned long long)atoll(ev) * 1000000ull;
  int rc = FSGPU_OK;
  auto body = [&]() -> int {
    if (c->stream == 0) {  // kernels that wait for peers must not sit on the legacy default stream
      FS_CUDA(cudaStreamCreateWithFlags(&D->own_stream, cudaStreamNonBlocking));
      FS_CUDA(cudaStreamSynchronize(0));
      c->stream = D->own_stream;
    }
    cudaStream_t st = c->stream;
    // owner rank of every local column (-1: own)
    std::vector<signed char> owner((size_t)nfl);
    for (int64_t k = 0; k < nfl; ++k) {
      if (k >= row_lo && k < row_hi) {
        owner[k] = -1;
        continue;
      }
      const int64_t g = loc2glob ? loc2glob[k] - 1 : k;
      const int p = (int)(std::upper_bound(bounds, bounds + world + 1, g) - bounds) - 1;
      FS_REQUIRE(p >= 0 && p < world && p != rank, FSGPU_ERR_ARG, "local row %lld (global %lld) has no owner rank", (long long)k,
                 (long long)g);
      owner[k] = (signed char)p;
    }
    // CSR of the local matrix, own rows sliced out
    DBuf<int32_t> rp, cv;
    DBuf<double> vv;
    FS_TRY(csc_to_csr(c, c->colptr.p, c->rowval.p, c->nzval.p, c->rrows, c->rcols, c->rnnz, rp, cv, vv));
    int32_t ends[2];
    FS_CUDA(cudaMemcpy(&ends[0], rp.p + row_lo, sizeof(int32_t), cudaMemcpyDeviceToHost));
    FS_CUDA(cudaMemcpy(&ends[1], rp.p + row_hi, sizeof(int32_t), cudaMemcpyDeviceToHost));
    const int64_t nnz = (int64_t)ends[1] - ends[0];
    h->n = n_own;
    h->row0 = bounds[rank];
    h->nnz = nnz;
    h->dt = dt;
    h->c_scale = c_scale;
    FS_TRY(h->rowptr.ensure((size_t)n_own + 1));
    FS_TRY(h->colval.ensure((size_t)nnz + 1));
    FS_TRY(h->val.ensure((size_t)nnz + 1));
    FS_TRY(h->M.ensure((size_t)n_own + 1));
    XL(h, k_shift_i32, n_own + 1, rp.p + row_lo, h->rowptr.p, ends[0], n_own + 1);
    FS_CUDA(cudaMemcpyAsync(h->colval.p, cv.p + ends[0], (size_t)nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    FS_CUDA(cudaMemcpyAsync(h->val.p, vv.p + ends[0], (size_t)nnz * sizeof(double), cudaMemcpyDeviceToDevice, st));
    FS_CUDA(cudaMemcpyAsync(h->M.p, c->vec.p + row_lo, (size_t)n_own * sizeof(double), cudaMemcpyDeviceToDevice, st));
    // which foreign columns the own rows reference, and which ranks every own row couples to
    DBuf<signed char> d_owner;
    DBuf<unsigned char> d_used;
    DBuf<uint32_t> d_mask;
    FS_TRY(d_owner.ensure((size_t)nfl));
    FS_TRY(d_used.ensure((size_t)nfl));
    FS_TRY(d_mask.ensure((size_t)n_own));
    FS_TRY(upload(c, d_owner.p, owner.data(), (size_t)nfl));
    FS_CUDA(cudaMemsetAsync(d_used.p, 0, (size_t)nfl, st));
    XL(h, k_dist_scan, n_own * LPR, h->rowptr.p, h->colval.p, d_owner.p, n_own, d_used.p, d_mask.p);
    std::vector<unsigned char> used((size_t)nfl);
    std::vector<uint32_t> mask((size_t)n_own);
    FS_CUDA(cudaMemcpyAsync(used.data(), d_used.p, (size_t)nfl, cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaMemcpyAsync(mask.data(), d_mask.p, (size_t)n_own * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaStreamSynchronize(st));
    // halo layout: [own | pad to 16 doubles | entries of rank 0's rows, ascending | rank 1's | ...]
    D->n_own = n_own;
    D->halo0 = (n_own + 15) / 16 * 16;
    for (int64_t k = 0; k < nfl; ++k)
      if (used[k]) D->halo_cnt[(int)owner[k]]++;
    int64_t run = D->halo0;
    for (int p = 0; p < world; ++p) {
      D->halo_base[p] = run;
      run += D->halo_cnt[p];
    }
    D->n_halo = run - D->halo0;
    D->ext = (run + 31) / 32 * 32;
    {
      std::vector<int32_t> colmap((size_t)nfl, -1);
      int64_t next[kMaxWorld];
      for (int p = 0; p < world; ++p) next[p] = D->halo_base[p];
      for (int64_t k = 0; k < nfl; ++k) {
        if (owner[k] < 0)
          colmap[k] = (int32_t)(k - row_lo);
        else if (used[k])
          colmap[k] = (int32_t)next[(int)owner[k]]++;
      }
      DBuf<int32_t> d_map;
      FS_TRY(d_map.ensure((size_t)nfl));
      FS_TRY(upload(c, d_map.p, colmap.data(), (size_t)nfl * sizeof(int32_t)));
      XL(h, k_remap, nnz, h->colval.p, d_map.p, nnz);
      FS_CUDA(cudaStreamSynchronize(st));
    }
    // send lists: own rows a peer's rows reference (= the peer's halo entries of this rank, by the symmetry of the
    // pattern; the counts are cross-checked at connect), ascending
    for (int p = 0; p < world; ++p) {
      if (p == rank) continue;
      for (int64_t r = 0; r < n_own; ++r)
        if (mask[r] >> p & 1u) {
          D->push_src.push_back((int32_t)r);
          D->push_rank.push_back(p);
          D->send_cnt[p]++;
        }
      if (D->send_cnt[p] > 0 || D->halo_cnt[p] > 0) {
        FS_REQUIRE(D->npeers < kMaxPeers, FSGPU_ERR_ARG, "more than %d neighbouring ranks", kMaxPeers);
        D->peer_rank[D->npeers++] = p;
      }
    }
    D->n_push = (int)D->push_src.size();
    D->push_peer.resize(D->push_src.size());
    for (size_t i = 0; i < D->push_src.size(); ++i)
      for (int k = 0; k < D->npeers; ++k)
        if (D->peer_rank[k] == D->push_rank[i]) D->push_peer[i] = k;
    // the window
    D->win_bytes = kCtrlBytes + (size_t)kNVec * (size_t)D->ext * sizeof(double);
    FS_CUDA(cudaMalloc((void**)&D->win, D->win_bytes));
    FS_CUDA(cudaMemsetAsync(D->win, 0, D->win_bytes, st));
    D->ctrl[rank] = D->win;
    D->peer_ext[rank] = D->ext;
    FS_TRY(D->done.ensure(1));
    FS_TRY(D->err.ensure(1));
    FS_TRY(D->red.ensure(2 * kRedN));
    FS_CUDA(cudaHostAlloc((void**)&D->hpin, 32 * sizeof(double), cudaHostAllocDefault));
    FS_CUDA(cudaMemsetAsync(D->done.p, 0, sizeof(unsigned int), st));
    FS_CUDA(cudaMemsetAsync(D->err.p, 0, sizeof(int), st));
    FS_TRY(D->d_push_src.ensure((size_t)D->n_push + 1));
    FS_TRY(D->d_push_dst.ensure((size_t)D->n_push + 1));
    FS_TRY(D->d_push_peer.ensure((size_t)D->n_push + 1));
    FS_TRY(upload(c, D->d_push_src.p, D->push_src.data(), (size_t)D->n_push * sizeof(int32_t)));
    FS_TRY(upload(c, D->d_push_peer.p, D->push_peer.data(), (size_t)D->n_push * sizeof(int32_t)));
    FS_CUDA(cudaStreamSynchronize(st));
    FS_TRY(alloc_vectors(h));
    // runs that read halo entries go first (CTAs [0, nb_ctas) of the fused step)
    {
      std::vector<int32_t> runs((size_t)h->nruns + 1), order;
      FS_CUDA(cudaMemcpy(runs.data(), h->runs.p, ((size_t)h->nruns + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost));
      order.reserve((size_t)h->nruns);
      std::vector<int32_t> inner;
      inner.reserve((size_t)h->nruns);
      for (int64_t k = 0; k < h->nruns; ++k) {
        uint32_t m = 0;
        for (int32_t r = runs[k]; r < runs[k + 1]; ++r) m |= mask[r];
        (m ? order : inner).push_back((int32_t)k);
      }
      D->n_bruns = (int64_t)order.size();
      D->nb_ctas = (int)((D->n_bruns * LPR + 255) / 256);
      order.insert(order.end(), inner.begin(), inner.end());
      FS_TRY(D->order.ensure((size_t)h->nruns + 1));
      FS_TRY(upload(c, D->order.p, order.data(), (size_t)h->nruns * sizeof(int32_t)));
      FS_CUDA(cudaStreamSynchronize(st));
    }
    // CUDA loads kernels lazily, and loading may need the device idle: a first launch issued while a peer rank of
    // the same process spins on this device would never return.  Load everything the collective calls launch now.
    {
      cudaFuncAttributes fa;
      const void* fns[] = {(const void*)k_dist_barrier, (const void*)k_dist_push, (const void*)k_dist_wait,
                           (const void*)k_dist_allreduce, (const void*)k_spmv_step<true>, (const void*)k_spmv_step<false>,
                           (const void*)k_spmv, (const void*)k_update_u, (const void*)k_start, (const void*)k_div,
                           (const void*)k_scale, (const void*)k_fill_start, (const void*)k_wdot, (const void*)k_setup_damping,
                           (const void*)k_finish_step};
      for (const void* fn : fns) FS_CUDA(cudaFuncGetAttributes(&fa, fn));
    }
    FS_TRY(h->Y.ensure((size_t)n_own + 1));  // no allocation inside the collective calls either
    FS_TRY(c->flag.ensure(8));
    if (world == 1) D->connected = true;
    return FSGPU_OK;
  };
  rc = body();
  if (rc != FSGPU_OK) {
    dist_release(h);
    delete h;
    return rc;
  }
  *out = h;
  return FSGPU_OK;
}

extern "C" int fsgpu_explicit_export(fsgpu_explicit* h, void* blob) {
  FS_REQUIRE(h && blob, FSGPU_ERR_ARG, "null argument");
  FS_REQUIRE(h->dist, FSGPU_ERR_STATE, "not a row-partitioned run (fsgpu_explicit_create_dist)");
  FS_TRY(check_ctx(h->ctx));
  Dist* D = h->dist;
  FS_CUDA(cudaStreamSynchronize(h->ctx->stream));  // the window is zeroed before anybody can map it
  DistBlob b;
  memset(&b, 0, sizeof b);
  b.magic = kBlobMagic;
  b.rank = D->rank;
  b.world = D->world;
  b.pid = (int64_t)getpid();
  b.proc_token = process_token();
  b.device = h->ctx->device;
  b.devptr = (uint64_t)(uintptr_t)D->win;
  b.win_bytes = D->win_bytes;
  b.n_own = D->n_own;
  b.ext = D->ext;
  // the handle is only opened by OTHER processes; inside one process the pointer is used as it is
  cudaError_t e = cudaIpcGetMemHandle(&b.ipc, D->win);
  if (e != cudaSuccess) {
    cudaGetLastError();
    memset(&b.ipc, 0, sizeof b.ipc);
  }
  for (int p = 0; p < D->world; ++p) {
    b.halo_base[p] = D->halo_base[p];
    b.halo_cnt[p] = D->halo_cnt[p];
  }
  memset(blob, 0, FSGPU_EXPLICIT_BLOB_BYTES);
  memcpy(blob, &b, sizeof b);
  return FSGPU_OK;
}

extern "C" int fsgpu_explicit_connect(fsgpu_explicit* h, const void* blobs) {
  FS_REQUIRE(h && blobs, FSGPU_ERR_ARG, "null argument");
  FS_REQUIRE(h->dist, FSGPU_ERR_STATE, "not a row-partitioned run (fsgpu_explicit_create_dist)");
  FS_TRY(check_ctx(h->ctx));
  Dist* D = h->dist;
  FS_REQUIRE(!D->connected || D->world == 1, FSGPU_ERR_STATE, "already connected");
  std::vector<int32_t> dst((size_t)D->n_push);
  for (int p = 0; p < D->world; ++p) {
    DistBlob b;
    memcpy(&b, (const unsigned char*)blobs + (size_t)p * FSGPU_EXPLICIT_BLOB_BYTES, sizeof b);
    FS_REQUIRE(b.magic == kBlobMagic && b.rank == p && b.world == D->world, FSGPU_ERR_ARG,
               "blob %d is not the export of rank %d of %d", p, p, D->world);
    D->peer_ext[p] = b.ext;
    if (p == D->rank) continue;
    FS_REQUIRE(b.halo_cnt[D->rank] == D->send_cnt[p], FSGPU_ERR_ARG,
               "rank %d expects %lld halo entries of rank %d, which would send %lld: the matrix pattern is not symmetric "
               "or the ranks disagree about the row bounds",
               p, (long long)b.halo_cnt[D->rank], D->rank, (long long)D->send_cnt[p]);
    if (b.proc_token == process_token() && b.pid == (int64_t)getpid()) {
      D->ctrl[p] = (unsigned char*)(uintptr_t)b.devptr;
      if (b.device != h->ctx->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) FS_CUDA(e);
        cudaGetLastError();
      }
    } else {
      void* q = nullptr;
      FS_CUDA(cudaIpcOpenMemHandle(&q, b.ipc, cudaIpcMemLazyEnablePeerAccess));
      D->ctrl[p] = (unsigned char*)q;
      D->ipc_opened[p] = true;
    }
  }
  // destination of every pushed entry: the peer's halo segment of this rank's rows, same (ascending) order
  {
    int64_t pos[kMaxWorld] = {};
    for (int i = 0; i < D->n_push; ++i) {
      const int p = D->push_rank[i];
      DistBlob b;
      memcpy(&b, (const unsigned char*)blobs + (size_t)p * FSGPU_EXPLICIT_BLOB_BYTES, sizeof b);
      dst[i] = (int32_t)(b.halo_base[D->rank] + pos[p]++);
    }
    FS_TRY(upload(h->ctx, D->d_push_dst.p, dst.data(), (size_t)D->n_push * sizeof(int32_t)));
    FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  }
  D->connected = true;
  // every rank has mapped every window before anybody pushes
  FS_TRY(dist_barrier(h));
  FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return dist_check(h);
}

extern "C" int fsgpu_explicit_dist_info(fsgpu_explicit* h, int64_t* n_own, int64_t* n_halo, int64_t* n_push,
                                        int64_t* boundary_runs, int32_t* npeers) {
  FS_REQUIRE(h && h->dist, FSGPU_ERR_STATE, "not a row-partitioned run");
  if (n_own) *n_own = h->dist->n_own;
  if (n_halo) *n_halo = h->dist->n_halo;
  if (n_push) *n_push = h->dist->n_push;
  if (boundary_runs) *boundary_runs = h->dist->n_bruns;
  if (npeers) *npeers = h->dist->npeers;
  return FSGPU_OK;
}

extern "C" int fsgpu_explicit_set_state(fsgpu_explicit* h, const double* U0, const double* V0) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  const size_t b = (size_t)h->n * sizeof(double);
  h->u_ahead = false;
  if (h->dist) FS_CUDA(cudaStreamSynchronize(h->ctx->stream));  // pageable copies only on an idle stream (see Dist::hpin)
  if (U0) FS_TRY(upload(h->ctx, h->Uc, U0, b));
  if (V0) FS_TRY(upload(h->ctx, h->V.p, V0, b));
  FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_set_timestep(fsgpu_explicit* h, double c_scale, double dt) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  FS_REQUIRE(dt >= 0.0, FSGPU_ERR_ARG, "negative time step");
  h->dt = dt;
  h->c_scale = c_scale;
  h->u_ahead = false;  // the displacements written ahead were formed with the old step
  XL(h, k_setup_damping, h->n, h->M.p, h->c_scale, h->dt, h->C.p, h->invMC.p, h->n);
  FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_set_load(fsgpu_explicit* h, const double* F0) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  if (F0) {
    if (h->dist) FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
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
namespace {
void swap_u(fsgpu_explicit* h) {
  std::swap(h->Uc, h->Unx);
  if (h->dist) h->dist->cur ^= 1;
}
}  // namespace
extern "C" int fsgpu_explicit_step(fsgpu_explicit* h, int64_t nsteps, const double* fscale) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  FS_TRY(dist_require(h));
  const double dt = h->dt;
  Dist* D = h->dist;
  if (!D) {
    FS_TRY(h->Un.ensure((size_t)h->n + 1));
    if (!h->Unx) h->Unx = h->Un.p;
  }
  for (int64_t s = 0; s < nsteps; ++s) {
    if (h->u_ahead) {
      swap_u(h);
    } else {
      XL(h, k_update_u, h->n, h->Uc, h->V.p, h->A.p, dt, (dt * dt) / 2, h->n);
      if (D) FS_TRY(dist_sync_halo(h, D->cur));
    }
    const double fs_ = fscale ? fscale[s] : 1.0;
    if (D) {
      // waits for the halo entries announced by step flag `epoch`, pushes the next displacements with flag epoch + 1
      const DistDev dd = dist_dev(h, D->cur ^ 1);
      k_spmv_step<true><<<grid_for(h->nruns * LPR, 256), 256, 0, h->ctx->stream>>>(
          h->runs.p, h->nruns, h->rowptr.p, h->colval.p, h->val.p, h->Uc, h->have_load ? h->F0.p : nullptr, fs_, h->C.p,
          h->invMC.p, h->V.p, h->A.p, h->E.p, dt / 2, dt, (dt * dt) / 2, h->Unx, dd);
      h->ctx->launches++;
      if (D->world > 1) D->epoch++;
    } else {
      DistDev dd;
      memset(&dd, 0, sizeof dd);
      XL(h, k_spmv_step<false>, h->nruns * LPR, h->runs.p, h->nruns, h->rowptr.p, h->colval.p, h->val.p, h->Uc,
         h->have_load ? h->F0.p : nullptr, fs_, h->C.p, h->invMC.p, h->V.p, h->A.p, h->E.p, dt / 2, dt, (dt * dt) / 2, h->Unx,
         dd);
    }
    h->u_ahead = true;
  }
  FS_CUDA(cudaGetLastError());
  FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  if (D) FS_TRY(dist_check(h));
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_step_begin(fsgpu_explicit* h) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  FS_REQUIRE(!h->dist, FSGPU_ERR_STATE, "row-partitioned runs exchange inside fsgpu_explicit_step");
  const double dt = h->dt;
  if (h->u_ahead) {
    swap_u(h);
    h->u_ahead = false;
  } else {
    XL(h, k_update_u, h->n, h->Uc, h->V.p, h->A.p, dt, (dt * dt) / 2, h->n);
  }
  XL(h, k_spmv, h->nruns * LPR, h->runs.p, h->nruns, h->rowptr.p, h->colval.p, h->val.p, h->Uc, h->E.p);
  FS_CUDA(cudaGetLastError());
  return FSGPU_OK;  // asynchronous: the host's exchange is enqueued on the same stream
}
extern "C" int fsgpu_explicit_step_end(fsgpu_explicit* h, double fscale) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  FS_REQUIRE(!h->dist, FSGPU_ERR_STATE, "row-partitioned runs exchange inside fsgpu_explicit_step");
  XL(h, k_finish_step, h->n, h->E.p, h->have_load ? h->F0.p : nullptr, fscale, h->C.p, h->invMC.p, h->V.p, h->A.p,
     h->dt / 2, h->n);
  FS_CUDA(cudaGetLastError());
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_get_state(fsgpu_explicit* h, double* U, double* V, double* A) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  FS_TRY(check_ctx(h->ctx));
  const size_t b = (size_t)h->n * sizeof(double);
  if (h->dist) FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  if (U) FS_TRY(download(h->ctx, U, h->Uc, b));
  if (V) FS_TRY(download(h->ctx, V, h->V.p, b));
  if (A) FS_TRY(download(h->ctx, A, h->A.p, b));
  FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_device_state(fsgpu_explicit* h, double** U, double** V, double** A, double** E) {
  FS_REQUIRE(h, FSGPU_ERR_ARG, "null handle");
  // U is the CURRENT displacement buffer: fsgpu_explicit_step / _step_begin alternate between two buffers, so the
  // pointer must be asked for again after every call that advances the state
  if (U) *U = h->Uc;
  if (V) *V = h->V.p;
  if (A) *A = h->A.p;
  if (E) *E = h->E.p;
  return FSGPU_OK;
}
namespace {
// y = K x with x in the device vector xv (own entries); row-partitioned: xv is window vector 2, halo entries fetched first
int spmv_any(fsgpu_explicit* h, double* xv, double* y) {
  if (h->dist) FS_TRY(dist_sync_halo(h, 2));
  XL(h, k_spmv, h->nruns * LPR, h->runs.p, h->nruns, h->rowptr.p, h->colval.p, h->val.p, xv, y);
  return FSGPU_OK;
}
}  // namespace
extern "C" int fsgpu_explicit_spmv(fsgpu_explicit* h, const double* x, double* y) {
  FS_REQUIRE(h && x && y, FSGPU_ERR_ARG, "null argument");
  FS_TRY(check_ctx(h->ctx));
  FS_TRY(dist_require(h));
  double* xv;
  if (h->dist) {
    xv = h->dist->vec(h->dist->rank, 2);
  } else {
    FS_TRY(h->X.ensure((size_t)h->n + 1));
    xv = h->X.p;
  }
  FS_TRY(h->Y.ensure((size_t)h->n + 1));
  if (h->dist) FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  FS_TRY(upload(h->ctx, xv, x, (size_t)h->n * sizeof(double)));
  if (h->dist) FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  FS_TRY(spmv_any(h, xv, h->Y.p));
  if (h->dist) FS_TRY(dist_check(h));  // the kernels that wait for peers are done before the pageable copy is queued
  FS_TRY(download(h->ctx, y, h->Y.p, (size_t)h->n * sizeof(double)));
  FS_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_omega_max(fsgpu_explicit* h, int32_t maxit, double* lambda_max) {
  FS_REQUIRE(h && lambda_max, FSGPU_ERR_ARG, "null argument");
  FS_TRY(check_ctx(h->ctx));
  FS_TRY(dist_require(h));
  double* xv;
  if (h->dist) {
    xv = h->dist->vec(h->dist->rank, 2);
  } else {
    FS_TRY(h->X.ensure((size_t)h->n + 1));
    xv = h->X.p;
  }
  FS_TRY(h->Y.ensure((size_t)h->n + 1));
  // start vector: a function of the GLOBAL row, so a partitioned run iterates on the same vector
  XL(h, k_fill_start, h->n, xv, h->n, h->dist ? h->row0 : 0);
  double lam = 0.0;
  for (int it = 0; it < maxit; ++it) {
    // y = M^-1 K x ; lambda = (x' M y) / (x' M x) ; x = y / |y|
    FS_TRY(spmv_any(h, xv, h->Y.p));
    XL(h, k_div, h->n, h->Y.p, h->M.p, h->Y.p, h->n);
    double xmy, xmx, yy;
    FS_TRY(wdot(h, xv, h->Y.p, h->M.p, &xmy));
    FS_TRY(wdot(h, xv, xv, h->M.p, &xmx));
    FS_TRY(wdot(h, h->Y.p, h->Y.p, nullptr, &yy));
    lam = xmy / xmx;
    XL(h, k_scale, h->n, xv, h->Y.p, 1.0 / sqrt(yy), h->n);
  }
  *lambda_max = lam;
  return FSGPU_OK;
}
extern "C" int fsgpu_explicit_kinetic_energy(fsgpu_explicit* h, double* ke) {
  FS_REQUIRE(h && ke, FSGPU_ERR_ARG, "null argument");
  FS_TRY(check_ctx(h->ctx));
  FS_TRY(dist_require(h));
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
