// Sync-free subspace eigensolver: the k leading eigenpairs of a PSD Gram matrix (n > 256) as ONE stream-ordered chain of
// kernels with no host round trip.  Same mathematics as eig.cuh (Chebyshev-filtered subspace iteration, Cholesky-QR
// orthonormalisation in fp64, Rayleigh-Ritz with an fp32-exact product), different control plane:
//
//   * everything the host used to decide between kernels — filter bounds and degree, convergence, Cholesky breakdown —
//     is decided on the device and kept in a control block (ChfsiCtrl, cheb_filter.cuh); the host enqueues a fixed
//     number of stages and every kernel of a stage beyond convergence returns at once (common.cuh::tnb_skip);
//   * per stage: resident filter (cheb_filter_kernel, 1 launch) -> Gram of the filtered block (1) -> reduce +
//     Cholesky + triangular inverse (1 CTA) -> X = Y T (1) -> W = G X exact (2) -> X^T W (1) -> reduce + two-sided
//     Jacobi + Ritz values + stopping rule + next filter (1 CTA) -> X <- X Q (1): 9 launches, was ~60 + 2 host syncs;
//   * stopping rule on the captured energy cap = sum_{i<k} theta_i (the quantity the relative error depends on):
//     with delta_s = cap_s - cap_{s-1} and rho = delta_s / delta_{s-1}, the energy still missing is estimated by the
//     geometric tail delta_s rho / (1 - rho) (conservative: the filters sharpen from stage to stage) and compared with
//     the change of the relative error it could cause: stop when it is below 5 tol sqrt(tail / trace) trace (tol = 1e-6:
//     d(relerr) <= 2.5e-6, a quarter of the 1e-5 parity bar), or below the resolution of the block precision.
//     On the config-2 spectrum (2048^2, Marchenko-Pastur flat) this stops after 2 filters (48 products, energy deficit
//     5e-7 of the trace) where the growth-below-tol rule of eig.cuh needed 3-4 (80-110 products, deficit 1e-10).
//
// The caller learns the outcome from the control block at its own final synchronisation (sweep.cuh): `done` must be set
// and `error` clear, otherwise it repeats the step on the host-driven solver of eig.cuh.
#pragma once
#include "cheb_filter.cuh"
#include "common.cuh"
#include "eig.cuh"
#include "gemm_generic.cuh"
#include "jacobi2.cuh"
#include "small_kernels.cuh"

namespace tnb {

constexpr int CD_MAX_STAGES = 6;  // filters enqueued per solve; the chain stops itself at convergence
constexpr int CD_MAX_B = 96;      // block width limit (shared memory of the one-CTA kernels)
constexpr int CD_GRAM_ROWS = 64;  // rows of the block per CTA of the Gram kernel (fp64 FMA rate is ~16 / clk / SM: spread it)

template <typename TB>
struct CdWork {
  TB* ring[3];   // filter ring (n x b each)
  TB* Xo;        // orthonormalised block
  TB* Wb;        // G * Xo
  TB* T1;        // b x b Cholesky-QR transform
  TB* T2;        // b x b Ritz rotation
  double* gpart; // [P][b][b] partial Gram matrices
  double* Sg;    // b x b
  double* Qd;    // b x b
  double* lam;   // b Ritz values, descending
  double* cscr;  // Cholesky scratch when L, L^-1 do not fit shared memory (unused for b <= 96)
  ChfsiCtrl* ctrl;
  void* partial;       // split-K scratch of the exact product
  size_t partial_bytes;
  void* fws;           // resident filter workspace
  size_t fws_bytes;
  int P;
};

// Block width of the sync-free solver: k wanted pairs + max(16, k/2) guard vectors, a multiple of 4.  Narrower than
// eig.cuh's 2k: the one-CTA steps (Cholesky, Jacobi) cost ~b^3 and on the config-2 spectrum a 48-wide block needs 52
// products where a 64-wide one needs 48 (tests/sweep_model.py replayed with the new stopping rule).
inline int chfsi_dev_block(int n, int k) {
  int g = k / 2 > 16 ? k / 2 : 16;
  int b = (k + g + 3) / 4 * 4;
  if (b > n) b = n / 4 * 4;
  return b;
}

inline bool chfsi_dev_ok(int n, int b) {
  return b >= 8 && b <= CD_MAX_B && b % 4 == 0 && tc_path_available() && cheb_filter_shape_ok(n, b) &&
         jacobi2_ok(b, true);
}

template <typename TB, class ArenaT>
inline void chfsi_dev_carve(ArenaT& ar, int n, int b, CdWork<TB>& w) {
  const size_t nb = (size_t)n * b;
  for (int i = 0; i < 3; ++i) w.ring[i] = ar.template take<TB>(nb);
  w.Xo = ar.template take<TB>(nb);
  w.Wb = ar.template take<TB>(nb);
  w.T1 = ar.template take<TB>((size_t)b * b);
  w.T2 = ar.template take<TB>((size_t)b * b);
  w.P = (n + CD_GRAM_ROWS - 1) / CD_GRAM_ROWS;
  w.gpart = ar.template take<double>((size_t)w.P * b * b);
  w.Sg = ar.template take<double>((size_t)b * b);
  w.Qd = ar.template take<double>((size_t)b * b);
  w.lam = ar.template take<double>(b);
  w.cscr = ar.template take<double>((size_t)2 * b * (b | 1));
  w.ctrl = ar.template take<ChfsiCtrl>(1);
  GemmPlan pl = plan_gemm(n, b, n, false);
  w.partial_bytes = pl.partial_elems * sizeof(TB);
  w.partial = ar.template take<char>(w.partial_bytes);
  w.fws_bytes = cheb_filter_workspace_bytes(n, b);
  w.fws = ar.template take<char>(w.fws_bytes);
}

// ---------------------------------------------------------------------------------------------------------------
// partial[p] = A_p^T B_p over the row slab p (fp64 accumulation of exact products), 4x4 register tiles
// ---------------------------------------------------------------------------------------------------------------
template <typename TB>
__global__ void __launch_bounds__(256) cd_gram_partial_kernel(const TB* __restrict__ A, const TB* __restrict__ B, int n,
                                                              int b, double* __restrict__ partial, const int* skip,
                                                              int stage) {
  if (tnb_skip(skip, stage)) return;
  __shared__ TB As[16][CD_MAX_B + 4];
  __shared__ TB Bs[16][CD_MAX_B + 4];
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * CD_GRAM_ROWS;
  const int r1 = min(n, r0 + CD_GRAM_ROWS);
  const int nt4 = (b + 3) / 4, ntiles = nt4 * nt4;
  double acc[3][4][4];
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.0;
  for (int rc = r0; rc < r1; rc += 16) {
    for (int idx = tid; idx < 16 * b; idx += 256) {
      const int rr = idx / b, c = idx - rr * b;
      const bool ok = rc + rr < r1;
      As[rr][c] = ok ? A[(size_t)(rc + rr) * b + c] : (TB)0;
      Bs[rr][c] = ok ? B[(size_t)(rc + rr) * b + c] : (TB)0;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const int tile = tid + t * 256;
      if (tile < ntiles) {
        const int ti = tile / nt4, tj = tile - ti * nt4;
#pragma unroll 4
        for (int kk = 0; kk < 16; ++kk) {
          double a[4], bb[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = (double)As[kk][ti * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) bb[j] = (double)Bs[kk][tj * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[t][i][j] = fma(a[i], bb[j], acc[t][i][j]);
        }
      }
    }
    __syncthreads();
  }
  double* out = partial + (size_t)blockIdx.x * b * b;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int tile = tid + t * 256;
    if (tile < ntiles) {
      const int ti = tile / nt4, tj = tile - ti * nt4;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ti * 4 + i < b && tj * 4 + j < b) out[(size_t)(ti * 4 + i) * b + tj * 4 + j] = acc[t][i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// S = sum_p partial[p];  D S D = L L^T;  T1 = D L^-T  (one CTA; a dependent block raises ctrl->error = 1)
// ---------------------------------------------------------------------------------------------------------------
template <typename TB>
__global__ void __launch_bounds__(1024) cd_chol_kernel(const double* __restrict__ partial, int P, int b, double* Sg,
                                                       double* scratch, TB* __restrict__ T1, ChfsiCtrl* ctrl, int stage,
                                                       int use_smem) {
  extern __shared__ __align__(16) unsigned char cd_chol_smem[];
  if (tnb_skip(&ctrl->done, stage)) return;
  __shared__ int s_flag;
  const int tid = threadIdx.x;
  if (tid == 0) s_flag = 0;
  for (int idx = tid; idx < b * b; idx += blockDim.x) {
    double s = 0.0;
    for (int p = 0; p < P; ++p) s += partial[(size_t)p * b * b + idx];
    Sg[idx] = s;
  }
  __syncthreads();
  chol_orth_device<TB>(Sg, b, scratch, T1, &s_flag, use_smem, nullptr, cd_chol_smem);
  __syncthreads();
  if (tid == 0 && s_flag) ctrl->error = 1;
}

// ---------------------------------------------------------------------------------------------------------------
// X = Y * T  (n x b times b x b), one CTA per 32 rows, T and the row chunk in shared memory
// ---------------------------------------------------------------------------------------------------------------
template <typename TB>
struct CdRing {
  TB* p[3];
};

template <typename TB>
__global__ void __launch_bounds__(256) cd_rotate_kernel(const TB* __restrict__ Y, const TB* __restrict__ T, TB* Xfixed,
                                                        const CdRing<TB> ring, int n, int b, const ChfsiCtrl* ctrl,
                                                        int stage, int out_ring) {
  extern __shared__ __align__(16) unsigned char cd_rot_smem[];
  if (tnb_skip(&ctrl->done, stage)) return;
  TB* Ts = reinterpret_cast<TB*>(cd_rot_smem);       // b x b
  TB* Ys = Ts + (size_t)b * b;                       // 32 x (b + 1)
  TB* X = Xfixed;
  if (out_ring) {                                    // the ring position the NEXT filter will read from
    const int xi = __ldcg(&ctrl->xin);
    X = xi == 0 ? ring.p[0] : (xi == 1 ? ring.p[1] : ring.p[2]);
  }
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * 32;
  for (int idx = tid; idx < b * b; idx += 256) Ts[idx] = T[idx];
  for (int idx = tid; idx < 32 * b; idx += 256) {
    const int rr = idx / b, c = idx - rr * b;
    Ys[rr * (b + 1) + c] = (r0 + rr < n) ? Y[(size_t)(r0 + rr) * b + c] : (TB)0;
  }
  __syncthreads();
  for (int idx = tid; idx < 32 * b; idx += 256) {
    const int rr = idx / b, j = idx - rr * b;
    if (r0 + rr >= n) continue;
    TB acc = (TB)0;
    const TB* yr = Ys + rr * (b + 1);
#pragma unroll 8
    for (int c = 0; c < b; ++c) acc = fma(yr[c], Ts[c * b + j], acc);
    X[(size_t)(r0 + rr) * b + j] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Rayleigh-Ritz on the orthonormal block: S = sum_p partial[p] (= Xo^T G Xo), two-sided Jacobi in the block precision,
// Ritz values as fp64 Rayleigh quotients against S, descending sort, T2 = Q[:, order]; then the control decisions.
// ---------------------------------------------------------------------------------------------------------------
struct CdRule {
  int k;            // wanted eigenpairs
  int mmax;         // largest filter degree
  double spread;    // largest amplification ratio of a filter (dynamic range the block precision tolerates)
  double tol;       // relative-error resolution asked for (1e-6)
  double floor_tol; // resolution of Ritz-value sums in the block precision, relative to the trace
  double jac_tol;   // Jacobi threshold
  int last_stage;   // no filter is enqueued after this stage: not converged there = error 2
};

template <typename TB>
__global__ void __launch_bounds__(1024) cd_rr_kernel(const double* __restrict__ partial, int P, int b, double* Sg, double* Qd,
                                                     TB* __restrict__ T2, double* __restrict__ lam,
                                                     const double* __restrict__ d_trace, ChfsiCtrl* ctrl, int stage,
                                                     CdRule rule) {
  extern __shared__ __align__(16) unsigned char cd_rr_smem[];
  if (tnb_skip(&ctrl->done, stage)) return;
  __shared__ double s_w[CD_MAX_B + 2];
  __shared__ int s_rank[CD_MAX_B + 2];
  __shared__ double s_gmax;
  __shared__ int s_sweeps;
  typedef TB R;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  Jac2<R> J;
  jac2_carve<R>(cd_rr_smem, b, (R)rule.jac_tol, J);
  const int lds = J.lds;  // b is even (multiple of 4): np == b
  if (tid == 0) s_gmax = 0.0;
  for (int idx = tid; idx < b * b; idx += nt) {
    double s = 0.0;
    for (int p = 0; p < P; ++p) s += partial[(size_t)p * b * b + idx];
    Sg[idx] = s;
  }
  __syncthreads();
  double dmax = 0.0;
  for (int i = tid; i < b; i += nt) dmax = fmax(dmax, fabs(Sg[(size_t)i * b + i]));
  for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(&s_gmax), (unsigned long long)__double_as_longlong(dmax));
  __syncthreads();
  const double gscale = s_gmax > 0.0 ? s_gmax : 1.0;
  const double ginv = 1.0 / gscale;
  for (int idx = tid; idx < b * b; idx += nt) {
    const int r = idx / b, c = idx - r * b;
    const double v = 0.5 * (Sg[idx] + Sg[(size_t)c * b + r]);
    Qd[idx] = v;  // symmetrised copy (Qd is free until the solve is over)
    if (c >= r) J.S[0][r * lds + c] = (R)(v * ginv);  // canonical upper storage
    J.V[0][r * lds + c] = (r == c) ? (R)1 : (R)0;
  }
  __syncthreads();
  for (int idx = tid; idx < b * b; idx += nt) Sg[idx] = Qd[idx];  // Sg := its symmetric part, read coalesced below
  __syncthreads();
  const int cur = jac2_solve(J, 30, &s_sweeps);
  const R* V = J.V[cur];
  // Q in fp64; for an fp32 solve one Newton-Schulz step Q <- Q (1.5 I - 0.5 Q^T Q) restores orthogonality from ~1e-5
  // (rounding of ~500 fp32 rotations per column) to ~1e-10, so the rotated block stays orthonormal to fp32 level
  for (int idx = tid; idx < b * b; idx += nt) Qd[idx] = (double)V[(idx / b) * lds + (idx % b)];
  __syncthreads();
  if (sizeof(R) == 4) {
    // in fp32 (the B200 issues ~16 fp64 FMAs per clock and SM, 128 fp32 ones): E = 1.5 I - 0.5 Q^T Q with Q in shared
    // memory (the other V buffer is dead), the residual Q^T Q - I is ~1e-5, so fp32 leaves ~1e-7
    const R* Vs = V;                                   // b x lds, shared
    R* E = J.S[0];                                     // b x lds, shared (S is dead)
    for (int idx = tid; idx < b * b; idx += nt) {
      const int i = idx / b, j = idx - i * b;
      R s = (R)0;
      for (int c = 0; c < b; ++c) s = fma(Vs[c * lds + i], Vs[c * lds + j], s);
      E[i * lds + j] = (i == j ? (R)1.5 : (R)0) - (R)0.5 * s;
    }
    __syncthreads();
    for (int idx = tid; idx < b * b; idx += nt) {
      const int i = idx / b, j = idx - i * b;
      R s = (R)0;
      for (int c = 0; c < b; ++c) s = fma(Vs[i * lds + c], E[c * lds + j], s);
      Qd[idx] = (double)s;
    }
    __syncthreads();
  }
  // Ritz values theta_j = q_j^T S q_j in fp64 (one warp per column)
  for (int j = warp; j < b; j += nwarps) {
    double acc = 0.0;
    for (int r = lane; r < b; r += 32) {
      double t = 0.0;
      for (int c = 0; c < b; ++c) t = fma(Sg[(size_t)c * b + r], Qd[(size_t)c * b + j], t);
      acc = fma(Qd[(size_t)r * b + j], t, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_w[j] = acc;
  }
  __syncthreads();
  for (int i = tid; i < b; i += nt) {
    const double wi = s_w[i];
    int r = 0;
    for (int j = 0; j < b; ++j) {
      const double wj = s_w[j];
      r += (wj > wi) || (wj == wi && j < i);
    }
    s_rank[i] = r;
    lam[r] = wi;
  }
  __syncthreads();
  for (int idx = tid; idx < b * b; idx += nt) {
    const int i = idx / b, j = idx - i * b;
    T2[(size_t)i * b + s_rank[j]] = (TB)Qd[idx];
  }
  __syncthreads();  // lam[] complete (global, same CTA)
  if (tid != 0) return;
  // ---- control: convergence of the captured energy, then bounds / degree / coefficients of the next filter ----
  ctrl->outer += 1;
  ctrl->jac_sweeps += s_sweeps < 0 ? -s_sweeps : s_sweeps;
  double cap = 0.0;
  bool finite = true;
  for (int i = 0; i < b; ++i) {
    const double v = lam[i];
    if (!(v == v) || fabs(v) > 1e300) finite = false;
    if (i < rule.k) cap += v;
  }
  if (!finite) { ctrl->error = 3; return; }
  double trace = d_trace ? *d_trace : 0.0;
  if (!(trace > 0.0)) {
    trace = 0.0;
    for (int i = 0; i < b; ++i) trace += lam[i] > 0.0 ? lam[i] : 0.0;
  }
  ctrl->trace = trace;
  ctrl->cap = cap;
  const double top = lam[0];
  bool conv = false;
  if (!(top > 0.0)) conv = true;  // numerically zero matrix: nothing to iterate on
  if (stage >= 1 && !conv) {
    const double tail = trace - cap > 0.0 ? trace - cap : 0.0;
    double need = 5.0 * rule.tol * sqrt(tail * trace);      // = 5 tol * relerr_est * trace
    const double fl = rule.floor_tol * trace;
    if (need < fl) need = fl;
    const double delta = cap - ctrl->prev_cap;
    if (delta <= need) conv = true;
    else if (ctrl->prev_delta > 0.0) {
      double rho = delta / ctrl->prev_delta;
      if (rho > 0.5) rho = 0.5;
      if (rho < 0.0) rho = 0.0;
      if (delta * rho / (1.0 - rho) <= need) conv = true;
    }
    ctrl->prev_delta = delta;
  } else if (stage == 0) {
    ctrl->prev_delta = -1.0;
  }
  ctrl->prev_cap = cap;
  if (conv) {
    ctrl->conv_stage = stage;
    __threadfence();
    ctrl->done = 1;
    ctrl->xin = 0;  // the final rotation of this stage writes ring[0]
    return;
  }
  if (stage >= rule.last_stage) { ctrl->xin = 0; ctrl->conv_stage = stage; ctrl->error = 2; return; }
  // scaled Chebyshev filter damping [0, hi], hi = smallest Ritz value of the block (eig.cuh)
  const double cut = lam[b - 1] > 0.0 ? lam[b - 1] : 0.0;
  double hi = cut;
  const double tiny = 1e-30 * top + 1e-300;
  if (hi < tiny) hi = tiny;
  const double e = 0.5 * hi, c = 0.5 * hi;
  double x1 = (top - c) / e;
  if (x1 < 1.0) x1 = 1.0;
  double ac = acosh(x1);
  if (ac < 1e-12) ac = 1e-12;
  int m = (int)floor(log(2.0 * rule.spread) / ac);
  if (m < 1) m = 1;
  if (m > rule.mmax) m = rule.mmax;
  const double sigma1 = e / (top - c);
  double sg = sigma1;
  ctrl->a[0] = (float)(sigma1 / e);
  ctrl->bc[0] = (float)(-c * sigma1 / e);
  ctrl->g[0] = 0.f;
  for (int i = 2; i <= m; ++i) {
    const double sigma2 = 1.0 / (2.0 / sigma1 - sg);
    ctrl->a[i - 1] = (float)(2.0 * sigma2 / e);
    ctrl->bc[i - 1] = (float)(-2.0 * sigma2 * c / e);
    ctrl->g[i - 1] = (float)(-sg * sigma2);
    sg = sigma2;
  }
  ctrl->steps = m;
  ctrl->products += m + 1;
  ctrl->xin = (3 - m % 3) % 3;  // the filter's result then lands in ring[0]
}

// theta_out / X_out in fp64 for the factor extraction; a chain that ran out of stages reports error 2
template <typename TB>
__global__ void cd_finish_kernel(const TB* __restrict__ X /* ring[0]: where converged chains leave their block */,
                                 const double* __restrict__ lam, int n, int b, double* theta_out, double* X_out,
                                 ChfsiCtrl* ctrl, int* sweep_flags) {
  const int64_t total = (int64_t)n * b;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    X_out[i] = (double)X[i];
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < b; i += blockDim.x) theta_out[i] = lam[i];
    if (threadIdx.x == 0) {
      if (!ctrl->done && !ctrl->error) ctrl->error = 2;
      if (sweep_flags) {
        if (ctrl->error) atomicOr(sweep_flags, 1 << (ctrl->error > 3 ? 3 : ctrl->error));
        atomicAdd(sweep_flags + 1, ctrl->products);   // diagnostics of the whole sweep (info_host[2], [30], [29])
        atomicAdd(sweep_flags + 2, ctrl->outer);
        atomicAdd(sweep_flags + 3, ctrl->jac_sweeps);
      }
    }
  }
}

__global__ void cd_init_ctrl_kernel(ChfsiCtrl* ctrl) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    ctrl->done = 0; ctrl->error = 0; ctrl->conv_stage = -1; ctrl->outer = 0; ctrl->products = 1; ctrl->steps = 0;
    ctrl->xin = 0; ctrl->jac_sweeps = 0; ctrl->prev_cap = 0.0; ctrl->prev_delta = -1.0; ctrl->trace = 0.0; ctrl->cap = 0.0;
  }
}

// One solve, enqueued in pieces so that a batch can interleave the stages of several solves (sweep.cuh: the resident
// filter kernels of all streams run one after the other, so tensor A's Rayleigh-Ritz step should be in flight while
// tensor B's filter runs): cd_begin, cd_stage(0 .. CD_MAX_STAGES), cd_end.  G: n x n fp32 (TF32 filter products,
// fp32-exact Rayleigh-Ritz product), d_trace: device scalar trace(G) (may be null), theta_out: b doubles, X_out: n x b
// doubles.  sweep_flags (device int[4], may be null): bit (1 << error) is OR-ed into [0] when the chain fails.
struct CdRun {
  const float* G = nullptr;
  int n = 0, k = 0, b = 0;
  const double* d_trace = nullptr;
  CdWork<float>* w = nullptr;
  double *theta_out = nullptr, *X_out = nullptr;
  int* sweep_flags = nullptr;
  cudaStream_t st = 0;
  CdRule rule;
  CdRing<float> ring;
  GemmPlan pl;
  size_t chol_smem = 0, rot_smem = 0, rr_smem = 0;
};

inline int cd_begin(CdRun& r, const float* G, int n, int k, int b, const double* d_trace, double tol, CdWork<float>& w,
                    double* theta_out, double* X_out, int* sweep_flags, cudaStream_t st) {
  typedef float TB;
  if (!chfsi_dev_ok(n, b) || b < k) return fail(TNB_ERR_UNSUPPORTED, "chfsi_dev: n=%d b=%d k=%d outside the envelope", n, b, k);
  r.G = G; r.n = n; r.k = k; r.b = b; r.d_trace = d_trace; r.w = &w; r.theta_out = theta_out; r.X_out = X_out;
  r.sweep_flags = sweep_flags; r.st = st;
  for (int i = 0; i < 3; ++i) r.ring.p[i] = w.ring[i];
  r.rule.k = k;
  r.rule.mmax = 40;
  r.rule.spread = 1e4;
  r.rule.tol = tol;
  r.rule.floor_tol = 1e-7;
  r.rule.jac_tol = 1e-4;  // the Ritz basis only needs ~1e-4: the captured energy is second order in it, Q stays orthogonal
  r.rule.last_stage = CD_MAX_STAGES;
  r.chol_smem = (size_t)2 * b * (b | 1) * sizeof(double);
  r.rot_smem = ((size_t)b * b + (size_t)32 * (b + 1)) * sizeof(TB);
  r.rr_smem = jac2_smem_bytes<TB>(b) > (size_t)b * b * sizeof(double) ? jac2_smem_bytes<TB>(b) : (size_t)b * b * sizeof(double);
  static PerDeviceFlag attr_done[3];
  TNB_CUDA(ensure_dyn_smem(attr_done[0], cd_chol_kernel<TB>, 180 * 1024));
  TNB_CUDA(ensure_dyn_smem(attr_done[1], cd_rotate_kernel<TB>, 100 * 1024));
  TNB_CUDA(ensure_dyn_smem(attr_done[2], cd_rr_kernel<TB>, (int)jac2_smem_bytes<TB>(JAC2_MAX_N_F32)));
  r.pl = plan_gemm(n, b, n, false);
  cd_init_ctrl_kernel<<<1, 32, 0, st>>>(w.ctrl);
  TNB_LAUNCH_CHECK();
  random_fill_kernel<TB><<<grid_for((int64_t)n * b), 256, 0, st>>>(w.ring[0], (int64_t)n * b, 0x1234567u);
  TNB_LAUNCH_CHECK();
  return TNB_OK;
}

inline int cd_stage(CdRun& r, int stage) {
  typedef float TB;
  CdWork<float>& w = *r.w;
  const int n = r.n, b = r.b;
  cudaStream_t st = r.st;
  const int* skip = &w.ctrl->done;
  if (stage >= 1) {
    float* fb[3] = {w.ring[0], w.ring[1], w.ring[2]};
    TNB_TRY(cheb_filter_f32(r.G, n, b, fb, 1, nullptr, nullptr, nullptr, w.fws, w.fws_bytes, st, w.ctrl, stage));
  }
  // the block to orthonormalise is ring[0] (random start at stage 0, the filter's result afterwards)
  cd_gram_partial_kernel<TB><<<w.P, 256, 0, st>>>(w.ring[0], w.ring[0], n, b, w.gpart, skip, stage);
  TNB_LAUNCH_CHECK();
  cd_chol_kernel<TB><<<1, 1024, r.chol_smem, st>>>(w.gpart, w.P, b, w.Sg, w.cscr, w.T1, w.ctrl, stage, 1);
  TNB_LAUNCH_CHECK();
  cd_rotate_kernel<TB><<<(n + 31) / 32, 256, r.rot_smem, st>>>(w.ring[0], w.T1, w.Xo, r.ring, n, b, w.ctrl, stage, 0);
  TNB_LAUNCH_CHECK();
  TNB_TRY((gemm_splitk<TB, TB, TB, TB, TB>(r.pl, n, b, n, r.G, n, false, w.Xo, b, false, reinterpret_cast<TB*>(w.partial), w.Wb, b,
                                           (TB)1, nullptr, 0, (TB)0, nullptr, 0, (TB)0, false, (TB*)nullptr, 0, st, skip, stage)));
  cd_gram_partial_kernel<TB><<<w.P, 256, 0, st>>>(w.Xo, w.Wb, n, b, w.gpart, skip, stage);
  TNB_LAUNCH_CHECK();
  cd_rr_kernel<TB><<<1, 1024, r.rr_smem, st>>>(w.gpart, w.P, b, w.Sg, w.Qd, w.T2, w.lam, r.d_trace, w.ctrl, stage, r.rule);
  TNB_LAUNCH_CHECK();
  cd_rotate_kernel<TB><<<(n + 31) / 32, 256, r.rot_smem, st>>>(w.Xo, w.T2, nullptr, r.ring, n, b, w.ctrl, stage, 1);
  TNB_LAUNCH_CHECK();
  return TNB_OK;
}

inline int cd_end(CdRun& r) {
  CdWork<float>& w = *r.w;
  cd_finish_kernel<float><<<grid_for((int64_t)r.n * r.b), 256, 0, r.st>>>(w.ring[0], w.lam, r.n, r.b, r.theta_out, r.X_out, w.ctrl,
                                                                           r.sweep_flags);
  TNB_LAUNCH_CHECK();
  return TNB_OK;
}

// The whole solve on one stream; returns without waiting.
inline int eig_topk_chfsi_dev(const float* G, int n, int k, int b, const double* d_trace, double tol, CdWork<float>& w,
                              double* theta_out, double* X_out, int* sweep_flags, cudaStream_t st) {
  CdRun r;
  TNB_TRY(cd_begin(r, G, n, k, b, d_trace, tol, w, theta_out, X_out, sweep_flags, st));
  for (int stage = 0; stage <= CD_MAX_STAGES; ++stage) TNB_TRY(cd_stage(r, stage));
  return cd_end(r);
}

}  // namespace tnb
