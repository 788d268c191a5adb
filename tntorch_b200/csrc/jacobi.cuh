// Small dense symmetric eigenproblems on one CTA: parallel-order one-sided (Hestenes) Jacobi with
// warp-shuffle reductions, fp64 throughout.  Replaces torch.linalg.eigh (round.py:114) and the
// U,S part of torch.linalg.svd (round.py:96) for Gram matrices up to JACOBI_MAX_N.
#pragma once
#include "common.cuh"

namespace tnb {

constexpr int JACOBI_MAX_N = 256;      // one-CTA solver limit (matrix in L1/L2-backed global memory above ~100)
constexpr int JACOBI_SMEM_MAX_N = 104; // A and V both in shared memory up to this size
constexpr int JACOBI_THREADS = 1024;

__device__ __forceinline__ double block_reduce_sum(double v, double* red /*>=32 doubles smem*/) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (w == 0) {
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  const double out = red[0];
  __syncthreads();
  return out;
}

// Eigen-decomposition of the symmetric PSD n x n matrix Gin (leading dimension ldg) by ONE-SIDED
// (Hestenes) Jacobi: W = G V is kept column by column; each warp owns one column pair per round,
// takes the three inner products with shuffle reductions, rotates its two columns of W and V, and the
// only block-wide synchronisation is one barrier per round (np-1 rounds per sweep, circle ordering).
// At convergence the columns of W are orthogonal: W = V diag(lambda), lambda_c = v_c . w_c.
//   w_out[0..n)   eigenvalues, descending
//   V_out[n x n]  row-major, column j = eigenvector of w_out[j]
//   scratch       2*np*np doubles when !SMEM (np = n rounded up to even)
//   info[0]       number of sweeps used (negative if max_sweeps hit without convergence)
template <bool SMEM>
__global__ void __launch_bounds__(JACOBI_THREADS) jacobi_eigh_kernel(const double* __restrict__ Gin, int n, int ldg,
                                                                     double* __restrict__ w_out,
                                                                     double* __restrict__ V_out,
                                                                     double* __restrict__ scratch, int max_sweeps,
                                                                     double tol, int* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char jac_smem_raw[];
  const int np = n + (n & 1);
  const int m = np >> 1;
  double* Wt;  // column-major: Wt[c*np + r] = W[r][c]
  double* Vt;
  if (SMEM) {
    Wt = reinterpret_cast<double*>(jac_smem_raw);
    Vt = Wt + (size_t)np * np;
  } else {
    Wt = scratch;
    Vt = scratch + (size_t)np * np;
  }
  __shared__ double s_w[JACOBI_MAX_N + 2];
  __shared__ int s_rank[JACOBI_MAX_N + 2];
  __shared__ double s_gmax;
  __shared__ int s_nrot;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;

  for (int idx = tid; idx < np * np; idx += nt) {
    const int c = idx / np, r = idx % np;
    double v = 0.0;
    if (r < n && c < n) v = 0.5 * (Gin[(size_t)r * ldg + c] + Gin[(size_t)c * ldg + r]);  // symmetrise
    Wt[idx] = v;
    Vt[idx] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  // scale reference: largest squared column norm (~ lambda_max^2)
  for (int c = warp; c < np; c += nwarps) {
    double s = 0.0;
    for (int r = lane; r < np; r += 32) { const double a = Wt[(size_t)c * np + r]; s += a * a; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_w[c] = s;
  }
  __syncthreads();
  if (tid == 0) {
    double g = 0.0;
    for (int c = 0; c < np; ++c) g = s_w[c] > g ? s_w[c] : g;
    s_gmax = g;
  }
  __syncthreads();
  const double floor2 = 1e-30 * s_gmax;  // pairs of columns both at the noise level are left alone

  int sweeps_done = 0;
  bool converged = false;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    if (tid == 0) s_nrot = 0;
    __syncthreads();
    int rot = 0;
    for (int step = 0; step < np - 1; ++step) {
      for (int pi = warp; pi < m; pi += nwarps) {
        int p, q;
        if (pi == 0) {
          p = np - 1;
          q = step % (np - 1);
        } else {
          p = (step + pi) % (np - 1);
          q = (step - pi + (np - 1)) % (np - 1);
        }
        if (p > q) { const int t = p; p = q; q = t; }
        double* wp = Wt + (size_t)p * np;
        double* wq = Wt + (size_t)q * np;
        double dpp = 0.0, dqq = 0.0, dpq = 0.0;
        for (int r = lane; r < np; r += 32) {
          const double a = wp[r], b = wq[r];
          dpp = fma(a, a, dpp);
          dqq = fma(b, b, dqq);
          dpq = fma(a, b, dpq);
        }
        for (int o = 16; o > 0; o >>= 1) {
          dpp += __shfl_xor_sync(0xffffffffu, dpp, o);
          dqq += __shfl_xor_sync(0xffffffffu, dqq, o);
          dpq += __shfl_xor_sync(0xffffffffu, dpq, o);
        }
        const double scale = sqrt(dpp * dqq);
        if (fabs(dpq) > tol * scale && scale > floor2) {  // warp-uniform
          rot = 1;
          const double tau = (dqq - dpp) / (2.0 * dpq);
          const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          const double c = 1.0 / sqrt(1.0 + t * t);
          const double s = t * c;
          double* vp = Vt + (size_t)p * np;
          double* vq = Vt + (size_t)q * np;
          for (int r = lane; r < np; r += 32) {
            const double a = wp[r], b = wq[r];
            wp[r] = c * a - s * b;
            wq[r] = s * a + c * b;
            const double x = vp[r], y = vq[r];
            vp[r] = c * x - s * y;
            vq[r] = s * x + c * y;
          }
        }
      }
      __syncthreads();
    }
    if (lane == 0 && rot) atomicAdd(&s_nrot, 1);
    __syncthreads();
    sweeps_done = sweep + 1;
    const int nrot = s_nrot;
    __syncthreads();
    if (nrot == 0) {  // a full sweep without a single rotation: converged
      converged = true;
      break;
    }
  }

  // eigenvalues as Rayleigh quotients lambda_c = v_c . (G v_c) = v_c . w_c ; sort descending; write out
  for (int c = warp; c < n; c += nwarps) {
    double s = 0.0;
    for (int r = lane; r < np; r += 32) s = fma(Vt[(size_t)c * np + r], Wt[(size_t)c * np + r], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_w[c] = s;
  }
  __syncthreads();
  for (int i = tid; i < n; i += nt) {
    const double wi = s_w[i];
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const double wj = s_w[j];
      r += (wj > wi) || (wj == wi && j < i);
    }
    s_rank[i] = r;
    w_out[r] = wi;
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += nt) {
    const int k = idx / n, i = idx % n;
    V_out[(size_t)k * n + s_rank[i]] = Vt[(size_t)i * np + k];
  }
  if (tid == 0 && info) info[0] = converged ? sweeps_done : -sweeps_done;
}

inline size_t jacobi_scratch_doubles(int n) {
  const int np = n + (n & 1);
  return (size_t)2 * np * np;
}

// G: n x n fp64 (ld = ldg). w: n, V: n x n. scratch: jacobi_scratch_doubles(n) doubles (+ 1 int info at the end).
inline int jacobi_eigh(const double* G, int n, int ldg, double* w, double* V, double* scratch, int* info,
                       cudaStream_t st) {
  if (n < 1 || n > JACOBI_MAX_N) return fail(TNB_ERR_UNSUPPORTED, "jacobi_eigh: n=%d outside [1,%d]", n, JACOBI_MAX_N);
  const int np = n + (n & 1);
  const double tol = 1e-14;  // relative off-diagonal threshold |a_pq| <= tol*sqrt(a_pp*a_qq)
  const int max_sweeps = 30;
  if (n <= JACOBI_SMEM_MAX_N) {
    const size_t smem = (size_t)2 * np * np * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
      TNB_CUDA(cudaFuncSetAttribute(jacobi_eigh_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    2 * (JACOBI_SMEM_MAX_N) * (JACOBI_SMEM_MAX_N) * (int)sizeof(double)));
      attr_set = true;
    }
    jacobi_eigh_kernel<true><<<1, JACOBI_THREADS, smem, st>>>(G, n, ldg, w, V, scratch, max_sweeps, tol, info);
  } else {
    jacobi_eigh_kernel<false><<<1, JACOBI_THREADS, 0, st>>>(G, n, ldg, w, V, scratch, max_sweeps, tol, info);
  }
  TNB_LAUNCH_CHECK();
  return TNB_OK;
}

}  // namespace tnb
