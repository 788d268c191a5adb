// Small dense symmetric eigenproblems on one CTA: parallel-order two-sided Jacobi with
// warp/block reductions, fp64 throughout.  Replaces torch.linalg.eigh (round.py:114) and the
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

// Eigen-decomposition of the symmetric n x n matrix Gin (leading dimension ldg).
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
  double* A;
  double* V;
  if (SMEM) {
    A = reinterpret_cast<double*>(jac_smem_raw);
    V = A + (size_t)np * np;
  } else {
    A = scratch;
    V = scratch + (size_t)np * np;
  }
  __shared__ double s_c[JACOBI_MAX_N / 2 + 1], s_s[JACOBI_MAX_N / 2 + 1];
  __shared__ int s_p[JACOBI_MAX_N / 2 + 1], s_q[JACOBI_MAX_N / 2 + 1];
  __shared__ double s_red[32];
  __shared__ double s_w[JACOBI_MAX_N + 2];
  __shared__ int s_rank[JACOBI_MAX_N + 2];
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int idx = tid; idx < np * np; idx += nt) {
    const int i = idx / np, j = idx % np;
    double v = 0.0;
    if (i < n && j < n) v = 0.5 * (Gin[(size_t)i * ldg + j] + Gin[(size_t)j * ldg + i]);  // symmetrise
    A[idx] = v;
    V[idx] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();

  __shared__ int s_nrot;
  int sweeps_done = 0;
  bool converged = false;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    if (tid == 0) s_nrot = 0;
    __syncthreads();
    for (int step = 0; step < np - 1; ++step) {
      // phase 1: the m disjoint pairs of this round (circle method) and their rotations
      if (tid < m) {
        int p, q;
        if (tid == 0) {
          p = np - 1;
          q = step % (np - 1);
        } else {
          p = (step + tid) % (np - 1);
          q = (step - tid + (np - 1)) % (np - 1);
        }
        if (p > q) { const int t = p; p = q; q = t; }
        const double app = A[(size_t)p * np + p], aqq = A[(size_t)q * np + q], apq = A[(size_t)p * np + q];
        double c = 1.0, s = 0.0;
        // relative (Demmel-Veselic) rotation threshold: keeps small eigenvalues of PSD matrices accurate
        if (fabs(apq) > tol * sqrt(fabs(app * aqq)) && fabs(apq) > 1e-290) {
          atomicAdd(&s_nrot, 1);
          const double tau = (aqq - app) / (2.0 * apq);
          const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + t * t);
          s = t * c;
        }
        s_p[tid] = p; s_q[tid] = q; s_c[tid] = c; s_s[tid] = s;
      }
      __syncthreads();
      // phase 2: A <- A J, V <- V J  (columns p,q of every row)
      for (int idx = tid; idx < m * np; idx += nt) {
        const int i = idx % m, k = idx / m;
        const double c = s_c[i], s = s_s[i];
        if (s == 0.0) continue;
        const int p = s_p[i], q = s_q[i];
        const size_t kp = (size_t)k * np + p, kq = (size_t)k * np + q;
        const double a1 = A[kp], a2 = A[kq];
        A[kp] = c * a1 - s * a2;
        A[kq] = s * a1 + c * a2;
        const double v1 = V[kp], v2 = V[kq];
        V[kp] = c * v1 - s * v2;
        V[kq] = s * v1 + c * v2;
      }
      __syncthreads();
      // phase 3: A <- J^T A  (rows p,q across every column)
      for (int idx = tid; idx < m * np; idx += nt) {
        const int i = idx / np, k = idx % np;
        const double c = s_c[i], s = s_s[i];
        if (s == 0.0) continue;
        const int p = s_p[i], q = s_q[i];
        const size_t pk = (size_t)p * np + k, qk = (size_t)q * np + k;
        const double a1 = A[pk], a2 = A[qk];
        A[pk] = c * a1 - s * a2;
        A[qk] = s * a1 + c * a2;
      }
      __syncthreads();
      if (tid < m && s_s[tid] != 0.0) {  // the annihilated pair: exact zeros, symmetric
        const int p = s_p[tid], q = s_q[tid];
        A[(size_t)p * np + q] = 0.0;
        A[(size_t)q * np + p] = 0.0;
      }
      __syncthreads();
    }
    sweeps_done = sweep + 1;
    const int nrot = s_nrot;
    __syncthreads();
    if (nrot == 0) {  // a full sweep without a single rotation: converged
      converged = true;
      break;
    }
  }

  // sort descending (rank by counting), write out
  for (int i = tid; i < n; i += nt) s_w[i] = A[(size_t)i * np + i];
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
    V_out[(size_t)k * n + s_rank[i]] = V[(size_t)k * np + i];
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
