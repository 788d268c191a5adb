// maxvol on the device: dominant r x r submatrix of a tall N x r matrix, one CTA per problem (batched).
// Same algorithm as the reference's py_maxvol (tntorch/maxvol.py:114-170): LU with partial pivoting of the
// N x r matrix (getrf, :135) -> pivot rows (:137-141) -> C = A inv(A[pivots]) by two triangular solves
// (:145-148) -> repeat { (i,j) = argmax |C| ; stop if <= tol ; swap row j into slot i, rank-1 update (:160-169) }.
// The reference runs this on host NumPy after a forced device->host copy per TT core (cross.py:400-402);
// here the coefficient matrix never leaves the GPU and is returned: it IS the interpolation core
// Q inv(Q[local]) that cross.py:403 recomputes with lstsq.
//
// Index arithmetic is exact (int32); ties in both argmax searches resolve to the first element in the
// reference's scan order (LAPACK idamax: lowest row; NumPy argmax over the r x N array: lowest i*N + j).
#pragma once
#include "common.cuh"

namespace tnb {

struct ArgMax {
  double v;
  long long key;
};
__device__ __forceinline__ ArgMax argmax_better(ArgMax a, ArgMax b) {
  return (b.v > a.v || (b.v == a.v && b.key < a.key)) ? b : a;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax x, ArgMax* red) {
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.key = __shfl_xor_sync(0xffffffffu, x.key, o);
    x = argmax_better(x, y);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = x;
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    ArgMax t = lane < nw ? red[lane] : ArgMax{-1.0, 0x7fffffffffffffffLL};
    for (int o = 16; o > 0; o >>= 1) {
      ArgMax y;
      y.v = __shfl_xor_sync(0xffffffffu, t.v, o);
      y.key = __shfl_xor_sync(0xffffffffu, t.key, o);
      t = argmax_better(t, y);
    }
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  const ArgMax out = red[0];
  __syncthreads();
  return out;
}

// A: [nbatch][N][r] fp64 (read only).  work: [nbatch][N][r] scratch.  Cout: [nbatch][N][r].  index_out: [nbatch][r].
// iters_out: [nbatch] number of swap iterations performed.
__global__ void __launch_bounds__(256) maxvol_kernel(const double* __restrict__ A_all, int N, int r, double tol,
                                                     int max_iters, double* __restrict__ work_all,
                                                     double* __restrict__ C_all, int* __restrict__ index_all,
                                                     int* __restrict__ perm_all, int* __restrict__ iters_out) {
  __shared__ ArgMax red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t boff = (size_t)blockIdx.x * N * r;
  const double* A = A_all + boff;
  double* B = work_all + (size_t)blockIdx.x * ((size_t)N * r + N + r);
  double* C = C_all + boff;
  int* index = index_all + (size_t)blockIdx.x * r;
  int* perm = perm_all + (size_t)blockIdx.x * N;  // row permutation of the LU: perm[k] = original row now at position k

  if (N <= r) {  // maxvol.py:126-127: all rows, identity coefficients
    for (int i = tid; i < N; i += nt) index[i] = i;
    for (int i = N + tid; i < r; i += nt) index[i] = -1;
    for (int idx = tid; idx < N * r; idx += nt) C[idx] = (idx / r == idx % r) ? 1.0 : 0.0;
    if (tid == 0 && iters_out) iters_out[blockIdx.x] = 0;
    return;
  }
  if (tol < 1.0) tol = 1.0;
  for (int idx = tid; idx < N * r; idx += nt) B[idx] = A[idx];
  for (int i = tid; i < N; i += nt) perm[i] = i;
  __syncthreads();

  // ---- LU with partial pivoting (unblocked getrf) ----
  for (int k = 0; k < r; ++k) {
    ArgMax best{-1.0, 0x7fffffffffffffffLL};
    for (int i = k + tid; i < N; i += nt) best = argmax_better(best, ArgMax{fabs(B[(size_t)i * r + k]), (long long)i});
    best = block_argmax(best, red);
    const int p = (int)best.key;
    if (p != k) {  // swap rows k and p (whole rows, like LAPACK's laswp)
      for (int j = tid; j < r; j += nt) {
        const double t = B[(size_t)k * r + j];
        B[(size_t)k * r + j] = B[(size_t)p * r + j];
        B[(size_t)p * r + j] = t;
      }
      if (tid == 0) { const int t = perm[k]; perm[k] = perm[p]; perm[p] = t; }
    }
    __syncthreads();
    const double piv = B[(size_t)k * r + k];
    const double inv = piv != 0.0 ? 1.0 / piv : 0.0;
    for (int i = k + 1 + tid; i < N; i += nt) B[(size_t)i * r + k] *= inv;
    __syncthreads();
    const int rem = r - k - 1;
    if (rem > 0) {
      for (long long idx = tid; idx < (long long)(N - k - 1) * rem; idx += nt) {
        const int i = k + 1 + (int)(idx / rem), j = k + 1 + (int)(idx % rem);
        B[(size_t)i * r + j] -= B[(size_t)i * r + k] * B[(size_t)k * r + j];
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < r; i += nt) index[i] = perm[i];
  // ---- C = A inv(A[index]) : c_i L1 U = a_i, one row per thread (H = B[0:r] holds L1 \ U) ----
  for (int i = tid; i < N; i += nt) {
    double* c = C + (size_t)i * r;
    const double* a = A + (size_t)i * r;
    for (int j = 0; j < r; ++j) {  // y U = a
      double s = a[j];
      for (int k = 0; k < j; ++k) s -= c[k] * B[(size_t)k * r + j];
      c[j] = s / B[(size_t)j * r + j];
    }
    for (int j = r - 1; j >= 0; --j) {  // c L1 = y (unit diagonal)
      double s = c[j];
      for (int k = j + 1; k < r; ++k) s -= c[k] * B[(size_t)k * r + j];
      c[j] = s;
    }
  }
  __syncthreads();
  // ---- greedy swaps (maxvol.py:150-169) ----
  double* tmp_row = B + (size_t)N * r;  // length N  (C[:, i] of the N x r layout = reference C[i])
  double* tmp_col = tmp_row + N;        // length r  (C[j, :] = reference C[:, j])
  int iters = 0;
  while (true) {
    ArgMax best{-1.0, 0x7fffffffffffffffLL};
    for (long long idx = tid; idx < (long long)N * r; idx += nt) {
      const int j = (int)(idx / r), i = (int)(idx % r);  // element C[j][i] <-> reference C[i, j]
      best = argmax_better(best, ArgMax{fabs(C[idx]), (long long)i * N + j});
    }
    best = block_argmax(best, red);
    if (!(best.v > tol) || iters >= max_iters) break;
    const int i = (int)(best.key / N), j = (int)(best.key % N);
    if (tid == 0) index[i] = j;
    const double cij = C[(size_t)j * r + i];
    const double alpha = -1.0 / cij;
    for (int b = tid; b < N; b += nt) tmp_row[b] = C[(size_t)b * r + i];
    for (int a = tid; a < r; a += nt) tmp_col[a] = C[(size_t)j * r + a] - (a == i ? 1.0 : 0.0);
    __syncthreads();
    for (long long idx = tid; idx < (long long)N * r; idx += nt) {
      const int b = (int)(idx / r), a = (int)(idx % r);
      C[idx] += alpha * tmp_col[a] * tmp_row[b];
    }
    ++iters;
    __syncthreads();
  }
  if (tid == 0 && iters_out) iters_out[blockIdx.x] = iters;
}

inline size_t maxvol_workspace_bytes(int nbatch, int N, int r) {
  return align_up((size_t)nbatch * ((size_t)N * r + N + r) * sizeof(double)) + align_up((size_t)nbatch * N * sizeof(int)) +
         align_up((size_t)nbatch * sizeof(int));
}

inline int maxvol_run(const double* A, int nbatch, int N, int r, double tol, int max_iters, void* ws, size_t ws_bytes,
                      int* index_out, double* C_out, int* iters_host, cudaStream_t st) {
  if (nbatch < 1 || N < 1 || r < 1) return fail(TNB_ERR_INVALID, "maxvol: bad shape nbatch=%d N=%d r=%d", nbatch, N, r);
  if (ws_bytes < maxvol_workspace_bytes(nbatch, N, r)) return fail(TNB_ERR_WORKSPACE, "maxvol: workspace too small");
  Arena ar(ws, ws_bytes);
  double* work = ar.take<double>((size_t)nbatch * ((size_t)N * r + N + r));
  int* perm = ar.take<int>((size_t)nbatch * N);
  int* iters = ar.take<int>(nbatch);
  maxvol_kernel<<<nbatch, 256, 0, st>>>(A, N, r, tol, max_iters, work, C_out, index_out, perm, iters);
  TNB_LAUNCH_CHECK();
  if (iters_host) {
    TNB_CUDA(cudaMemcpyAsync(iters_host, iters, sizeof(int) * nbatch, cudaMemcpyDeviceToHost, st));
    TNB_CUDA(cudaStreamSynchronize(st));
  }
  return TNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// rect_maxvol: the rectangular extension (tntorch/maxvol.py:30-111).  After maxvol (start_maxvol_iters swaps) the
// row with the largest squared 2-norm in the coefficient matrix C is added to the index set while that norm exceeds
// tol^2 (and K < maxK), or while K < minK; each addition is the Sherman-Woodbury-Morrison update of maxvol.py:94-103:
//   c = C[i];  v = C c;  l = 1 / (1 + v_i);  C <- [C - l v c^T,  l v];  norms -= l v^2;  norms[chosen] = 0.
// One CTA per problem.  C lives in a caller buffer of N x maxK doubles (leading dimension maxK); on return K[b] columns
// are valid and, like the reference with identity_submatrix=True, the rows of the index set hold the identity.
// Ties in the argmax resolve to the lowest row, NumPy's argmax order.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rect_maxvol_extend_kernel(const double* __restrict__ C0_all, const int* __restrict__ idx0_all,
                                                                 int N, int r, double tol2, int minK, int maxK,
                                                                 double* __restrict__ C_all, int* __restrict__ index_all,
                                                                 int* __restrict__ K_all, double* __restrict__ scratch_all) {
  __shared__ ArgMax red[32];
  __shared__ double s_l;
  const int tid = threadIdx.x, nt = blockDim.x;
  const double* C0 = C0_all + (size_t)blockIdx.x * N * r;
  const int* idx0 = idx0_all + (size_t)blockIdx.x * r;
  double* C = C_all + (size_t)blockIdx.x * N * maxK;
  int* index = index_all + (size_t)blockIdx.x * maxK;
  double* norms = scratch_all + (size_t)blockIdx.x * (2 * (size_t)N + maxK);  // N
  double* v = norms + N;                                                       // N
  double* c = v + N;                                                           // maxK
  for (long long i = tid; i < (long long)N * maxK; i += nt) {
    const int row = (int)(i / maxK), col = (int)(i % maxK);
    C[i] = col < r ? C0[(size_t)row * r + col] : 0.0;
  }
  for (int a = tid; a < r; a += nt) index[a] = idx0[a];
  __syncthreads();
  for (int row = tid; row < N; row += nt) {
    double s = 0.0;
    for (int a = 0; a < r; ++a) { const double x = C[(size_t)row * maxK + a]; s += x * x; }
    norms[row] = s;
  }
  __syncthreads();
  for (int a = tid; a < r; a += nt) norms[idx0[a]] = 0.0;  // chosen rows do not compete (maxvol.py:76-79)
  __syncthreads();
  int K = r;
  for (;;) {
    ArgMax best{-1.0, 0x7fffffffffffffffLL};
    for (int row = tid; row < N; row += nt) best = argmax_better(best, ArgMax{norms[row], (long long)row});
    best = block_argmax(best, red);
    const int i = (int)best.key;
    if (!((best.v > tol2 && K < maxK) || K < minK)) break;
    for (int a = tid; a < K; a += nt) c[a] = C[(size_t)i * maxK + a];
    if (tid == 0) index[K] = i;
    __syncthreads();
    for (int row = tid; row < N; row += nt) {
      double s = 0.0;
      for (int a = 0; a < K; ++a) s = fma(C[(size_t)row * maxK + a], c[a], s);
      v[row] = s;
    }
    __syncthreads();
    if (tid == 0) s_l = 1.0 / (1.0 + v[i]);
    __syncthreads();
    const double l = s_l;
    for (long long e = tid; e < (long long)N * (K + 1); e += nt) {
      const int row = (int)(e / (K + 1)), a = (int)(e % (K + 1));
      if (a < K) C[(size_t)row * maxK + a] -= l * v[row] * c[a];
      else C[(size_t)row * maxK + K] = l * v[row];
    }
    for (int row = tid; row < N; row += nt) norms[row] -= l * v[row] * v[row];
    __syncthreads();
    if (tid == 0) norms[i] = 0.0;
    for (int a = tid; a < K; a += nt) norms[index[a]] = 0.0;
    ++K;
    __syncthreads();
  }
  // identity_submatrix=True (maxvol.py:107-109)
  for (int e = tid; e < K * K; e += nt) {
    const int a = e / K, col = e % K;
    C[(size_t)index[a] * maxK + col] = (a == col) ? 1.0 : 0.0;
  }
  if (tid == 0) K_all[blockIdx.x] = K;
}

inline size_t rect_maxvol_workspace_bytes(int nbatch, int N, int r, int maxK) {
  return maxvol_workspace_bytes(nbatch, N, r) + align_up((size_t)nbatch * N * r * sizeof(double)) +
         align_up((size_t)nbatch * r * sizeof(int)) + align_up((size_t)nbatch * (2 * (size_t)N + maxK) * sizeof(double));
}

// A: [nbatch][N][r].  index_out: [nbatch][maxK] int32, C_out: [nbatch][N][maxK] (ld maxK), K_out: [nbatch] (device).
inline int rect_maxvol_run(const double* A, int nbatch, int N, int r, double tol, int minK, int maxK, int start_iters,
                           void* ws, size_t ws_bytes, int* index_out, double* C_out, int* K_out, cudaStream_t st) {
  if (nbatch < 1 || N < 1 || r < 1 || N <= r) return fail(TNB_ERR_INVALID, "rect_maxvol: needs N > r (N=%d r=%d)", N, r);
  if (maxK < r || maxK > N || minK < r || minK > maxK)
    return fail(TNB_ERR_INVALID, "rect_maxvol: need r <= minK <= maxK <= N (r=%d minK=%d maxK=%d N=%d)", r, minK, maxK, N);
  if (ws_bytes < rect_maxvol_workspace_bytes(nbatch, N, r, maxK)) return fail(TNB_ERR_WORKSPACE, "rect_maxvol: workspace too small");
  char* base = static_cast<char*>(ws);
  const size_t mv = maxvol_workspace_bytes(nbatch, N, r);
  Arena ar(base + mv, ws_bytes - mv);
  double* C0 = ar.take<double>((size_t)nbatch * N * r);
  int* idx0 = ar.take<int>((size_t)nbatch * r);
  double* scratch = ar.take<double>((size_t)nbatch * (2 * (size_t)N + maxK));
  TNB_TRY(maxvol_run(A, nbatch, N, r, 1.05, start_iters, base, mv, idx0, C0, nullptr, st));  // maxvol.py:73
  rect_maxvol_extend_kernel<<<nbatch, 256, 0, st>>>(C0, idx0, N, r, tol * tol, minK, maxK, C_out, index_out, K_out, scratch);
  TNB_LAUNCH_CHECK();
  return TNB_OK;
}

}  // namespace tnb
