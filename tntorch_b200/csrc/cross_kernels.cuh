// Batched TT-cross plumbing on the device (tn.cross, tntorch/cross.py:316-455, for B independent problems that share the
// grid and the rank profile — BASELINE.json config 5).  The reference has no batch support (cross.py:256-258) and keeps
// its index sets on the host; its sampling step forms the coordinates through interface products
// (einsum 'ai,ibj,jc->abc', cross.py:318-321).  For a tensor-product grid those products only SELECT grid values, so
// here the multi-indices themselves live on the device and three small kernels do the work:
//
//   cross_gather_coords   (b, a, i, c) -> the N coordinates of sample (lsets[j][b,a,:], i, rsets[j][b,c,:])
//   cross_update_lsets / cross_update_rsets   the nested index sets after a maxvol step (cross.py:405-411, 437-443)
//   cross_tt_eval         values of B TT tensors at P multi-indices each (the validation error, cross.py:457-459)
//
// Index arithmetic is int32 and exact.  QR and maxvol of the B sample matrices are one batched launch each
// (qr.cuh, maxvol.cuh).
#pragma once
#include "common.cuh"

namespace tnb {

// lsets: [B][Rl][j] (multi-index prefixes of modes 0..j-1), rsets: [B][Rr][N-j-1] (suffixes of modes j+1..N-1).
// grid: [N][Imax] coordinates (fp64), Is: [N] mode sizes.  X: [N][B*Rl*I*Rr] output, sample p = (a*I + i)*Rr + c.
__global__ void cross_gather_coords_kernel(const int* __restrict__ lsets, const int* __restrict__ rsets,
                                           const double* __restrict__ grid, int Imax, int B, int N, int j, int Rl, int I,
                                           int Rr, double* __restrict__ X) {
  const int64_t P = (int64_t)Rl * I * Rr, total = (int64_t)B * P;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / P);
    const int64_t p = t - (int64_t)b * P;
    const int c = (int)(p % Rr);
    const int i = (int)((p / Rr) % I);
    const int a = (int)(p / ((int64_t)Rr * I));
    const int* ls = lsets + ((size_t)b * Rl + a) * j;
    const int* rs = rsets + ((size_t)b * Rr + c) * (N - j - 1);
    for (int k = 0; k < j; ++k) X[(size_t)k * total + t] = grid[(size_t)k * Imax + ls[k]];
    X[(size_t)j * total + t] = grid[(size_t)j * Imax + i];
    for (int k = j + 1; k < N; ++k) X[(size_t)k * total + t] = grid[(size_t)k * Imax + rs[k - j - 1]];
  }
}

// After the left-to-right maxvol of core j: local[b][s] indexes the rows (a, i) of the (Rl*I) x Rn sample matrix.
// lnext[b][s][0..j-1] = lsets[b][a][:], lnext[b][s][j] = i.   active[b] == 0: problem b is frozen (already converged).
__global__ void cross_update_lsets_kernel(const int* __restrict__ lsets, const int* __restrict__ local, int B, int j, int Rl,
                                          int I, int Rn, const int* __restrict__ active, int* __restrict__ lnext) {
  const int total = B * Rn;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int b = t / Rn;
    if (active && !active[b]) continue;
    const int row = local[t];
    const int a = row / I, i = row - a * I;
    const int* src = lsets + ((size_t)b * Rl + a) * j;
    int* dst = lnext + (size_t)t * (j + 1);
    for (int k = 0; k < j; ++k) dst[k] = src[k];
    dst[j] = i;
  }
}

// After the right-to-left maxvol of core j: local[b][s] indexes the rows (i, c) of the (I*Rr) x Rp transposed sample
// matrix.  rprev[b][s][0] = i, rprev[b][s][1..] = rsets[b][c][:]  (suffix of modes j..N-1, length N-j).
__global__ void cross_update_rsets_kernel(const int* __restrict__ rsets, const int* __restrict__ local, int B, int N, int j,
                                          int I, int Rr, int Rp, const int* __restrict__ active, int* __restrict__ rprev) {
  const int total = B * Rp;
  const int len = N - j - 1;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int b = t / Rp;
    if (active && !active[b]) continue;
    const int row = local[t];
    const int i = row / Rr, c = row - i * Rr;
    const int* src = rsets + ((size_t)b * Rr + c) * len;
    int* dst = rprev + (size_t)t * (len + 1);
    dst[0] = i;
    for (int k = 0; k < len; ++k) dst[k + 1] = src[k];
  }
}

constexpr int CROSS_EVAL_MAX_R = 64;
struct CrossEvalArgs {
  const double* cores[32];  // core n: [B][R_n][I_n][R_{n+1}]
  int R[33];
  int I[32];
  int N;
};
// out[b][p] = TT_b(idx[p][0..N-1]); idx is shared by the problems ([P][N]) or per problem ([B][P][N], per_problem = 1).
__global__ void cross_tt_eval_kernel(const CrossEvalArgs args, const int* __restrict__ idx, int B, int P, int per_problem,
                                     double* __restrict__ out) {
  const int total = B * P;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int b = t / P, p = t - b * P;
    const int* ix = idx + ((size_t)(per_problem ? b : 0) * P + p) * args.N;
    double v[CROSS_EVAL_MAX_R], w[CROSS_EVAL_MAX_R];
    v[0] = 1.0;
    for (int n = 0; n < args.N; ++n) {
      const int r0 = args.R[n], r1 = args.R[n + 1], In = args.I[n];
      const double* c = args.cores[n] + (size_t)b * r0 * In * r1 + (size_t)ix[n] * r1;
      for (int q = 0; q < r1; ++q) {
        double s = 0.0;
        for (int a = 0; a < r0; ++a) s = fma(v[a], c[(size_t)a * In * r1 + q], s);
        w[q] = s;
      }
      for (int q = 0; q < r1; ++q) v[q] = w[q];
    }
    out[t] = v[0];
  }
}

}  // namespace tnb
