// Two-sided (classical) Jacobi for small symmetric matrices, parallel ordering, one 2x2 block per thread.
//
// Replaces torch.linalg.eigh (round.py:114) / the U,S part of torch.linalg.svd (round.py:96) on the Gram matrices of the
// sweep when n <= JAC2_MAX_N, and the b x b Rayleigh-Ritz problems of the subspace eigensolver.
//
// Why not the one-sided kernel of jacobi.cuh: there every column pair needs three length-n inner products per round
// (shuffle reductions) and rotates W = G V and V; per round that is ~150 instructions per thread on 2 warps per
// scheduler (measured 0.58 ms for a 64 x 64 fp64 problem).  Here S itself is rotated, S <- J^T S J: a round needs only
// the three numbers (S_pp, S_qq, S_pq) per pair — no reductions — and the update of S splits into (n/2)^2 independent
// 2x2 blocks B_ab <- J_a^T B_ab J_b (one thread each, upper triangle only) plus the column rotations of V.
//
// Parallel ordering without index bookkeeping: the pairs are always (2k, 2k+1); after every round rows/columns are
// PHYSICALLY permuted by the round-robin map sigma (folded into the store of the updated blocks, double-buffered), so
// that n-1 rounds visit every pair once (Brent-Luk tournament).  V's columns follow the same relabelling, its rows
// (original coordinates) do not, hence (S_jj, V[:, j]) stays a consistent eigenpair estimate.
//
// The per-thread phase functions are plain C++ over pointers (TNB_HD), so the index logic is unit-tested on the host
// (tests/host_emul/) by looping "threads" between the two barriers of a round; the CUDA wrappers are in jacobi2.cuh.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define TNB_HD __host__ __device__ __forceinline__
#else
#define TNB_HD inline
#endif

namespace tnb {

constexpr int JAC2_MAX_N = 128;

// round-robin relabelling: index 2k = "top k", 2k+1 = "bottom k"; top 0 stays, the others rotate one seat
TNB_HD int jac2_sigma(int i, int m) {
  if (m <= 1 || i == 0) return i;
  const int k = i >> 1;
  if (i & 1) return k == 0 ? 2 : 2 * k - 1;       // bottom k -> bottom k-1 (bottom 0 -> top 1)
  return k == m - 1 ? 2 * m - 1 : 2 * k + 2;      // top k -> top k+1 (top m-1 -> bottom m-1)
}

template <typename R>
struct Jac2 {
  R* S[2];        // np x lds symmetric matrix, double-buffered
  R* V[2];        // np x lds accumulated rotations
  R* cs;          // 3 * m: (c, s, t) of each pair for the current round
  unsigned short* blk;  // m(m+1)/2 upper-triangle block list: a | b << 8
  int* flag;      // [0]: a rotation happened in this sweep
  int np, m, lds;
  R tol2;         // rotate iff S_pq^2 > tol2 * |S_pp S_qq|
  R big2;         // a rotation with S_pq^2 > big2 * |S_pp S_qq| (big2 = tol) asks for another sweep: convergence is
                  // quadratic, so a sweep whose largest relative off-diagonal was below sqrt(tol) ends below tol
  R floor_abs;    // ... and |S_pq| > floor_abs (~eps of R relative to the largest diagonal entry: skipping such a pair
                  // is a backward error below the rounding level; noise-level pairs of a rank-deficient Gram matrix
                  // would otherwise keep rotating among themselves)
};

TNB_HD void jac2_rotation_plain(double spp, double sqq, double spq, double& c, double& s, double& t) {
  const double tau = (sqq - spp) / (2.0 * spq);
  t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
  c = 1.0 / sqrt(1.0 + t * t);
  s = t * c;
}
TNB_HD void jac2_rotation_plain(float spp, float sqq, float spq, float& c, float& s, float& t) {
  const float tau = (sqq - spp) / (2.0f * spq);
  t = (tau >= 0.0f ? 1.0f : -1.0f) / (fabsf(tau) + sqrtf(1.0f + tau * tau));
  c = 1.0f / sqrtf(1.0f + t * t);
  s = t * c;
}

#ifdef __CUDA_ARCH__
// device: t from fp32 MUFU on exponent-normalised inputs, c refined by two Newton steps (same construction as
// jacobi.cuh::jacobi_rotation: any t gives an exactly orthogonal rotation as long as c and s are consistent)
__device__ __forceinline__ void jac2_rotation(double spp, double sqq, double spq, double& c, double& s, double& t) {
  const double h = sqq - spp, b2 = 2.0 * spq;
  const double big = fmax(fabs(h), fabs(b2));
  const int ebits = (__double2hiint(big) >> 20) & 0x7ff;
  int sbits = 2046 - ebits;
  sbits = sbits < 1 ? 1 : (sbits > 2046 ? 2046 : sbits);
  const double scale = __hiloint2double(sbits << 20, 0);
  const float hf = (float)(h * scale), bf = (float)(b2 * scale);
  const float radf = sqrtf(fmaf(hf, hf, bf * bf));
  float tf = __fdividef(bf, fabsf(hf) + radf);
  if (hf < 0.f) tf = -tf;
  t = (double)tf;  // fp32-accurate root: S_pq drops by ~1e-7 per rotation, the next sweep finishes (still superlinear)
  const double x = fma(t, t, 1.0);
  double y = (double)rsqrtf((float)x);
  y = y * fma(-0.5 * x, y * y, 1.5);
  y = y * fma(-0.5 * x, y * y, 1.5);
  c = y;
  s = t * y;
}
__device__ __forceinline__ void jac2_rotation(float spp, float sqq, float spq, float& c, float& s, float& t) {
  const float h = sqq - spp, b2 = 2.f * spq;
  const float radf = sqrtf(fmaf(h, h, b2 * b2));
  t = __fdividef(b2, fabsf(h) + radf);
  if (h < 0.f) t = -t;
  c = rsqrtf(fmaf(t, t, 1.f));
  s = t * c;
}
#else
template <typename R>
inline void jac2_rotation(R spp, R sqq, R spq, R& c, R& s, R& t) { jac2_rotation_plain(spp, sqq, spq, c, s, t); }
#endif

// Block list (a <= b), built once: item w -> (a, b).  Row-major over the upper triangle so that a warp's items share `a`.
template <typename R>
TNB_HD void jac2_build_blocks(const Jac2<R>& J, int tid, int nthreads) {
  const int m = J.m;
  for (int a = tid; a < m; a += nthreads) {
    const int base = a * m - a * (a - 1) / 2;  // number of blocks in rows < a
    for (int b = a; b < m; ++b) J.blk[base + (b - a)] = (unsigned short)(a | (b << 8));
  }
}

// Phase A (threads tid < m): rotation of pair (2 tid, 2 tid + 1) from buffer `cur`.
template <typename R>
TNB_HD void jac2_phase_a(const Jac2<R>& J, int cur, int tid) {
  if (tid >= J.m) return;
  const R* S = J.S[cur];
  const int p = 2 * tid, q = p + 1;
  const R spp = S[p * J.lds + p], sqq = S[q * J.lds + q], spq = S[p * J.lds + q];
  R c = (R)1, s = (R)0, t = (R)0;
  const R prod = spp * sqq;
  const R aprod = prod < (R)0 ? -prod : prod;
  if (spq * spq > J.tol2 * aprod && (spq < (R)0 ? -spq : spq) > J.floor_abs) {
    jac2_rotation(spp, sqq, spq, c, s, t);
    if (spq * spq > J.big2 * aprod) J.flag[0] = 1;  // benign race: every writer stores 1
  }
  J.cs[3 * tid] = c;
  J.cs[3 * tid + 1] = s;
  J.cs[3 * tid + 2] = t;
}

// Phase B: work items [0, U) are the upper-triangle 2x2 blocks of S, items [U, U + np*m) the (row, pair) column
// rotations of V.  Reads buffer `cur`, writes buffer `cur ^ 1` at the relabelled positions.
template <typename R>
TNB_HD void jac2_phase_b(const Jac2<R>& J, int cur, int tid, int nthreads) {
  const int m = J.m, lds = J.lds, np = J.np;
  const int U = m * (m + 1) / 2;
  const int total = U + np * m;
  const R* S = J.S[cur];
  R* So = J.S[cur ^ 1];
  const R* V = J.V[cur];
  R* Vo = J.V[cur ^ 1];
  for (int w = tid; w < total; w += nthreads) {
    if (w < U) {
      const int ab = J.blk[w];
      const int a = ab & 255, b = ab >> 8;
      const int pa = 2 * a, qa = pa + 1, pb = 2 * b, qb = pb + 1;
      const int spa = jac2_sigma(pa, m), sqa = jac2_sigma(qa, m), spb = jac2_sigma(pb, m), sqb = jac2_sigma(qb, m);
      const R ca = J.cs[3 * a], sa = J.cs[3 * a + 1];
      if (a == b) {
        // exact similarity for ANY consistent (c, s) — t is only fp32-accurate on the device, so S_pq is not forced
        // to zero (that would be a backward error of 1e-7 |S_pq| per rotation) but carried at its true small value
        const R spp = S[pa * lds + pa], sqq = S[qa * lds + qa], spq = S[pa * lds + qa];
        const R cc = ca * ca, ss = sa * sa, cs2 = (R)2 * ca * sa;
        const R npp = cc * spp - cs2 * spq + ss * sqq;
        const R nqq = ss * spp + cs2 * spq + cc * sqq;
        const R npq = ca * sa * (spp - sqq) + (cc - ss) * spq;
        So[spa * lds + spa] = npp;
        So[sqa * lds + sqa] = nqq;
        So[spa * lds + sqa] = npq;
        So[sqa * lds + spa] = npq;
      } else {
        const R cb = J.cs[3 * b], sb = J.cs[3 * b + 1];
        const R b00 = S[pa * lds + pb], b01 = S[pa * lds + qb], b10 = S[qa * lds + pb], b11 = S[qa * lds + qb];
        // T = B J_b,  J_b = [c s; -s c]
        const R t00 = cb * b00 - sb * b01, t01 = sb * b00 + cb * b01;
        const R t10 = cb * b10 - sb * b11, t11 = sb * b10 + cb * b11;
        // B' = J_a^T T,  J_a^T = [c -s; s c]
        const R n00 = ca * t00 - sa * t10, n01 = ca * t01 - sa * t11;
        const R n10 = sa * t00 + ca * t10, n11 = sa * t01 + ca * t11;
        So[spa * lds + spb] = n00;
        So[spa * lds + sqb] = n01;
        So[sqa * lds + spb] = n10;
        So[sqa * lds + sqb] = n11;
        So[spb * lds + spa] = n00;
        So[sqb * lds + spa] = n01;
        So[spb * lds + sqa] = n10;
        So[sqb * lds + sqa] = n11;
      }
    } else {
      const int idx = w - U;
      const int r = idx / m, pr = idx - r * m;
      const int p = 2 * pr, q = p + 1;
      const R c = J.cs[3 * pr], s = J.cs[3 * pr + 1];
      const R x = V[r * lds + p], y = V[r * lds + q];
      Vo[r * lds + jac2_sigma(p, m)] = c * x - s * y;
      Vo[r * lds + jac2_sigma(q, m)] = s * x + c * y;
    }
  }
}

}  // namespace tnb
