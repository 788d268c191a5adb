// Two-sided (classical) Jacobi for small symmetric matrices, parallel ordering, one 2x2 block per work item.
//
// Replaces torch.linalg.eigh (round.py:114) / the U,S part of torch.linalg.svd (round.py:96) on the Gram matrices of the
// sweep when they fit one CTA's shared memory, and the b x b Rayleigh-Ritz problems of the subspace eigensolver.
//
// Why not the one-sided kernel of jacobi.cuh: there every column pair needs three length-n inner products per round
// (shuffle reductions) and rotates W = G V and V.  Here S itself is rotated, S <- J^T S J: a round needs only the three
// numbers (S_pp, S_qq, S_pq) per pair — no reductions — and the update of S splits into independent 2x2 blocks
// B_ab <- J_a^T B_ab J_b (upper triangle only, canonical storage S[min][max]) plus the column rotations of V.
//
// Parallel ordering without index bookkeeping: the pairs are always (2k, 2k+1); after every round rows/columns are
// PHYSICALLY relabelled by the round-robin map sigma (folded into the stores, double-buffered), so n-1 rounds visit
// every pair once (Brent-Luk tournament).  Because the pairing never changes, the work of a thread is the SAME every
// round: its read and write offsets are computed once (Jac2Item, kept in registers) and a round costs a thread a handful
// of loads, ~16 flops and four stores per block — the first version recomputed the index algebra every round and was
// issue-bound at ~2000 cycles per round (measured: 1.1 ms for 64 x 64 fp64).
//
// One barrier per round: the rotation of the NEXT round's pair k needs only the three new elements of that pair, which
// a dedicated "pair thread" recomputes from the old blocks (ten loads) while the other warps update the matrix, so the
// ~250-cycle scalar chain of the rotation (sqrt, division, two Newton steps) is off the critical path.
//
// The per-thread functions are plain C++ over pointers (TNB_HD): the index logic is unit-tested on the host
// (tests/host_emul/) by looping "threads" between barriers; the CUDA wrappers are in jacobi2.cuh.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define TNB_HD __host__ __device__ __forceinline__
#else
#define TNB_HD inline
#endif

namespace tnb {

constexpr int JAC2_MAX_N = 128;
constexpr int JAC2_ITEMS = 3;  // work items per worker thread kept in registers

// round-robin relabelling: index 2k = "top k", 2k+1 = "bottom k"; top 0 stays, the others rotate one seat
TNB_HD int jac2_sigma(int i, int m) {
  if (m <= 1 || i == 0) return i;
  const int k = i >> 1;
  if (i & 1) return k == 0 ? 2 : 2 * k - 1;       // bottom k -> bottom k-1 (bottom 0 -> top 1)
  return k == m - 1 ? 2 * m - 1 : 2 * k + 2;      // top k -> top k+1 (top m-1 -> bottom m-1)
}
TNB_HD int jac2_sigma_inv(int j, int m) {
  if (m <= 1 || j == 0) return j;
  if (j == 2) return 1;                            // top 1 <- bottom 0
  if (j == 2 * m - 1) return 2 * m - 2;            // bottom m-1 <- top m-1
  if (j & 1) return j + 2;                         // bottom k-1 <- bottom k
  return j - 2;                                    // top k+1 <- top k
}

template <typename R>
struct Jac2 {
  R* S[2];        // np x lds, canonical upper storage (element (i, j) lives at [min(i,j)][max(i,j)]), double-buffered
  R* V[2];        // np x lds accumulated rotations
  R* cs[2];       // 2 * m: (c, s) of each pair, double-buffered (a round reads one, the pair threads fill the other)
  int* flag;      // [0], [1]: a "big" rotation was chosen in this (even / odd) sweep
  int np, m, lds;
  R tol2;         // rotate iff S_pq^2 > tol2 * |S_pp S_qq|
  R big2;         // a rotation with S_pq^2 > big2 * |S_pp S_qq| (big2 = tol) asks for another sweep: convergence is
                  // quadratic, so a sweep whose largest relative off-diagonal was below sqrt(tol) ends below tol
  R floor_abs;    // ... and |S_pq| > floor_abs (~eps of R relative to the largest diagonal entry: skipping such a pair
                  // is a backward error below the rounding level; noise-level pairs of a rank-deficient Gram matrix
                  // would otherwise keep rotating among themselves)
};

TNB_HD void jac2_rotation_plain(double spp, double sqq, double spq, double& c, double& s) {
  const double tau = (sqq - spp) / (2.0 * spq);
  const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
  c = 1.0 / sqrt(1.0 + t * t);
  s = t * c;
}
TNB_HD void jac2_rotation_plain(float spp, float sqq, float spq, float& c, float& s) {
  const float tau = (sqq - spp) / (2.0f * spq);
  const float t = (tau >= 0.0f ? 1.0f : -1.0f) / (fabsf(tau) + sqrtf(1.0f + tau * tau));
  c = 1.0f / sqrtf(1.0f + t * t);
  s = t * c;
}

#ifdef __CUDA_ARCH__
// device: t from fp32 MUFU on exponent-normalised inputs, c refined by two Newton steps (same construction as
// jacobi.cuh::jacobi_rotation: any t gives an exactly orthogonal rotation as long as c and s are consistent; the
// diagonal block is updated with the exact similarity formulas, so an fp32-accurate t only means that S_pq drops by
// ~1e-7 per rotation instead of to zero — the next sweep finishes, convergence stays superlinear)
__device__ __forceinline__ void jac2_rotation(double spp, double sqq, double spq, double& c, double& s) {
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
  const double t = (double)tf;
  const double x = fma(t, t, 1.0);
  double y = (double)rsqrtf((float)x);
  y = y * fma(-0.5 * x, y * y, 1.5);
  y = y * fma(-0.5 * x, y * y, 1.5);
  c = y;
  s = t * y;
}
__device__ __forceinline__ void jac2_rotation(float spp, float sqq, float spq, float& c, float& s) {
  const float h = sqq - spp, b2 = 2.f * spq;
  const float radf = sqrtf(fmaf(h, h, b2 * b2));
  float t = __fdividef(b2, fabsf(h) + radf);
  if (h < 0.f) t = -t;
  c = rsqrtf(fmaf(t, t, 1.f));
  s = t * c;
}
#else
template <typename R>
inline void jac2_rotation(R spp, R sqq, R spq, R& c, R& s) { jac2_rotation_plain(spp, sqq, spq, c, s); }
#endif

// Decide and compute the rotation of a pair from its three elements; raises the sweep flag for "big" rotations.
template <typename R>
TNB_HD void jac2_decide(const Jac2<R>& J, R spp, R sqq, R spq, R& c, R& s, int* flag) {
  c = (R)1;
  s = (R)0;
  const R prod = spp * sqq;
  const R aprod = prod < (R)0 ? -prod : prod;
  if (spq * spq > J.tol2 * aprod && (spq < (R)0 ? -spq : spq) > J.floor_abs) {
    jac2_rotation(spp, sqq, spq, c, s);
    if (spq * spq > J.big2 * aprod) *flag = 1;  // benign race: every writer stores 1
  }
}

TNB_HD int jac2_off(int i, int j, int lds) { return i <= j ? i * lds + j : j * lds + i; }

// ---- static work description of one thread ---------------------------------------------------------------------------
struct Jac2Item {
  int kind;        // 0 none, 1 off-diagonal block (a < b), 2 diagonal block, 3 column rotation of V (row, pair)
  int r0, r1;      // read offsets: (pa,pb) and (qa,pb) [kind 1]; (pa,pa) and (qa,qa) [kind 2]; V[r][2p] [kind 3]
  int w0, w1, w2, w3;  // write offsets in the other buffer
  int ia, ib;      // pair indices whose (c, s) are needed
};
struct Jac2Pair {  // what a pair thread needs to form the NEXT round's pair k = its three new elements
  int a, b;        // old pairs holding the two members i = sigma^-1(2k), j = sigma^-1(2k+1)
  int di, dj;      // their positions (0 top / 1 bottom) inside those pairs
  int oa, ob;      // offsets of the diagonal blocks (pa,pa) of a and b
  int oab0, oab1;  // offsets of the two rows of the off-diagonal block (min(a,b), max(a,b))
  int swap;        // 1 when a > b (the stored block is (b, a): the wanted element sits at the transposed position)
  int same;        // 1 when both members come from the same old pair (only for m == 1: the single pair again)
};

template <typename R>
TNB_HD int jac2_total_items(const Jac2<R>& J) { return J.m * (J.m + 1) / 2 + J.np * J.m; }

template <typename R>
TNB_HD void jac2_make_item(const Jac2<R>& J, int w, Jac2Item& it) {
  const int m = J.m, lds = J.lds, U = m * (m + 1) / 2;
  it.kind = 0;
  it.r0 = it.r1 = it.w0 = it.w1 = it.w2 = it.w3 = it.ia = it.ib = 0;
  if (w < 0 || w >= U + J.np * m) return;
  if (w < U) {
    // row-major enumeration of the upper triangle: base(a) = a*m - a(a-1)/2 blocks precede row a
    int a = 0;
    while ((a + 1) * m - (a + 1) * a / 2 <= w) ++a;
    const int b = a + (w - (a * m - a * (a - 1) / 2));
    const int pa = 2 * a, qa = pa + 1, pb = 2 * b, qb = pb + 1;
    const int spa = jac2_sigma(pa, m), sqa = jac2_sigma(qa, m), spb = jac2_sigma(pb, m), sqb = jac2_sigma(qb, m);
    it.ia = a;
    it.ib = b;
    if (a == b) {
      it.kind = 2;
      it.r0 = pa * lds + pa;                 // (pa,pa), (pa,qa) contiguous
      it.r1 = qa * lds + qa;
      it.w0 = spa * lds + spa;
      it.w1 = sqa * lds + sqa;
      it.w2 = jac2_off(spa, sqa, lds);
      it.w3 = it.w2;
    } else {
      it.kind = 1;
      it.r0 = pa * lds + pb;                 // (pa,pb), (pa,qb) contiguous
      it.r1 = qa * lds + pb;                 // (qa,pb), (qa,qb) contiguous
      it.w0 = jac2_off(spa, spb, lds);
      it.w1 = jac2_off(spa, sqb, lds);
      it.w2 = jac2_off(sqa, spb, lds);
      it.w3 = jac2_off(sqa, sqb, lds);
    }
  } else {
    const int idx = w - U;
    const int r = idx / m, pr = idx - r * m;
    it.kind = 3;
    it.ia = pr;
    it.r0 = r * lds + 2 * pr;                // V[r][2p], V[r][2p+1] contiguous
    it.w0 = r * lds + jac2_sigma(2 * pr, m);
    it.w1 = r * lds + jac2_sigma(2 * pr + 1, m);
  }
}

template <typename R>
TNB_HD void jac2_make_pair(const Jac2<R>& J, int k, Jac2Pair& p) {
  const int m = J.m, lds = J.lds;
  const int i = jac2_sigma_inv(2 * k, m), j = jac2_sigma_inv(2 * k + 1, m);
  p.a = i >> 1; p.b = j >> 1; p.di = i & 1; p.dj = j & 1;
  p.oa = (2 * p.a) * lds + 2 * p.a;
  p.ob = (2 * p.b) * lds + 2 * p.b;
  const int lo = p.a < p.b ? p.a : p.b, hi = p.a < p.b ? p.b : p.a;
  p.swap = p.a > p.b ? 1 : 0;
  p.same = p.a == p.b ? 1 : 0;
  p.oab0 = (2 * lo) * lds + 2 * hi;
  p.oab1 = (2 * lo + 1) * lds + 2 * hi;
}

// One work item of a round: reads buffer `cur` with the rotations cs[cur_cs], writes buffer `cur ^ 1`.
template <typename R>
TNB_HD void jac2_do_item(const Jac2<R>& J, int cur, int cur_cs, const Jac2Item& it) {
  if (it.kind == 0) return;
  const R* cs = J.cs[cur_cs];
  if (it.kind == 3) {
    const R* V = J.V[cur];
    R* Vo = J.V[cur ^ 1];
    const R c = cs[2 * it.ia], s = cs[2 * it.ia + 1];
    const R x = V[it.r0], y = V[it.r0 + 1];
    Vo[it.w0] = c * x - s * y;
    Vo[it.w1] = s * x + c * y;
    return;
  }
  const R* S = J.S[cur];
  R* So = J.S[cur ^ 1];
  const R ca = cs[2 * it.ia], sa = cs[2 * it.ia + 1];
  if (it.kind == 2) {
    // exact similarity for ANY consistent (c, s): S_pq is carried at its true (small) value, not forced to zero
    const R spp = S[it.r0], spq = S[it.r0 + 1], sqq = S[it.r1];
    const R cc = ca * ca, ss = sa * sa, cs2 = (R)2 * ca * sa;
    So[it.w0] = cc * spp - cs2 * spq + ss * sqq;
    So[it.w1] = ss * spp + cs2 * spq + cc * sqq;
    So[it.w2] = ca * sa * (spp - sqq) + (cc - ss) * spq;
    return;
  }
  const R cb = cs[2 * it.ib], sb = cs[2 * it.ib + 1];
  const R b00 = S[it.r0], b01 = S[it.r0 + 1], b10 = S[it.r1], b11 = S[it.r1 + 1];
  // T = B J_b,  J_b = [c s; -s c];   B' = J_a^T T,  J_a^T = [c -s; s c]
  const R t00 = cb * b00 - sb * b01, t01 = sb * b00 + cb * b01;
  const R t10 = cb * b10 - sb * b11, t11 = sb * b10 + cb * b11;
  So[it.w0] = ca * t00 - sa * t10;
  So[it.w1] = ca * t01 - sa * t11;
  So[it.w2] = sa * t00 + ca * t10;
  So[it.w3] = sa * t01 + ca * t11;
}

// Pair thread k: the three elements of the NEXT round's pair k from the old blocks, then its rotation into cs[cur_cs ^ 1].
template <typename R>
TNB_HD void jac2_do_pair(const Jac2<R>& J, int cur, int cur_cs, int k, const Jac2Pair& p, int* flag) {
  const R* S = J.S[cur];
  const R* cs = J.cs[cur_cs];
  R* cso = J.cs[cur_cs ^ 1];
  const R ca = cs[2 * p.a], sa = cs[2 * p.a + 1], cb = cs[2 * p.b], sb = cs[2 * p.b + 1];
  R nii, njj, nij;
  if (p.same) {  // m == 1: the pair meets itself again; all three elements come from its own diagonal block
    const R spp = S[p.oa], spq = S[p.oa + 1], sqq = S[p.oa + J.lds + 1];
    const R cc = ca * ca, ss = sa * sa, cs2 = (R)2 * ca * sa;
    nii = cc * spp - cs2 * spq + ss * sqq;
    njj = ss * spp + cs2 * spq + cc * sqq;
    nij = ca * sa * (spp - sqq) + (cc - ss) * spq;
    R c, s;
    jac2_decide(J, nii, njj, nij, c, s, flag);
    cso[2 * k] = c;
    cso[2 * k + 1] = s;
    return;
  }
  {
    const R spp = S[p.oa], spq = S[p.oa + 1], sqq = S[p.oa + J.lds + 1];
    const R cc = ca * ca, ss = sa * sa, cs2 = (R)2 * ca * sa;
    nii = p.di == 0 ? cc * spp - cs2 * spq + ss * sqq : ss * spp + cs2 * spq + cc * sqq;
  }
  {
    const R spp = S[p.ob], spq = S[p.ob + 1], sqq = S[p.ob + J.lds + 1];
    const R cc = cb * cb, ss = sb * sb, cs2 = (R)2 * cb * sb;
    njj = p.dj == 0 ? cc * spp - cs2 * spq + ss * sqq : ss * spp + cs2 * spq + cc * sqq;
  }
  {
    // stored block is (lo, hi); rows belong to pair lo, columns to pair hi
    const R b00 = S[p.oab0], b01 = S[p.oab0 + 1], b10 = S[p.oab1], b11 = S[p.oab1 + 1];
    const R cl = p.swap ? cb : ca, sl = p.swap ? sb : sa;  // rotation of the row pair (lo)
    const R ch = p.swap ? ca : cb, sh = p.swap ? sa : sb;  // rotation of the column pair (hi)
    const int x = p.swap ? p.dj : p.di;                    // wanted row inside lo
    const int y = p.swap ? p.di : p.dj;                    // wanted column inside hi
    // column y of T = B J_hi
    const R t0 = y == 0 ? ch * b00 - sh * b01 : sh * b00 + ch * b01;
    const R t1 = y == 0 ? ch * b10 - sh * b11 : sh * b10 + ch * b11;
    nij = x == 0 ? cl * t0 - sl * t1 : sl * t0 + cl * t1;
  }
  R c, s;
  jac2_decide(J, nii, njj, nij, c, s, flag);
  cso[2 * k] = c;
  cso[2 * k + 1] = s;
}

// The very first rotations (before round 0): plain read of the pair elements.
template <typename R>
TNB_HD void jac2_first_pair(const Jac2<R>& J, int cur, int cur_cs, int k, int* flag) {
  const R* S = J.S[cur];
  const int p = 2 * k;
  R c, s;
  jac2_decide(J, S[p * J.lds + p], S[(p + 1) * J.lds + p + 1], S[p * J.lds + p + 1], c, s, flag);
  J.cs[cur_cs][2 * k] = c;
  J.cs[cur_cs][2 * k + 1] = s;
}

}  // namespace tnb
