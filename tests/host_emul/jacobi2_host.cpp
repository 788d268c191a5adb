// Host emulation of the two-sided Jacobi work functions (tntorch_b200/csrc/jacobi2_core.h): the "threads" of a CTA are
// looped between the barriers of a round, exactly as jacobi2.cuh::jac2_solve schedules them (pair threads + workers with
// static work items).  Checks the round-robin relabelling (every pair met exactly once per sweep), the eigen-decomposition
// through the residual ||G V - V diag(w)||, orthogonality, and odd / tiny / rank-deficient sizes.
// Built and run by tests/test_host_emul.py (g++, no CUDA).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <set>
#include <utility>
#include "../../tntorch_b200/csrc/jacobi2_core.h"

using namespace tnb;

static unsigned long long rs = 88172645463325252ULL;
static double urand() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (double)(rs >> 11) / 9007199254740992.0 - 0.5; }

template <typename R>
static int run_case(int n, int nthreads, double tol, double want_resid, int rank_def) {
  const int np = n + (n & 1), m = np / 2, lds = np + 2;
  std::vector<double> G((size_t)n * n, 0.0);
  {
    const int k = rank_def ? n / 2 + 1 : n + 3;
    std::vector<double> A((size_t)k * n);
    for (auto& a : A) a = urand();
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int t = 0; t < k; ++t) s += A[(size_t)t * n + i] * A[(size_t)t * n + j];
        G[(size_t)i * n + j] = s;
      }
  }
  std::vector<R> S0((size_t)np * lds, 0), S1((size_t)np * lds, 0), V0((size_t)np * lds, 0), V1((size_t)np * lds, 0), c0(2 * m), c1(2 * m);
  int flag[2] = {0, 0};
  double gmax = 0;
  for (int i = 0; i < n; ++i) gmax = std::fmax(gmax, std::fabs(G[(size_t)i * n + i]));
  if (gmax == 0) gmax = 1;
  for (int i = 0; i < np; ++i)
    for (int j = 0; j < np; ++j) {
      if (j >= i) S0[i * lds + j] = (i < n && j < n) ? (R)(0.5 * (G[(size_t)i * n + j] + G[(size_t)j * n + i]) / gmax) : (R)0;
      V0[i * lds + j] = (i == j) ? (R)1 : (R)0;
    }
  Jac2<R> J;
  J.S[0] = S0.data(); J.S[1] = S1.data(); J.V[0] = V0.data(); J.V[1] = V1.data();
  J.cs[0] = c0.data(); J.cs[1] = c1.data(); J.flag = flag; J.np = np; J.m = m; J.lds = lds;
  J.tol2 = (R)(tol * tol); J.big2 = (R)tol; J.floor_abs = (R)(sizeof(R) == 8 ? 2.3e-16 : 1.2e-7);
  // thread roles as in jac2_solve
  const int pw = (m + 31) / 32 * 32;
  const int workers = nthreads - pw;
  if (workers < 1) { printf("FAIL: no workers\n"); return 1; }
  const int total = jac2_total_items(J);
  const int per = (total + workers - 1) / workers;
  std::vector<std::vector<Jac2Item>> items(workers, std::vector<Jac2Item>(per));
  for (int w = 0; w < workers; ++w)
    for (int q = 0; q < per; ++q) jac2_make_item(J, w + q * workers, items[w][q]);
  std::vector<Jac2Pair> pairs(m);
  for (int k = 0; k < m; ++k) jac2_make_pair(J, k, pairs[k]);
  std::vector<int> label(np);
  for (int i = 0; i < np; ++i) label[i] = i;
  std::set<std::pair<int, int>> met;
  int cur = 0, ccs = 0, sweeps = 0;
  for (int k = 0; k < m; ++k) jac2_first_pair(J, 0, 0, k, &flag[0]);
  bool conv = false;
  const int rounds = np > 2 ? np - 1 : 1;
  for (; sweeps < 40 && !conv; ++sweeps) {
    flag[(sweeps + 1) & 1] = 0;
    for (int round = 0; round < rounds; ++round) {
      if (sweeps == 0)
        for (int k = 0; k < m; ++k) {
          int a = label[2 * k], b = label[2 * k + 1];
          if (a > b) std::swap(a, b);
          if (!met.insert({a, b}).second) { printf("FAIL n=%d: pair (%d,%d) met twice in a sweep\n", n, a, b); return 1; }
        }
      // pair threads and workers run concurrently on the device: neither reads what the other writes in this round
      for (int k = 0; k < m; ++k) {
        jac2_do_pair(J, cur, ccs, k, pairs[k], &flag[sweeps & 1]);
      }
      for (int w = 0; w < workers; ++w)
        for (int q = 0; q < per; ++q) jac2_do_item(J, cur, ccs, items[w][q]);
      cur ^= 1;
      ccs ^= 1;
      std::vector<int> nl(np);
      for (int i = 0; i < np; ++i) nl[jac2_sigma(i, m)] = label[i];
      label = nl;
    }
    if (sweeps == 0 && (int)met.size() != np * (np - 1) / 2) { printf("FAIL n=%d: %zu pairs of %d\n", n, met.size(), np * (np - 1) / 2); return 1; }
    conv = flag[sweeps & 1] == 0;
  }
  const R* S = J.S[cur];
  const R* V = J.V[cur];
  double resid = 0, orth = 0, off = 0;
  for (int j = 0; j < np; ++j) {
    if (np != n && std::fabs((double)V[n * lds + j]) > 0.5) continue;  // pad column
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int c = 0; c < n; ++c) s += G[(size_t)i * n + c] / gmax * (double)V[c * lds + j];
      resid = std::fmax(resid, std::fabs(s - (double)S[j * lds + j] * (double)V[i * lds + j]));
    }
    for (int k = 0; k < np; ++k) {
      double s = 0;
      for (int c = 0; c < np; ++c) s += (double)V[c * lds + j] * (double)V[c * lds + k];
      orth = std::fmax(orth, std::fabs(s - (j == k ? 1.0 : 0.0)));
      if (k > j) off = std::fmax(off, std::fabs((double)S[j * lds + k]));
    }
  }
  const bool ok = resid <= want_resid && orth <= want_resid && conv;
  printf("%s n=%3d R=%s threads=%4d items/worker=%d sweeps=%2d resid=%.2e orth=%.2e offdiag=%.2e%s\n", ok ? "ok  " : "FAIL", n,
         sizeof(R) == 8 ? "f64" : "f32", nthreads, per, sweeps, resid, orth, off, rank_def ? " (rank deficient)" : "");
  return ok ? 0 : 1;
}

int main() {
  int bad = 0;
  for (int n : {1, 2, 3, 4, 5, 8, 17, 32, 37, 64}) bad += run_case<double>(n, n <= 8 ? 64 : 256, 1e-14, 1e-12, 0);
  bad += run_case<double>(64, 1024, 1e-14, 1e-12, 1);
  bad += run_case<double>(33, 1024, 1e-14, 1e-12, 1);
  bad += run_case<double>(80, 1024, 1e-14, 1e-12, 0);
  for (int n : {2, 7, 32, 64, 96, 112}) bad += run_case<float>(n, 1024, 2e-6, 3e-5, 0);
  bad += run_case<float>(64, 512, 2e-5, 1e-4, 1);
  printf(bad ? "FAILED %d\n" : "all ok\n", bad);
  return bad ? 1 : 0;
}
