// The n x n decisions of eigsort (reference eigsort.py:66-122, 142-160; focusr.py:459-490) for one (target, source)
// pair, shared by the CUDA kernel (eigsort.cu: one thread per pair) and the host test double (tests/hostsim):
//   c_lambda, c = c_spatial * c_lambda * c_hist and its flipped twin, Q = min(c, c_f), S = c > c_f,
//   scipy.optimize.linear_sum_assignment(Q) (or Q^T when the source is the reference),
//   the flip list, the column moves, and the spectral weights exp(-w^2 / (2 mean(w)^2)), w = Q_pair * max(lambda).
// With n <= ~70 this is microseconds of scalar work; it moved to the device so that the batched pipeline has no host
// visit between the upload of the vertices and the download of the correspondences.
#pragma once
#include <math.h>

#include "rowops.h"   // FB_HD, FB_MUL / FB_DIV (no FMA contraction)

namespace fb {

constexpr int LSAP_MAX = 96;

// scipy.optimize.linear_sum_assignment for a square n x n cost matrix (row-major, finite entries): the shortest
// augmenting path algorithm of Crouse (IEEE TAES 2016) in the form scipy's rectangular_lsap.cpp gives it -- rows are
// added in order, the unscanned columns are visited from the highest index down, a column is the new minimum if its
// reduced cost is lower OR equal and it is still unassigned -- so that ties resolve exactly as in scipy.  Returns 0, or
// -1 for an infeasible matrix.  col4row[i] = column assigned to row i.
FB_HD int lsap_square(const double* cost, int n, int* col4row) {
  double u[LSAP_MAX], v[LSAP_MAX], spc[LSAP_MAX];
  int path[LSAP_MAX], row4col[LSAP_MAX], remaining[LSAP_MAX];
  bool sr[LSAP_MAX], sc[LSAP_MAX];
  for (int i = 0; i < n; ++i) {
    u[i] = 0.0;
    v[i] = 0.0;
    path[i] = -1;
    col4row[i] = -1;
    row4col[i] = -1;
  }
  for (int cur = 0; cur < n; ++cur) {
    double min_val = 0.0;
    int num_remaining = n;
    for (int it = 0; it < n; ++it) {
      remaining[it] = n - it - 1;
      sr[it] = false;
      sc[it] = false;
      spc[it] = INFINITY;
    }
    int sink = -1, i = cur;
    while (sink == -1) {
      int index = -1;
      double lowest = INFINITY;
      sr[i] = true;
      for (int it = 0; it < num_remaining; ++it) {
        const int j = remaining[it];
        const double r = min_val + cost[i * n + j] - u[i] - v[j];
        if (r < spc[j]) {
          path[j] = i;
          spc[j] = r;
        }
        if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) {
          lowest = spc[j];
          index = it;
        }
      }
      min_val = lowest;
      if (!(min_val < INFINITY)) return -1;
      const int j = remaining[index];
      if (row4col[j] == -1)
        sink = j;
      else
        i = row4col[j];
      sc[j] = true;
      remaining[index] = remaining[--num_remaining];
    }
    u[cur] += min_val;
    for (int r = 0; r < n; ++r)
      if (sr[r] && r != cur) u[r] += min_val - spc[col4row[r]];
    for (int j = 0; j < n; ++j)
      if (sc[j]) v[j] -= min_val - spc[j];
    int j = sink;
    for (;;) {
      const int r = path[j];
      row4col[j] = r;
      const int t = col4row[r];
      col4row[r] = j;
      j = t;
      if (r == cur) break;
    }
  }
  return 0;
}

// One pair.  vals_t / vals_s: the eigenvalues each graph returned (nf_t / nf_s of them, ascending); the four cost
// matrices are n x n row-major [target i][source j].  Outputs: q[n] (the cost of each matched pair, in the order of the
// reference's match list), moves of the graph that gets permuted -- new[:, dst[k]] = sign[k] * old[:, src[k]] --, and
// w[ns] (all ones when !weighted).  scratch: 2 n^2 doubles.
FB_HD int eigsort_decide_pair(const double* vals_t, int nf_t, const double* vals_s, int nf_s, const double* c_hist,
                              const double* c_hist_f, const double* c_spatial, const double* c_spatial_f, int n, int ns,
                              bool target_as_reference, bool weighted, double* q, int* dst, int* src, int* sign, double* w,
                              double* scratch) {
  // eigsort.py:142-160: the gap averages over ALL eigenvalues each graph returned
  double gt = 0.0, gs = 0.0;
  for (int i = 1; i < nf_t; ++i) gt += vals_t[i] - vals_t[i - 1];
  for (int i = 1; i < nf_s; ++i) gs += vals_s[i] - vals_s[i - 1];
  const double gap = (gt / (nf_t - 1) + gs / (nf_s - 1)) / 2;
  double* qm = scratch;          // min(c, c_f), laid out for the assignment (transposed if the source is the reference)
  double* sm = scratch + n * n;  // 1.0 where c > c_f  (indexed [t][s])
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      const double dl = vals_t[i] - vals_s[j];
      const double cl = exp(FB_DIV(FB_MUL(dl, dl), FB_MUL(2.0, FB_MUL(gap, gap))));
      const double c = FB_MUL(FB_MUL(c_spatial[i * n + j], cl), c_hist[i * n + j]);
      const double cf = FB_MUL(FB_MUL(c_spatial_f[i * n + j], cl), c_hist_f[i * n + j]);
      const double m = cf < c ? cf : c;
      qm[target_as_reference ? i * n + j : j * n + i] = m;
      sm[i * n + j] = c > cf ? 1.0 : 0.0;
    }
  int assign[LSAP_MAX];
  if (lsap_square(qm, n, assign) != 0) return -1;
  for (int k = 0; k < n; ++k) {
    // target_as_reference: target k <-> source assign[k]; else: source k <-> target assign[k]
    const int t = target_as_reference ? k : assign[k], s = target_as_reference ? assign[k] : k;
    q[k] = qm[k * n + assign[k]];
    dst[k] = target_as_reference ? t : s;
    src[k] = target_as_reference ? s : t;
    sign[k] = sm[t * n + s] != 0.0 ? -1 : 1;
  }
  // focusr.py:481-490
  if (weighted) {
    double mean = 0.0;
    for (int k = 0; k < ns; ++k) {
      w[k] = FB_MUL(q[k], vals_s[k] > vals_t[k] ? vals_s[k] : vals_t[k]);
      mean += w[k];
    }
    mean /= ns;
    for (int k = 0; k < ns; ++k) w[k] = exp(FB_DIV(-FB_MUL(w[k], w[k]), FB_MUL(2.0, FB_MUL(mean, mean))));
  } else {
    for (int k = 0; k < ns; ++k) w[k] = 1.0;
  }
  return 0;
}

}  // namespace fb
