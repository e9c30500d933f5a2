// The small NON-symmetric Rayleigh-Ritz step (b x b, b <= 64) written once for host and device, like dense_small.h.
//
// Open / non-manifold meshes -- the reference's own shipped 15k pair -- have a structurally non-symmetric adjacency
// (reference graph.py:158-178 writes one direction per cell edge), so the block iteration projects in the Euclidean inner
// product and needs the eigen-decomposition of a small general real matrix per mesh and outer iteration: what ARPACK's
// dneupd / dlahqr do inside scipy `eigs` (reference graph.py:372).  Round 1 did this on the host (nonsym_host.hpp): 3.4 ms
// per 48 x 48 problem, 8 outer iterations, three host round trips each -- more than the GPU work of the solve for a single
// pair, and serial over the meshes of a batch.  Here one CTA per mesh does it in shared memory:
//   Cholesky of G, whitening of H, Householder reduction to Hessenberg form, explicit single-shift (Wilkinson) QR in
//   complex arithmetic to a complex Schur form, eigenvectors by back substitution, the driver's selection (real Ritz
//   values below `cut` first, ascending; the rest with complex pairs carried as (Re, Im)), and W = R^-1 Y.
// Same algorithm and same selection rules as nonsym_host.hpp (which stays as the path for b > 64, where the three complex
// b x b arrays no longer fit one SM's shared memory).  Parallel structure: every `par.for_n` body writes data no other
// index of the same loop reads; scalar decisions (deflation, shifts, selection) are taken redundantly by all threads from
// values that are stable between two `par.sync()`.  tests/hostsim runs this very code with the sequential `SeqPar`.
#pragma once
#include "dense_small.h"

namespace fb {

struct Cd {
  double re, im;
};
FB_HD Cd cd_make(double r, double i) {
  Cd c;
  c.re = r;
  c.im = i;
  return c;
}
FB_HD Cd cd_add(Cd a, Cd b) { return cd_make(a.re + b.re, a.im + b.im); }
FB_HD Cd cd_sub(Cd a, Cd b) { return cd_make(a.re - b.re, a.im - b.im); }
FB_HD Cd cd_mul(Cd a, Cd b) { return cd_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
FB_HD Cd cd_scale(Cd a, double s) { return cd_make(a.re * s, a.im * s); }
FB_HD Cd cd_conj(Cd a) { return cd_make(a.re, -a.im); }
FB_HD Cd cd_neg(Cd a) { return cd_make(-a.re, -a.im); }
FB_HD double cd_norm(Cd a) { return a.re * a.re + a.im * a.im; }
FB_HD double cd_abs(Cd a) { return hypot(a.re, a.im); }
FB_HD double cd_abs1(Cd a) { return fabs(a.re) + fabs(a.im); }  // LAPACK's CABS1: the cheap magnitude of the deflation test
FB_HD Cd cd_div(Cd a, Cd b) {  // Smith's algorithm
  if (fabs(b.re) >= fabs(b.im)) {
    const double r = b.im / b.re, d = b.re + b.im * r;
    return cd_make((a.re + a.im * r) / d, (a.im - a.re * r) / d);
  }
  const double r = b.re / b.im, d = b.re * r + b.im;
  return cd_make((a.re * r + a.im) / d, (a.im * r - a.re) / d);
}
FB_HD Cd cd_sqrt(Cd a) {  // principal branch
  if (a.re == 0.0 && a.im == 0.0) return cd_make(0.0, 0.0);
  const double t = sqrt(0.5 * (fabs(a.re) + hypot(a.re, a.im)));
  if (a.re >= 0.0) return cd_make(t, a.im / (2.0 * t));
  return cd_make(fabs(a.im) / (2.0 * t), a.im >= 0.0 ? t : -t);
}

// scratch of rr_nonsym_small: 4 b complex, 2 b doubles, 6 b + 2 ints
FB_HD size_t nonsym_small_scratch_bytes(int b) {
  return sizeof(Cd) * 4 * (size_t)b + sizeof(double) * 2 * (size_t)b + sizeof(int) * (6 * (size_t)b + 2);
}

// Complex Schur form of the complex matrix H (n x n, row-major): H <- T upper triangular, Q <- the unitary factor.
// v, cs, sn: n complex of scratch each.  Returns the number of eigenvalues deflated by force (0 = converged).
// `small` : n ints of scratch (deflation flags).
template <class Par>
FB_HD int schur_complex(Cd* H, Cd* Q, int n, Cd* v, Cd* cs, Cd* sn, int* small_sub, double* hnorm_out, const Par& par) {
#define FB_H(i, j) H[(i) * n + (j)]
  par.for_n(n * n, [&](int e) { Q[e] = cd_make((e / n == e % n) ? 1.0 : 0.0, 0.0); });
  par.sync();
  // --- Householder reduction to Hessenberg form, Q accumulates the reflectors
  for (int k = 0; k + 2 < n; ++k) {
    double acc = 0.0;
    par.for_n(n - k - 1, [&](int t) { acc += cd_norm(FB_H(k + 1 + t, k)); });
    const double nrm = sqrt(par.sum(acc));
    if (nrm == 0.0) continue;
    const Cd x0 = FB_H(k + 1, k);
    const double ax0 = cd_abs(x0);
    const Cd phase = ax0 > 0.0 ? cd_scale(x0, 1.0 / ax0) : cd_make(1.0, 0.0);
    const Cd alpha = cd_scale(phase, -nrm);
    par.sync();  // everyone has read x0 before column k changes hands
    acc = 0.0;
    par.for_n(n - k - 1, [&](int t) {
      const int i = k + 1 + t;
      Cd vi = FB_H(i, k);
      if (i == k + 1) vi = cd_sub(vi, alpha);
      v[i] = vi;
      acc += cd_norm(vi);
    });
    const double vn = sqrt(par.sum(acc));
    if (vn == 0.0) continue;
    const double vinv = 1.0 / vn;
    par.for_n(n - k - 1, [&](int t) { v[k + 1 + t] = cd_scale(v[k + 1 + t], vinv); });
    par.sync();
    // H <- (I - 2 v v^H) H : a thread owns a column
    par.for_n(n, [&](int j) {
      Cd s = cd_make(0.0, 0.0);
      for (int i = k + 1; i < n; ++i) s = cd_add(s, cd_mul(cd_conj(v[i]), FB_H(i, j)));
      s = cd_scale(s, 2.0);
      for (int i = k + 1; i < n; ++i) FB_H(i, j) = cd_sub(FB_H(i, j), cd_mul(v[i], s));
    });
    par.sync();
    // H <- H (I - 2 v v^H), Q <- Q (I - 2 v v^H) : a thread owns a row of H or of Q
    par.for_n(2 * n, [&](int t) {
      Cd* row = t < n ? H + (size_t)t * n : Q + (size_t)(t - n) * n;
      Cd s = cd_make(0.0, 0.0);
      for (int j = k + 1; j < n; ++j) s = cd_add(s, cd_mul(row[j], v[j]));
      s = cd_scale(s, 2.0);
      for (int j = k + 1; j < n; ++j) row[j] = cd_sub(row[j], cd_mul(s, cd_conj(v[j])));
    });
    par.sync();
    par.for_n(n - k - 2, [&](int t) { FB_H(k + 2 + t, k) = cd_make(0.0, 0.0); });
    par.sync();
  }
  // --- shifted QR.  hnorm: Frobenius norm (the scale of "zero")
  const double eps = 2.220446049250313e-16;
  double acc = 0.0;
  par.for_n(n * n, [&](int e) { acc += cd_norm(H[e]); });
  double hnorm = sqrt(par.sum(acc));
  if (hnorm == 0.0) hnorm = 1.0;
  *hnorm_out = hnorm;
  int failed = 0, ihi = n - 1, iter = 0;
  while (ihi > 0) {
    par.sync();
    // deflation test of every subdiagonal entry of the active window at once (zlahqr's criterion with CABS1: a hypot
    // per entry and thread, scanned serially, was most of the kernel's time), then the scan over the flags
    par.for_n(ihi, [&](int t) {
      const int l2 = t + 1;
      double s = cd_abs1(FB_H(l2 - 1, l2 - 1)) + cd_abs1(FB_H(l2, l2));
      if (s == 0.0) s = hnorm;
      small_sub[l2] = cd_abs1(FB_H(l2, l2 - 1)) <= eps * s;
    });
    par.sync();
    int l = ihi;
    while (l > 0 && !small_sub[l]) --l;
    bool force = false;
    if (l != ihi && ++iter > 60) {  // give up on this eigenvalue, deflate by force
      ++failed;
      force = true;
    }
    par.sync();  // every thread has finished its scan
    if (par.lane() == 0) {
      if (l > 0) FB_H(l, l - 1) = cd_make(0.0, 0.0);
      if (force) FB_H(ihi, ihi - 1) = cd_make(0.0, 0.0);
    }
    par.sync();  // ... and sees the deflated entries before it forms the shift
    if (l == ihi || force) {
      --ihi;
      iter = 0;
      continue;
    }
    Cd shift;
    if (iter == 10 || iter == 20 || iter == 30) {
      shift = cd_add(FB_H(ihi, ihi),
                     cd_make(cd_abs(FB_H(ihi, ihi - 1)) + (ihi > 1 ? cd_abs(FB_H(ihi - 1, ihi - 2)) : 0.0), 0.0));
    } else {
      const Cd aa = FB_H(ihi - 1, ihi - 1), bb = FB_H(ihi - 1, ihi), cc = FB_H(ihi, ihi - 1), dd = FB_H(ihi, ihi);
      const Cd half = cd_scale(cd_add(aa, dd), 0.5);
      const Cd dif = cd_sub(aa, dd);
      const Cd disc = cd_sqrt(cd_add(cd_scale(cd_mul(dif, dif), 0.25), cd_mul(bb, cc)));
      const Cd m1 = cd_add(half, disc), m2 = cd_sub(half, disc);
      shift = (cd_abs(cd_sub(m1, dd)) < cd_abs(cd_sub(m2, dd))) ? m1 : m2;
    }
    par.sync();  // the shift has been formed by everyone from the unshifted matrix
    par.for_n(ihi - l + 1, [&](int t) { FB_H(l + t, l + t) = cd_sub(FB_H(l + t, l + t), shift); });
    par.sync();
    // Givens rotations from the left, one after the other; a thread owns a column j > k.  Column k itself becomes
    // (r, 0) analytically and is written after the sweep, so nobody writes what the others are still reading.
    for (int k = l; k < ihi; ++k) {
      const Cd x = FB_H(k, k), y = FB_H(k + 1, k);
      const double r2 = cd_norm(x) + cd_norm(y);
      const double r = sqrt(r2);
      Cd c = cd_make(1.0, 0.0), s = cd_make(0.0, 0.0);
      if (r > 0.0) {
        const double rinv = 1.0 / r;
        c = cd_scale(x, rinv);
        s = cd_scale(y, rinv);
      }
      if (par.lane() == 0) {
        cs[k] = c;
        sn[k] = s;
        v[k] = cd_make(r > 0.0 ? r : x.re, r > 0.0 ? 0.0 : x.im);
      }
      par.for_n(n - k - 1, [&](int t) {
        const int j = k + 1 + t;
        const Cd t0 = FB_H(k, j), t1 = FB_H(k + 1, j);
        FB_H(k, j) = cd_add(cd_mul(cd_conj(c), t0), cd_mul(cd_conj(s), t1));
        FB_H(k + 1, j) = cd_add(cd_mul(cd_neg(s), t0), cd_mul(c, t1));
      });
      par.sync();
    }
    par.for_n(ihi - l, [&](int t) {
      const int k = l + t;
      FB_H(k, k) = v[k];
      FB_H(k + 1, k) = cd_make(0.0, 0.0);
    });
    par.sync();
    // the same rotations from the right: a thread owns a row of H (rows <= ihi) or of Q and walks along it
    par.for_n(2 * n, [&](int t) {
      const bool is_h = t < n;
      const int i = is_h ? t : t - n;
      if (is_h && i > ihi) return;
      Cd* row = is_h ? H + (size_t)i * n : Q + (size_t)i * n;
      int k0 = l;
      if (is_h && i - 2 > k0) k0 = i - 2;
      for (int k = k0; k < ihi; ++k) {
        const Cd c = cs[k], s = sn[k];
        const Cd t0 = row[k], t1 = row[k + 1];
        row[k] = cd_add(cd_mul(t0, c), cd_mul(t1, s));
        row[k + 1] = cd_add(cd_mul(cd_neg(t0), cd_conj(s)), cd_mul(t1, cd_conj(c)));
      }
    });
    par.sync();
    par.for_n(ihi - l + 1, [&](int t) { FB_H(l + t, l + t) = cd_add(FB_H(l + t, l + t), shift); });
  }
  par.sync();
#undef FB_H
  return failed;
}

// g (b x b, = X^T X) and h (b x b, = X^T L X), row-major, contiguous (h == g + b*b); both are destroyed.
// Hc, Qc: b*b complex each.  r_save: b*b doubles that survive (the Cholesky factor is parked there; may be global memory).
// The b*b complex array of back-substituted vectors aliases (g, h).  Writes w_out (b x b: columns = real basis of Ritz
// vectors), theta_out (b), *n_low_out; returns the number of clamped Cholesky pivots, or -1 if the QR iteration failed.
template <class Par>
FB_HD int rr_nonsym_small(double* g, double* h, Cd* Hc, Cd* Qc, double* r_save, int b, double cut, double* w_out,
                          double* theta_out, int* n_low_out, void* scratch, const Par& par) {
  Cd* v = static_cast<Cd*>(scratch);
  Cd* cs = v + b;
  Cd* sn = cs + b;
  Cd* ev = sn + b;
  double* diag0 = reinterpret_cast<double*>(ev + b);
  double* theta = diag0 + b;
  int* is_real = reinterpret_cast<int*>(theta + b);
  int* low = is_real + b;
  int* rest = low + b;
  int* used = rest + b;
  int* sel_idx = used + b;
  int* sel_im = sel_idx + b;
  int* counts = sel_im + b;  // [0] = n_low, [1] = n_rest
  const int n = b;
  const int bad = cholesky_upper(g, diag0, b, par);
  // h <- R^-T h R^-1 without symmetrising: forward substitution down every column, then along every row
  par.for_n(b, [&](int j) {
    for (int i = 0; i < b; ++i) {
      double val = h[i * b + j];
      for (int k = 0; k < i; ++k) val -= g[k * b + i] * h[k * b + j];
      h[i * b + j] = val / g[i * b + i];
    }
  });
  par.sync();
  par.for_n(b, [&](int i) {
    for (int j = 0; j < b; ++j) {
      double val = h[i * b + j];
      for (int k = 0; k < j; ++k) val -= h[i * b + k] * g[k * b + j];
      h[i * b + j] = val / g[j * b + j];
    }
  });
  par.sync();
  par.for_n(b * b, [&](int e) {
    r_save[e] = g[e];
    Hc[e] = cd_make(h[e], 0.0);
  });
  par.sync();
  double hnorm = 1.0;
  const int failed = schur_complex(Hc, Qc, n, v, cs, sn, is_real, &hnorm, par);
  if (failed != 0) return -1;
  // --- eigenvectors of the triangular factor (a thread owns an eigenvalue; vector k lives in column k of Y)
  Cd* Y = reinterpret_cast<Cd*>(g);
  const double small = 2.220446049250313e-16 * hnorm;
  par.for_n(n, [&](int k) {
    const Cd lam = Hc[k * n + k];
    ev[k] = lam;
    Y[k * n + k] = cd_make(1.0, 0.0);
    for (int i = k - 1; i >= 0; --i) {
      Cd s = cd_make(0.0, 0.0);
      for (int j = i + 1; j <= k; ++j) s = cd_add(s, cd_mul(Hc[i * n + j], Y[j * n + k]));
      Cd d = cd_sub(Hc[i * n + i], lam);
      if (cd_abs(d) < small) d = cd_make(small, 0.0);
      Y[i * n + k] = cd_div(cd_neg(s), d);
    }
  });
  par.sync();
  // back-transformed: E = Q Y, stored over the (now dead) triangular factor, columns normalised
  Cd* E = Hc;
  par.for_n(n * n, [&](int e) {
    const int i = e / n, k = e % n;
    Cd s = cd_make(0.0, 0.0);
    for (int j = 0; j <= k; ++j) s = cd_add(s, cd_mul(Qc[i * n + j], Y[j * n + k]));
    E[e] = s;
  });
  par.sync();
  // --- selection (the rules of rr_nonsym_host): scalar work, one thread
  if (par.lane() == 0) {
    int nl = 0, nr = 0;
    for (int i = 0; i < n; ++i) {
      const double mag = cd_abs(ev[i]);
      const double lim = 1e-8 * mag > 1e-13 ? 1e-8 * mag : 1e-13;
      is_real[i] = fabs(ev[i].im) <= lim;
      if (is_real[i] && ev[i].re <= cut)
        low[nl++] = i;
      else
        rest[nr++] = i;
      used[i] = 0;
    }
    for (int a = 1; a < nl; ++a) {  // ascending real part
      const int x = low[a];
      int p = a - 1;
      while (p >= 0 && ev[low[p]].re > ev[x].re) {
        low[p + 1] = low[p];
        --p;
      }
      low[p + 1] = x;
    }
    for (int a = 1; a < nr; ++a) {  // ascending real part, then descending imaginary part
      const int x = rest[a];
      int p = a - 1;
      while (p >= 0 && (ev[rest[p]].re > ev[x].re || (ev[rest[p]].re == ev[x].re && ev[rest[p]].im < ev[x].im))) {
        rest[p + 1] = rest[p];
        --p;
      }
      rest[p + 1] = x;
    }
    int col = 0;
    for (int a = 0; a < nl; ++a) {
      sel_idx[col] = low[a];
      sel_im[col] = 0;
      used[low[a]] = 1;
      ++col;
    }
    counts[0] = col;
    for (int a = 0; a < nr; ++a) {
      const int idx = rest[a];
      if (used[idx] || col >= n) continue;
      used[idx] = 1;
      if (is_real[idx]) {
        sel_idx[col] = idx;
        sel_im[col] = 0;
        ++col;
      } else {
        // conjugate partner: closest unused complex eigenvalue to conj(ev[idx])
        int partner = -1;
        double best = 1e300;
        for (int c2 = 0; c2 < nr; ++c2) {
          const int j = rest[c2];
          if (!used[j] && !is_real[j]) {
            const double d = cd_abs(cd_sub(ev[j], cd_conj(ev[idx])));
            if (d < best) {
              best = d;
              partner = j;
            }
          }
        }
        sel_idx[col] = idx;
        sel_im[col] = 0;
        ++col;
        if (col < n && partner >= 0 && best <= 1e-6 * cd_abs(ev[idx])) {
          used[partner] = 1;
          sel_idx[col] = idx;
          sel_im[col] = 1;
          ++col;
        }
      }
    }
    for (int idx = 0; idx < n && col < n; ++idx)  // numerical leftovers (unpaired complex values)
      if (!used[idx]) {
        used[idx] = 1;
        sel_idx[col] = idx;
        sel_im[col] = 0;
        ++col;
      }
    counts[1] = col;
  }
  par.sync();
  // --- yr (over the dead Q): column c = Re or Im of eigenvector sel_idx[c], unit 2-norm; then w = R^-1 yr
  double* yr = reinterpret_cast<double*>(Qc);
  double* w = yr + (size_t)b * b;
  const int ncol = counts[1];
  par.for_n(n, [&](int c) {
    if (c >= ncol) {
      for (int i = 0; i < n; ++i) yr[i * n + c] = (i == c) ? 1.0 : 0.0;
      theta[c] = 0.0;
      return;
    }
    const int idx = sel_idx[c];
    const bool im = sel_im[c] != 0;
    double nrm = 0.0;
    for (int i = 0; i < n; ++i) {
      const double val = im ? E[i * n + idx].im : E[i * n + idx].re;
      nrm += val * val;
    }
    nrm = sqrt(nrm);
    const double inv = nrm > 0.0 ? 1.0 / nrm : 1.0;
    for (int i = 0; i < n; ++i) yr[i * n + c] = (im ? E[i * n + idx].im : E[i * n + idx].re) * inv;
    theta[c] = ev[idx].re;
  });
  par.sync();
  par.for_n(n, [&](int j) {
    for (int i = n - 1; i >= 0; --i) {
      double val = yr[i * n + j];
      for (int k = i + 1; k < n; ++k) val -= r_save[i * n + k] * w[k * n + j];
      w[i * n + j] = val / r_save[i * n + i];
    }
  });
  par.sync();
  par.for_n(n * n, [&](int e) { w_out[e] = w[e]; });
  par.for_n(n, [&](int j) { theta_out[j] = theta[j]; });
  if (par.lane() == 0) *n_low_out = counts[0];
  par.sync();
  return bad;
}

}  // namespace fb
