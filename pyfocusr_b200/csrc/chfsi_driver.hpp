// Host driver of the block eigen-solver that replaces scipy `eigs(L, k, sigma=1e-10, which="LM",
// ncv=4k)` + the retry logic of `recursive_eig` (reference graph.py:357-389) on the B200.
//
// Method: Chebyshev-filtered subspace iteration (ChFSI) on L = D~^-1 (D - A) itself, no
// factorisation.  One outer iteration =
//     Z = (D - A) X                       (CSR SpMM, GPU)
//     G = X^T g X,  H = X^T h Z           (tall-skinny Gram, GPU, FP64 tensor cores)
//     Rayleigh-Ritz on (G, H) -> W, theta (b x b: GPU warp-Jacobi if A is symmetric,
//                                          host complex-Schur if not, see nonsym_host.hpp)
//     X <- X W, residuals                 (GPU)
//     X <- p_m(L) X                       (m fused SpMM+axpby Chebyshev steps, GPU: the hot loop)
// with (g, h) = (D~, 1) for a symmetric adjacency (L is self-adjoint in the D~ inner product, so
// H is symmetric) and (1, D~^-1) otherwise (Euclidean projection, H general).
// The filter damps [a, beta], a = largest Ritz value of the block, beta = 2 (Gershgorin bound of
// the random-walk Laplacian).  Zero-degree rows (unreferenced vertices) are exact null vectors
// e_i; they are pinned to zero in X (an invariant of every step) and accounted for analytically
// in the retry count, reproducing `k_final` of SURVEY.md section 7.3-2.
//
// Mixed precision: rounding in a filter step perturbs the block by eps relative to its current size, and what
// matters is the part of that perturbation outside the wanted subspace, which the remaining steps do not
// amplify -- so a pass in fp32 can bring residuals down to ~1e-6 (measured floor) but not below.  A pass whose
// predicted landing stays above `lowp_floor` is therefore run with fp32 blocks (47% fewer bytes per step, with the
// fp32 copy of the matrix); the Rayleigh-Ritz steps and every residual that is tested stay fp64.
// Below that floor fp32 still works in CORRECTION form: with (x_j, theta_j) a Ritz pair and r_j = L x_j - theta_j x_j
// its fp64 residual, p(L) x_j = x_j + z with z_{k+1} = alpha_kj ((L - c) z_k + r_j) - gamma_kj z_{k-1}, z_0 = 0, when the
// polynomial of column j is normalised to 1 at theta_j.  z is as small as the error of x_j, so fp32 rounding in z is
// relative to that error, not to x_j: the pass behaves like the fp64 one down to LOWP_CORR_NOISE times the residual it
// started from.  Only x_j, r_j and the final x_j + z are fp64.  16 b N vector bytes per step instead of 24 b N.
// The driver picks the form per pass (PassKind below); a batch takes one form per pass, the common denominator of its meshes.
//
// This header is pure C++ (no CUDA types): the N-sized work is behind the `Backend` concept.
// The product backend is CudaBackend (eigs.cu); tests/hostsim has a plain-loop backend used
// only to exercise this driver logic on machines without a GPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

#include "dense_small.h"
#include "nonsym_host.hpp"

namespace fb {

enum SolveStatus {
  SOLVE_OK = 0,
  SOLVE_NOT_CONVERGED = 1,
  SOLVE_BLOCK_TOO_SMALL = 2,
  SOLVE_BREAKDOWN = 3,
  SOLVE_OUTPUT_TOO_SMALL = 4
};

struct SolveParams {
  int k0;            // `k` of recursive_eig (graph.py:245: n_spectral_features + 1)
  int n_needed;      // `n_k_needed`
  int k_buffer;      // `k_buffer`
  double min_eig;    // MIN_EIG_VAL = 1e-10 (graph.py:369)
  double tol;        // residual ||L v - theta v||_2 / ||v||_2 of every returned pair
  int max_outer;     // outer (Rayleigh-Ritz) iterations
  double amp_target; // filter amplification of the slowest wanted pair per outer iteration
  int max_degree;    // cap of the Chebyshev degree per outer iteration
  double beta;       // guaranteed upper bound of the spectrum (Gershgorin: 2)
  int ldv;           // capacity (columns) of the output arrays
  int probe_degree;  // > 0: tighten beta per mesh with a top-of-spectrum probe of this many filter steps
                     // (symmetric batches only); 0: filter up to `beta` as given
  double land;       // the sized pass aims at land * tol
  double lowp_floor; // > 0: fp32 filter passes are allowed (on a backend that offers them; symmetric and non-symmetric
                     // adjacencies alike: the shipped open meshes converge in the same number of steps); a pass on
                     // fp32 blocks must be predicted to leave every residual above this value.  0: fp64 only
  double lowp_aim;   // residual a sized plain-fp32 pass aims at (>= lowp_floor), from where the fp32 correction
                     // form can reach the tolerance
};

struct MeshResult {
  int status;
  int n_out;       // eigenpairs written (theta > min_eig among the k_final smallest)
  int k_final;     // the k the reference's recursion would have ended at
  int outer_iters;
  int total_degree;
  int block;
  double max_residual;
  double beta;     // upper edge of the filter interval that was used
  int lowp_degree; // filter steps (of total_degree, plus the probe) that ran in fp32
};

enum PassKind { PASS_FP64 = 0, PASS_FP32 = 1, PASS_FP32_CORR = 2 };

// Noise floor of an fp32 correction pass relative to the residual it starts from (measured 1.5e-5 on 15k-vertex
// meshes: 2.5e-6 -> 3.7e-11 of rounding noise next to the 2.8e-11 the polynomial leaves), with a margin.
constexpr double LOWP_CORR_NOISE = 2.3e-5;

inline void cheb_table(double a, double a_low, double beta, int m, double* alpha, double* gamma,
                       double* center) {
  const double e = 0.5 * (beta - a), c = 0.5 * (beta + a);
  double sigma = e / (a_low - c);
  const double sigma1 = sigma;
  alpha[0] = sigma1 / e;
  gamma[0] = 0.0;
  for (int s = 1; s < m; ++s) {
    const double sigma2 = 1.0 / (2.0 / sigma1 - sigma);
    alpha[s] = 2.0 * sigma2 / e;
    gamma[s] = sigma * sigma2;
    sigma = sigma2;
  }
  *center = c;
}

// Column j of the correction pass: cheb_table(a, min(theta_j, a), beta) step by step, handed to `put(step, alpha,
// gamma)`.  Same operations in the same order as cheb_table, so host and device tables are bit-identical.
template <class Put>
FB_HD void corr_table_column(double a, double theta_j, double beta, int m, Put put) {
  const double a_low = theta_j < a ? theta_j : a;
  const double e = 0.5 * (beta - a), c = 0.5 * (beta + a);
  double sigma = e / (a_low - c);
  const double sigma1 = sigma;
  put(0, sigma1 / e, 0.0);
  for (int s = 1; s < m; ++s) {
    const double sigma2 = 1.0 / (2.0 / sigma1 - sigma);
    put(s, 2.0 * sigma2 / e, sigma * sigma2);
    sigma = sigma2;
  }
}

inline int cheb_degree(double a, double theta_k, double beta, double amp, int max_degree) {
  const double e = 0.5 * (beta - a);
  const double eps = std::max((a - theta_k) / e, 1e-14);
  const double rate = std::acosh(1.0 + eps);
  double m = std::ceil(std::log(2.0 * amp) / rate);
  if (!(m < (double)max_degree)) m = max_degree;
  return std::max(4, (int)m);
}

// Non-symmetric Rayleigh-Ritz on the host.  g = X^T X, h = X^T L X (b x b row-major, both
// overwritten).  Produces w (b x b; columns = real basis of Ritz vectors: the `n_low` real Ritz
// values below `cut` first, ascending, then the rest with complex pairs as (Re, Im)) and theta
// (real parts, same order).  Returns #clamped Cholesky pivots, or -1 if the QR iteration failed.
inline int rr_nonsym_host(double* g, double* h, int b, double cut, double* w, double* theta,
                          int* n_low_out) {
  SeqPar par;
  std::vector<double> diag0(b);
  const int bad = cholesky_upper(g, diag0.data(), b, par);
  // h <- R^-T h R^-1 without symmetrising
  for (int i = 0; i < b; ++i) {
    const double dinv = 1.0 / g[i * b + i];
    for (int j = 0; j < b; ++j) {
      double v = h[i * b + j];
      for (int k = 0; k < i; ++k) v -= g[k * b + i] * h[k * b + j];
      h[i * b + j] = v * dinv;
    }
  }
  for (int j = 0; j < b; ++j) {
    const double dinv = 1.0 / g[j * b + j];
    for (int i = 0; i < b; ++i) {
      double v = h[i * b + j];
      for (int k = 0; k < j; ++k) v -= h[i * b + k] * g[k * b + j];
      h[i * b + j] = v * dinv;
    }
  }
  std::vector<cplx> ev(b), evec((size_t)b * b);
  if (eig_general(h, b, ev.data(), evec.data()) != 0) return -1;

  std::vector<int> low, rest;
  std::vector<char> is_real(b);
  for (int i = 0; i < b; ++i) {
    const double mag = std::abs(ev[i]);
    is_real[i] = std::fabs(ev[i].imag()) <= std::max(1e-8 * mag, 1e-13);
    if (is_real[i] && ev[i].real() <= cut)
      low.push_back(i);
    else
      rest.push_back(i);
  }
  std::sort(low.begin(), low.end(), [&](int x, int y) { return ev[x].real() < ev[y].real(); });
  std::sort(rest.begin(), rest.end(), [&](int x, int y) {
    if (ev[x].real() != ev[y].real()) return ev[x].real() < ev[y].real();
    return ev[x].imag() > ev[y].imag();
  });
  std::vector<double> yr((size_t)b * b, 0.0);
  int col = 0;
  auto put = [&](int idx, bool imag_part) {
    double nrm = 0.0;
    for (int i = 0; i < b; ++i) {
      const double v = imag_part ? evec[i * b + idx].imag() : evec[i * b + idx].real();
      yr[i * b + col] = v;
      nrm += v * v;
    }
    nrm = std::sqrt(nrm);
    if (nrm > 0.0)
      for (int i = 0; i < b; ++i) yr[i * b + col] /= nrm;
    theta[col] = ev[idx].real();
    ++col;
  };
  for (int idx : low) put(idx, false);
  *n_low_out = col;
  std::vector<char> used(b, 0);
  for (int idx : low) used[idx] = 1;
  for (int idx : rest) {
    if (used[idx] || col >= b) continue;
    used[idx] = 1;
    if (is_real[idx]) {
      put(idx, false);
    } else {
      // conjugate partner: closest unused eigenvalue to conj(ev[idx])
      int partner = -1;
      double best = 1e300;
      for (int j : rest)
        if (!used[j] && !is_real[j]) {
          const double d = std::abs(ev[j] - std::conj(ev[idx]));
          if (d < best) {
            best = d;
            partner = j;
          }
        }
      put(idx, false);
      if (col < b && partner >= 0 && best <= 1e-6 * std::abs(ev[idx])) {
        used[partner] = 1;
        put(idx, true);
      }
    }
  }
  // numerical leftovers (unpaired complex values): fill with remaining real parts
  for (int idx = 0; idx < b && col < b; ++idx)
    if (!used[idx]) {
      used[idx] = 1;
      put(idx, false);
    }
  // w = R^-1 yr
  for (int j = 0; j < b; ++j)
    for (int i = b - 1; i >= 0; --i) {
      double v = yr[i * b + j];
      for (int k = i + 1; k < b; ++k) v -= g[i * b + k] * w[k * b + j];
      w[i * b + j] = v / g[i * b + i];
    }
  return bad;
}

// Retry contract of recursive_eig (graph.py:374-379) given the pinned operator's Ritz values
// (ascending) and `zr` analytically known zero eigenvalues.  Returns the number `kp` of pinned
// pairs that must be converged, and k_final through *k_final; kp > n_avail means "need a bigger
// block".
inline int retry_contract(const double* theta, int n_avail, int zr, const SolveParams& p,
                          int* k_final) {
  int k_cur = p.k0;
  for (;;) {
    const int kp = k_cur - std::min(zr, k_cur);
    if (kp > n_avail) {
      *k_final = k_cur;
      return kp;
    }
    int good = 0;
    for (int j = 0; j < kp; ++j) good += theta[j] > p.min_eig;
    if (good >= p.n_needed) {
      *k_final = k_cur;
      return kp;
    }
    k_cur += p.k_buffer + p.n_needed;
  }
}

// Top-of-spectrum probe (symmetric adjacency).  The Gershgorin bound 2 is only attained by bipartite
// graphs; triangle meshes sit near 1.5, and the filter degree scales with sqrt(beta).  The start block
// is taken through `probe_degree` steps of the polynomial that damps [0, 1] (lambda_max > 1 always:
// trace(L)/N ~ 1), then Rayleigh-Ritz gives the largest Ritz value theta_max <= lambda_max and the
// residual r of its vector; some eigenvalue lies within r of theta_max, and with the dense top end of a
// mesh spectrum theta_max + r has covered lambda_max on every mesh tried (the standard safeguarded bound
// of Chebyshev-filtered iterations).  beta_m = min(beta, theta_max + 1.5 r + 0.01).  An underestimate is
// not silent: what the filter then amplifies shows up as Ritz values near beta_m and the driver falls
// back to `beta` for that mesh (see `leak` below).
template <class BE>
void probe_upper_bound(BE& be, const SolveParams& p, double* beta_m) {
  const int M = be.n_meshes();
  const int B = be.block();
  const int deg = p.probe_degree;
  std::vector<double> alpha((size_t)M * deg), gamma((size_t)M * deg), center(M);
  be.init_block();
  for (int m = 0; m < M; ++m)
    cheb_table(0.0, p.beta, 1.0, deg, &alpha[(size_t)m * deg], &gamma[(size_t)m * deg], &center[m]);
  be.filter(deg, alpha.data(), gamma.data(), center.data(), p.lowp_floor > 0.0 && be.lowp_available());
  be.apply_DmA();
  be.gram();
  be.rr_sym();
  be.rotate_and_residual();
  std::vector<double> theta((size_t)M * B), res((size_t)M * B);
  be.get_theta_res(theta.data(), res.data());
  for (int m = 0; m < M; ++m) {
    const double th = theta[(size_t)m * B + B - 1], r = res[(size_t)m * B + B - 1];
    const double est = th + 1.5 * r + 0.01;
    beta_m[m] = (th == th && r == r && th > 1.0 && est < p.beta) ? est : p.beta;
  }
}

template <class BE>
int chfsi_solve(BE& be, const SolveParams& p, MeshResult* out) {
  const int M = be.n_meshes();
  const int B = be.block();
  const bool sym = be.symmetric();
  const int guard = sym ? 2 : 4;  // Ritz values kept between the wanted set and the filter edge
  std::vector<double> theta((size_t)M * B), res((size_t)M * B);
  std::vector<double> a_prev(M, -1.0), last_a(M, 0.5 * p.beta), last_alow(M, 0.0);
  std::vector<int> done(M, 0), n_low(M, B), flags(M), sel((size_t)M * B), n_out(M), last_kp(M, 1);
  std::vector<double> gh, wbuf, thbuf;
  std::vector<int> rr_rc(M, 0);
  bool rr_on_backend = false;
  if (!sym) {
    gh.resize((size_t)2 * M * B * B);
    wbuf.resize((size_t)M * B * B);
    thbuf.resize((size_t)M * B);
  }
  for (int m = 0; m < M; ++m) {
    out[m].status = SOLVE_NOT_CONVERGED;
    out[m].n_out = 0;
    out[m].k_final = p.k0;
    out[m].outer_iters = 0;
    out[m].total_degree = 0;
    out[m].block = B;
    out[m].max_residual = -1.0;
    out[m].lowp_degree = 0;
  }
  std::vector<double> alpha, gamma, center(M), beta_m(M, p.beta);
  // (non-symmetric runs take the fp32 forms too: the correction form is an algebraic identity for any (x, theta), columns
  // whose Ritz value lies above the filter edge -- the carried complex pairs among them -- are normalised at the edge)
  const bool lowp_on = p.lowp_floor > 0.0 && be.lowp_available();
  if (sym && p.probe_degree > 0) {
    probe_upper_bound(be, p, beta_m.data());
    if (lowp_on)
      for (int m = 0; m < M; ++m) out[m].lowp_degree += p.probe_degree;
  }
  for (int m = 0; m < M; ++m) out[m].beta = beta_m[m];
  be.init_block();
  int n_done = 0;
  int rc = SOLVE_OK;
  for (int outer = 0; outer < p.max_outer && n_done < M; ++outer) {
    be.apply_DmA();
    be.gram();
    if (sym) {
      be.rr_sym();
    } else {
      // The meshes' small general eigenproblems: on the backend when it can (CudaBackend: one CTA per mesh,
      // nonsym_small.h, b <= 64 -- nothing leaves the device until the residuals do), else here on the host.
      std::vector<double> cut(M);
      for (int m = 0; m < M; ++m) cut[m] = a_prev[m] < 0.0 ? 0.5 * p.beta : 2.0 * a_prev[m];
      rr_on_backend = be.rr_nonsym_device(cut.data());
      if (!rr_on_backend) {
        be.get_GH(gh.data(), gh.data() + (size_t)M * B * B);
        // independent problems: a small OpenMP team (8 threads at most: one process per GPU times the steps in flight
        // must not oversubscribe the host) -- from 8 meshes up only: for a single pair the team start-up costs more
        // than it saves (Focusr() 38 -> 79 ms measured)
#if defined(_OPENMP)
#pragma omp parallel for schedule(dynamic) if (M >= 8) num_threads(8)
#endif
        for (int m = 0; m < M; ++m)
          rr_rc[m] = rr_nonsym_host(gh.data() + (size_t)m * B * B, gh.data() + (size_t)(M + m) * B * B, B, cut[m],
                                    wbuf.data() + (size_t)m * B * B, thbuf.data() + (size_t)m * B, &n_low[m]);
        be.set_W_theta(wbuf.data(), thbuf.data());
      }
    }
    be.rotate_and_residual();
    be.get_theta_res(theta.data(), res.data());
    if (!sym) {
      if (rr_on_backend) be.get_nonsym_info(rr_rc.data(), n_low.data());
      for (int m = 0; m < M; ++m)
        if (rr_rc[m] < 0 && !done[m]) {  // the QR iteration failed
          out[m].status = SOLVE_BREAKDOWN;
          done[m] = 1;
          ++n_done;
          rc = SOLVE_BREAKDOWN;
        }
    }

    bool any_flag = false;
    for (int m = 0; m < M; ++m) {
      flags[m] = 0;
      if (done[m]) continue;
      const double* th = &theta[(size_t)m * B];
      const double* rs = &res[(size_t)m * B];
      out[m].outer_iters = outer + 1;
      const int n_avail = (sym ? B : n_low[m]) - guard;
      int k_final = p.k0;
      const int kp = retry_contract(th, std::max(n_avail, 0), be.zero_rows(m), p, &k_final);
      if (kp > n_avail || !(th[0] == th[0])) {
        out[m].status = (th[0] == th[0]) ? SOLVE_BLOCK_TOO_SMALL : SOLVE_BREAKDOWN;
        out[m].k_final = k_final;
        done[m] = 1;
        ++n_done;
        rc = std::max(rc, out[m].status);
        continue;
      }
      last_kp[m] = std::max(kp, 1);
      double worst = 0.0;
      for (int j = 0; j < kp; ++j) worst = std::max(worst, rs[j]);
      out[m].max_residual = worst;
      if (worst <= p.tol) {
        int cnt = 0;
        for (int j = 0; j < kp; ++j)
          if (th[j] > p.min_eig) sel[(size_t)m * B + cnt++] = j;
        for (int j = cnt; j < B; ++j) sel[(size_t)m * B + j] = -1;
        out[m].k_final = k_final;
        out[m].n_out = cnt;
        if (cnt > p.ldv) {
          out[m].status = SOLVE_OUTPUT_TOO_SMALL;
          rc = std::max(rc, (int)SOLVE_OUTPUT_TOO_SMALL);
        } else {
          out[m].status = SOLVE_OK;
          n_out[m] = cnt;
          flags[m] = 1;
          any_flag = true;
        }
        done[m] = 1;
        ++n_done;
      }
    }
    if (any_flag) be.finalize(flags.data(), sel.data(), n_out.data());
    if (n_done >= M) break;

    // --- next filter: per-mesh interval, one common degree
    int deg = 0;
    int kind = -1;  // PASS_* common to the batch
    for (int m = 0; m < M; ++m) {
      if (done[m]) continue;
      const double* th = &theta[(size_t)m * B];
      const int top = (sym ? B : n_low[m]) - 1;
      double a = th[top];
      if (outer >= 1 && beta_m[m] < p.beta && a > 0.8 * beta_m[m]) {
        // leak: the filter amplified something above the probed bound; back to the guaranteed one
        beta_m[m] = p.beta;
        out[m].beta = p.beta;
      }
      const double beta = beta_m[m];
      if (!(a < beta)) a = 0.5 * beta;
      if (!(a > 0.0)) a = 1e-3 * beta;
      const int kp = last_kp[m];
      const double thk = th[std::max(0, std::min(kp, top) - 1)];
      if (outer >= 3 && (a - thk) <= 1e-3 * a) {
        // the wanted set touches the filter edge (a multiplet cut by the block): more columns needed
        out[m].status = SOLVE_BLOCK_TOO_SMALL;
        done[m] = 1;
        ++n_done;
        rc = std::max(rc, (int)SOLVE_BLOCK_TOO_SMALL);
        continue;
      }
      last_a[m] = a;
      last_alow[m] = std::min(th[0], 0.0);
      a_prev[m] = a;
      // Amplification of the slowest wanted pair for this pass.  The residual of a Ritz pair drops by
      // about that factor per pass, so once the block has settled (outer >= 1) the pass is sized to
      // land a little below the tolerance in one go instead of the fixed default; the cap keeps the
      // Gram matrix of the filtered (Ritz-rotated, hence graded) block factorisable.
      double amp = p.amp_target;
      const double worst = out[m].max_residual;
      const double cap = sym ? 1e7 : 1e4;
      if (outer >= 1 && worst > 0.0 && worst < 1e-2) {
        const double need = worst / (p.land * p.tol);
        if (need <= cap) amp = std::max(need, 30.0);
      }
      // Precision of the pass (see the header comment).  With fp32 available every pass takes one of its two forms:
      // blocks that are still far from converged are filtered on fp32 blocks, sized to land at lowp_aim; from 1e-3 down
      // the correction form takes over, which cannot gain more than ~1 / LOWP_CORR_NOISE in one pass (its rounding
      // noise is that fraction of the residual it starts from) -- a pass that needs more is followed by a short second one.
      int k = PASS_FP64;
      if (lowp_on && worst > 0.0) {
        if (worst < 1e-3) {
          k = PASS_FP32_CORR;
          const double need = worst / (p.land * p.tol);
          amp = std::max(std::min(need, std::min(cap, 0.5 / LOWP_CORR_NOISE)), 30.0);
        } else if (worst < 1e-1) {
          k = PASS_FP32;
          amp = std::min(std::max(worst / p.lowp_aim, 30.0), cap);
        } else if (worst / amp >= p.lowp_floor) {
          k = PASS_FP32;  // unsettled block: default amplification, far above the fp32 floor
        }
      }
      if (kind < 0) kind = k;
      else if (kind != k)  // mixed batch: the correction form serves every fp32 case, fp64 serves all
        kind = (kind == PASS_FP64 || k == PASS_FP64) ? PASS_FP64 : PASS_FP32_CORR;
      deg = std::max(deg, cheb_degree(a, thk, beta, amp, p.max_degree));
    }
    if (n_done >= M) break;
    if (kind < 0 || deg < 3) kind = PASS_FP64;
    if (kind == PASS_FP32_CORR) {
      // y_k = x + z_k with the polynomial of column j normalised to 1 at theta_j: z_{k+1} = alpha_kj ((L - c) z_k + r_j) -
      // gamma_kj z_{k-1}, z_0 = 0, r = L x - theta x from the Rayleigh-Ritz step (fp64).  The M * B per-column tables
      // (corr_table_column above) are built by the backend from the Ritz values it already holds -- a kernel on the
      // GPU, so that nothing N- or degree-sized is computed or uploaded by the host between two filter passes.
      be.filter_correction(deg, last_a.data(), beta_m.data());
    } else {
      alpha.resize((size_t)M * deg);
      gamma.resize((size_t)M * deg);
      for (int m = 0; m < M; ++m)
        cheb_table(last_a[m], last_alow[m], beta_m[m], deg, &alpha[(size_t)m * deg],
                   &gamma[(size_t)m * deg], &center[m]);
      be.filter(deg, alpha.data(), gamma.data(), center.data(), kind == PASS_FP32);
    }
    for (int m = 0; m < M; ++m)
      if (!done[m]) {
        out[m].total_degree += deg;
        if (kind != PASS_FP64) out[m].lowp_degree += deg;
      }
  }
  for (int m = 0; m < M; ++m)
    if (!done[m]) rc = std::max(rc, (int)SOLVE_NOT_CONVERGED);
  return rc;
}

}  // namespace fb
