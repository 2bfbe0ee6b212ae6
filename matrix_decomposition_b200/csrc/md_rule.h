// md_rule.h -- the per-column scalar rule of the modified LDL^T, FP64, host + device.
//
// Restates matrix/approximation/positive_semidefinite/Reimer.py:42-176 (_minimal_change),
// :25-39 (_difference_frobenius_norm) and matrix/_util/roots.py:8-68 (solver_depressed_cubic, real
// roots) in double precision.  The reference evaluates these in x87 long double; B200 has no such
// type, so decisions can differ from the reference only when gamma - alpha is within rounding of a
// bound (documented near-tie classes (ii)/(iii) in DESIGN.md).
//
// The same header is compiled by nvcc into the kernels and by g++ into the host test shim
// (tests/host_rule_shim.cpp), so the rule is unit-tested on CPU against the reference's golden vectors.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define MD_HD __host__ __device__ __forceinline__
#else
#define MD_HD inline
#endif

#define MD_RULE_REIMER 0  // approximate: clamp d / shrink row by omega (Reimer.py)
#define MD_RULE_EXACT 1   // decompose: d = gamma - alpha, omega = 1, fail when d <= 0 (dense/calculate.py:108-119)

struct MdRuleParams {
  int rule;                // MD_RULE_*
  double min_diag_D;       // NaN = "None": per-index default of Reimer.py:488-491
  double max_diag_D;       // +inf = none
  double min_abs_value_D;  // Reimer.py:398-402 (default sqrt(eps(float128)) rounded to double)
  double min_abs_value_L;  // Reimer.py:404-408
  double min_diag_B;       // scalar bounds, used when the per-index vectors are null
  double max_diag_B;
};

struct MdDecision {
  double d;
  double omega;
  double f;
  int clamped;  // 1 when the rule left the unclamped branch (Reimer.py:68)
};

MD_HD double md_pymax(double a, double b) { return (b > a) ? b : a; }  // Python max(a, b)
MD_HD double md_pymin(double a, double b) { return (b < a) ? b : a; }  // Python min(a, b)

// Real roots of x^3 + p x + q (roots.py:8-68).  Returns the count; order as in the reference.
//
// The reference evaluates D = (p/3)^3 + (q/2)^2 in long double, picks the branch by its sign, and then
// pushes D through a DOUBLE (math.sqrt / cmath.sqrt, roots.py:15,36).  In the explosive regime
// (alpha ~ 1e100+) that double underflows to a denormal or to zero and the reference's roots carry the
// artifact (e.g. u = v = cbrt(-q/2) once sqrt(D) == 0).  The artifact decides how rows are shrunk, so
// it is part of the behaviour to reproduce: magnitudes below use D as an FP64 value (same underflow),
// the branch uses the true sign (recovered from rescaled coefficients when the FP64 D is exactly zero),
// and the D < 0 roots are written as the reference forms them, 2 |u|^(1/3) cos(arg(u)/3 + 2 pi k/3)
// with u = -q/2 + i sqrt(-D) and the exponent 1/3 rounded to double.
MD_HD int md_cubic_real_roots(double p, double q, double* roots) {
  const double P = p / 3.0, Q = q / 2.0;
  const double D = P * P * P + Q * Q;  // roots.py:10, FP64 range
  if (!isfinite(D)) {
    // FP64 overflow (the reference's math.sqrt would see inf and produce NaN roots): solve the
    // rescaled cubic x = s y instead, with the cancellation-free pairing.
    const double scale = md_pymax(sqrt(fabs(p)), cbrt(fabs(q)));
    const double ps = p / (scale * scale), qs = (q / scale) / (scale * scale);
    const double Ps = ps / 3.0, Qs = qs / 2.0;
    const double Ds = Ps * Ps * Ps + Qs * Qs;
    if (Ds > 0) {
      const double s = sqrt(Ds);
      const double u = cbrt(-Qs - copysign(s, Qs));
      const double v = (u != 0) ? -Ps / u : 0.0;
      // u + v = (u^3 + v^3) / (u^2 - u v + v^2) = -2Q / (u^2 + P + v^2): no cancellation when P > 0
      roots[0] = ((Ps > 0) ? -2.0 * Qs / (u * u + Ps + v * v) : (u + v)) * scale;
      return 1;
    } else if (Ds < 0) {
      const double m = 2.0 * sqrt(-Ps), theta = atan2(sqrt(-Ds), -Qs);
      const double two_pi_3 = 2.0943951023931954923;
      roots[0] = m * cos(theta / 3.0) * scale;
      roots[1] = m * cos(theta / 3.0 + two_pi_3) * scale;
      roots[2] = m * cos(theta / 3.0 - two_pi_3) * scale;
      return 3;
    }
    roots[0] = 3 * qs / ps * scale;
    roots[1] = -0.5 * roots[0];
    return 2;
  }
  int sgn = (D > 0) - (D < 0);
  if (sgn == 0 && (p != 0 || q != 0)) {
    // P^3 and Q^2 both underflowed (or cancelled exactly): recover the sign of the true D
    const double scale = md_pymax(sqrt(fabs(p)), cbrt(fabs(q)));
    const double Ps = p / (scale * scale) / 3.0, Qs = (q / scale) / (scale * scale) / 2.0;
    const double Ds = Ps * Ps * Ps + Qs * Qs;
    sgn = (Ds > 0) - (Ds < 0);
  }
  if (sgn > 0) {  // roots.py:13-31
    // the reference's own (cancellation-prone) form: mirroring it tracks the reference's roots to
    // ~1e-12, the mathematically stable form only to ~1e-9 (the reference itself is that inexact)
    const double s = sqrt(D);
    const double u = cbrt(-Q + s);
    const double v = cbrt(-Q - s);
    roots[0] = u + v;
    return 1;
  } else if (sgn < 0) {  // roots.py:34-51
    const double s = sqrt(-D);
    const double third = 1.0 / 3.0;
    const double m = 2.0 * pow(hypot(Q, s), third);
    const double theta = atan2(s, -Q) * third;
    // cos(theta +- 2 pi / 3) from one sin / cos pair
    const double ct = cos(theta), st = sin(theta);
    const double h3 = 0.86602540378443864676;  // sqrt(3) / 2
    roots[0] = m * ct;
    roots[1] = m * (-0.5 * ct - h3 * st);
    roots[2] = m * (-0.5 * ct + h3 * st);
    return 3;
  } else {  // roots.py:54-64
    if (p == 0) { roots[0] = 0; return 1; }
    roots[0] = 3 * q / p;
    roots[1] = -0.5 * roots[0];
    return 2;
  }
}

// Reimer.py:25-39
MD_HD double md_difference_frobenius_norm(double d, double omega, double alpha, double beta, double gamma) {
  if (omega == 0) return (d - gamma) * (d - gamma) + beta;
  if (omega == 1) return (d + alpha - gamma) * (d + alpha - gamma);
  double t = d + omega * omega * alpha - gamma;
  return t * t + (omega - 1) * (omega - 1) * beta;
}

// Candidate order of the reference: min over the list by the key (f, -d, omega), Reimer.py:156-160.
MD_HD bool md_candidate_better(double cf, double cd, double cw, double bf, double bd, double bw) {
  if (cf != bf) return cf < bf;
  if (cd != bd) return cd > bd;
  return cw < bw;
}

// Best candidate of a subset of the candidate list of Reimer.py:42-176 (clamped branch only):
//   part bit 0: (a, 1) / (b, 1) (:82-88) and the cubic roots for d = lo (:95-96, :103-135)
//   part bit 1: the cubic roots for d = max_diag_D (:97-98), the omega = 0 fallback (:146-148), (0, 0) (:151-152)
// have = 0 when the subset is empty.  min over the union of both parts = the reference's choice (the key is a
// total order when no f is NaN), so the two cubic solves can run in two threads and be merged afterwards.
struct MdBest {
  double d, w, f;
  int have;
};

MD_HD MdBest md_merge_best(const MdBest& x, const MdBest& y) {  // y is "later in the list" than x
  if (!x.have) return y;
  if (!y.have) return x;
  return md_candidate_better(y.f, y.d, y.w, x.f, x.d, x.w) ? y : x;
}

// The candidates that one value dd of d contributes (Reimer.py:103-135): the real roots of the cubic in omega,
// clipped to [omega_lower, omega_upper].  *fallback is raised when p or q is not finite (:105-110).  Branch-free
// in its arguments apart from the cubic itself, so that two lanes of a warp can evaluate the two values of dd
// (lo and max_diag_D) side by side without diverging.
MD_HD MdBest md_candidates_for_d(bool enabled, double dd, double alpha, double beta, double gamma, double min_diag_B,
                                 double max_diag_B, int* fallback) {
  double best_d = 0, best_w = 0, best_f = 0;
  int have = 0;
#define MD_CONSIDER(dd_, ww_)                                                        \
  do {                                                                               \
    const double cd_ = (dd_), cw_ = (ww_);                                           \
    const double cf_ = md_difference_frobenius_norm(cd_, cw_, alpha, beta, gamma);   \
    if (!have || md_candidate_better(cf_, cd_, cw_, best_f, best_d, best_w)) {       \
      best_d = cd_; best_w = cw_; best_f = cf_; have = 1;                            \
    }                                                                                \
  } while (0)
  if (enabled) {
    const double q = (-0.5 * beta / alpha) / alpha;  // :103
    const double p = (dd - gamma) / alpha - q;       // :104
    if (!isfinite(q) || !isfinite(p)) {              // :105-110
      *fallback = 1;
    } else {
      double roots[3];
      const int nr = md_cubic_real_roots(p, q, roots);
      const double omega_lower = sqrt(md_pymax(min_diag_B - dd, 0.0) / alpha);        // :118
      const double omega_upper = md_pymin(sqrt((max_diag_B - dd) / alpha), 1.0);      // :120
      int add_lower = 0, add_upper = 0;
      for (int k = 0; k < nr; ++k) {  // :125-131
        const double w = roots[k];
        if (w <= omega_lower) add_lower = 1;
        else if (w >= omega_upper) add_upper = 1;
        else MD_CONSIDER(dd, w);
      }
      if (add_lower) MD_CONSIDER(dd, omega_lower);
      if (add_upper) MD_CONSIDER(dd, omega_upper);
    }
  }
#undef MD_CONSIDER
  MdBest r;
  r.d = best_d; r.w = best_w; r.f = best_f; r.have = have;
  return r;
}

// One candidate (d, omega) as an MdBest
MD_HD MdBest md_single_candidate(double d, double omega, double alpha, double beta, double gamma) {
  MdBest r;
  r.d = d; r.w = omega; r.f = md_difference_frobenius_norm(d, omega, alpha, beta, gamma); r.have = 1;
  return r;
}

// The clamped branch of Reimer.py:42-176 assembled from its pieces, in the reference's candidate order:
//   (a, 1) / (b, 1) :82-88;  roots for d = lo (`from_lo`) :95-96;  roots for d = max_diag_D (`from_max`) :97-98;
//   the omega = 0 fallback :146-148;  (0, 0) :151-152.
// `from_lo` / `from_max` are md_candidates_for_d results (have = 0 when that d does not apply), `fallback` the OR of
// their flags.  The kernels evaluate the two md_candidates_for_d calls in two lanes and call this in one.
MD_HD MdBest md_assemble_minimal_change(const MdBest& from_lo, const MdBest& from_max, int fallback, double alpha,
                                        double beta, double gamma, double min_diag_D, double max_diag_D,
                                        double min_diag_B, double max_diag_B, double min_abs_value_D) {
  const double lo = md_pymax(min_diag_D, min_abs_value_D);  // :64
  const double a = md_pymax(lo, min_diag_B - alpha);        // :65
  const double b = md_pymin(max_diag_D, max_diag_B - alpha);  // :66
  const double dstar = gamma - alpha;                       // :67
  MdBest best;
  best.d = best.w = best.f = 0;
  best.have = 0;
  if (isfinite(alpha)) {
    if (a <= b) {  // :82-88
      if (dstar < a) best = md_single_candidate(a, 1.0, alpha, beta, gamma);
      else if (dstar > b) best = md_single_candidate(b, 1.0, alpha, beta, gamma);
    }
    if (isfinite(beta)) {
      best = md_merge_best(best, from_lo);
      best = md_merge_best(best, from_max);
    } else {
      fallback = 1;  // :136-139
    }
  } else {
    fallback = 1;  // :140-143
  }
  if (fallback) {  // :146-148
    const double dd = md_pymax(md_pymax(lo, min_diag_B), md_pymin(md_pymin(gamma, max_diag_D), max_diag_B));
    best = md_merge_best(best, md_single_candidate(dd, 0.0, alpha, beta, gamma));
  }
  if (min_diag_D == 0 && min_diag_B <= 0 && 2 * gamma <= min_abs_value_D)  // :151-152
    best = md_merge_best(best, md_single_candidate(0.0, 0.0, alpha, beta, gamma));
  return best;
}

// Which values of d contribute cubic candidates (Reimer.py:89-98): index 0 = lo, 1 = max_diag_D.
MD_HD bool md_d_candidate(int which, double alpha, double beta, double min_diag_D, double max_diag_D, double min_diag_B,
                          double max_diag_B, double min_abs_value_D, double* dd) {
  const double lo = md_pymax(min_diag_D, min_abs_value_D);
  const bool base = isfinite(alpha) && isfinite(beta) && alpha != 0;
  if (which == 0) {
    *dd = lo;
    return base && (min_diag_B - alpha <= lo);                                 // :95-96
  }
  *dd = max_diag_D;
  return base && isfinite(max_diag_D) && max_diag_D <= max_diag_B;            // :97-98
}

MD_HD MdBest md_minimal_change_part(int part, double alpha, double beta, double gamma, double min_diag_D,
                                    double max_diag_D, double min_diag_B, double max_diag_B, double min_abs_value_D) {
  // kept for the host tests: the subset `part` of the candidate list (bit 0: (a|b, 1) and d = lo; bit 1: the rest)
  MdBest none;
  none.d = none.w = none.f = 0;
  none.have = 0;
  int fb0 = 0, fb1 = 0;
  double dd0, dd1;
  const bool e0 = md_d_candidate(0, alpha, beta, min_diag_D, max_diag_D, min_diag_B, max_diag_B, min_abs_value_D, &dd0);
  const bool e1 = md_d_candidate(1, alpha, beta, min_diag_D, max_diag_D, min_diag_B, max_diag_B, min_abs_value_D, &dd1);
  const MdBest c0 = (part & 1) ? md_candidates_for_d(e0, dd0, alpha, beta, gamma, min_diag_B, max_diag_B, &fb0) : none;
  const MdBest c1 = (part & 2) ? md_candidates_for_d(e1, dd1, alpha, beta, gamma, min_diag_B, max_diag_B, &fb1) : none;
  if (part == 3)
    return md_assemble_minimal_change(c0, c1, fb0 | fb1, alpha, beta, gamma, min_diag_D, max_diag_D, min_diag_B,
                                      max_diag_B, min_abs_value_D);
  // partial lists: the fixed candidates belong to bit 0 ((a|b, 1)) and bit 1 (fallback, (0, 0)) respectively
  const double lo = md_pymax(min_diag_D, min_abs_value_D);
  const double a = md_pymax(lo, min_diag_B - alpha);
  const double b = md_pymin(max_diag_D, max_diag_B - alpha);
  const double dstar = gamma - alpha;
  MdBest best = none;
  int fallback = !isfinite(alpha) || !isfinite(beta);
  if (part & 1) {
    if (isfinite(alpha) && a <= b) {
      if (dstar < a) best = md_single_candidate(a, 1.0, alpha, beta, gamma);
      else if (dstar > b) best = md_single_candidate(b, 1.0, alpha, beta, gamma);
    }
    if (isfinite(alpha) && isfinite(beta)) best = md_merge_best(best, c0);
    fallback |= fb0;
  }
  if (part & 2) {
    if (isfinite(alpha) && isfinite(beta)) best = md_merge_best(best, c1);
    fallback |= fb1;
  }
  if (fallback) {
    const double dd = md_pymax(md_pymax(lo, min_diag_B), md_pymin(md_pymin(gamma, max_diag_D), max_diag_B));
    best = md_merge_best(best, md_single_candidate(dd, 0.0, alpha, beta, gamma));
  }
  if ((part & 2) && min_diag_D == 0 && min_diag_B <= 0 && 2 * gamma <= min_abs_value_D)
    best = md_merge_best(best, md_single_candidate(0.0, 0.0, alpha, beta, gamma));
  return best;
}

// Upper clamp without the two cubic solves.  When the unclamped value d* = gamma - alpha lies ABOVE the admissible
// interval [a, b] and b is max_diag_D itself, the candidate (b, 1) of Reimer.py:87-88 wins the minimisation of
// :156-160 whatever the cubic roots are, so they need not be computed:
//   every other candidate of the list is (dd, omega) with dd in {lo, max_diag_D} and omega in [0, 1] (the roots are
//   clipped to [omega_lower, omega_upper], omega_lower <= 1 whenever dd = lo is admitted (:95) and = 0 for
//   dd = max_diag_D because min_diag_B <= max_diag_D);  for those, with alpha > 0,
//       g = dd + omega^2 alpha - gamma <= dd + alpha - gamma <= b + alpha - gamma = -(d* - b) < 0,
//   hence f = g^2 + (omega - 1)^2 beta >= (b + alpha - gamma)^2 = f(b, 1), with equality only for dd = b and omega
//   within rounding of 1 -- the same point.  For dd = lo < b the gap is (d* - lo)^2 - (d* - b)^2 > 0; the shortcut
//   asks for lo == b or a gap far above any rounding of the reference's float128 evaluation.  It does not apply when
//   the omega = 0 fallback (:136-148) or the (0, 0) candidate (:151-152) are in play, or when a > b.
// Returns true and fills (d, omega, f) exactly as md_single_candidate(b, 1) would.  This covers every clamp of the
// benign regime (family F2 with s > 1: "all clamps hit the upper bound") at the cost of a few comparisons.
MD_HD bool md_upper_clamp_shortcut(double alpha, double beta, double gamma, double min_diag_D, double max_diag_D,
                                   double min_diag_B, double max_diag_B, double min_abs_value_D, MdBest* out) {
  const double lo = md_pymax(min_diag_D, min_abs_value_D);
  const double a = md_pymax(lo, min_diag_B - alpha);
  const double b = md_pymin(max_diag_D, max_diag_B - alpha);
  const double dstar = gamma - alpha;
  if (!(isfinite(alpha) && isfinite(beta) && isfinite(gamma) && alpha > 0 && beta >= 0)) return false;
  if (!(a <= b && dstar > b && b == max_diag_D)) return false;
  if (!(lo == b || (b - lo) > 1e-9 * (fabs(alpha) + fabs(gamma) + fabs(b)))) return false;
  if (min_diag_D == 0 && min_diag_B <= 0 && 2 * gamma <= min_abs_value_D) return false;   // (0, 0) is a candidate
  // the fallback candidate appears when p or q of either cubic is not finite (:105-110)
  const double q = (-0.5 * beta / alpha) / alpha;
  if (!isfinite(q) || !isfinite((lo - gamma) / alpha - q) || !isfinite((max_diag_D - gamma) / alpha - q)) return false;
  *out = md_single_candidate(b, 1.0, alpha, beta, gamma);
  return true;
}

// The same test without its three divisions, for the dependency chain of the diagonal-block kernel: a SUFFICIENT
// condition (true only where md_upper_clamp_shortcut is true; the result is then (b, 1) with b = max_diag_D).  The
// divisions of the full test only detect a non-finite p or q of the cubics; with alpha in [1e-100, 1e100] and beta,
// |gamma|, lo, max_diag_D <= 1e100 they are finite: |q| <= 0.5e100 / 1e-200, |(dd - gamma) / alpha| <= 2e100 / 1e-100.
// Columns that fail this cheap test go through md_decide (which tries the full shortcut first).
MD_HD bool md_upper_clamp_shortcut_cheap(double alpha, double beta, double gamma, double min_diag_D, double max_diag_D,
                                         double min_diag_B, double max_diag_B, double min_abs_value_D) {
  const double lo = md_pymax(min_diag_D, min_abs_value_D);
  const double a = md_pymax(lo, min_diag_B - alpha);
  const double b = md_pymin(max_diag_D, max_diag_B - alpha);
  const double dstar = gamma - alpha;
  const double R = 1e100;
  const bool range = alpha >= 1e-100 && alpha <= R && beta >= 0 && beta <= R && fabs(gamma) <= R && fabs(lo) <= R &&
                     fabs(max_diag_D) <= R;
  const bool geom = a <= b && dstar > b && b == max_diag_D;
  const bool gap = lo == b || (b - lo) > 1e-9 * (fabs(alpha) + fabs(gamma) + fabs(b));
  const bool zero_cand = min_diag_D == 0 && min_diag_B <= 0 && 2 * gamma <= min_abs_value_D;
  return range && geom && gap && !zero_cand;
}

// The cheap test once more, with everything that does not depend on alpha / beta folded into per-row constants (the
// diagonal-block kernel evaluates it after every column on its dependency chain): a SUFFICIENT condition for
// md_upper_clamp_shortcut_cheap, hence for md_upper_clamp_shortcut.  The gap test (b - lo) > 1e-9 (|alpha| + |gamma| + |b|)
// with b = max_diag_D becomes alpha <= alpha_hi with a factor 2 of margin, so rounding cannot make it fire where the
// original does not.
struct MdRowConst {
  double lo;        // max(min_diag_D, min_abs_value_D), Reimer.py:64
  double alpha_hi;  // the shortcut needs 1e-100 <= alpha <= alpha_hi
  int ok;           // the alpha-independent conditions hold
};

MD_HD MdRowConst md_row_const(double gamma, double min_diag_D, double max_diag_D, double min_diag_B,
                              double min_abs_value_D) {
  MdRowConst c;
  const double R = 1e100;
  c.lo = md_pymax(min_diag_D, min_abs_value_D);
  const bool zero_cand = min_diag_D == 0 && min_diag_B <= 0 && 2 * gamma <= min_abs_value_D;
  c.ok = (fabs(gamma) <= R && fabs(c.lo) <= R && fabs(max_diag_D) <= R && !zero_cand) ? 1 : 0;
  c.alpha_hi = R;
  if (c.lo != max_diag_D) {
    const double t = 0.5 * ((max_diag_D - c.lo) / 1e-9) - fabs(gamma) - fabs(max_diag_D);
    c.alpha_hi = (t < R) ? t : R;   // NaN -> stays NaN -> every comparison below is false
  }
  return c;
}

// t2 = max_diag_B - alpha and a_lim = max(lo, min_diag_B - alpha) are passed in because the caller needs them anyway
MD_HD bool md_upper_clamp_shortcut_row(const MdRowConst& c, double alpha, double beta, double dstar, double a_lim,
                                       double t2, double max_diag_D) {
  return c.ok && alpha >= 1e-100 && alpha <= c.alpha_hi && beta >= 0 && beta <= 1e100 && t2 >= max_diag_D && a_lim <= max_diag_D &&
         dstar > max_diag_D;
}

// Reimer.py:68-72: the unclamped branch
MD_HD bool md_unclamped(double alpha, double gamma, double min_diag_D, double max_diag_D, double min_diag_B,
                        double max_diag_B, double min_abs_value_D) {
  const double lo = md_pymax(min_diag_D, min_abs_value_D);
  const double a = md_pymax(lo, min_diag_B - alpha);
  const double b = md_pymin(max_diag_D, max_diag_B - alpha);
  const double dstar = gamma - alpha;
  return a <= dstar && dstar <= b;
}

MD_HD MdDecision md_decision_from_best(const MdBest& m) {
  MdDecision r;
  if (m.have) { r.d = m.d; r.omega = m.w; r.f = m.f; }
  else { r.d = NAN; r.omega = NAN; r.f = NAN; }
  r.clamped = 1;
  return r;
}

// Reimer.py:42-176
MD_HD MdDecision md_minimal_change(double alpha, double beta, double gamma, double min_diag_D, double max_diag_D,
                                   double min_diag_B, double max_diag_B, double min_abs_value_D) {
  if (md_unclamped(alpha, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, min_abs_value_D)) {  // :68-72
    MdDecision r;
    r.d = gamma - alpha; r.omega = 1; r.f = 0; r.clamped = 0;
    return r;
  }
  MdBest quick;
  if (md_upper_clamp_shortcut(alpha, beta, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, min_abs_value_D, &quick))
    return md_decision_from_best(quick);
  return md_decision_from_best(
      md_minimal_change_part(3, alpha, beta, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, min_abs_value_D));
}

// The same without the shortcut (the reference's full candidate list): kept for the tests that prove the shortcut.
MD_HD MdDecision md_minimal_change_full(double alpha, double beta, double gamma, double min_diag_D, double max_diag_D,
                                        double min_diag_B, double max_diag_B, double min_abs_value_D) {
  if (md_unclamped(alpha, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, min_abs_value_D)) {
    MdDecision r;
    r.d = gamma - alpha; r.omega = 1; r.f = 0; r.clamped = 0;
    return r;
  }
  return md_decision_from_best(
      md_minimal_change_part(3, alpha, beta, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, min_abs_value_D));
}

// effective min_diag_D of Reimer.py:488-491 (NaN = "None": per-index default)
MD_HD double md_effective_min_diag_D(const MdRuleParams& prm, double gamma, double min_diag_B_i, double max_diag_B_i) {
  double mD = prm.min_diag_D;
  if (isnan(mD)) mD = md_pymin(0.5 * md_pymin(md_pymax(gamma, min_diag_B_i), max_diag_B_i), prm.max_diag_D);
  return mD;
}

// Decision for one column given its accumulated (alpha, beta, gamma) and per-index B bounds.
// For MD_RULE_EXACT, clamped = 1 flags "not positive definite at this pivot" (d <= 0 or NaN).
MD_HD MdDecision md_decide(const MdRuleParams& prm, double alpha, double beta, double gamma, double min_diag_B_i,
                           double max_diag_B_i) {
  if (prm.rule == MD_RULE_EXACT) {
    MdDecision r;
    r.d = gamma - alpha; r.omega = 1; r.f = 0;
    r.clamped = !(r.d > 0);
    return r;
  }
  const double mD = md_effective_min_diag_D(prm, gamma, min_diag_B_i, max_diag_B_i);  // Reimer.py:488-491
  return md_minimal_change(alpha, beta, gamma, mD, prm.max_diag_D, min_diag_B_i, max_diag_B_i, prm.min_abs_value_D);
}
