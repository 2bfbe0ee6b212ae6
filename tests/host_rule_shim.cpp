// Test-only host build of csrc/md_rule.h (g++), so the FP64 scalar rule that the kernels use can be
// checked on CPU against the reference's golden vectors and driven by the numpy model of the blocked
// algorithm (oracle/blocked_model.py).  Not part of the product library.
#include "../matrix_decomposition_b200/csrc/md_rule.h"

extern "C" {
int mdh_cubic(double p, double q, double* roots) { return md_cubic_real_roots(p, q, roots); }
int mdh_minimal_change(double alpha, double beta, double gamma, double min_diag_D, double max_diag_D, double min_diag_B,
                       double max_diag_B, double mav, double* out3) {
  MdDecision r = md_minimal_change(alpha, beta, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, mav);
  out3[0] = r.d; out3[1] = r.omega; out3[2] = r.f;
  return r.clamped;
}
// the full candidate list (no upper-clamp shortcut)
int mdh_minimal_change_full(double alpha, double beta, double gamma, double min_diag_D, double max_diag_D, double min_diag_B,
                            double max_diag_B, double mav, double* out3) {
  MdDecision r = md_minimal_change_full(alpha, beta, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, mav);
  out3[0] = r.d; out3[1] = r.omega; out3[2] = r.f;
  return r.clamped;
}
int mdh_upper_clamp_shortcut(double alpha, double beta, double gamma, double min_diag_D, double max_diag_D, double min_diag_B,
                             double max_diag_B, double mav, double* out3) {
  MdBest b;
  const bool hit = md_upper_clamp_shortcut(alpha, beta, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, mav, &b);
  if (hit) { out3[0] = b.d; out3[1] = b.w; out3[2] = b.f; }
  return hit ? 1 : 0;
}
// the division-free sufficient condition used on the dependency chain of k_diag_block
int mdh_upper_clamp_shortcut_cheap(double alpha, double beta, double gamma, double min_diag_D, double max_diag_D,
                                   double min_diag_B, double max_diag_B, double mav) {
  return md_upper_clamp_shortcut_cheap(alpha, beta, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, mav) ? 1 : 0;
}
// ... and its per-row precomputed form (what the chain of k_diag_block evaluates after every column)
int mdh_upper_clamp_shortcut_row(double alpha, double beta, double gamma, double min_diag_D, double max_diag_D,
                                 double min_diag_B, double max_diag_B, double mav) {
  const MdRowConst c = md_row_const(gamma, min_diag_D, max_diag_D, min_diag_B, mav);
  const double a_lim = md_pymax(c.lo, min_diag_B - alpha);
  return md_upper_clamp_shortcut_row(c, alpha, beta, gamma - alpha, a_lim, max_diag_B - alpha, max_diag_D) ? 1 : 0;
}
// the two-thread evaluation of the kernels: both halves of the candidate list, merged
int mdh_minimal_change_split(double alpha, double beta, double gamma, double min_diag_D, double max_diag_D,
                             double min_diag_B, double max_diag_B, double mav, double* out3) {
  if (md_unclamped(alpha, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, mav)) {
    out3[0] = gamma - alpha; out3[1] = 1; out3[2] = 0;
    return 0;
  }
  const MdBest x = md_minimal_change_part(1, alpha, beta, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, mav);
  const MdBest y = md_minimal_change_part(2, alpha, beta, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, mav);
  MdDecision r = md_decision_from_best(md_merge_best(x, y));
  out3[0] = r.d; out3[1] = r.omega; out3[2] = r.f;
  return r.clamped;
}
int mdh_decide(int rule, double min_diag_D, double max_diag_D, double mavD, double alpha, double beta, double gamma,
               double mB, double MB, double* out3) {
  MdRuleParams prm;
  prm.rule = rule; prm.min_diag_D = min_diag_D; prm.max_diag_D = max_diag_D; prm.min_abs_value_D = mavD;
  prm.min_abs_value_L = 0; prm.min_diag_B = mB; prm.max_diag_B = MB;
  MdDecision r = md_decide(prm, alpha, beta, gamma, mB, MB);
  out3[0] = r.d; out3[1] = r.omega; out3[2] = r.f;
  return r.clamped;
}
}
