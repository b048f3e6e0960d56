/* vtrace_c.c -- plain-C (float64) restatement of the reference's optimizer/vtrace.py -- TEST INFRASTRUCTURE ONLY.
 *
 * A second, independent CPU restatement next to oracle/vtrace_np.py: same functions, scalar loops instead of array
 * expressions, so that the two can pin each other (tests/test_oracle_vtrace.py) -- PARITY UNPINNED like the rest of the
 * oracle (TensorFlow 1.14 is not installable; the reference ships no tests).  Built by oracle/c/Makefile (gcc) into
 * oracle/_build/libvtrace_oracle.so and loaded with ctypes by oracle/vtrace_c.py.  Only tests/ may use it.
 *
 *   vtrace_from_importance_weights   optimizer/vtrace.py:71-103  (time-major [T, B]; clip_pg_rho_threshold unused, :72)
 *   vtrace_from_softmax              optimizer/vtrace.py:29-69   (batch-major [B, T, A] / [B, T])
 *   vtrace_losses                    optimizer/vtrace.py:105-126 (policy-gradient, baseline, entropy sums)
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* rho = exp(log_rho); rho_bar = min(clip, rho) (clip < 0: no clipping, :75-78); c = min(1, rho) (:80);
 * delta_t = rho_bar_t (r_t + g_t V_{t+1} - V_t) with V_T = bootstrap (:81-84);
 * acc_t = delta_t + g_t c_t acc_{t+1}, acc_T = 0 (:88-100); vs = acc + V (:101). */
void vtrace_from_importance_weights(const double* log_rhos, const double* discounts, const double* rewards,
                                    const double* values, const double* bootstrap, int T, int B, double clip_rho,
                                    double* vs, double* clipped_rhos) {
  for (int b = 0; b < B; ++b) {
    double acc = 0.0;
    for (int t = T - 1; t >= 0; --t) {
      const size_t i = (size_t)t * B + b;
      const double rho = exp(log_rhos[i]);
      const double rho_bar = (clip_rho >= 0.0 && clip_rho < rho) ? clip_rho : rho;
      const double c = rho < 1.0 ? rho : 1.0;
      const double v_next = (t + 1 < T) ? values[i + B] : bootstrap[b];
      const double delta = rho_bar * (rewards[i] + discounts[i] * v_next - values[i]);
      acc = delta + discounts[i] * c * acc;
      vs[i] = acc + values[i];
      clipped_rhos[i] = rho_bar;
    }
  }
}

/* log pi(a) - log mu(a) with the probability selected by one-hot sum (no epsilon, :16-27,:49-51); then the recursion
 * above per trajectory with bootstrap = next_values[b, T-1] (:62). */
void vtrace_from_softmax(const double* behavior, const double* target, const int32_t* actions, const double* discounts,
                         const double* rewards, const double* values, const double* next_values, int B, int T, int A,
                         double clip_rho, double* vs, double* clipped_rhos) {
  for (int b = 0; b < B; ++b) {
    double acc = 0.0;
    const double boot = next_values[(size_t)b * T + (T - 1)];
    for (int t = T - 1; t >= 0; --t) {
      const size_t i = (size_t)b * T + t;
      const int a = actions[i];
      double pt = 0.0, pb = 0.0;                       /* tf.one_hot: out-of-range action -> all-zero row */
      if (a >= 0 && a < A) { pt = target[i * A + a]; pb = behavior[i * A + a]; }
      const double rho = exp(log(pt) - log(pb));
      const double rho_bar = (clip_rho >= 0.0 && clip_rho < rho) ? clip_rho : rho;
      const double c = rho < 1.0 ? rho : 1.0;
      const double v_next = (t + 1 < T) ? values[i + 1] : boot;
      const double delta = rho_bar * (rewards[i] + discounts[i] * v_next - values[i]);
      acc = delta + discounts[i] * c * acc;
      vs[i] = acc + values[i];
      clipped_rhos[i] = rho_bar;
    }
  }
}

/* out[0] = -sum log(pi(a) + 1e-8) adv (:105-112); out[1] = 0.5 sum (vs - V)^2 (:114-118); out[2] = sum pi log pi (:120-126) */
void vtrace_losses(const double* softmax, const int32_t* actions, const double* advantages, const double* vs,
                   const double* value, int B, int T, int A, double* out) {
  double pg = 0.0, bl = 0.0, en = 0.0;
  for (size_t i = 0; i < (size_t)B * T; ++i) {
    const int a = actions[i];
    const double sel = (a >= 0 && a < A) ? softmax[i * A + a] : 0.0;
    pg -= log(sel + 1e-8) * advantages[i];
    const double e = vs[i] - value[i];
    bl += e * e;
    for (int k = 0; k < A; ++k) {
      const double p = softmax[i * A + k];
      en += p * log(p);
    }
  }
  out[0] = pg;
  out[1] = 0.5 * bl;
  out[2] = en;
}
