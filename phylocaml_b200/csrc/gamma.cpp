// Discrete-Gamma rate classes on the host: the reference builds them with
// Pareto.Distributions.Gamma.quantile (GSL's gsl_cdf_gamma_Pinv; lib/mlModel.ml:93-99, :676-694),
// a dependency that is neither vendored nor version-pinned. This file is self-contained:
// regularised incomplete gamma P(a, x) by series / continued fraction, its inverse by a safeguarded
// Newton iteration. Two definitions are offered because the reference is not consistent with its own
// documentation: mode 0 is what the code does, mode 1 what lib/mlModel.mli:12 says.
#include <cmath>
#include <limits>

#include "phylo_engine.h"

namespace {

// P(a, x) = gamma(a, x) / Gamma(a), a > 0, x >= 0
double gamma_p(double a, double x) {
  if (!(x > 0.0)) return 0.0;
  if (std::isinf(x)) return 1.0;
  const double lg = std::lgamma(a);
  if (x < a + 1.0) {  // series
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 100000; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
    }
    return sum * std::exp(-x + a * std::log(x) - lg);
  }
  // continued fraction for Q(a, x) (modified Lentz)
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 100000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-17) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - lg) * h;
}

// x with P(a, x) = p
double gamma_p_inv(double a, double p) {
  if (!(p > 0.0)) return 0.0;
  if (p >= 1.0) return std::numeric_limits<double>::infinity();
  const double lg = std::lgamma(a);
  // bracket [lo, hi] with P(lo) < p <= P(hi)
  double lo = 0.0, hi = a > 1.0 ? a : 1.0;
  while (gamma_p(a, hi) < p) { lo = hi; hi *= 2.0; }
  // start: small-x expansion P ~ x^a / Gamma(a + 1) when it lands inside the bracket, else the midpoint
  double x = std::exp((std::log(p) + std::lgamma(a + 1.0)) / a);
  if (!(x > lo && x < hi)) x = 0.5 * (lo + hi);
  for (int it = 0; it < 200; ++it) {
    const double f = gamma_p(a, x) - p;
    if (f > 0.0) hi = x; else lo = x;
    const double dens = std::exp(-x + (a - 1.0) * std::log(x) - lg);  // dP/dx
    double nx = (dens > 0.0 && std::isfinite(dens)) ? x - f / dens : -1.0;
    if (!(nx > lo && nx < hi)) nx = 0.5 * (lo + hi);  // Newton left the bracket: bisect
    if (std::fabs(nx - x) <= 1e-16 * std::fabs(x) || hi - lo <= 1e-16 * hi) { x = nx; break; }
    x = nx;
  }
  return x;
}

}  // namespace

extern "C" int phylo_gamma_rates(double alpha, int k, int mode, double *rates, double *probs) {
  if (!(alpha > 0.0) || !std::isfinite(alpha) || k < 1 || !rates || (mode != 0 && mode != 1)) return PHYLO_ERR_ARG;
  if (probs)
    for (int i = 0; i < k; ++i) probs[i] = 1.0 / k;
  if (mode == 0) {
    // lib/mlModel.ml:93-99 called as `gamma_rates y y x` (:679): quantile at p = i/k, i = 0..k-1, of
    // Gamma(shape = alpha, scale = alpha); rates[0] = 0
    for (int i = 0; i < k; ++i) rates[i] = alpha * gamma_p_inv(alpha, (double)i / k);
    return PHYLO_OK;
  }
  // Yang 1994: mean of Gamma(shape = alpha, rate = alpha) within each of k equal-probability
  // classes; (1/k) sum rates = 1. With cut points c_i (quantiles at i/k):
  // mean_i = k (P(alpha + 1, alpha c_{i+1}) - P(alpha + 1, alpha c_i)), alpha c_i = P^-1(alpha, i/k)
  double prev = 0.0;
  for (int i = 0; i < k; ++i) {
    const double up = (i + 1 == k) ? 1.0 : gamma_p(alpha + 1.0, gamma_p_inv(alpha, (double)(i + 1) / k));
    rates[i] = (up - prev) * k;
    prev = up;
  }
  return PHYLO_OK;
}
