// Host-side pieces of MlModel that sit next to the path: the bridge from a
// likelihood model to the integer cost matrix of the parsimony kernels.
#include <cmath>

#include "phylo_engine.h"

// MlModel.integerized_model (lib/mlModel.ml:639-660): from P = P(t) (phylo_compose_sym / _gtr),
//   cost[i][j] = -trunc(10^sigma * ln(P[i][j]))               (JC69, K2P: priors == NULL)
//   cost[i][j] = -trunc(10^sigma * ln(priors[i] * P[i][j]))   (every other model)
// `int_of_float` truncates toward zero, as the conversion below does. An entry whose logarithm is
// not finite (P[i][j] <= 0) or whose cost does not fit an int32 is an error; the reference's result
// is unspecified there (int_of_float of -infinity).
extern "C" int phylo_integerize_matrix(const double *P, const double *priors, int n, int sigma, int32_t *cost_out) {
  if (!P || !cost_out || n < 1 || sigma < 0 || sigma > 15) return PHYLO_ERR_ARG;
  const double scale = std::pow(10.0, (double)sigma);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      const double m = priors ? priors[i] * P[(size_t)i * n + j] : P[(size_t)i * n + j];
      const double x = scale * std::log(m);
      if (!std::isfinite(x) || std::fabs(x) >= 2147483647.0) return PHYLO_ERR_NUMERIC;
      cost_out[(size_t)i * n + j] = -(int32_t)x;
    }
  return PHYLO_OK;
}
