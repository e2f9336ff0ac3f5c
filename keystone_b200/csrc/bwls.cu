// BlockWeightedLeastSquaresEstimator on the device (K/nodes/learning/BlockWeightedLeastSquares.scala:102-321).
#include "engine.h"

namespace ks {

int64_t fit_bwls(Ctx& c, FeatSrc& src, Matrix& Y, int bs, int num_iter, double lam, double w, int64_t nf_opt) {
  (void)c; (void)src; (void)Y; (void)bs; (void)num_iter; (void)lam; (void)w; (void)nf_opt;
  throw KsError{KS_ERR_INVALID, "ks_blockwls_fit: not implemented yet"};
}

}  // namespace ks
