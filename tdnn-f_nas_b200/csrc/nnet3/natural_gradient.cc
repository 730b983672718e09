// OnlineNaturalGradient (kaldi: nnet3/natural-gradient-online.{h,cc}) -- SURVEY.md "next" row N1.
//
// Status: the class carries the configuration the components read/write (rank, alpha,
// num-samples-history, update period 4), so config and model I/O are faithful, but
// PreconditionDirections is the IDENTITY (directions untouched, scale 1): the parameter update is
// the raw-gradient path that BASELINE.md section 3 and the parity tests pin.  A one-time warning says so.
//
// Planned B200 form (so that the R x (n*D_in+1) spliced input never has to be materialised):
// with X' = X - (X W^T) M W and O' = O - (O V^T) N V (rank-r projections), the update
// O'^T X' expands to O^T X minus rank-r corrections built from the skinny products X W^T, O V^T,
// which are epilogue reductions of the same GEMM pre-pass; the r x r eigen-update stays on the host.
#include <atomic>

#include "components.h"

namespace tdnnf {
namespace nnet3 {

static std::atomic<bool> g_warned(false);

BaseFloat OnlineNaturalGradient::PreconditionDirectionsScale() const {
  if (!g_warned.exchange(true))
    KaldiWarn("OnlineNaturalGradient::PreconditionDirections is the identity in this build (SURVEY.md N1): "
              "updates use the un-preconditioned gradient");
  return 1.0;
}

}  // namespace nnet3
}  // namespace tdnnf
