// multibox_b200 -- instantiations of the register-resident matching kernel for 1-warp CTAs.
#include "mbx_match_reg.cuh"

namespace mbx {
template int launch_cols<1>(const MatchParams &, int, cudaStream_t);
}  // namespace mbx
